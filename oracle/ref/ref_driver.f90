! ORACLE - TEST INFRASTRUCTURE ONLY.
! Runs the UNMODIFIED reference (module_mp_thompson09n.f90 "M:", mphys_thompson09n.f90 "I:") on inputs written by
! oracle/ref/ref.py and dumps what it computes, so that the C++ restatement (oracle/thompson_oracle.cpp) and the CUDA
! path can be pinned to the reference itself.  Usage:
!     ref_driver <input file> <output file>
! Input (stream, native endianness): int32 mode, nx, nz, iiwarm, l_sediment; real32 dt, set_Nc; then
!   mode 1 "columns"  : qv qc qi qr qs qg ni nr t p, each (nz, nx), then dz (nz).  For each column: call mp_thompson with
!                       nc1d, nwfa1d, nifa1d as mp_gt_driver sets them (M:957-964; decision U1 of SURVEY.md section 8c) and w1d = 0.
!   mode 2 "interface": theta dtheta_adv dtheta_div exner qv dqv_adv dqv_div (nz, nx), dz (nz), then for the seven prognostic
!                       moments qc qr nr qi ni qs qg: value, advective and divergence tendency (nz, nx) each.  One call of
!                       mphys_thompson09_interfacen (I:28).
!   mode 3 "aerosol"  : mode 1 with is_aerosol_aware = .true. (M:28 is a module VARIABLE: set before thompson_init): after p, the
!                       planes nc nwfa nifa w (nz, nx) are read, passed as nc1d nwfa1d nifa1d w1d, and nc nwfa nifa are written
!                       after ppt.
! Output: mode 1: the nine fields after the step, ppt (4, nx) rain ice snow graupel.  mode 2: dtheta_mphys, dqv_mphys, the
!   seven moment tendencies.  Both: 64 table words = sum and a strided checksum of every lookup table (real64), and the
!   number of save_dg calls.
program ref_driver
  use module_mp_thompson09n
  use mphys_thompson09n, only : mphys_thompson09_interfacen
  use namelists, only : iiwarm, set_Nc
  use switches, only : l_sediment, l_reuse_thompson_lookup
  use parameters, only : nx, nz, dt
  use column_variables
  use diagnostics, only : n_save_dg
  implicit none
  character(512) :: fin, fout
  integer :: mode, inx, inz, iwarm, ised, i, k, m, ih, im
  real :: rdt, rnc
  real, allocatable :: f(:, :, :), p(:, :), dzq(:), ppt(:, :), plane(:, :)
  real, allocatable :: nc1d(:), nwfa1d(:), nifa1d(:), w1d(:), col(:, :), ae(:, :, :)
  real :: rho, pptrain, pptsnow, pptgraul, pptice
  real(8) :: tw(64)
  integer, parameter :: mh(7) = (/1, 2, 2, 3, 3, 4, 5/), mm(7) = (/1, 1, 2, 1, 2, 1, 1/)

  call get_command_argument(1, fin)
  call get_command_argument(2, fout)
  open(21, file=trim(fin), access='stream', form='unformatted', status='old')
  read(21) mode, inx, inz, iwarm, ised
  read(21) rdt, rnc
  nx = inx; nz = inz; dt = rdt
  iiwarm = iwarm /= 0; l_sediment = ised /= 0; set_Nc = rnc; l_reuse_thompson_lookup = .false.
  open(22, file=trim(fout), access='stream', form='unformatted', status='replace')

  if (mode == 1 .or. mode == 3) then
     allocate(f(nz, nx, 9), p(nz, nx), dzq(nz), ppt(4, nx), nc1d(nz), nwfa1d(nz), nifa1d(nz), w1d(nz), col(nz, 9))
     read(21) f
     read(21) p
     if (mode == 3) then
        allocate(ae(nz, nx, 4))
        read(21) ae
        is_aerosol_aware = .true.
     end if
     read(21) dzq
     call thompson_init                                     ! I:100-103
     do i = 1, nx
        pptrain = 0.; pptsnow = 0.; pptgraul = 0.; pptice = 0.      ! I:55-58
        col = f(:, i, :)
        do k = 1, nz                                        ! M:957-964
           rho = 0.622 * p(k, i) / (287.04 * col(k, 9) * (col(k, 1) + 0.622))
           nc1d(k) = (set_Nc * 1.e6) / rho
           nwfa1d(k) = 11.1E6 / rho
           nifa1d(k) = 0.5E6 * 0.01 / rho
           w1d(k) = 0.
        end do
        if (mode == 3) then
           nc1d = ae(:, i, 1); nwfa1d = ae(:, i, 2); nifa1d = ae(:, i, 3); w1d = ae(:, i, 4)
        end if
        ! field order of the file: qv qc qi qr qs qg ni nr t
        call mp_thompson(col(:, 1), col(:, 2), col(:, 3), col(:, 4), col(:, 5), col(:, 6), col(:, 7), col(:, 8), nc1d, nwfa1d, &
             nifa1d, col(:, 9), p(:, i), w1d, dzq, pptrain, pptsnow, pptgraul, pptice, 1, nz, dt, i, 1)
        f(:, i, :) = col
        ppt(1, i) = pptrain; ppt(2, i) = pptice; ppt(3, i) = pptsnow; ppt(4, i) = pptgraul
        if (mode == 3) then
           ae(:, i, 1) = nc1d; ae(:, i, 2) = nwfa1d; ae(:, i, 3) = nifa1d
        end if
     end do
     write(22) f
     write(22) ppt
     if (mode == 3) write(22) ae(:, :, 1:3)
  else
     call allocate_columns(nz, nx)
     allocate(plane(nz, nx))
     read(21) theta
     read(21) dtheta_adv
     read(21) dtheta_div
     read(21) exner
     read(21) qv
     read(21) dqv_adv
     read(21) dqv_div
     read(21) dz
     do m = 1, 7
        ih = mh(m); im = mm(m)
        read(21) plane
        hydrometeors(:, :, ih)%moments(1, im) = plane
        read(21) plane
        dhydrometeors_adv(:, :, ih)%moments(1, im) = plane
        read(21) plane
        dhydrometeors_div(:, :, ih)%moments(1, im) = plane
     end do
     call mphys_thompson09_interfacen
     write(22) dtheta_mphys
     write(22) dqv_mphys
     do m = 1, 7
        plane = dhydrometeors_mphys(:, :, mh(m))%moments(1, mm(m))
        write(22) plane
     end do
  end if

  tw = 0.d0
  call chk(1, reshape(t_Efrw, (/size(t_Efrw)/)))
  if (.not. iiwarm) then
     call chk(3, reshape(tcg_racg, (/size(tcg_racg)/)));   call chk(5, reshape(tmr_racg, (/size(tmr_racg)/)))
     call chk(7, reshape(tcr_gacr, (/size(tcr_gacr)/)));   call chk(9, reshape(tmg_gacr, (/size(tmg_gacr)/)))
     call chk(11, reshape(tnr_racg, (/size(tnr_racg)/)));  call chk(13, reshape(tnr_gacr, (/size(tnr_gacr)/)))
     call chk(15, reshape(tcs_racs1, (/size(tcs_racs1)/))); call chk(17, reshape(tmr_racs1, (/size(tmr_racs1)/)))
     call chk(19, reshape(tcs_racs2, (/size(tcs_racs2)/))); call chk(21, reshape(tmr_racs2, (/size(tmr_racs2)/)))
     call chk(23, reshape(tcr_sacr1, (/size(tcr_sacr1)/))); call chk(25, reshape(tms_sacr1, (/size(tms_sacr1)/)))
     call chk(27, reshape(tcr_sacr2, (/size(tcr_sacr2)/))); call chk(29, reshape(tms_sacr2, (/size(tms_sacr2)/)))
     call chk(31, reshape(tnr_racs1, (/size(tnr_racs1)/))); call chk(33, reshape(tnr_racs2, (/size(tnr_racs2)/)))
     call chk(35, reshape(tnr_sacr1, (/size(tnr_sacr1)/))); call chk(37, reshape(tnr_sacr2, (/size(tnr_sacr2)/)))
     call chk(39, reshape(tpi_qcfz, (/size(tpi_qcfz)/)));  call chk(41, reshape(tni_qcfz, (/size(tni_qcfz)/)))
     call chk(43, reshape(tpi_qrfz, (/size(tpi_qrfz)/)));  call chk(45, reshape(tpg_qrfz, (/size(tpg_qrfz)/)))
     call chk(47, reshape(tni_qrfz, (/size(tni_qrfz)/)));  call chk(49, reshape(tnr_qrfz, (/size(tnr_qrfz)/)))
     call chk(51, reshape(tps_iaus, (/size(tps_iaus)/)));  call chk(53, reshape(tni_iaus, (/size(tni_iaus)/)))
     call chk(55, reshape(tpi_ide, (/size(tpi_ide)/)));    call chk(57, reshape(t_Efsw, (/size(t_Efsw)/)))
  end if
  tw(64) = dble(n_save_dg)
  write(22) tw
  close(22)
contains
  subroutine chk(at, v)          ! sum of the table and of every 997th element weighted by its position
    integer, intent(in) :: at
    real(8), intent(in) :: v(:)
    integer :: j
    tw(at) = sum(v)
    do j = 1, size(v), 997
       tw(at + 1) = tw(at + 1) + v(j) * dble(mod(j, 1009) + 1)
    end do
  end subroutine chk
end program ref_driver
