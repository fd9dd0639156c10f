! ORACLE - TEST INFRASTRUCTURE ONLY.
! Minimal stand-ins for the seven KiD host modules that the reference scheme imports but /root/reference does not hold
! (module_mp_thompson09n.f90 lines 19-23, mphys_thompson09n.f90 lines 11-17).  They carry exactly the names the two
! reference files use, with KiD's meaning, and nothing else: no advection, no output (save_dg discards its arguments).
! With them the UNMODIFIED reference sources compile into oracle/_ref/ (see Makefile); they are never part of the product.
module typeKind
  implicit none
  integer, parameter :: wp = kind(1.0)          ! KiD's working precision; single in the builds this scheme is used in (U5)
  integer, parameter :: sp = kind(1.0), dp = kind(1.0d0)
end module typeKind

module switches
  implicit none
  logical :: l_sediment = .true.                ! gates ice / snow / graupel sedimentation (M:3449, M:3506, M:3555)
  logical :: l_reuse_thompson_lookup = .false.  ! read run_data/rac[gs]_thompson09.data when present (M:3720, M:3867)
end module switches

module namelists
  implicit none
  logical :: iiwarm = .false.                   ! warm rain only (M:773, M:1545, M:1749 ...)
  real :: set_Nc = 100.                         ! cloud droplets per cm^3 (M:381)
end module namelists

module parameters
  implicit none
  integer, parameter :: max_char_len = 200
  integer, parameter :: nspecies = 5, num_h_moments(5) = (/1, 2, 2, 1, 1/), num_h_bins(5) = 1
  integer :: nx = 1, nz = 1                     ! set by the driver before the first call
  real :: dt = 1.
  character(10) :: h_names(5) = (/ 'cloud     ', 'rain      ', 'ice       ', 'snow      ', 'graupel   ' /)
  character(10) :: mom_units(3) = (/ 'kg/kg     ', '/kg       ', 'm3        ' /)
end module parameters

module physconst
  implicit none
  real, parameter :: p0 = 100000., r_on_cp = 287.05 / 1005., pi = 3.14159265358979
end module physconst

module diagnostics
  implicit none
  integer :: i_dgtime = 1
  integer :: n_save_dg = 0                      ! calls seen (the driver reports it)
  interface save_dg
     module procedure save_dg_scalar, save_dg_level_r4, save_dg_level_r8, save_dg_levelx_r4, save_dg_levelx_r8, &
          save_dg_1d, save_dg_2d
  end interface save_dg
contains
  subroutine save_dg_scalar(value, name, itime, units, dim)          ! I:162
    real, intent(in) :: value
    character(*), intent(in) :: name
    integer, intent(in) :: itime
    character(*), intent(in), optional :: units, dim
    n_save_dg = n_save_dg + 1
  end subroutine save_dg_scalar
  subroutine save_dg_level_r8(k, value, name, itime, units, dim)     ! M:2967: one level of a profile (the rates are DOUBLE PRECISION)
    integer, intent(in) :: k
    double precision, intent(in) :: value
    character(*), intent(in) :: name
    integer, intent(in) :: itime
    character(*), intent(in), optional :: units, dim
    n_save_dg = n_save_dg + 1
  end subroutine save_dg_level_r8
  subroutine save_dg_level_r4(k, value, name, itime, units, dim)     ! the same from a REAL caller (the kidmp shim)
    integer, intent(in) :: k
    real, intent(in) :: value
    character(*), intent(in) :: name
    integer, intent(in) :: itime
    character(*), intent(in), optional :: units, dim
    n_save_dg = n_save_dg + 1
  end subroutine save_dg_level_r4
  subroutine save_dg_levelx_r8(k, i, value, name, itime, units, dim) ! M:3046: one level of one column (nx > 1)
    integer, intent(in) :: k, i
    double precision, intent(in) :: value
    character(*), intent(in) :: name
    integer, intent(in) :: itime
    character(*), intent(in), optional :: units, dim
    n_save_dg = n_save_dg + 1
  end subroutine save_dg_levelx_r8
  subroutine save_dg_levelx_r4(k, i, value, name, itime, units, dim)
    integer, intent(in) :: k, i
    real, intent(in) :: value
    character(*), intent(in) :: name
    integer, intent(in) :: itime
    character(*), intent(in), optional :: units, dim
    n_save_dg = n_save_dg + 1
  end subroutine save_dg_levelx_r4
  subroutine save_dg_1d(field, name, itime, units, dim)              ! I:255
    real, intent(in) :: field(:)
    character(*), intent(in) :: name
    integer, intent(in) :: itime
    character(*), intent(in), optional :: units, dim
    n_save_dg = n_save_dg + 1
  end subroutine save_dg_1d
  subroutine save_dg_2d(field, name, itime, units, dim)              ! I:307
    real, intent(in) :: field(:, :)
    character(*), intent(in) :: name
    integer, intent(in) :: itime
    character(*), intent(in), optional :: units, dim
    n_save_dg = n_save_dg + 1
  end subroutine save_dg_2d
end module diagnostics

module column_variables
  use parameters, only : nspecies
  implicit none
  type species                                  ! KiD: moments(bin, moment) of one hydrometeor species at one grid point
     real :: moments(1, 3) = 0.
  end type species
  real, allocatable :: theta(:, :), dtheta_adv(:, :), dtheta_div(:, :), dtheta_mphys(:, :), exner(:, :)
  real, allocatable :: qv(:, :), dqv_adv(:, :), dqv_div(:, :), dqv_mphys(:, :)
  real, allocatable :: dz(:)
  type(species), allocatable :: hydrometeors(:, :, :), dhydrometeors_adv(:, :, :), dhydrometeors_div(:, :, :), &
       dhydrometeors_mphys(:, :, :)
contains
  subroutine allocate_columns(nz, nx)
    integer, intent(in) :: nz, nx
    allocate(theta(nz, nx), dtheta_adv(nz, nx), dtheta_div(nz, nx), dtheta_mphys(nz, nx), exner(nz, nx))
    allocate(qv(nz, nx), dqv_adv(nz, nx), dqv_div(nz, nx), dqv_mphys(nz, nx), dz(nz))
    allocate(hydrometeors(nz, nx, nspecies), dhydrometeors_adv(nz, nx, nspecies), dhydrometeors_div(nz, nx, nspecies), &
         dhydrometeors_mphys(nz, nx, nspecies))
  end subroutine allocate_columns
end module column_variables
