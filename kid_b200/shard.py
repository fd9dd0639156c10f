"""Host-side logic of the multi-GPU path: columns are independent
(I:54, M:1156-1177), so a domain is cut into contiguous column ranges, one per rank, with no halo
and no data-path collective.  The only exchange is one all-reduce of the eight f64 domain
diagnostics (64 bytes) per output interval - the device-side equivalent of the column means
KiD saves at I:255-275."""
import numpy as np

NDIAG = 8


def shard_range(ncol_total, rank, world):
    """Columns [c0, c1) owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(int(ncol_total), int(world))
    c0 = rank * base + min(rank, rem)
    return c0, c0 + base + (1 if rank < rem else 0)


def diag_from_state(state, p, dz, ppt):
    """The eight domain sums of kidmp_diag, from host arrays in COL_FASTEST layout (nz, ncol):
    0..3 surface precipitation rain/ice/snow/graupel, 4 liquid water path, 5 ice water path
    [kg m^-2 summed over columns], 6 columns with any hydrometeor or precipitation, 7 columns."""
    f = lambda k: np.asarray(state[k], np.float32)
    rho = np.float32(0.622) * p / (np.float32(287.04) * f("t") * (f("qv") + np.float32(0.622)))
    dzc = np.asarray(dz, np.float32)[:, None]
    lwp = ((f("qc") + f("qr")) * rho * dzc).astype(np.float64).sum()
    iwp = ((f("qi") + f("qs") + f("qg")) * rho * dzc).astype(np.float64).sum()
    any_h = ((f("qc") + f("qr") + f("qi") + f("qs") + f("qg")) > 0).any(0) | (np.asarray(ppt) > 0).any(0)
    out = np.zeros(NDIAG, np.float64)
    out[:4] = np.asarray(ppt, np.float64).sum(1)
    out[4], out[5], out[6], out[7] = lwp, iwp, float(any_h.sum()), float(np.asarray(ppt).shape[1])
    return out


def allreduce_diag(local, device=None):
    """Sum the eight diagnostics over all ranks (NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.asarray(local, np.float64).copy())
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t.cpu().numpy()
