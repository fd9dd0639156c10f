"""Host-side logic of the multi-GPU path: columns are independent
(I:54, M:1156-1177), so a domain is cut into contiguous column ranges, one per rank, with no halo
and no data-path collective.  The only exchange is one all-reduce of the eight f64 domain
diagnostics (64 bytes) per output interval - the device-side equivalent of the column means
KiD saves at I:255-275."""
import ctypes
import os

import numpy as np

NDIAG = 8


def _parse_list(text):
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11]"""
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def gpu_numa_node(pci_bus_id):
    """NUMA node of a GPU from sysfs (-1: the platform exposes none)."""
    try:
        with open("/sys/bus/pci/devices/%s/numa_node" % pci_bus_id.lower()) as f:
            return int(f.read().strip())
    except (OSError, ValueError):
        return -1


def bind_to_gpu_numa(device_index):
    """Run this process, and place the pages it allocates from now on (the pinned staging arrays of the end-to-end
    path), on the NUMA node the GPU hangs off: eight ranks that all sit on node 0 push every host<->device byte of
    four of the GPUs across the socket link.  Does what the container allows and reports it: CPU affinity to the
    node's cores that are in the allowed set, preferred-node memory policy if the node is in Mems_allowed."""
    info = {"gpu": int(device_index), "node": -1, "cpus_bound": 0, "mem_policy": "default"}
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    except Exception as e:            # an older torch without the PCI ids: ask NVML
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[device_index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else device_index
            b = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
            b = b.decode() if isinstance(b, bytes) else b
            bus = b[-12:] if len(b) > 12 else b
        except Exception as e2:
            info["error"] = (str(e) + " / " + str(e2))[:120]
            return info
    node = gpu_numa_node(bus)
    info["pci"] = bus
    info["node"] = node
    try:
        info["nodes_online"] = open("/sys/devices/system/node/online").read().strip()
    except OSError:
        pass
    if node < 0:
        return info
    try:
        cpus = set(_parse_list(open("/sys/devices/system/node/node%d/cpulist" % node).read()))
        allowed = os.sched_getaffinity(0)
        use = sorted(cpus & allowed)
        if use:
            os.sched_setaffinity(0, use)
            info["cpus_bound"] = len(use)
        else:
            info["cpus_note"] = "none of node %d's cores is in the allowed set %s" % (node, sorted(allowed)[:2] + ["..."])
    except OSError as e:
        info["cpus_note"] = str(e)[:80]
    try:
        mems = []
        for line in open("/proc/self/status"):
            if line.startswith("Mems_allowed_list:"):
                mems = _parse_list(line.split(":", 1)[1])
        if node in mems:
            libc = ctypes.CDLL(None, use_errno=True)
            mask = (ctypes.c_ulong * 16)()
            mask[node // 64] = 1 << (node % 64)
            MPOL_PREFERRED, SYS_set_mempolicy = 1, 238          # x86-64
            rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), 16 * 64 + 1)
            info["mem_policy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy errno %d" % ctypes.get_errno()
        else:
            info["mem_policy"] = "node %d not in Mems_allowed %s" % (node, mems)
    except Exception as e:
        info["mem_policy"] = "unavailable: %s" % str(e)[:60]
    return info


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
