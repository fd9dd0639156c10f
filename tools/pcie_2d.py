"""H2D bandwidth of the copy pattern of kidmp_step's chunk pipeline: cudaMemcpy2DAsync of [60] rows x `chunk` columns out of pinned
[60][ncol] arrays, ten arrays per chunk, against one contiguous copy of the same bytes."""
import ctypes as C
import glob
import os
import sys
import time

import torch

cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")) + \
        glob.glob("/usr/local/cuda/lib64/libcudart.so*")
rt = C.CDLL(cands[0])
ncol, nz = 1 << 20, 60
host = [torch.empty((nz, ncol), dtype=torch.float32).pin_memory() for _ in range(10)]
for h in host:
    h.fill_(1.0)
s = torch.cuda.Stream()
H2D = 1
for chunk in (65536, 262144, 1048576):
    dev = [torch.empty((nz, chunk), dtype=torch.float32, device="cuda") for _ in range(10 * 3)]
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for it, c0 in enumerate(range(0, ncol, chunk)):
            for q in range(10):
                d = dev[(it % 3) * 10 + q]
                rc = rt.cudaMemcpy2DAsync(C.c_void_p(d.data_ptr()), C.c_size_t(chunk * 4), C.c_void_p(host[q].data_ptr() + c0 * 4),
                                          C.c_size_t(ncol * 4), C.c_size_t(chunk * 4), C.c_size_t(nz), C.c_int(H2D), C.c_void_p(s.cuda_stream))
                assert rc == 0, rc
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
    print("2D chunk %8d: %.1f ms  %.1f GB/s" % (chunk, el * 1e3, 10 * nz * ncol * 4 / el / 1e9), flush=True)
    del dev
dev = [torch.empty((nz, ncol), dtype=torch.float32, device="cuda") for _ in range(10)]
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s):
        for q in range(10):
            dev[q].copy_(host[q], non_blocking=True)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
print("contiguous: %.1f ms  %.1f GB/s" % (el * 1e3, 10 * nz * ncol * 4 / el / 1e9))
