"""tools/hbm_probe.py -- what this box's HBM does for the access mixes of the bench kernels: pure copy (the
MEASURED_PEAKS.json method), pure write, and a 1:8.7 read:write mix like the kinematics kernel (48 B in, 416 B out)."""
import torch, time
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best * 1e-3
n = 1 << 30
x = torch.empty(n, dtype=torch.bfloat16, device="cuda"); y = torch.empty_like(x)
s = t(lambda: y.copy_(x)); print(f"copy  (read+write) {2*n*2/s/1e9:8.1f} GB/s")
z = torch.empty(n * 2, dtype=torch.float32, device="cuda")   # 8 GiB
s = t(lambda: z.zero_()); print(f"write only        {z.numel()*4/s/1e9:8.1f} GB/s")
s = t(lambda: z.fill_(1.5)); print(f"fill  only        {z.numel()*4/s/1e9:8.1f} GB/s")
r = torch.empty(n, dtype=torch.float32, device="cuda")
s = t(lambda: r.sum()); print(f"read only (sum)   {r.numel()*4/s/1e9:8.1f} GB/s")

# the store pattern of the SoA batch kernels in isolation
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matrix_b200._lib import lib, check
n = 1 << 24
for planes in (1, 9, 16, 36, 52):
    buf = torch.empty((planes, n), dtype=torch.float64, device="cuda")
    for mode in (0, 1):
        s = t(lambda: check(lib.ppx_cuda_probe_write(C.c_void_p(buf.data_ptr()), planes, n, n, mode,
                                                     C.c_void_p(torch.cuda.current_stream().cuda_stream))))
        print(f"SoA write, {planes:2d} planes, {'128' if mode else ' 64'}-bit stores: {planes*n*8/s/1e9:8.1f} GB/s")
    del buf
