import sys, ctypes, torch, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiplanarunet_b200 import _C
from multiplanarunet_b200._C import lib, check
V, C, N = 6, 5, 256**3
dev = torch.device("cuda")
X = torch.rand(N, V, C, device=dev); y = torch.randint(0, C, (N,), device=dev, dtype=torch.uint8)
W = torch.ones(V, C, device=dev); b = torch.zeros(C, device=dev)
acc = torch.zeros(V*C+C+1, dtype=torch.float64, device=dev)
st = _C.current_stream()
def run(idx, n, reps=50):
    def k():
        check(lib.mpu_fusion_grad_indexed(_C.ptr(X), _C.ptr(y), _C.ptr(idx), ctypes.c_longlong(n), V, C, _C.ptr(W), _C.ptr(b), _C.ptr(acc), st))
    for _ in range(5): k()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): k()
    e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) / reps * 1e3
bs = 2**17
perm = torch.randperm(N, device=dev)
print("no index, n=2^17: %.1f us" % run(None, bs))
print("identity index:    %.1f us" % run(torch.arange(bs, device=dev), bs))
print("random index:      %.1f us" % run(perm[:bs].contiguous(), bs))
print("sorted random idx: %.1f us" % run(perm[:bs].sort().values.contiguous(), bs))
print("random idx within first 2^20 rows: %.1f us" % run(torch.randperm(2**20, device=dev)[:bs].contiguous(), bs))
print("no index, n=2^24: %.1f us" % run(None, N, 5))
