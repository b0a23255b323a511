import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from renet_b200 import _lib
L = _lib.lib(); L.renet_debug_gemm.argtypes = [ctypes.c_int]
dev = 'cuda:0'
A = torch.randn(23033, 200, device=dev); B = torch.randn(200, 200, device=dev) * 0.1
idx = torch.randint(0, 23033, (34483,), device=dev, dtype=torch.int32)
out = torch.empty(34483, 200, device=dev)
L.renet_set_gemm_engine(1)
n0 = _lib.launch_count()
for flags, name in ((0, 'full'), (7, 'nothing'), (1, 'no MMA'), (4, 'no stores'), (32, 'full, no pack kernel')):
    L.renet_debug_gemm(flags)
    for _ in range(3):
        L.renet_selfloop_gemm(_lib.ptr(A), _lib.ptr(idx), _lib.ptr(B), _lib.ptr(out), 34483, 200, 200, _lib.stream())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        L.renet_selfloop_gemm(_lib.ptr(A), _lib.ptr(idx), _lib.ptr(B), _lib.ptr(out), 34483, 200, 200, _lib.stream())
    b.record(); torch.cuda.synchronize()
    print('%-22s %.1f us  (launches per call %.1f)' % (name, a.elapsed_time(b) / 20 * 1e3, (_lib.launch_count() - n0) / 23)); n0 = _lib.launch_count()
