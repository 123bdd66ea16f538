"""Prints the pipeline timeline (cycles) of CTA 0 of the persistent tcgen05 kernel. NUDF_TC_DEBUG=16 python tools/tc_trace.py"""
import ctypes, os, sys
os.environ.setdefault("NUDF_TC_DEBUG", "16")
import torch
sys.path.insert(0, ".")
from neuraludf_b200 import _lib as L
lib = L.lib(); dev = torch.device("cuda"); P = 65536
X = torch.randn(P, 256, device=dev) * 0.1; W = torch.randn(256, 256, device=dev) * 0.06
b = torch.zeros(256, device=dev); Y = torch.empty(P, 256, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
im = torch.zeros(lib.nudf_tc_image_elems(256, 256, 2), dtype=torch.int16, device=dev)
lib.nudf_tc_prepare_weights(L.ptr(W), 256, 256, 256, 0, 2, L.ptr(im), st)
for _ in range(3):
    lib.nudf_dense_forward_tc(L.ptr(X), 256, L.ptr(im), 2, L.ptr(b), L.ptr(Y), 256, P, 256, 256, 2, st)
buf = (ctypes.c_longlong * 1024)()
lib.nudf_tc_read_trace(ctypes.cast(buf, ctypes.c_void_p))
tr = [[buf[r * 256 + i] for i in range(256)] for r in range(4)]
t0 = tr[3][0]
print("kernel body cycles:", tr[3][1] - t0)
print("producer (per slice): start, got_empty, stored, arrived   [cycles since setup]")
for c in range(16):
    print("  slice %2d:" % c, [tr[0][4 * c + k] - t0 for k in range(4)])
print("mma (per slice): wait_start, got_full, committed")
for c in range(16):
    print("  slice %2d:" % c, [tr[1][3 * c + k] - t0 for k in range(3)])
print("epilogue (per tile): wait_start, got_tfull, done")
for i in range(4):
    print("  tile %d:" % i, [tr[2][3 * i + k] - t0 for k in range(3)])
