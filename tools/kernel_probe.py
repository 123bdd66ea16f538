"""Launch each GEMM engine once or twice at the C2 layer size (for ncu captures)."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from neuraludf_b200 import _lib as L

lib = L.lib()
dev = torch.device("cuda")
P = 65536
X = torch.randn(P, 256, device=dev) * 0.1
W = torch.randn(256, 256, device=dev) * 0.06
b = torch.zeros(256, device=dev)
Y = torch.empty(P, 256, device=dev)
dW = torch.zeros(256, 256, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
imgs = {}
for npl in (2, 3):
    im = torch.zeros(lib.nudf_tc_image_elems(256, 256, npl), dtype=torch.int16, device=dev)
    lib.nudf_tc_prepare_weights(L.ptr(W), 256, 256, 256, 0, npl, L.ptr(im), st)
    imgs[npl] = im
for _ in range(2):
    lib.nudf_dense_forward(L.ptr(X), 256, L.ptr(W), 256, L.ptr(b), L.ptr(Y), 256, P, 256, 256, 2, st)
    lib.nudf_dense_forward_tc(L.ptr(X), 256, L.ptr(imgs[2]), 2, L.ptr(b), L.ptr(Y), 256, P, 256, 256, 2, st)
    lib.nudf_dense_forward_tc(L.ptr(X), 256, L.ptr(imgs[3]), 3, L.ptr(b), L.ptr(Y), 256, P, 256, 256, 2, st)
    lib.nudf_wgrad(L.ptr(Y), 256, L.ptr(X), 256, 256, 256, P, L.ptr(dW), 256, 1, st)
    lib.nudf_wgrad(L.ptr(Y), 256, L.ptr(X), 256, 256, 256, P, L.ptr(dW), 256, 0, st)
torch.cuda.synchronize()
print("done")
