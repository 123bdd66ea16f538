"""Times the persistent tcgen05 dense kernel with parts disabled (NUDF_TC_DEBUG) to locate its bottleneck."""
import ctypes, os, subprocess, sys, json
if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, ".")
    from neuraludf_b200 import _lib as L
    lib = L.lib(); dev = torch.device("cuda"); P = 65536
    X = torch.randn(P, 256, device=dev) * 0.1; W = torch.randn(256, 256, device=dev) * 0.06
    b = torch.zeros(256, device=dev); Y = torch.empty(P, 256, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    im = torch.zeros(lib.nudf_tc_image_elems(256, 256, 2), dtype=torch.int16, device=dev)
    lib.nudf_tc_prepare_weights(L.ptr(W), 256, 256, 256, 0, 2, L.ptr(im), st)
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    call = lambda: lib.nudf_dense_forward_tc(L.ptr(X), 256, L.ptr(im), 2, L.ptr(b), L.ptr(Y), 256, P, 256, 256, int(sys.argv[2]), st)
    for _ in range(3): call()
    t = 0
    for _ in range(10):
        flush.zero_(); a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(); e.record(); torch.cuda.synchronize(); t += a.elapsed_time(e)
    print("dbg=%s act=%s: %.1f us" % (os.environ.get("NUDF_TC_DEBUG", "0"), sys.argv[2], t / 10 * 1e3))
else:
    for dbg in (0, 1, 2, 4, 8, 3, 5, 6, 7, 15, 9, 10, 12):
        for act in ((2, 0) if dbg in (0, 2) else (2,)):
            env = dict(os.environ, NUDF_TC_DEBUG=str(dbg))
            print(subprocess.run([sys.executable, __file__, "run", str(act)], env=env, capture_output=True, text=True).stdout.strip())
