"""Stand-alone launches of the depthwise 7x7 kernels (forward and weight gradient): TMA-staged tiles (default) against the
LDGSTS-staged kernels and the one-tile-per-block kernels, at the shapes of the config-3 network (or one shape from argv)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cold_diffusion_models_b200._lib import call, ptr, stream, lib

SHAPES = [(32, 128, 128, 64), (32, 64, 64, 64), (32, 64, 64, 128), (32, 64, 64, 256), (32, 32, 32, 128), (32, 32, 32, 256),
          (32, 32, 32, 512), (32, 16, 16, 256), (32, 16, 16, 512), (32, 16, 16, 1024)]
if len(sys.argv) > 4:
    SHAPES = [tuple(int(v) for v in sys.argv[1:5])]
NULL = C.c_void_p(0)
flush = torch.empty(64 * 1024 * 1024, device='cuda')          # 256 MB > L2
print("%-22s %28s %28s" % ("(B,H,W,C)", "fwd us: tma / ldgsts / tile", "wgrad us: tma / ldgsts / tile"))
for (B, H, W, Cc) in SHAPES:
    x = torch.randn(B, H, W, Cc, device='cuda')
    dh = torch.randn(B, H, W, Cc, device='cuda')
    w = torch.randn(Cc, 1, 7, 7, device='cuda')
    b = torch.randn(Cc, device='cuda')
    out = torch.empty_like(x)
    dw = torch.zeros(Cc, 49, device='cuda')
    res = []
    for (pipe, tma) in ((1, 1), (1, 0), (0, 0)):
        lib.cd_dwconv7_set_pipe(pipe); lib.cd_dwconv7_set_tma(tma)
        tf, tw = [], []
        for it in range(6):
            flush.zero_()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            call('cd_dwconv7_fwd', ptr(x), Cc, B, H, W, Cc, ptr(w), ptr(b), NULL, 0, ptr(out), Cc, 0, NULL, 0, stream())
            e1.record()
            call('cd_dwconv7_wgrad', ptr(dh), Cc, ptr(x), Cc, B, H, W, Cc, ptr(dw), stream())
            e2.record()
            torch.cuda.synchronize()
            if it >= 2:
                tf.append(e0.elapsed_time(e1) * 1e3); tw.append(e1.elapsed_time(e2) * 1e3)
        res.append((sorted(tf)[len(tf) // 2], sorted(tw)[len(tw) // 2]))
    lib.cd_dwconv7_set_pipe(1); lib.cd_dwconv7_set_tma(1)
    print("%-22s %8.1f / %6.1f / %6.1f     %8.1f / %6.1f / %6.1f" % ((str((B, H, W, Cc)),) + tuple(r[0] for r in res) + tuple(r[1] for r in res)))
