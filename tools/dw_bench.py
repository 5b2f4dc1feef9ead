"""Stand-alone launches of the depthwise 7x7 kernels (forward and weight gradient) at one shape; used under ncu."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cold_diffusion_models_b200._lib import call, ptr, stream, lib

B, H, W, Cc = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (32, 128, 128, 64))]
x = torch.randn(B, H, W, Cc, device='cuda')
dh = torch.randn(B, H, W, Cc, device='cuda')
w = torch.randn(Cc, 1, 7, 7, device='cuda')
b = torch.randn(Cc, device='cuda')
out = torch.empty_like(x)
dw = torch.zeros(Cc, 49, device='cuda')
NULL = C.c_void_p(0)
outs = {}
for pipe in (1, 0):
    lib.cd_dwconv7_set_pipe(pipe)
    for it in range(3):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        call('cd_dwconv7_fwd', ptr(x), Cc, B, H, W, Cc, ptr(w), ptr(b), NULL, 0, ptr(out), Cc, 0, NULL, 0, stream())
        e1.record()
        call('cd_dwconv7_wgrad', ptr(dh), Cc, ptr(x), Cc, B, H, W, Cc, ptr(dw), stream())
        e2.record()
        torch.cuda.synchronize()
    outs[pipe] = out.clone()
    print("pipe=%d fwd %.1f us   wgrad %.1f us" % (pipe, e0.elapsed_time(e1) * 1e3, e1.elapsed_time(e2) * 1e3))
lib.cd_dwconv7_set_pipe(1)
