"""Time the Unet forward (config 3 network) at a given batch; used under ncu for launch lists."""
import sys, io, contextlib, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cold_diffusion_models_b200 as cdm

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
with contextlib.redirect_stdout(io.StringIO()):
    u = cdm.Unet(dim=64, dim_mults=(1, 2, 4, 8), channels=3).cuda()
x = torch.rand(B, 3, 128, 128, device='cuda') * 2 - 1
t = torch.randint(0, 200, (B,), device='cuda')
with torch.no_grad():
    for _ in range(3):
        y = u(x, t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(iters):
        y = u(x, t)
    e1.record()
    host = time.time() - t0
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print("B=%d fwd %.3f ms/iter (host enqueue %.3f ms/iter) -> %.1f img/s, %.1f TFLOP/s" % (B, ms, host / iters * 1e3, B / ms * 1e3, 67.41e9 * B / ms / 1e9))
