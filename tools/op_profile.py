"""In-situ time of every libcolddiff entry point during one training micro-step (config 3 network, forward + backward)
and one optimizer step, measured with CUDA events around each C-ABI call (warm caches, real launch order).  ncu launch
lists replay each kernel with cold caches, which overstates the small latency-bound kernels; this table does not."""
import sys, io, contextlib, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cold_diffusion_models_b200 as cdm
from cold_diffusion_models_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
with contextlib.redirect_stdout(io.StringIO()):
    u = cdm.Unet(dim=64, dim_mults=(1, 2, 4, 8), channels=3).cuda()
    gd = cdm.GaussianDiffusion(u, image_size=128, device_of_kernel='cuda', channels=3, timesteps=200, kernel_std=0.01,
                               kernel_size=15, blur_routine='Exponential_reflect', sampling_routine='x0_step_down').cuda()
    tr = cdm.Trainer(gd, None, image_size=128, train_batch_size=B, gradient_accumulate_every=2, results_folder='/tmp/opprof',
                     dataset='synthetic')
xs = [torch.rand(B, 3, 128, 128, device='cuda') * 2 - 1 for _ in range(2)]
for _ in range(3):
    tr.train_step(batches=xs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
tr.train_step(batches=xs)
e1.record()
torch.cuda.synchronize()
print("optimizer step (2 micro-batches of %d) without per-call events: %.2f ms" % (B, e0.elapsed_time(e1)))
_lib.profile_start()
tr.train_step(batches=xs)
prof = _lib.profile_stop()
tot = sum(t for c, t in prof.values())
print("sum over calls with events: %.2f ms" % tot)
agg = {}
for name, (c, t) in prof.items():
    base = name.split('(')[0]
    c0, t0 = agg.get(base, (0, 0.0))
    agg[base] = (c0 + c, t0 + t)
print("%-28s %6s %10s %7s %10s" % ("entry point", "calls", "total ms", "share", "us/call"))
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-28s %6d %10.3f %6.1f%% %10.1f" % (name, c, t, 100 * t / tot, t / c * 1e3))
print("\nper shape (HBM-bound kernels):")
for name, (c, t) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    if '(' in name:
        print("%-44s %6d %10.3f %10.1f" % (name, c, t, t / c * 1e3))
