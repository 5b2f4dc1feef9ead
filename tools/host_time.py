"""How long does the HOST need to enqueue one optimizer step (Python + ctypes + 1133 kernel launches), against the GPU time of the
step?  Ten steps are enqueued back to back after one synchronize; the host clock is read when the last launch has been issued
(before synchronizing) and after the device has drained."""
import sys, io, contextlib, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cold_diffusion_models_b200 as cdm

B = 32
with contextlib.redirect_stdout(io.StringIO()):
    u = cdm.Unet(dim=64, dim_mults=(1, 2, 4, 8), channels=3).cuda()
    gd = cdm.GaussianDiffusion(u, image_size=128, device_of_kernel='cuda', channels=3, timesteps=200, kernel_std=0.01,
                               kernel_size=15, blur_routine='Exponential_reflect', sampling_routine='x0_step_down').cuda()
    tr = cdm.Trainer(gd, None, image_size=128, train_batch_size=B, gradient_accumulate_every=2, results_folder='/tmp/onestep',
                     dataset='synthetic')
xs = [torch.rand(B, 3, 128, 128, device='cuda') * 2 - 1 for _ in range(2)]
for _ in range(3):
    tr.train_step(batches=xs)
torch.cuda.synchronize()
N = 10
t0 = time.perf_counter()
for _ in range(N):
    tr.train_step(batches=xs)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.1f ms per step; step (host + device drained) %.1f ms per step; the host is %.1f ms ahead of the device after %d steps"
      % ((t1 - t0) / N * 1e3, (t2 - t0) / N * 1e3, (t2 - t1) * 1e3, N))
# single step from an idle device: what the queue depth hides
torch.cuda.synchronize(); t0 = time.perf_counter(); tr.train_step(batches=xs); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("one step from an idle device: host %.1f ms, total %.1f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
print("cpu count", os.cpu_count())
