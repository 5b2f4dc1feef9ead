"""ORACLE (test infrastructure only).  CPU fp32 restatement of resolution_diffusion_pytorch.GaussianDiffusion
(RS = resolution-diffusion-pytorch/resolution_diffusion_pytorch/resolution_diffusion_pytorch.py): transform_func /
get_funcs (RS:354-414) applied SEQUENTIALLY with torch's interpolate exactly as the reference does, q_sample
(RS:630-652), sample (RS:417-459), p_losses 'Final' (RS:655-667).  Pinned by tests/test_oracle_golden.py."""
import torch
import torch.nn.functional as F
from deblur_oracle import gaussian_2d


class ResolutionOracle:
    def __init__(self, denoise_fn, *, image_size, channels=3, timesteps=1000, loss_type='l1',
                 resolution_routine='Incremental', sampling_routine='default'):
        self.denoise_fn, self.image_size, self.channels = denoise_fn, image_size, channels
        self.num_timesteps, self.loss_type = int(timesteps), loss_type
        self.routine, self.sampling_routine = resolution_routine, sampling_routine

    def _spec(self, i):
        r, S = self.routine, self.image_size
        mode = 'bilinear' if 'bilinear' in r else ('area' if 'area' in r else 'bicubic')
        d = S - S // 2 ** (i + 1) if 'factor_2' in r else i
        return d, mode, ('with_blur' in r)

    def _blur(self, x):
        w = gaussian_2d(3, 0.5)[None, None].repeat(self.channels, 1, 1, 1)
        return F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), w, groups=self.channels)

    def func(self, i, img):
        d, mode, blur = self._spec(i)
        if blur:
            img = self._blur(img)
        x = F.interpolate(img, size=img.shape[2] - d, mode=mode, antialias=False)
        x = F.interpolate(x, size=img.shape[2], mode='nearest-exact', antialias=False)
        return self._blur(x) if blur else x

    def q_sample(self, x_start, t):
        xs, x = [], x_start
        for i in range(int(torch.max(t)) + 1):
            x = self.func(i, x)
            xs.append(x)
        xs = torch.stack(xs)
        return torch.stack([xs[int(t[b]), b] for b in range(t.shape[0])])

    def _loss(self, a, b):
        return (a - b).abs().mean() if self.loss_type == 'l1' else F.mse_loss(a, b)

    def p_losses(self, x_start, t, train_routine='Final'):
        """RS:655-761: 'Final' and the five research routines (the random ones draw from torch's global generator like the
        reference: same seed, same numbers)"""
        r = train_routine
        if r == 'Final':
            return self._loss(x_start, self.denoise_fn(self.q_sample(x_start, t), t))
        if r == 'Final_small_noise':
            x_start = x_start + 0.001 * torch.randn_like(x_start)
            return self._loss(x_start, self.denoise_fn(self.q_sample(x_start, t), t))
        if r in ('Final_random_mean', 'Final_random_mean_and_actual'):
            loss1 = self._loss(x_start, self.denoise_fn(self.q_sample(x_start, t), t)) if r.endswith('actual') else 0.0
            new_mean = torch.randn_like(torch.mean(x_start, [2, 3]))[:, :, None, None]
            x_start = x_start - torch.mean(x_start, [2, 3], keepdim=True) + new_mean
            return loss1 + self._loss(x_start, self.denoise_fn(self.q_sample(x_start, t), t))
        if r == 'Gradient_norm':          # the reference raises here (torch.linalg.norm rejects dim=(1,2,3), RS:738): intended semantics
            x_blur = self.q_sample(x_start, t)
            gradient = x_blur - x_start
            norm = torch.linalg.norm(gradient.flatten(1), dim=1).reshape(-1, 1, 1, 1)
            return self._loss(gradient / (norm + 1e-5), self.denoise_fn(x_blur, t))
        if r == 'Step':
            return self._loss(self.q_sample(x_start, t - 1), self.denoise_fn(self.q_sample(x_start, t), t))
        raise UnboundLocalError(r)

    @torch.no_grad()
    def sample(self, batch_size, img, t=None):
        t = t or self.num_timesteps
        for i in range(t):
            img = self.func(i, img)
        xt, direct = img, None
        while t:
            x = self.denoise_fn(img, torch.full((batch_size,), t - 1, dtype=torch.long))
            if direct is None:
                direct = x
            if self.sampling_routine == 'default':
                for i in range(t - 1):
                    x = self.func(i, x)
            elif self.sampling_routine == 'x0_step_down':
                a = x
                for i in range(t):
                    a = self.func(i, a)
                b = x
                for i in range(t - 1):
                    b = self.func(i, b)
                x = img - a + b
            img = x
            t -= 1
        return xt, direct, img
