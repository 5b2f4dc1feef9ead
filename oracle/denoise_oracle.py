"""ORACLE (test infrastructure only).  CPU fp32 restatement of the Gaussian-noise baseline
(DN = denoising-diffusion-pytorch/denoising_diffusion_pytorch/denoising_diffusion_pytorch.py): cosine schedule
(DN:295-305), q_sample (DN:517-522), get_x2_bar_from_xt (DN:377-381), sample (DN:342-375), gen_sample
(DN:383-434), p_losses (DN:524-536).  Pinned against the unmodified reference by tests/test_oracle_golden.py
(tests/golden/denoise_small.npz)."""
import torch
import torch.nn.functional as F


def cosine_beta_schedule(timesteps, s=0.008):
    steps = timesteps + 1
    x = torch.linspace(0, steps, steps)
    ac = torch.cos(((x / steps) + s) / (1 + s) * torch.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return torch.clip(betas, 0, 0.999)


class DenoiseOracle:
    def __init__(self, denoise_fn, *, image_size, channels=3, timesteps=1000, loss_type='l1', sampling_routine='default'):
        self.denoise_fn = denoise_fn
        self.num_timesteps = int(timesteps)
        self.loss_type, self.sampling_routine = loss_type, sampling_routine
        ac = torch.cumprod(1. - cosine_beta_schedule(timesteps), axis=0)
        self.sa, self.sb = torch.sqrt(ac), torch.sqrt(1. - ac)

    def _e(self, a, t):
        return a.gather(-1, t).reshape(-1, 1, 1, 1)

    def q_sample(self, x_start, x_end, t):
        return self._e(self.sa, t) * x_start + self._e(self.sb, t) * x_end

    def get_x2_bar_from_xt(self, x1_bar, xt, t):
        return (xt - self._e(self.sa, t) * x1_bar) / self._e(self.sb, t)

    def p_losses(self, x_start, x_end, t):
        x_recon = self.denoise_fn(self.q_sample(x_start, x_end, t), t)
        if self.loss_type == 'l1':
            return (x_start - x_recon).abs().mean()
        elif self.loss_type == 'l2':
            return F.mse_loss(x_start, x_recon)
        raise NotImplementedError()

    @torch.no_grad()
    def _reverse(self, batch_size, img, t, fixed_noise):
        direct = None
        while t:
            step = torch.full((batch_size,), t - 1, dtype=torch.long)
            x1_bar = self.denoise_fn(img, step)
            x2_bar = fixed_noise if fixed_noise is not None else self.get_x2_bar_from_xt(x1_bar, img, step)
            if direct is None:
                direct = x1_bar
            xt_bar = self.q_sample(x1_bar, x2_bar, step)
            xt_sub1 = x1_bar
            if t - 1 != 0:
                xt_sub1 = self.q_sample(x1_bar, x2_bar, torch.full((batch_size,), t - 2, dtype=torch.long))
            img = img - xt_bar + xt_sub1
            t -= 1
        return direct, img

    def sample(self, batch_size, img, t=None):
        d, out = self._reverse(batch_size, img, t or self.num_timesteps, None)
        return img, d, out

    def gen_sample(self, batch_size, img, t=None):
        t = t or self.num_timesteps
        if self.sampling_routine == 'ddim':
            d, out = self._reverse(batch_size, img, t, None)
        elif self.sampling_routine == 'x0_step_down':
            d, out = self._reverse(batch_size, img, t, img)
        else:
            d, out = None, img
        return img, d, out
