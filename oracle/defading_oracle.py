"""ORACLE (test infrastructure only).  CPU fp32 restatement of defading_diffusion_pytorch.GaussianDiffusion
(DFG = defading-diffusion-pytorch/defading_diffusion_pytorch/defading_diffusion_gaussian.py): fade kernels
(DFG:328-352), q_sample (DFG:495-533), p_losses (DFG:535-546), sample (DFG:355-424), with the per-sample window
offsets of the 'Random_*' routines passed in explicitly.  Pinned by tests/test_oracle_golden.py."""
import torch
import torch.nn.functional as F
from deblur_oracle import gaussian_2d


class DefadeOracle:
    def __init__(self, defade_fn, *, image_size, channels=3, timesteps=1000, loss_type='l1', kernel_std=0.1,
                 initial_mask=11, fade_routine='Incremental', sampling_routine='default', discrete=False):
        self.defade_fn, self.S, self.T = defade_fn, image_size, int(timesteps)
        self.loss_type, self.routine, self.sampling_routine, self.discrete = loss_type, fade_routine, sampling_routine, discrete
        ks = []
        for i in range(self.T):
            if fade_routine == 'Incremental':
                ks.append(self._k(image_size + 1, kernel_std * (i + initial_mask)))
            elif fade_routine == 'Constant':
                ks.append(self._k(image_size + 1, kernel_std))
            elif fade_routine == 'Random_Incremental':
                ks.append(self._k(2 * image_size + 1, kernel_std * (i + initial_mask)))
        self.fade_kernels = torch.stack(ks)

    @staticmethod
    def _k(dim, std):
        k = gaussian_2d(dim, std)
        k = k / torch.max(k)
        return (torch.ones_like(k) - k)[1:, 1:]

    def _kern(self, i, rx, ry):
        if rx is None:
            return self.fade_kernels[i]
        S = self.S
        return torch.stack([self.fade_kernels[i][int(rx[b]):int(rx[b]) + S, int(ry[b]):int(ry[b]) + S] for b in range(len(rx))])[:, None]

    def _quant(self, x):
        x = (x + 1) * 0.5
        x = x * 255
        x = x.int().float() / 255
        return x * 2 - 1

    def q_sample(self, x_start, t, rx=None, ry=None):
        xs, x = [], x_start
        for i in range(int(torch.max(t)) + 1):
            x = self._kern(i, rx, ry) * x
            xs.append(x)
        xs = torch.stack(xs)
        out = torch.stack([xs[int(t[b]), b] for b in range(t.shape[0])])
        return self._quant(out) if self.discrete else out

    def p_losses(self, x_start, t, rx=None, ry=None):
        x_recon = self.defade_fn(self.q_sample(x_start, t, rx, ry), t)
        return (x_start - x_recon).abs().mean() if self.loss_type == 'l1' else F.mse_loss(x_start, x_recon)

    @torch.no_grad()
    def sample(self, batch_size, x, t=None, rx=None, ry=None):
        t = t or self.T
        for i in range(t):
            x = self._kern(i, rx, ry) * x
        if self.discrete:
            x = self._quant(x)
        xt, direct, recon = x, None, None
        while t:
            recon = self.defade_fn(x, torch.full((batch_size,), t - 1, dtype=torch.long))
            if direct is None:
                direct = recon
            if self.sampling_routine == 'default':
                for i in range(t - 1):
                    recon = self._kern(i, rx, ry) * recon
                x = recon
            elif self.sampling_routine == 'x0_step_down':
                sub = recon
                for i in range(t):
                    sub = recon
                    recon = self._kern(i, rx, ry) * recon
                x = x - recon + sub
            recon = x
            t -= 1
        return xt, direct, recon
