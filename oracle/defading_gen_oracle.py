"""ORACLE (test infrastructure only).  CPU fp32 restatement of the defading-GENERATION package
(DFGEN = defading-generation-diffusion-pytorch/defading_diffusion_pytorch/defading_diffusion_pytorch.py): per-pixel
fade schedule (DFGEN:313-344), q_sample (DFGEN:543-548), p_losses (DFGEN:550-562), sample (DFGEN:385-418),
gen_sample (DFGEN:427-457), forward_and_backward (DFGEN:459-504), all_sample (DFGEN:506-541).  The Gaussian taps are
the torchgeometry restatement of deblur_oracle.gaussian_1d (parity unpinned at that boundary only); everything else is
pinned against the unmodified reference by tests/test_oracle_golden.py (tests/golden/defading_gen_small.npz)."""
import torch
import torch.nn.functional as F

from deblur_oracle import gaussian_1d


def fade_kernel(size, std):
    g = gaussian_1d(size, std)
    k = torch.matmul(g.unsqueeze(-1), g.unsqueeze(-1).t())
    k = k / torch.max(k)
    return (torch.ones_like(k) - k)[1:, 1:]


def schedule(timesteps, size, kernel_std, initial_mask, reverse):
    kers, cum, cum_rev = torch.ones((1, size, size)), [], []
    for i in range(timesteps):
        cum_rev.append(kers)
        kers = kers * fade_kernel(size + 1, kernel_std * (i + initial_mask))
        cum.append(kers)
    if reverse:
        cum_rev.reverse()
        one_minus = torch.stack(cum_rev)
        return 1. - one_minus, one_minus
    alphas = torch.stack(cum)
    return alphas, 1. - alphas


class DefadingGenOracle:
    def __init__(self, denoise_fn, *, image_size, channels=3, timesteps=1000, loss_type='l1', reverse=False,
                 kernel_std=0.15, initial_mask=11):
        self.denoise_fn = denoise_fn
        self.num_timesteps = int(timesteps)
        self.loss_type = loss_type
        self.alphas, self.one_minus_alphas = schedule(self.num_timesteps, image_size, kernel_std, initial_mask, reverse)

    def q_sample(self, x_start, x_end, t):
        return self.alphas[t] * x_start + self.one_minus_alphas[t] * x_end          # (B,1,S,S) planes, DFGEN:285-294

    def p_losses(self, x_start, x_end, t):
        x_recon = self.denoise_fn(self.q_sample(x_start, x_end, t), t)
        if self.loss_type == 'l1':
            return (x_start - x_recon).abs().mean()
        elif self.loss_type == 'l2':
            return F.mse_loss(x_start, x_recon)
        raise NotImplementedError()

    @torch.no_grad()
    def _reverse(self, batch_size, img, x2, t, X1=None, Xt=None):
        direct = None
        while t:
            step = torch.full((batch_size,), t - 1, dtype=torch.long)
            x1_bar = self.denoise_fn(img, step)
            if X1 is not None:
                X1.append(x1_bar)
            if Xt is not None:
                Xt.append(img)
            if direct is None:
                direct = x1_bar
            xt_bar = self.q_sample(x1_bar, x2, step)
            xt_sub1 = x1_bar
            if t - 1 != 0:
                xt_sub1 = self.q_sample(x1_bar, x2, torch.full((batch_size,), t - 2, dtype=torch.long))
            img = img - xt_bar + xt_sub1
            t -= 1
        return direct, img

    def sample(self, batch_size, img, t=None):
        d, out = self._reverse(batch_size, img, img, t or self.num_timesteps)
        return img, d, out

    def gen_sample(self, batch_size, img, noise=None, t=None):
        start = img if noise is None else img + noise
        d, out = self._reverse(batch_size, start, img, t or self.num_timesteps)
        return img, d, out

    def forward_and_backward(self, batch_size, img1, img2, t=None):
        Forward = [img1]
        for i in range(self.num_timesteps):
            Forward.append(self.q_sample(img1, img2, torch.full((batch_size,), i, dtype=torch.long)))
        Backward = []
        _, out = self._reverse(batch_size, img2, img2, t or self.num_timesteps, Xt=Backward)
        return Forward, Backward, out

    def all_sample(self, batch_size, img, t=None):
        X1, Xt = [], []
        self._reverse(batch_size, img, img, t or self.num_timesteps, X1=X1, Xt=Xt)
        return X1, Xt
