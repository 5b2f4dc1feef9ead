"""ORACLE (test infrastructure only).  CPU fp32 restatement of the snowification/decolor package
(FP = snowification/diffusion/forward_process_impl.py, SN = snowification/diffusion/diffusion.py): DeColorization
steps (FP:150-195) and Snow.forward (FP:361-372) applied SEQUENTIALLY with the reference's masked stepping loops
(q_sample SN:344-388, sample_one_step SN:195-245, sample SN:259-295).  kornia's rgb_to_grayscale (absent, un-pinned
upstream) is restated as 0.299 R + 0.587 G + 0.114 B: parity unpinned at that single constant triple.
Snow layers are taken from the engine-independent generator restated in cold_diffusion_models_b200.snowification
(host code, pinned against the reference's layers by tests/test_oracle_golden.py)."""
import torch
import torch.nn.functional as F


class DecolorFP:
    def __init__(self, factors, channels=3):
        eye, ones = torch.eye(channels), torch.ones((channels, channels)) / float(channels)
        self.w = [(f * eye + (1.0 - f) * ones)[:, :, None, None] for f in factors]

    def forward(self, x, i, og=None):
        return F.conv2d(x, self.w[i])


class SnowFP:
    def __init__(self, snow, br_coef, fix_brightness=False):
        self.snow, self.br, self.fix = snow, br_coef, fix_brightness           # snow: [T][SB][3][H][W]

    def forward(self, x, i, og=None):
        og_r = (og + 1.) / 2.
        gray = (0.299 * og_r[:, 0:1] + 0.587 * og_r[:, 1:2] + 0.114 * og_r[:, 2:3]) * 1.5 + 0.5
        gray = torch.maximum(og_r, gray)
        base = og_r if self.fix else self.br[i] * og_r + (1 - self.br[i]) * gray
        s = self.snow[i]
        return torch.clip(base + s + torch.rot90(s, k=2, dims=[2, 3]), 0.0, 1.0) * 2. - 1.


class SnowOracle:
    def __init__(self, denoise_fn, fp, *, timesteps, loss_type='l1', sampling_routine='default'):
        self.denoise_fn, self.fp, self.T = denoise_fn, fp, int(timesteps)
        self.loss_type, self.sampling_routine = loss_type, sampling_routine

    def q_sample(self, x_start, t):
        final = x_start.clone()
        sel = torch.where(t != -1)
        x = x_start[sel]
        if x.shape[0] == 0:
            return final
        xs = []
        for i in range(int(torch.max(t)) + 1):
            x = self.fp.forward(x, i, og=final[sel])
            xs.append(x)
        xs = torch.stack(xs)
        # NB (SN:373-378): the reference indexes the UNFILTERED t with the filtered row number, and a -1 there means
        # Python's last element, i.e. the fully degraded image
        final[sel] = torch.stack([xs[int(t[b]), b] for b in range(x.shape[0])])
        return final

    def p_losses(self, x_start, t):
        x_recon = self.denoise_fn(self.q_sample(x_start, t), t)
        return (x_start - x_recon).abs().mean() if self.loss_type == 'l1' else F.mse_loss(x_start, x_recon)

    @torch.no_grad()
    def sample_one_step(self, img, t):
        x = self.denoise_fn(img, t)
        direct = x.clone()
        if self.sampling_routine == 'default':
            a = x.clone()
            cur = torch.zeros_like(t)
            idx = torch.where(cur < t - 1)[0]
            for i in range(int(t.max()) - 1):
                a[idx] = self.fp.forward(a[idx], i, og=x[idx])
                cur += 1
                idx = torch.where(cur < t - 1)[0]
            x = a
        elif self.sampling_routine == 'x0_step_down':
            a = x.clone()
            b = a.clone()
            cur = torch.zeros_like(t)
            idx = torch.where(cur < t)[0]
            for i in range(int(t.max())):
                b = a.clone()
                a[idx] = self.fp.forward(a[idx], i, og=x[idx])
                cur += 1
                idx = torch.where(cur < t)[0]
            x = img - a + b
        return x, direct

    @torch.no_grad()
    def sample(self, batch_size, img, t=None):
        t = t or self.T
        og = img.clone()
        for i in range(t):
            img = self.fp.forward(img, i, og=og)
        xt, direct = img, None
        while t:
            x, cur = self.sample_one_step(img, torch.full((batch_size,), t - 1, dtype=torch.long))
            if direct is None:
                direct = cur
            img = x
            t -= 1
        return {'xt': xt, 'direct_recons': direct, 'recon': img}
