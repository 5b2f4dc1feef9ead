"""ORACLE (test infrastructure only).  CPU fp32 restatement of the demixing ("animorphosis") package
(DM = demixing-diffusion-pytorch/demixing_diffusion_pytorch/demixing_diffusion_pytorch.py): the cosine alpha-bar
lerp towards an image of a second dataset; q_sample (DM:497-502), p_losses (DM:504-516), sample (DM:343-375),
gen_sample (DM:384-414), forward_and_backward (DM:416-460), all_sample (DM:462-495).  Pinned against the unmodified
reference by tests/test_oracle_golden.py (tests/golden/demixing_small.npz)."""
import torch

from denoise_oracle import DenoiseOracle


class DemixingOracle(DenoiseOracle):
    def gen_sample(self, batch_size, img, noise=None, t=None):
        """`noise` = randn_like(img) * noise_level drawn by the caller (the reference draws it inside)"""
        t = t or self.num_timesteps
        start = img if noise is None else img + noise
        d, out = self._reverse(batch_size, start, t, img)
        return img, d, out

    @torch.no_grad()
    def forward_and_backward(self, batch_size, img1, img2, t=None):
        t = t or self.num_timesteps
        Forward = [img1]
        n_img = img1
        for i in range(t):
            n_img = self.q_sample(img1, img2, torch.full((batch_size,), i, dtype=torch.long))
            Forward.append(n_img)
        Backward, img = [], n_img
        while t:
            step = torch.full((batch_size,), t - 1, dtype=torch.long)
            x1_bar = self.denoise_fn(img, step)
            Backward.append(img)
            xt_bar = self.q_sample(x1_bar, img2, step)
            xt_sub1 = x1_bar
            if t - 1 != 0:
                xt_sub1 = self.q_sample(x1_bar, img2, torch.full((batch_size,), t - 2, dtype=torch.long))
            img = img - xt_bar + xt_sub1
            t -= 1
        return Forward, Backward, img

    @torch.no_grad()
    def all_sample(self, batch_size, img, t=None):
        t = t or self.num_timesteps
        X1_0s, X_ts = [], []
        while t:
            step = torch.full((batch_size,), t - 1, dtype=torch.long)
            x1_bar = self.denoise_fn(img, step)
            x2_bar = self.get_x2_bar_from_xt(x1_bar, img, step)
            X1_0s.append(x1_bar)
            X_ts.append(img)
            xt_bar = self.q_sample(x1_bar, x2_bar, step)
            xt_sub1 = x1_bar
            if t - 1 != 0:
                xt_sub1 = self.q_sample(x1_bar, x2_bar, torch.full((batch_size,), t - 2, dtype=torch.long))
            img = img - xt_bar + xt_sub1
            t -= 1
        return X1_0s, X_ts
