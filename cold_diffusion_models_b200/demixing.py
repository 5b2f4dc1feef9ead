"""Drop-in `GaussianDiffusion` of demixing_diffusion_pytorch -- "animorphosis": the end state x2 is an image of a
second dataset instead of Gaussian noise (reference:
demixing-diffusion-pytorch/demixing_diffusion_pytorch/demixing_diffusion_pytorch.py:310-522, "DM").

The schedule, `q_sample`, `p_losses`, `forward(x1, x2)` and `sample` are those of the denoising package (cosine
alpha-bar lerp, DM:497-502); what differs is `gen_sample` (fixed end image + optional noise on the start, DM:384-414),
`forward_and_backward(img1, img2)` (DM:416-460) and the two-list return of `all_sample` (DM:462-495)."""
import torch

from .denoising import GaussianDiffusion as _NoiseDiffusion


class GaussianDiffusion(_NoiseDiffusion):
    @torch.no_grad()
    def gen_sample(self, batch_size=16, img=None, noise_level=0, t=None):
        """DM:384-414 -> (noise, direct_recons, img): Algorithm 2 with x2 fixed to the input image"""
        self.denoise_fn.eval()
        if t is None:
            t = self.num_timesteps
        noise = img.contiguous().float()
        start = noise + torch.randn_like(noise) * noise_level
        direct_recons, out = self._reverse(batch_size, start, t, 1, noise)
        return noise, direct_recons, out

    @torch.no_grad()
    def forward_and_backward(self, batch_size=16, img1=None, img2=None, t=None, times=None, eval=True):
        """DM:416-460 -> (Forward, Backward, img)"""
        self.denoise_fn.eval()
        if t is None:
            t = self.num_timesteps
        img = img1.contiguous().float()
        noise = img2.contiguous().float()
        Forward = [img]
        n_img = img
        for i in range(t):
            step = torch.full((batch_size,), i, dtype=torch.long, device=img.device)
            n_img = self.q_sample(x_start=img, x_end=noise, t=step)
            Forward.append(n_img)
        Backward = []
        _, out = self._reverse(batch_size, n_img, t, 1, noise, collect=lambda x1, cur, s: Backward.append(cur))
        return Forward, Backward, out

    @torch.no_grad()
    def all_sample(self, batch_size=16, img=None, t=None, times=None, eval=True):
        """DM:462-495 -> (X1_0s, X_ts) as CPU tensors (x2 estimated from x_t as in `sample`)"""
        X1_0s, _, X_ts = super().all_sample(batch_size=batch_size, img=img, t=t, times=times, eval=eval)
        return X1_0s, X_ts
