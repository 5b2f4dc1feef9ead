"""ORACLE (test infrastructure only).  CPU fp32 restatement of the snowification/decolor package
(FP = snowification/diffusion/forward_process_impl.py, SN = snowification/diffusion/diffusion.py): DeColorization
steps (FP:150-195) and Snow.forward (FP:361-372) applied SEQUENTIALLY with the reference's masked stepping loops
(q_sample SN:344-388, sample_one_step SN:195-245, sample SN:259-295).  kornia's rgb_to_grayscale (absent, un-pinned
upstream) is restated as 0.299 R + 0.587 G + 0.114 B: parity unpinned at that single constant triple.
Snow layers: `generate_snow_layers` below restates Snow.generate_snow_layer (FP:252-355) with the reference's own tools
(numpy generator, scipy.ndimage.zoom, torch CPU convolutions); it is pinned against layers recorded from the reference
(tests/golden/snow_small.npz) by tests/test_oracle_golden.py and is what the device generator (cd_snow_layers) is checked
against."""
import numpy as np
import torch
import torch.nn.functional as F

SNOW_LEVELS = {   # FP:261-293: c, (thres start,end), (motion-blur sigma start,end), (brightness start,end)
    1: ((0.1, 0.3, 3, 0.5, 5, 4, 0.8), (0.7, 0.3), (0.5, 5.0), (0.95, 0.7)),
    2: ((0.55, 0.3, 2.5, 0.85, 11, 12, 0.55), (1.15, 0.7), (0.05, 12), (0.95, 0.55)),
    3: ((0.55, 0.3, 2.5, 0.7, 11, 16, 0.4), (1.15, 0.7), (0.05, 16), (0.95, 0.4)),
    4: ((0.55, 0.3, 2.5, 0.55, 11, 20, 0.3), (1.15, 0.55), (0.05, 20), (0.95, 0.3)),
}


def clipped_zoom(img, zoom_factor):
    # FP:32-42 (scipy.ndimage.zoom, order 1, centre crop, centre trim)
    from scipy.ndimage import zoom as scizoom
    h = img.shape[0]
    ch = int(np.ceil(h / zoom_factor))
    top = (h - ch) // 2
    img = scizoom(img[top:top + ch, top:top + ch], (zoom_factor, zoom_factor, 1), order=1)
    trim = (img.shape[0] - h) // 2
    return img[trim:trim + h, trim:trim + h]


def gaussian_taps(ksize, sigma):
    # torchgeometry.image.get_gaussian_kernel: exponent in Python double -> fp32 exp -> fp32 normalise
    g = torch.stack([torch.exp(torch.tensor(-(x - ksize // 2) ** 2 / float(2 * sigma ** 2))) for x in range(ksize)])
    return g / g.sum()


def generate_snow_layers(image_size, snow_level=1, num_timesteps=50, random_snow=False, single_snow=False, batch_size=32):
    """FP:252-355 on the host -> (snow [T][SB][3][H][W], brightness coefficients [T]).  Consumes the numpy / torch global
    generators exactly as the reference does."""
    if not random_snow:
        rstate = np.random.get_state()
        np.random.seed(123321)
    c, thr, mbs, brc = SNOW_LEVELS[snow_level]
    T = num_timesteps
    thres = torch.linspace(thr[0], thr[1], T).tolist()
    sigmas = torch.linspace(mbs[0], mbs[1], T).tolist()
    br = torch.linspace(brc[0], brc[1], T).tolist()
    if single_snow:
        sb = [clipped_zoom(np.random.normal(size=image_size, loc=c[0], scale=c[1])[..., np.newaxis], c[2]) for _ in range(batch_size)]
        base = np.concatenate(sb, axis=2)
    else:
        base = clipped_zoom(np.random.normal(size=image_size, loc=c[0], scale=c[1])[..., np.newaxis], c[2])
    vertical_snow = bool(np.random.uniform() > 0.5)
    snow = []
    for i in range(T):
        layer = torch.Tensor(base).clone()
        layer[layer < thres[i]] = 0
        layer = torch.clip(layer, 0, 1).permute((2, 0, 1)).unsqueeze(1)          # [SB][1][H][W]
        mk = torch.zeros((c[4], c[4]))
        mk[int(c[4] / 2)] = gaussian_taps(c[4], sigmas[i])
        hk = mk[None, None, :].repeat(3, 1, 1, 1)
        vk = torch.rot90(mk, k=1, dims=[0, 1])[None, None, :].repeat(3, 1, 1, 1)
        vs = F.conv2d(layer, vk, padding='same')
        hs = F.conv2d(layer, hk, padding='same')
        if single_snow:
            vidx = torch.randperm(layer.shape[0])[:int(layer.shape[0] / 2)]
            layer = hs
            layer[vidx] = vs[vidx]
        elif vertical_snow:
            layer = vs
        else:
            layer = hs
        snow.append(layer)
    if not random_snow:
        np.random.set_state(rstate)
    return torch.stack(snow).contiguous(), br


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
