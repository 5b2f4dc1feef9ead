"""Host-side construction of the cumulative blur operators A_t (init-time only).

Reference: GaussianDiffusion.get_kernels / get_conv / blur (DB:348-389) build T depthwise nn.Conv2d
modules whose weight is outer(g_i, g_i), g_i = torchgeometry get_gaussian_kernel (restated below from
torchgeometry 0.1.2 image/gaussian.py; the package is absent from the image and un-pinned upstream).
Each step is separable with circular or reflect boundary handling, so applying steps 0..t to a plane X
equals A_t X A_t^T with A_t = K_t K_{t-1} ... K_0 and K_i the S x S 1-D blur matrix of step i.
The products are accumulated in float64 and stored as fp32 [T][S][S] for cd_blur_apply / cd_blur_step_down.
"""
import numpy as np
import torch


def gaussian_taps(ksize, sigma):
    """fp32 taps exactly as the reference obtains them (exponent in Python double -> fp32 exp -> fp32 normalise)."""
    if not isinstance(ksize, int) or ksize % 2 == 0 or ksize <= 0:
        raise TypeError("ksize must be an odd positive integer. Got {}".format(ksize))

    def gauss_fcn(x):
        return -(x - ksize // 2) ** 2 / float(2 * sigma ** 2)
    g = torch.stack([torch.exp(torch.tensor(gauss_fcn(x))) for x in range(ksize)])
    return g / g.sum()


def blur_schedule(blur_routine, timesteps, kernel_size, kernel_std):
    """(ksize, sigma, padding_mode) per step -- DB:363-389.  Unknown routine -> [] like the reference."""
    out = []
    for i in range(timesteps):
        if blur_routine == 'Incremental':
            out.append((kernel_size, kernel_std * (i + 1), 'circular'))
        elif blur_routine == 'Constant':
            out.append((kernel_size, kernel_std, 'circular'))
        elif blur_routine == 'Constant_reflect':
            out.append((kernel_size, kernel_std, 'reflect'))
        elif blur_routine == 'Exponential_reflect':
            out.append((kernel_size, np.exp(kernel_std * i), 'reflect'))
        elif blur_routine == 'Exponential':
            out.append((kernel_size, np.exp(kernel_std * i), 'circular'))
        elif blur_routine == 'Individual_Incremental':
            ks = 2 * i + 1
            out.append((ks, 2 * ks, 'circular'))
        elif blur_routine == 'Special_6_routine':
            out.append((11, i / 100 + 0.35, 'reflect'))
    return out


def blur_matrix(taps, S, mode):
    """S x S float64 matrix of the 1-D convolution with `taps` and nn.Conv2d padding_mode `mode`."""
    k = len(taps)
    pad = int((k - 1) / 2)
    if pad >= S and mode == 'reflect':
        raise ValueError("reflect padding needs kernel half-width < image size")
    K = np.zeros((S, S), dtype=np.float64)
    rows = np.arange(S)
    for a in range(k):
        j = rows + a - pad
        if mode == 'circular':
            j = np.mod(j, S)
        else:  # reflect (no edge repeat)
            j = np.where(j < 0, -j, j)
            j = np.where(j >= S, 2 * (S - 1) - j, j)
        np.add.at(K, (rows, j), float(taps[a]))
    return K


def build_blur_operators(blur_routine, timesteps, kernel_size, kernel_std, S):
    """-> (ops_cum fp32 [T][S][S], ops_single or None, taps, schedule).
    ops_cum[i] = K_i ... K_0 is the degradation after steps 0..i (q_sample DB:934-941; sample DB:405-407).
    ops_single[i] = K_i alone is what `sample` applies for 'Individual_Incremental' (DB:401-402, 429-430)."""
    sched = blur_schedule(blur_routine, timesteps, kernel_size, kernel_std)
    taps = [gaussian_taps(k, s) for (k, s, _) in sched]
    ops = np.zeros((len(sched), S, S), dtype=np.float32)
    single = np.zeros((len(sched), S, S), dtype=np.float32) if blur_routine == 'Individual_Incremental' else None
    A = np.eye(S, dtype=np.float64)
    for i, ((k, s, mode), g) in enumerate(zip(sched, taps)):
        K = blur_matrix(g.double().numpy(), S, mode)
        if single is not None:
            single[i] = K.astype(np.float32)
        A = K @ A
        ops[i] = A.astype(np.float32)
    return torch.from_numpy(ops), (torch.from_numpy(single) if single is not None else None), taps, sched
