"""Thin Python wrappers over the C ABI (include/colddiff.h): descriptor builders for the tap-list
convolution contract and one function per entry point.  No arithmetic happens here -- torch is
used for device memory and streams only."""
import ctypes as C
import torch

from . import _lib
from ._lib import lib, ptr, stream, call, ConvDesc, ConvSrc, CONV_SIMT, CONV_TC, ACT_NONE, ACT_GELU, ACT_GELU_BWD


# --------------------------------------------------------------------------------------------
# tap lists
# --------------------------------------------------------------------------------------------
def taps_conv(k, pad):
    """forward taps of nn.Conv2d(k, padding=pad): (ky, kx, dy, dx)"""
    return [(ky, kx, ky - pad, kx - pad) for ky in range(k) for kx in range(k)]


def taps_conv_dgrad(k, pad):
    """data-gradient of a stride-1 conv: a conv of dY with the flipped kernel"""
    return [(ky, kx, pad - ky, pad - kx) for ky in range(k) for kx in range(k)]


def taps_convT4_parity(py, px):
    """nn.ConvTranspose2d(4, stride 2, padding 1) (and the dgrad of Conv2d(4, 2, 1)) for the output
    parity class (Y % 2, X % 2) == (py, px): out[2g+p] = sum_k in[g + d] * W[k]"""
    ys = [(1, 0), (3, -1)] if py == 0 else [(0, 1), (2, 0)]
    xs = [(1, 0), (3, -1)] if px == 0 else [(0, 1), (2, 0)]
    return [(ky, kx, dy, dx) for (ky, dy) in ys for (kx, dx) in xs]


def pack_weight(w, taps, mode=0, transposed_conv=False, round_tf32=True, out=None):
    """reference-layout conv weight (O,I,KH,KW) [or ConvTranspose2d (I,O,KH,KW)] -> packed
    [tap][N][K] fp32 (mode 0: N=out ch, K=in ch; mode 1 (data-gradient operand): N=in, K=out)."""
    if transposed_conv:
        I, O, KH, KW = w.shape
    else:
        O, I, KH, KW = w.shape
    n, k = (O, I) if mode == 0 else (I, O)
    nt = len(taps)
    if out is None:
        out = torch.empty((nt, n, k), device=w.device, dtype=torch.float32)
    ky = (C.c_int32 * nt)(*[t[0] for t in taps])
    kx = (C.c_int32 * nt)(*[t[1] for t in taps])
    call('cd_pack_weight', ptr(w), O, I, KH, KW, int(transposed_conv), mode, ky, kx, nt, int(round_tf32),
         ptr(out), stream())
    return out


def unpack_wgrad(packed, taps, w_grad, transposed_conv=False, accumulate=True):
    if transposed_conv:
        I, O, KH, KW = w_grad.shape
    else:
        O, I, KH, KW = w_grad.shape
    nt = len(taps)
    ky = (C.c_int32 * nt)(*[t[0] for t in taps])
    kx = (C.c_int32 * nt)(*[t[1] for t in taps])
    call('cd_unpack_wgrad', ptr(packed), O, I, KH, KW, int(transposed_conv), ky, kx, nt, ptr(w_grad),
         int(accumulate), stream())


# --------------------------------------------------------------------------------------------
# NHWC views: (tensor, channel offset, channels).  tensor is [B, H, W, ld] contiguous.
# --------------------------------------------------------------------------------------------
class View:
    __slots__ = ('t', 'c0', 'C')

    def __init__(self, t, c0=0, C=None):
        self.t, self.c0 = t, c0
        self.C = (t.shape[-1] - c0) if C is None else C

    @property
    def ld(self):
        return self.t.shape[-1]

    @property
    def B(self):
        return self.t.shape[0]

    @property
    def H(self):
        return self.t.shape[1]

    @property
    def W(self):
        return self.t.shape[2]

    def addr(self):
        return self.t.data_ptr() + 4 * self.c0


def make_conv_desc(srcs, out, grid, *, stride=1, Cout, bias=None, resid=None, act=ACT_NONE,
                   round_tf32=False, out_map=(1, 1, 0, 0), out2=None, aux=None):
    """srcs: list of (View, taps[(ky,kx,dy,dx)], packed_w, w_per_batch); out/resid/out2: View;
    grid: (B, Hg, Wg)."""
    d = ConvDesc()
    d.B, d.Hg, d.Wg = grid
    d.sy = d.sx = stride
    d.Cout = Cout
    d.nsrc = len(srcs)
    keep = []
    for i, (v, taps, w, wpb) in enumerate(srcs):
        s = d.s[i]
        s.src = v.addr(); s.ld = v.ld; s.C = v.C; s.H = v.H; s.W = v.W
        s.ntaps = len(taps)
        for j, tp in enumerate(taps):
            s.dy[j] = tp[2]; s.dx[j] = tp[3]
        s.w = w.data_ptr(); s.w_per_batch = int(wpb)
        keep.append(w)
    d.out = out.addr(); d.out_ld = out.ld; d.Ho = out.H; d.Wo = out.W
    d.oys, d.oxs, d.oy0, d.ox0 = out_map
    d.bias = bias.data_ptr() if bias is not None else None
    if resid is not None:
        d.resid = resid.addr(); d.resid_ld = resid.ld
    d.act = act
    d.round_tf32 = int(round_tf32)
    if out2 is not None:
        d.out2 = out2.addr(); d.out2_ld = out2.ld
    if aux is not None:
        d.aux = aux.addr(); d.aux_ld = aux.ld
    d._keep = keep
    return d


def conv_fwd(desc, impl=CONV_TC):
    call('cd_conv_fwd', C.byref(desc), impl, stream())


def conv_wgrad(desc, dout, dw_packed, db=None, impl=CONV_SIMT):
    call('cd_conv_wgrad', C.byref(desc), C.c_void_p(dout.addr()), dout.ld, ptr(dw_packed), ptr(db), impl, stream())
