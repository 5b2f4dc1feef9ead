"""TEST INFRASTRUCTURE: a CPU emulation of the libcolddiff entry points used by the DDPM-style `Model` (numpy on raw addresses,
same argument lists as include/colddiff.h).  It lets the host-side schedules in cold_diffusion_models_b200/model2.py and
model2_train.py (buffer planning, gradient routing, tap lists, packed-weight layouts) run on CPU tensors, so their LOGIC can be
checked against the reference-generated goldens without a GPU.  It says nothing about the CUDA kernels themselves -- those are
checked by the `-m gpu` tests through the real library.  Never imported by the product."""
import ctypes as C
import math
import os
import numpy as np


TF32_EMULATION = False      # True: tap-list convolutions launched with impl = CD_CONV_TC round both operands to TF32 (RN), like the
                            # TFLOAT32 tensor maps do -- used to estimate the TF32-path error of a test before it runs on a GPU


def _tf32(a):
    i = np.ascontiguousarray(a, dtype=np.float32).view(np.int32)
    return ((i + 0x1000) & ~0x1FFF).view(np.float32)


def _v(a):
    if isinstance(a, (int, float)):
        return a
    if hasattr(a, 'value'):
        return 0 if a.value is None else a.value
    return a


def _arr(addr, shape, strides_elems):
    """numpy view of float32 memory at `addr` with the given shape / element strides"""
    addr = _v(addr)
    if not addr:
        return None
    n = 1 + sum((s - 1) * st for s, st in zip(shape, strides_elems))
    flat = np.ctypeslib.as_array((C.c_float * n).from_address(addr))
    return np.lib.stride_tricks.as_strided(flat, shape=shape, strides=tuple(4 * st for st in strides_elems))


def _rows(addr, rows, Cc, ld):
    return _arr(addr, (rows, Cc), (ld, 1))


def _nhwc(addr, B, H, W, Cc, ld):
    return _arr(addr, (B, H, W, Cc), (H * W * ld, W * ld, ld, 1))


def _i64(addr, n):
    return np.ctypeslib.as_array((C.c_int64 * n).from_address(_v(addr)))


def _gelu(x):
    from scipy.special import erf
    return 0.5 * x * (1.0 + erf(x / math.sqrt(2.0)))


def _gelu_grad(x):
    from scipy.special import erf
    return 0.5 * (1.0 + erf(x / math.sqrt(2.0))) + x * np.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


def _swish(z):
    return z / (1.0 + np.exp(-z))


def _swish_grad(z):
    s = 1.0 / (1.0 + np.exp(-z))
    return s * (1.0 + z * (1.0 - s))


# ------------------------------------------------------------------------------------------------------------------
# tap-list convolution (forward / data gradient) and its weight gradient
# ------------------------------------------------------------------------------------------------------------------
def _gather(src, B, Hg, Wg, sy, sx, dy, dx):
    """src[b, gy*sy+dy, gx*sx+dx, :] with zero fill outside the image -> [B, Hg, Wg, C] (float64)"""
    H, W = src.shape[1], src.shape[2]
    out = np.zeros((B, Hg, Wg, src.shape[3]), dtype=np.float64)
    iy = np.arange(Hg) * sy + dy
    ix = np.arange(Wg) * sx + dx
    vy = np.where((iy >= 0) & (iy < H))[0]
    vx = np.where((ix >= 0) & (ix < W))[0]
    if len(vy) and len(vx):
        out[np.ix_(np.arange(B), vy, vx)] = src[np.ix_(np.arange(B), iy[vy], ix[vx])]
    return out


def cd_conv_fwd(desc, impl, stream):
    d = desc._obj if hasattr(desc, '_obj') else desc
    B, Hg, Wg = d.B, d.Hg, d.Wg
    acc = np.zeros((B, Hg, Wg, d.Cout), dtype=np.float64)
    tf32 = TF32_EMULATION and _v(impl) == 1
    for si in range(d.nsrc):
        s = d.s[si]
        src = _nhwc(s.src, B, s.H, s.W, s.C, s.ld)
        if tf32 and s.C % 32 == 0:
            src = _tf32(src)
        if s.w_per_batch:
            w = _arr(s.w, (B, s.ntaps, d.Cout, s.C), (s.ntaps * d.Cout * s.C, d.Cout * s.C, s.C, 1))
        else:
            w = _arr(s.w, (s.ntaps, d.Cout, s.C), (d.Cout * s.C, s.C, 1))
        if tf32 and s.C % 32 == 0:
            w = _tf32(w)
        for t in range(s.ntaps):
            g = _gather(src, B, Hg, Wg, d.sy, d.sx, s.dy[t], s.dx[t])
            if s.w_per_batch:
                acc += np.einsum('bhwc,boc->bhwo', g, w[:, t].astype(np.float64))
            else:
                acc += np.einsum('bhwc,oc->bhwo', g, w[t].astype(np.float64))
    if d.bias:
        acc += _arr(d.bias, (d.Cout,), (1,)).astype(np.float64)
    out_full = _nhwc(d.out, B, d.Ho, d.Wo, d.Cout, d.out_ld)
    ys = np.arange(Hg) * d.oys + d.oy0
    xs = np.arange(Wg) * d.oxs + d.ox0
    sel = np.ix_(np.arange(B), ys, xs)
    if d.resid:
        acc += _nhwc(d.resid, B, d.Ho, d.Wo, d.Cout, d.resid_ld)[sel].astype(np.float64)
    if d.out2:
        _nhwc(d.out2, B, d.Ho, d.Wo, d.Cout, d.out2_ld)[sel] = acc.astype(np.float32)       # pre-activation, kept for the backward
    if d.act == 1:                                          # CD_ACT_GELU
        acc = _gelu(acc)
    elif d.act == 2:                                        # CD_ACT_GELU_BWD: multiply by gelu'(saved pre-activation)
        acc = acc * _gelu_grad(_nhwc(d.aux, B, d.Ho, d.Wo, d.Cout, d.aux_ld)[sel].astype(np.float64))
    out_full[sel] = acc.astype(np.float32)
    return 0


def cd_conv_wgrad(desc, dout, dout_ld, dw, db, impl, stream):
    d = desc._obj if hasattr(desc, '_obj') else desc
    B, Hg, Wg = d.B, d.Hg, d.Wg
    s = d.s[0]
    src = _nhwc(s.src, B, s.H, s.W, s.C, s.ld)
    dy_full = _nhwc(dout, B, d.Ho, d.Wo, d.Cout, _v(dout_ld))
    ys = np.arange(Hg) * d.oys + d.oy0
    xs = np.arange(Wg) * d.oxs + d.ox0
    dy = dy_full[np.ix_(np.arange(B), ys, xs)].astype(np.float64)
    if s.w_per_batch:
        w = _arr(dw, (B, s.ntaps, d.Cout, s.C), (s.ntaps * d.Cout * s.C, d.Cout * s.C, s.C, 1))
    else:
        w = _arr(dw, (s.ntaps, d.Cout, s.C), (d.Cout * s.C, s.C, 1))
    for t in range(s.ntaps):
        g = _gather(src, B, Hg, Wg, d.sy, d.sx, s.dy[t], s.dx[t])
        if s.w_per_batch:
            w[:, t] += np.einsum('bhwo,bhwc->boc', dy, g).astype(np.float32)
        else:
            w[t] += np.einsum('bhwo,bhwc->oc', dy, g).astype(np.float32)
    if _v(db):
        _arr(db, (d.Cout,), (1,))[:] += dy.sum(axis=(0, 1, 2)).astype(np.float32)
    return 0


def cd_pack_weight(w, O, I, KH, KW, transposed_conv, mode, ky, kx, ntaps, round_tf32, packed, stream):
    if transposed_conv:
        W = _arr(w, (I, O, KH, KW), (O * KH * KW, KH * KW, KW, 1)).transpose(1, 0, 2, 3)      # -> [o][i][ky][kx]
    else:
        W = _arr(w, (O, I, KH, KW), (I * KH * KW, KH * KW, KW, 1))
    N, K = (O, I) if mode == 0 else (I, O)
    P = _arr(packed, (ntaps, N, K), (N * K, K, 1))
    for t in range(ntaps):
        m = W[:, :, ky[t], kx[t]]
        P[t] = m if mode == 0 else m.T
    return 0


def cd_unpack_wgrad(packed, O, I, KH, KW, transposed_conv, ky, kx, ntaps, w_grad, accumulate, stream):
    P = _arr(packed, (ntaps, O, I), (O * I, I, 1))
    if transposed_conv:                                     # nn.ConvTranspose2d stores (I, O, KH, KW)
        G = _arr(w_grad, (I, O, KH, KW), (O * KH * KW, KH * KW, KW, 1)).transpose(1, 0, 2, 3)
    else:
        G = _arr(w_grad, (O, I, KH, KW), (I * KH * KW, KH * KW, KW, 1))
    for t in range(ntaps):
        if accumulate:
            G[:, :, ky[t], kx[t]] += P[t]
        else:
            G[:, :, ky[t], kx[t]] = P[t]
    return 0


class _RepackJob(C.Structure):          # CdRepackJob of include/colddiff.h, declared here independently of the package
    _fields_ = [('src', C.c_void_p), ('dst', C.c_void_p), ('O', C.c_int32), ('I', C.c_int32), ('KH', C.c_int32), ('KW', C.c_int32),
                ('transposed_conv', C.c_int32), ('mode', C.c_int32), ('ntaps', C.c_int32), ('round_tf32', C.c_int32),
                ('ky', C.c_int32 * 16), ('kx', C.c_int32 * 16), ('block0', C.c_int32), ('nblocks', C.c_int32)]


def _jobs(jobs, njobs, total_blocks):
    arr = (_RepackJob * njobs).from_address(_v(jobs))
    b0 = 0
    for j in arr:                       # the block ranges must tile [0, total_blocks) in job order
        assert j.block0 == b0 and j.nblocks >= 1
        b0 += j.nblocks
    assert b0 == total_blocks
    return arr


def cd_pack_weight_batched(jobs, njobs, total_blocks, stream):
    for j in _jobs(jobs, njobs, total_blocks):
        if not j.transposed_conv and j.mode == 0 and j.KH * j.KW <= 16:
            assert j.nblocks == (j.O * j.I + 255) // 256
        rc = cd_pack_weight(j.src, j.O, j.I, j.KH, j.KW, j.transposed_conv, j.mode, list(j.ky), list(j.kx), j.ntaps, j.round_tf32,
                            j.dst, stream)
        assert rc == 0
    return 0


def cd_unpack_wgrad_batched(jobs, njobs, total_blocks, accumulate, clear_src, stream):
    for j in _jobs(jobs, njobs, total_blocks):
        if not j.transposed_conv and j.KH * j.KW <= 16:
            assert j.nblocks == (j.O * j.I + 255) // 256
        rc = cd_unpack_wgrad(j.src, j.O, j.I, j.KH, j.KW, j.transposed_conv, list(j.ky), list(j.kx), j.ntaps, j.dst, accumulate, stream)
        assert rc == 0
        if clear_src:
            _arr(j.src, (j.ntaps * j.O * j.I,), (1,))[:] = 0.0
    return 0


def cd_transpose_taps_batched(jobs, njobs, total_blocks, stream):
    for j in _jobs(jobs, njobs, total_blocks):
        assert j.nblocks == j.ntaps * ((j.O + 31) // 32) * ((j.I + 31) // 32)
        ntap_src = max(j.ky[t] * j.KW + j.kx[t] for t in range(j.ntaps)) + 1
        S = _arr(j.src, (ntap_src, j.O, j.I), (j.O * j.I, j.I, 1))
        D = _arr(j.dst, (j.ntaps, j.I, j.O), (j.I * j.O, j.O, 1))
        for t in range(j.ntaps):
            D[t] = S[j.ky[t] * j.KW + j.kx[t]].T
    return 0


# ------------------------------------------------------------------------------------------------------------------
# GroupNorm, dropout, softmax, resampling, small dense pieces
# ------------------------------------------------------------------------------------------------------------------
def _gn_stats(x, B, HW, Cc, groups, eps):
    xg = x.reshape(B, HW, groups, Cc // groups)
    mean = xg.mean(axis=(1, 3), keepdims=True)
    var = xg.var(axis=(1, 3), keepdims=True)
    return xg, mean, 1.0 / np.sqrt(var + eps)


def cd_groupnorm_fwd(x, x_ld, B, HW, Cc, groups, cond, cond_ld, gamma, beta, eps, swish, y, y_ld, stream):
    HW = _v(HW); eps = _v(eps)
    X = _arr(x, (B, HW, Cc), (HW * x_ld, x_ld, 1)).astype(np.float64)
    if _v(cond):
        X = X + _arr(cond, (B, Cc), (cond_ld, 1)).astype(np.float64)[:, None, :]
    xg, mean, rstd = _gn_stats(X, B, HW, Cc, groups, eps)
    xh = ((xg - mean) * rstd).reshape(B, HW, Cc)
    z = xh * _arr(gamma, (Cc,), (1,)).astype(np.float64) + _arr(beta, (Cc,), (1,)).astype(np.float64)
    _arr(y, (B, HW, Cc), (HW * y_ld, y_ld, 1))[:] = (_swish(z) if swish else z).astype(np.float32)
    return 0


def cd_groupnorm_bwd(x, x_ld, B, HW, Cc, groups, cond, cond_ld, gamma, beta, eps, swish, dy, dy_ld, dx, dx_ld, dgamma, dbeta,
                     dcond, dcond_ld, stream):
    HW = _v(HW); eps = _v(eps)
    X = _arr(x, (B, HW, Cc), (HW * x_ld, x_ld, 1)).astype(np.float64)
    if _v(cond):
        X = X + _arr(cond, (B, Cc), (cond_ld, 1)).astype(np.float64)[:, None, :]
    g = _arr(gamma, (Cc,), (1,)).astype(np.float64); bt = _arr(beta, (Cc,), (1,)).astype(np.float64)
    xg, mean, rstd = _gn_stats(X, B, HW, Cc, groups, eps)
    xh = ((xg - mean) * rstd).reshape(B, HW, Cc)
    DY = _arr(dy, (B, HW, Cc), (HW * dy_ld, dy_ld, 1)).astype(np.float64)
    dz = DY * _swish_grad(xh * g + bt) if swish else DY
    _arr(dgamma, (Cc,), (1,))[:] += (dz * xh).sum(axis=(0, 1)).astype(np.float32)
    _arr(dbeta, (Cc,), (1,))[:] += dz.sum(axis=(0, 1)).astype(np.float32)
    dxh = (dz * g).reshape(B, HW, groups, Cc // groups)
    xhg = xh.reshape(B, HW, groups, Cc // groups)
    ma = dxh.mean(axis=(1, 3), keepdims=True)
    mb = (dxh * xhg).mean(axis=(1, 3), keepdims=True)
    DX = (rstd * (dxh - ma - xhg * mb)).reshape(B, HW, Cc)
    _arr(dx, (B, HW, Cc), (HW * dx_ld, dx_ld, 1))[:] = DX.astype(np.float32)
    if _v(dcond):
        _arr(dcond, (B, Cc), (dcond_ld, 1))[:] = DX.sum(axis=1).astype(np.float32)
    return 0


def _uniform01(seed, idx):
    with np.errstate(over='ignore'):
        h = np.uint64(seed) ^ (idx.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15))
        h ^= h >> np.uint64(33); h *= np.uint64(0xff51afd7ed558ccd); h ^= h >> np.uint64(33)
        h *= np.uint64(0xc4ceb9fe1a85ec53); h ^= h >> np.uint64(33)
    return (h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def cd_dropout(x, x_ld, npix, Cc, p, seed, y, y_ld, stream):
    npix, p, seed = _v(npix), np.float32(_v(p)), _v(seed)
    X = _rows(x, npix, Cc, x_ld).copy()
    keep = (_uniform01(seed, np.arange(npix * Cc)) >= p).reshape(npix, Cc)
    _rows(y, npix, Cc, y_ld)[:] = X * keep * np.float32(1.0 / (1.0 - p))
    return 0


def cd_softmax_rows(s, ld, rows, n, scale, stream):
    S = _rows(s, _v(rows), n, ld)
    z = S.astype(np.float64) * _v(scale)
    z = np.exp(z - z.max(axis=1, keepdims=True))
    S[:] = (z / z.sum(axis=1, keepdims=True)).astype(np.float32)
    return 0


def cd_softmax_bwd_rows(s, ds, ld, rows, n, scale, stream):
    S = _rows(s, _v(rows), n, ld).astype(np.float64)
    D = _rows(ds, _v(rows), n, ld)
    d = D.astype(np.float64)
    D[:] = (S * (d - (d * S).sum(axis=1, keepdims=True)) * _v(scale)).astype(np.float32)
    return 0


def cd_transpose_batched(src, ld, B, R, Cc, dst, stream):
    A = _arr(src, (B, R, Cc), (R * ld, ld, 1))
    _arr(dst, (B, Cc, R), (Cc * R, R, 1))[:] = A.transpose(0, 2, 1)
    return 0


def cd_upsample_nearest2x(x, x_ld, B, H, W, Cc, y, y_ld, stream):
    X = _nhwc(x, B, H, W, Cc, x_ld)
    _nhwc(y, B, 2 * H, 2 * W, Cc, y_ld)[:] = X.repeat(2, axis=1).repeat(2, axis=2)
    return 0


def cd_upsample_nearest2x_bwd(dy, dy_ld, B, H, W, Cc, dx, dx_ld, stream):
    D = _nhwc(dy, B, 2 * H, 2 * W, Cc, dy_ld).astype(np.float64)
    _nhwc(dx, B, H, W, Cc, dx_ld)[:] = (D[:, 0::2, 0::2] + D[:, 0::2, 1::2] + D[:, 1::2, 0::2] + D[:, 1::2, 1::2]).astype(np.float32)
    return 0


def cd_add(a, a_ld, b, b_ld, out, out_ld, npix, Cc, stream):
    npix = _v(npix)
    res = _rows(a, npix, Cc, a_ld) + _rows(b, npix, Cc, b_ld)
    _rows(out, npix, Cc, out_ld)[:] = res
    return 0


def cd_colsum(x, ld, rows, Cc, out, stream):
    _arr(out, (Cc,), (1,))[:] += _rows(x, _v(rows), Cc, ld).astype(np.float64).sum(axis=0).astype(np.float32)
    return 0


def cd_small_gemm(A, lda, transA, Bm, ldb, transB, Cm, ldc, M, N, K, accumulate, stream):
    a = _arr(A, (K, M), (lda, 1)).T if transA else _arr(A, (M, K), (lda, 1))
    b = _arr(Bm, (N, K), (ldb, 1)).T if transB else _arr(Bm, (K, N), (ldb, 1))
    c = _arr(Cm, (M, N), (ldc, 1))
    r = (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32)
    c[:] = c + r if accumulate else r
    return 0


def cd_swish(dy, pre, n, y, act_out, stream):
    n = _v(n)
    p = _arr(pre, (n,), (1,)).astype(np.float64)
    if _v(act_out):
        _arr(act_out, (n,), (1,))[:] = _swish(p).astype(np.float32)
    if _v(y):
        _arr(y, (n,), (1,))[:] = (_arr(dy, (n,), (1,)).astype(np.float64) * _swish_grad(p)).astype(np.float32)
    return 0


def cd_timestep_embedding(t, B, dim, emb, stream):
    tt = _i64(t, B).astype(np.float32)
    half = dim // 2
    f = np.exp(np.arange(half, dtype=np.float32) * np.float32(-(math.log(10000) / (half - 1))))
    a = tt[:, None] * f[None, :]
    E = _arr(emb, (B, dim), (dim, 1))
    E[:, :half] = np.sin(a); E[:, half:2 * half] = np.cos(a)
    if dim % 2:
        E[:, -1] = 0
    return 0


def cd_time_mlp2_fwd(t, B, dim, hid, tdim, act, w1, b1, w2, b2, wc, bc, sumC, temb, cond_all, stream):
    tt = _i64(t, B).astype(np.float64)
    half = dim // 2
    f = np.exp(-(math.log(10000.0) / (half - 1)) * np.arange(half))
    emb = np.zeros((B, dim))
    emb[:, :half] = np.sin(tt[:, None] * f); emb[:, half:2 * half] = np.cos(tt[:, None] * f)
    fn = _swish if act else _gelu
    h = fn(emb @ _arr(w1, (hid, dim), (dim, 1)).astype(np.float64).T + _arr(b1, (hid,), (1,)))
    te = h @ _arr(w2, (tdim, hid), (hid, 1)).astype(np.float64).T + _arr(b2, (tdim,), (1,))
    if _v(temb):
        _arr(temb, (B, tdim), (tdim, 1))[:] = te.astype(np.float32)
    ca = fn(te) @ _arr(wc, (sumC, tdim), (tdim, 1)).astype(np.float64).T + _arr(bc, (sumC,), (1,))
    _arr(cond_all, (B, sumC), (sumC, 1))[:] = ca.astype(np.float32)
    return 0


def cd_linear_fwd(x, K, w, bias, M, N, y, stream):
    r = _arr(x, (M, K), (K, 1)).astype(np.float64) @ _arr(w, (N, K), (K, 1)).astype(np.float64).T
    if _v(bias):
        r = r + _arr(bias, (N,), (1,)).astype(np.float64)
    _arr(y, (M, N), (N, 1))[:] = r.astype(np.float32)
    return 0


def cd_nchw_to_nhwc(x, B, Cc, H, W, out, ld, stream):
    X = _arr(x, (B, Cc, H, W), (Cc * H * W, H * W, W, 1))
    O = _nhwc(out, B, H, W, ld, ld)
    O[:] = 0
    O[..., :Cc] = X.transpose(0, 2, 3, 1)
    return 0


def cd_nhwc_to_nchw(x, ld, B, H, W, Cc, out, stream):
    _arr(out, (B, Cc, H, W), (Cc * H * W, H * W, W, 1))[:] = _nhwc(x, B, H, W, Cc, ld).transpose(0, 3, 1, 2)
    return 0



# ------------------------------------------------------------------------------------------------------------------
# ConvNeXt Unet entry points (engine.py / engine_bwd.py)
# ------------------------------------------------------------------------------------------------------------------
def cd_time_mlp_fwd(t, B, dim, w1, b1, w2, b2, wc, bc, sumC, sinemb, hid_pre, temb, cond_all, stream):
    tt = _i64(t, B).astype(np.float64)
    half = dim // 2
    f = np.exp(-(math.log(10000.0) / (half - 1)) * np.arange(half))
    emb = np.concatenate([np.sin(tt[:, None] * f), np.cos(tt[:, None] * f)], axis=1)
    if _v(sinemb):
        _arr(sinemb, (B, dim), (dim, 1))[:] = emb.astype(np.float32)
    hp = emb @ _arr(w1, (4 * dim, dim), (dim, 1)).astype(np.float64).T + _arr(b1, (4 * dim,), (1,))
    if _v(hid_pre):
        _arr(hid_pre, (B, 4 * dim), (4 * dim, 1))[:] = hp.astype(np.float32)
    te = _gelu(hp) @ _arr(w2, (dim, 4 * dim), (4 * dim, 1)).astype(np.float64).T + _arr(b2, (dim,), (1,))
    _arr(temb, (B, dim), (dim, 1))[:] = te.astype(np.float32)
    if sumC > 0:
        ca = _gelu(te) @ _arr(wc, (sumC, dim), (dim, 1)).astype(np.float64).T + _arr(bc, (sumC,), (1,))
        _arr(cond_all, (B, sumC), (sumC, 1))[:] = ca.astype(np.float32)
    return 0


def _dwconv7(x, x_ld, B, H, W, Cc, w_dw, b_dw, cond, cond_ld, flip, addend, addend_ld):
    X = _nhwc(x, B, H, W, Cc, x_ld)
    Wd = _arr(w_dw, (Cc, 49), (49, 1)).astype(np.float64)
    h = np.zeros((B, H, W, Cc), dtype=np.float64)
    for k in range(49):
        ky, kx = k // 7, k % 7
        h += _gather(X, B, H, W, 1, 1, ky - 3, kx - 3) * Wd[:, 48 - k if flip else k]
    if _v(b_dw):
        h += _arr(b_dw, (Cc,), (1,)).astype(np.float64)
    if _v(cond):
        h += _arr(cond, (B, Cc), (cond_ld, 1)).astype(np.float64)[:, None, None, :]
    if _v(addend):
        h += _nhwc(addend, B, H, W, Cc, addend_ld).astype(np.float64)
    return h


def cd_dwconv7_fwd(x, x_ld, B, H, W, Cc, w_dw, b_dw, cond, cond_ld, out, out_ld, flip, addend, addend_ld, stream):
    _nhwc(out, B, H, W, Cc, out_ld)[:] = _dwconv7(x, x_ld, B, H, W, Cc, w_dw, b_dw, cond, cond_ld, flip, addend, addend_ld).astype(np.float32)
    return 0


def _ln(h, g, beta, eps):
    mean = h.mean(axis=-1, keepdims=True)
    rstd = 1.0 / np.sqrt(h.var(axis=-1, keepdims=True) + eps)
    return (h - mean) * rstd * g + beta, mean, rstd


def cd_dwconv7_ln_fwd(x, x_ld, B, H, W, Cc, w_dw, b_dw, cond, cond_ld, g, beta, eps, y, y_ld, stats, hpre, hpre_ld, round_tf32, flip,
                      addend, addend_ld, stream):
    h = _dwconv7(x, x_ld, B, H, W, Cc, w_dw, b_dw, cond, cond_ld, flip, addend, addend_ld)
    if _v(hpre):
        _nhwc(hpre, B, H, W, Cc, hpre_ld)[:] = h.astype(np.float32)
    if _v(g):
        o, mean, rstd = _ln(h, _arr(g, (Cc,), (1,)).astype(np.float64), _arr(beta, (Cc,), (1,)).astype(np.float64), _v(eps))
        if _v(stats):
            st = _arr(stats, (B * H * W, 2), (2, 1))
            st[:, 0] = mean.reshape(-1); st[:, 1] = rstd.reshape(-1)
    else:
        o = h
    _nhwc(y, B, H, W, Cc, y_ld)[:] = o.astype(np.float32)
    return 0


def cd_layernorm_fwd(x, x_ld, npix, Cc, g, beta, eps, y, y_ld, stats, round_tf32, stream):
    npix = _v(npix)
    o, mean, rstd = _ln(_rows(x, npix, Cc, x_ld).astype(np.float64), _arr(g, (Cc,), (1,)).astype(np.float64),
                        _arr(beta, (Cc,), (1,)).astype(np.float64), _v(eps))
    if _v(stats):
        st = _arr(stats, (npix, 2), (2, 1))
        st[:, 0] = mean.reshape(-1); st[:, 1] = rstd.reshape(-1)
    _rows(y, npix, Cc, y_ld)[:] = o.astype(np.float32)
    return 0


def cd_layernorm_bwd(dy, dy_ld, h, h_ld, stats, g, npix, Cc, addend, addend_ld, dh, dh_ld, dg, dbeta, stream):
    npix = _v(npix)
    st = _arr(stats, (npix, 2), (2, 1)).astype(np.float64)
    xh = (_rows(h, npix, Cc, h_ld).astype(np.float64) - st[:, :1]) * st[:, 1:]
    DY = _rows(dy, npix, Cc, dy_ld).astype(np.float64)
    _arr(dg, (Cc,), (1,))[:] += (DY * xh).sum(axis=0).astype(np.float32)
    _arr(dbeta, (Cc,), (1,))[:] += DY.sum(axis=0).astype(np.float32)
    dv = DY * _arr(g, (Cc,), (1,)).astype(np.float64)
    o = st[:, 1:] * (dv - dv.mean(axis=1, keepdims=True) - xh * (dv * xh).mean(axis=1, keepdims=True))
    if _v(addend):
        o = o + _rows(addend, npix, Cc, addend_ld).astype(np.float64)
    _rows(dh, npix, Cc, dh_ld)[:] = o.astype(np.float32)
    return 0


def cd_linattn_context(qkv, ld, B, n, kmax, ksum, ctx, stream):
    Q = _arr(qkv, (B, n, 384), (n * ld, ld, 1)).astype(np.float64)
    k, v = Q[..., 128:256], Q[..., 256:384]
    km = k.max(axis=1)
    e = np.exp(k - km[:, None, :])
    _arr(kmax, (B, 128), (128, 1))[:] = km.astype(np.float32)
    _arr(ksum, (B, 128), (128, 1))[:] = e.sum(axis=1).astype(np.float32)
    c = np.einsum('bnhd,bnhe->bhde', e.reshape(B, n, 4, 32), v.reshape(B, n, 4, 32))
    _arr(ctx, (B, 4, 32, 32), (4096, 1024, 32, 1))[:] = c.astype(np.float32)
    return 0


def cd_linattn_context_det(qkv, ld, B, n, nblk, ppb, ws, kmax, ksum, ctx, stream):
    nblk, ppb = _v(nblk), _v(ppb)
    assert ppb % 32 == 0 and nblk == -(-_v(n) // ppb)
    _arr(ws, (_v(B), nblk, 4352), (nblk * 4352, 4352, 1))[:] = np.float32(0)        # the workspace is caller-owned and writable
    return cd_linattn_context(qkv, ld, B, n, kmax, ksum, ctx, stream)


def cd_linattn_weff(ctx, ksum, w_out, B, dim, scale, round_tf32, weff, stream):
    c = _arr(ctx, (B, 4, 32, 32), (4096, 1024, 32, 1)).astype(np.float64)
    ks = _arr(ksum, (B, 4, 32), (128, 32, 1)).astype(np.float64)
    wo = _arr(w_out, (dim, 4, 32), (128, 32, 1)).astype(np.float64)
    cn = c * _v(scale) / ks[..., None]
    _arr(weff, (B, dim, 4, 32), (dim * 128, 128, 32, 1))[:] = np.einsum('ohe,bhde->bohd', wo, cn).astype(np.float32)
    return 0


def cd_conv1x1_to_nchw(x, ld, B, H, W, Cc, w, b, Co, resid_nchw, out_nchw, stream):
    X = _arr(x, (B, H * W, Cc), (H * W * ld, ld, 1)).astype(np.float64)
    o = np.einsum('bpc,oc->bop', X, _arr(w, (Co, Cc), (Cc, 1)).astype(np.float64))
    if _v(b):
        o += _arr(b, (Co,), (1,)).astype(np.float64)[None, :, None]
    if _v(resid_nchw):
        o += _arr(resid_nchw, (B, Co, H * W), (Co * H * W, H * W, 1)).astype(np.float64)
    _arr(out_nchw, (B, Co, H * W), (Co * H * W, H * W, 1))[:] = o.astype(np.float32)
    return 0


def cd_conv1x1_to_nchw_bwd(dout_nchw, x, ld, B, H, W, Cc, w, Co, dx, dx_ld, dw, db, stream):
    D = _arr(dout_nchw, (B, Co, H * W), (Co * H * W, H * W, 1)).astype(np.float64)
    X = _arr(x, (B, H * W, Cc), (H * W * ld, ld, 1)).astype(np.float64)
    _arr(dx, (B, H * W, Cc), (H * W * dx_ld, dx_ld, 1))[:] = np.einsum('bop,oc->bpc', D, _arr(w, (Co, Cc), (Cc, 1)).astype(np.float64)).astype(np.float32)
    _arr(dw, (Co, Cc), (Cc, 1))[:] += np.einsum('bop,bpc->oc', D, X).astype(np.float32)
    _arr(db, (Co,), (1,))[:] += D.sum(axis=(0, 2)).astype(np.float32)
    return 0


def cd_dwconv7_wgrad(dh, dh_ld, x, x_ld, B, H, W, Cc, dw, stream):
    D = _nhwc(dh, B, H, W, Cc, dh_ld).astype(np.float64)
    X = _nhwc(x, B, H, W, Cc, x_ld)
    G = _arr(dw, (Cc, 49), (49, 1))
    for k in range(49):
        G[:, k] += (D * _gather(X, B, H, W, 1, 1, k // 7 - 3, k % 7 - 3)).sum(axis=(0, 1, 2)).astype(np.float32)
    return 0


def cd_colsum_batched(x, ld, B, rows, Cc, out, out_ld, stream):
    rows = _v(rows)
    _arr(out, (B, Cc), (out_ld, 1))[:] += _arr(x, (B, rows, Cc), (rows * ld, ld, 1)).astype(np.float64).sum(axis=1).astype(np.float32)
    return 0


def cd_linattn_bwd_small(dweff, ctx, ksum, w_out, B, dim, scale, dw_out, dctxn, rowdot, stream):
    scale = _v(scale)
    dwe = _arr(dweff, (B, dim, 4, 32), (dim * 128, 128, 32, 1)).astype(np.float64)
    c = _arr(ctx, (B, 4, 32, 32), (4096, 1024, 32, 1)).astype(np.float64)
    ks = _arr(ksum, (B, 4, 32), (128, 32, 1)).astype(np.float64)
    wo = _arr(w_out, (dim, 4, 32), (128, 32, 1)).astype(np.float64)
    cn = c / ks[..., None]
    _arr(dw_out, (dim, 4, 32), (128, 32, 1))[:] += (scale * np.einsum('bohd,bhde->ohe', dwe, cn)).astype(np.float32)
    dc = scale * np.einsum('bohd,ohe->bhde', dwe, wo)
    _arr(dctxn, (B, 4, 32, 32), (4096, 1024, 32, 1))[:] = dc.astype(np.float32)
    _arr(rowdot, (B, 4, 32), (128, 32, 1))[:] = (dc * cn).sum(axis=-1).astype(np.float32)
    return 0


def cd_linattn_bwd_kv(qkv, ld, B, n, kmax, ksum, dctxn, rowdot, dqkv, dld, stream):
    Q = _arr(qkv, (B, n, 384), (n * ld, ld, 1)).astype(np.float64)
    k, v = Q[..., 128:256].reshape(B, n, 4, 32), Q[..., 256:384].reshape(B, n, 4, 32)
    km = _arr(kmax, (B, 4, 32), (128, 32, 1)).astype(np.float64)
    ks = _arr(ksum, (B, 4, 32), (128, 32, 1)).astype(np.float64)
    dc = _arr(dctxn, (B, 4, 32, 32), (4096, 1024, 32, 1)).astype(np.float64)
    rd = _arr(rowdot, (B, 4, 32), (128, 32, 1)).astype(np.float64)
    P = np.exp(k - km[:, None]) / ks[:, None]
    dk = P * (np.einsum('bhde,bnhe->bnhd', dc, v) - rd[:, None])
    dv = np.einsum('bnhd,bhde->bnhe', P, dc)
    D = _arr(dqkv, (B, n, 384), (n * dld, dld, 1))
    D[..., 128:256] = dk.reshape(B, n, 128).astype(np.float32)
    D[..., 256:384] = dv.reshape(B, n, 128).astype(np.float32)
    return 0


def cd_transpose_weff(weff, B, dim, weff_t, stream):
    _arr(weff_t, (B, 128, dim), (128 * dim, dim, 1))[:] = _arr(weff, (B, dim, 128), (dim * 128, 128, 1)).transpose(0, 2, 1)
    return 0


def cd_gelu_bwd(dy, pre, n, y, act_out, stream):
    n = _v(n)
    p = _arr(pre, (n,), (1,)).astype(np.float64)
    if _v(act_out):
        _arr(act_out, (n,), (1,))[:] = _gelu(p).astype(np.float32)
    if _v(y):
        _arr(y, (n,), (1,))[:] = (_arr(dy, (n,), (1,)).astype(np.float64) * _gelu_grad(p)).astype(np.float32)
    return 0


def cd_loss_fwd_bwd(x0, xhat, n, mode, grad_scale, loss, dxhat, stream):
    n = _v(n)
    d = _arr(xhat, (n,), (1,)).astype(np.float64) - _arr(x0, (n,), (1,)).astype(np.float64)
    L = _arr(loss, (1,), (1,))
    if mode == 0:
        L[0] += np.float32(np.abs(d).mean())
        g = np.sign(d) / n
    else:
        L[0] += np.float32((d * d).mean())
        g = 2.0 * d / n
    if _v(dxhat):
        _arr(dxhat, (n,), (1,))[:] = (g * _v(grad_scale)).astype(np.float32)
    return 0


def cd_blur_apply(x, out, ops_, t, t_scalar, B, Cc, S, T, collapse_last, quantize, stream):
    X = _arr(x, (B, Cc, S, S), (Cc * S * S, S * S, S, 1)).astype(np.float64)
    A = _arr(ops_, (T, S, S), (S * S, S, 1)).astype(np.float64)
    O = _arr(out, (B, Cc, S, S), (Cc * S * S, S * S, S, 1))
    tt = _i64(t, B) if _v(t) else [t_scalar] * B
    assert not collapse_last and not quantize, "emulator: `discrete` options are not emulated"
    for b in range(B):
        ti = int(tt[b])
        O[b] = X[b].astype(np.float32) if ti < 0 else np.einsum('ij,cjk,lk->cil', A[ti], X[b], A[ti]).astype(np.float32)
    return 0


def cd_blur_step_down(xt, xhat, out, ops_, t_hi, t_lo, B, Cc, S, T, collapse_last, stream):
    assert not collapse_last, "emulator: `discrete` options are not emulated"
    shp, st = (B, Cc, S, S), (Cc * S * S, S * S, S, 1)
    X, Xh = _arr(xt, shp, st).astype(np.float64), _arr(xhat, shp, st).astype(np.float64)
    A = _arr(ops_, (T, S, S), (S * S, S, 1)).astype(np.float64)
    D = lambda i: Xh if i < 0 else np.einsum('ij,bcjk,lk->bcil', A[i], Xh, A[i])
    _arr(out, shp, st)[:] = (X - D(t_hi) + D(t_lo)).astype(np.float32)
    return 0


# ------------------------------------------------------------------------------------------------------------------
# degradations of the other packages (degrade.cu)
# ------------------------------------------------------------------------------------------------------------------
def _quantize8(v):
    q = (v.astype(np.float32) + np.float32(1)) * np.float32(0.5) * np.float32(255)
    return (np.trunc(q).astype(np.float32) / np.float32(255)) * np.float32(2) - np.float32(1)


def cd_noise_lerp(x1, x2, t, t_scalar, sa, sb, per_sample, n, out, stream):
    per_sample, n = _v(per_sample), _v(n)
    B = n // per_sample
    tt = _i64(t, B) if _v(t) else np.full(B, t_scalar)
    T = int(tt.max()) + 1
    a = _arr(sa, (T,), (1,))[tt][:, None]; b = _arr(sb, (T,), (1,))[tt][:, None]
    _arr(out, (B, per_sample), (per_sample, 1))[:] = a * _arr(x1, (B, per_sample), (per_sample, 1)) + b * _arr(x2, (B, per_sample), (per_sample, 1))
    return 0


def cd_noise_step(img, x1_bar, noise, mode, t, sa, sb, n, out, stream):
    n = _v(n)
    A, Bc = _arr(sa, (t,), (1,)), _arr(sb, (t,), (1,))
    im, xv = _arr(img, (n,), (1,)), _arr(x1_bar, (n,), (1,))
    a1, b1 = A[t - 1], Bc[t - 1]
    x2 = (im - a1 * xv) / b1 if mode == 0 else _arr(noise, (n,), (1,))
    xt_bar = a1 * xv + b1 * x2
    xt_sub1 = A[t - 2] * xv + Bc[t - 2] * x2 if t - 1 != 0 else xv
    _arr(out, (n,), (1,))[:] = im - xt_bar + xt_sub1
    return 0


def cd_fade_lerp(x1, x2, t, t_scalar, alphas, one_minus, B, Cc, HW, out, stream):
    tt = _i64(t, B) if _v(t) else np.full(B, t_scalar)
    T = int(tt.max()) + 1
    al = _arr(alphas, (T, HW), (HW, 1))[tt][:, None, :]; om = _arr(one_minus, (T, HW), (HW, 1))[tt][:, None, :]
    shp, st = (B, Cc, HW), (Cc * HW, HW, 1)
    _arr(out, shp, st)[:] = al * _arr(x1, shp, st) + om * _arr(x2, shp, st)
    return 0


def cd_fade_step(img, x1_bar, x2, t, alphas, one_minus, B, Cc, HW, out, stream):
    al, om = _arr(alphas, (t, HW), (HW, 1)), _arr(one_minus, (t, HW), (HW, 1))
    shp, st = (B, Cc, HW), (Cc * HW, HW, 1)
    im, xv, ev = _arr(img, shp, st), _arr(x1_bar, shp, st), _arr(x2, shp, st)
    xt_bar = al[t - 1] * xv + om[t - 1] * ev
    xt_sub1 = al[t - 2] * xv + om[t - 2] * ev if t - 1 != 0 else xv
    _arr(out, shp, st)[:] = im - xt_bar + xt_sub1
    return 0


def _mask_windows(masks, idx_per_b, rx, ry, B, S, MS, T):
    M = _arr(masks, (T, MS, MS), (MS * MS, MS, 1))
    out = np.ones((B, 1, S, S), dtype=np.float32)
    for b in range(B):
        i = int(idx_per_b[b])
        if i >= 0:
            oy = int(_i64(rx, B)[b]) if _v(rx) else 0
            ox = int(_i64(ry, B)[b]) if _v(ry) else 0
            out[b, 0] = M[i, oy:oy + S, ox:ox + S]
    return out


def cd_mask_apply(x, out, masks, t, t_scalar, rx, ry, B, Cc, S, MS, quantize, stream):
    tt = _i64(t, B) if _v(t) else np.full(B, t_scalar)
    Mw = _mask_windows(masks, tt, rx, ry, B, S, MS, max(int(tt.max()) + 1, 1))
    shp, st = (B, Cc, S, S), (Cc * S * S, S * S, S, 1)
    v = _arr(x, shp, st) * Mw
    _arr(out, shp, st)[:] = _quantize8(v) if quantize else v
    return 0


def cd_mask_step_down(xt, xhat, out, masks, idx_hi, idx_lo, rx, ry, B, Cc, S, MS, stream):
    T = max(idx_hi, idx_lo) + 1
    hi = _mask_windows(masks, [idx_hi] * B, rx, ry, B, S, MS, max(T, 1))
    lo = _mask_windows(masks, [idx_lo] * B, rx, ry, B, S, MS, max(T, 1))
    shp, st = (B, Cc, S, S), (Cc * S * S, S * S, S, 1)
    xv = _arr(xhat, shp, st)
    _arr(out, shp, st)[:] = _arr(xt, shp, st) - xv * hi + xv * lo
    return 0


def cd_chanmix(xt, xsrc, out, mats, t_hi, t_lo, hi_off, lo_off, B, Cc, HW, mode, stream):
    HW = _v(HW)
    th = _i64(t_hi, B) + hi_off
    tl = (_i64(t_lo, B) + lo_off) if mode else np.full(B, -1)
    T = int(max(th.max(), tl.max())) + 1
    M = _arr(mats, (max(T, 1), Cc, Cc), (Cc * Cc, Cc, 1)).astype(np.float64)
    shp, st = (B, Cc, HW), (Cc * HW, HW, 1)
    src, O = _arr(xsrc, shp, st).astype(np.float64), _arr(out, shp, st)
    for b in range(B):
        hi = src[b] if th[b] < 0 else M[th[b]] @ src[b]
        lo = src[b] if tl[b] < 0 else M[tl[b]] @ src[b]
        O[b] = (_arr(xt, shp, st)[b].astype(np.float64) - hi + lo if mode else hi).astype(np.float32)
    return 0


def cd_snow(xt, og, out, snow, br_coef, t_hi, t_lo, hi_off, lo_off, B, H, W, snow_batch, fix_brightness, mode, stream):
    HW = H * W
    th = _i64(t_hi, B) + hi_off
    tl = (_i64(t_lo, B) + lo_off) if mode else np.full(B, -1)
    T = int(max(th.max(), tl.max())) + 1
    S = _arr(snow, (max(T, 1), snow_batch, 3, HW), (snow_batch * 3 * HW, 3 * HW, HW, 1))
    br = _arr(br_coef, (max(T, 1),), (1,))
    shp, st = (B, 3, HW), (3 * HW, HW, 1)
    OG, O = _arr(og, shp, st), _arr(out, shp, st)
    for b in range(B):
        r = (OG[b] + np.float32(1)) / np.float32(2)
        gray = (np.float32(0.299) * r[0] + np.float32(0.587) * r[1] + np.float32(0.114) * r[2]) * np.float32(1.5) + np.float32(0.5)
        g3 = np.maximum(r, gray[None])
        res = []
        for i in (th[b], tl[b]):
            if i < 0:
                res.append(OG[b]); continue
            base = r if fix_brightness else br[i] * r + (np.float32(1) - br[i]) * g3
            sl = S[i, b if snow_batch > 1 else 0]
            res.append(np.clip(base + sl + sl[:, ::-1], 0, 1) * np.float32(2) - np.float32(1))
        O[b] = _arr(xt, shp, st)[b] - res[0] + res[1] if mode else res[0]
    return 0


def cd_snow_layers(noise, SB, ch, m, trim, H, thres, taps, k, vertical, T, base, snow, stream):
    N = np.ctypeslib.as_array((C.c_double * (SB * ch * ch)).from_address(_v(noise))).reshape(SB, ch, ch)
    V = np.ctypeslib.as_array((C.c_uint8 * (T * SB)).from_address(_v(vertical))).reshape(T, SB)
    TH, TP = _arr(thres, (T,), (1,)), _arr(taps, (T, k), (k, 1))
    Bs = _arr(base, (SB, H, H), (H * H, H, 1))
    S = _arr(snow, (T, SB, 3, H, H), (SB * 3 * H * H, 3 * H * H, H * H, H, 1))
    scale = (ch - 1) / (m - 1)
    cc = (np.arange(H, dtype=np.float64) + trim) * scale
    ok = cc <= ch - 1
    f = np.floor(cc)
    w0 = 1.0 - (cc - f); w1 = 1.0 - w0
    i0 = np.minimum(f.astype(np.int64), ch - 1); i1 = i0 + 1
    i1 = np.where(i1 > ch - 1, 2 * (ch - 1) - i1, i1)
    W, I = (w0, w1), (i0, i1)
    for s in range(SB):
        terms = [(N[s][np.ix_(I[a], I[b])] * W[a][:, None]) * W[b][None, :] for a in range(2) for b in range(2)]
        z = ((terms[0] + terms[1]) + terms[2]) + terms[3]
        z[~ok, :] = 0.0; z[:, ~ok] = 0.0
        Bs[s] = z.astype(np.float32)
    half = k // 2
    for t in range(T):
        L = np.where(Bs < TH[t], np.float32(0), np.clip(Bs, 0, 1)).astype(np.float32)
        P = np.zeros((SB, H + 2 * half, H + 2 * half), np.float32)
        P[:, half:half + H, half:half + H] = L
        for s in range(SB):
            acc = np.zeros((H, H), np.float32)
            for j in range(k):
                if V[t, s]:
                    acc += TP[t, k - 1 - j] * P[s, j:j + H, half:half + H]
                else:
                    acc += TP[t, j] * P[s, half:half + H, j:j + H]
            S[t, s, :] = acc[None]
    return 0


def cd_augment_u8(src, N, Hs, Ws, index, oy, ox, flip, B, S, out, stream):
    U = np.ctypeslib.as_array((C.c_uint8 * (N * Hs * Ws * 3)).from_address(_v(src))).reshape(N, Hs, Ws, 3)
    idx = _i64(index, B)
    i32 = lambda a: np.ctypeslib.as_array((C.c_int32 * B).from_address(_v(a)))
    OY, OX, FL = i32(oy), i32(ox), i32(flip)
    O = _arr(out, (B, 3, S, S), (3 * S * S, S * S, S, 1))
    for b in range(B):
        win = U[idx[b], OY[b]:OY[b] + S, OX[b]:OX[b] + S]
        if FL[b]:
            win = win[:, ::-1]
        O[b] = ((win.astype(np.float32) / np.float32(255)) * np.float32(2) - np.float32(1)).transpose(2, 0, 1)
    return 0


def cd_adam_ema_step(p, g, m, v, ema, n, lr, beta1, beta2, eps, step, ema_mode, ema_beta, grad_scale, stream):
    n = _v(n); lr, b1, b2, eps, eb, gs = (np.float32(_v(a)) for a in (lr, beta1, beta2, eps, ema_beta, grad_scale))
    P, Gr, M, V = (_arr(a, (n,), (1,)) for a in (p, g, m, v))
    gi = Gr * gs
    M[:] = b1 * M + (np.float32(1) - b1) * gi
    V[:] = b2 * V + (np.float32(1) - b2) * gi * gi
    bc1 = np.float32(1.0 - float(b1) ** step); bc2s = np.float32(math.sqrt(1.0 - float(b2) ** step))
    P[:] = P - (lr / bc1) * (M / (np.sqrt(V) / bc2s + eps))
    if ema_mode == 1:
        _arr(ema, (n,), (1,))[:] = P
    elif ema_mode == 2:
        E = _arr(ema, (n,), (1,))
        E[:] = E * eb + (np.float32(1) - eb) * P
    return 0


def cd_ema_update(ema, p, n, beta, mode, stream):
    n = _v(n); b = np.float32(_v(beta))
    E, P = _arr(ema, (n,), (1,)), _arr(p, (n,), (1,))
    E[:] = P if mode == 1 else E * b + (np.float32(1) - b) * P
    return 0

_TABLE = {k: v for k, v in globals().items() if k.startswith('cd_')}


def call(name, *args):
    fn = _TABLE.get(name)
    if fn is None:
        raise NotImplementedError("abi_emulator: %s is not emulated" % name)
    rc = fn(*args)
    assert rc == 0


_cpu_lib = None


def call_cuda_source(name, *args):
    """the same call executed by the library's own CUDA-core kernel SOURCES compiled for the CPU (tests/simt_cpu: every CUDA
    thread a fiber; tensor-core entry points defer to the fp32 CUDA-core convolutions)"""
    global _cpu_lib
    if _cpu_lib is None:
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'simt_cpu'))
        import build
        _cpu_lib = C.CDLL(build.build_all())
    rc = getattr(_cpu_lib, name)(*args)
    if rc != 0:
        buf = C.create_string_buffer(512)
        _cpu_lib.cd_last_error(buf, 512)
        raise RuntimeError("%s failed (%d): %s" % (name, rc, buf.value.decode()))


class patched:
    """context manager: route the host modules' `call` / `stream` to the emulator (CPU tensors).  backend 'numpy' (default) is
    the numpy statement of every entry point in this file; 'cuda_source' (or COLDDIFF_ABI_BACKEND=cuda_source) executes the
    CUDA kernel sources on the CPU instead."""

    def __init__(self, backend=None):
        self.backend = backend or os.environ.get('COLDDIFF_ABI_BACKEND', 'numpy')
        assert self.backend in ('numpy', 'cuda_source')

    def __enter__(self):
        from cold_diffusion_models_b200 import (ops, model2, model2_train, engine, engine_bwd, deblurring, trainer, denoising,
                                                resolution, defading, defading_generation, snowification)
        self._mods = (ops, model2, model2_train, engine, engine_bwd, deblurring, trainer, denoising, resolution, defading,
                      defading_generation, snowification)
        self._saved = [(m, m.call, m.stream) for m in self._mods]
        for m in self._mods:
            m.call = call if self.backend == 'numpy' else call_cuda_source
            m.stream = lambda: None
        return self

    def __exit__(self, *exc):
        for m, c, s in self._saved:
            m.call, m.stream = c, s
        return False
