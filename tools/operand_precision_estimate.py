"""CPU estimate of what the operand format of the tensor-core convolutions costs in accuracy: the small golden network's forward
output and every parameter gradient through tests/abi_emulator.py with both operands of every tensor-core convolution and weight
gradient rounded to TF32 (today's kernels), FP16 (same 10-bit mantissa; `kind::f16` MMAs run at twice the TF32 rate and read half
the shared-memory bytes per MAC) or BF16 (7-bit mantissa), against the reference goldens.  For FP16 the loss is scaled by 2^14 and the
gradients unscaled at the end (the loss scale an fp16 backward needs: a mean loss has ~1e-7 gradients per element, below fp16's range).
Usage: python tools/operand_precision_estimate.py"""
import os, sys, io, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('', 'tests'):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
import abi_emulator as E
torch.Tensor.is_cuda = property(lambda self: True)
import cold_diffusion_models_b200 as cdm

G = os.path.join(ROOT, 'tests', 'golden')
z = np.load(os.path.join(G, 'unet_small.npz'))
g = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


tf32 = E._tf32
ROUND = {
    'fp32 (no rounding)': lambda a: np.ascontiguousarray(a, dtype=np.float32),
    'tf32': tf32,
    'fp16': lambda a: np.ascontiguousarray(a, dtype=np.float32).astype(np.float16).astype(np.float32),
    'bf16': lambda a: ((np.ascontiguousarray(a, dtype=np.float32).view(np.int32) + 0x8000) & ~0xFFFF).view(np.float32),
}
real_wgrad = E.cd_conv_wgrad
LOSS_SCALE = 2.0 ** 14


def run(name):
    rnd = ROUND[name]
    E._tf32 = rnd
    E.TF32_EMULATION = name != 'fp32 (no rounding)'

    def wgrad(desc, dout, dout_ld, dw, db, impl, stream):
        d = desc._obj if hasattr(desc, '_obj') else desc
        if not E.TF32_EMULATION or E._v(impl) != 1 or d.s[0].C % 32:
            return real_wgrad(desc, dout, dout_ld, dw, db, impl, stream)
        s = d.s[0]
        src = E._nhwc(s.src, d.B, s.H, s.W, s.C, s.ld)
        dy = E._nhwc(dout, d.B, d.Ho, d.Wo, d.Cout, E._v(dout_ld))
        keep_s, keep_d = src.copy(), dy.copy()
        src[:] = rnd(keep_s)
        dy[:] = rnd(keep_d)
        try:
            return real_wgrad(desc, dout, dout_ld, dw, db, impl, stream)
        finally:
            src[:] = keep_s; dy[:] = keep_d
    E.cd_conv_wgrad = wgrad
    E._TABLE['cd_conv_wgrad'] = wgrad
    with contextlib.redirect_stdout(io.StringIO()):
        u = cdm.Unet(dim=32, dim_mults=(1, 2), channels=3)
    u.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith('sd:')})
    with E.patched():
        y = u(g['x'], g['t'])
        e_fwd = rel(y.detach(), g['y'])
        # smooth loss for the gradient comparison: the golden gradients are those of the L1 loss, so compare against our own fp32 run
        scale = LOSS_SCALE if name == 'fp16' else 1.0      # folded into the loss kernel, taken out of the gradients (exact: power of two)
        (((g['target'] - y) ** 2).mean() * scale).backward()
    grads = {n: p.grad.clone() / scale for n, p in u.named_parameters()}
    return e_fwd, grads


base_fwd, base = run('fp32 (no rounding)')
print('%-22s forward rel err vs reference golden %.2e' % ('fp32 (no rounding)', base_fwd))
for name in ('tf32', 'fp16', 'bf16'):
    e_fwd, gr = run(name)
    errs = sorted((rel(gr[n], base[n]), n) for n in base)
    print('%-22s forward %.2e   gradients vs fp32 run: median %.2e  worst %.2e (%s)' % (name, e_fwd, errs[len(errs) // 2][0], errs[-1][0], errs[-1][1]))
E._tf32 = tf32
