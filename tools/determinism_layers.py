"""Which engine buffer is the first to differ between two forwards of the same weights and input?

Runs the inference forward (and optionally the training forward + backward) N times and compares every engine buffer with the
first run, in first-use order.  Prints per buffer: number of runs that differ and the worst relative L2 difference.

    python tools/determinism_layers.py [small|full] [train]
"""
import sys, os, io, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch
import cold_diffusion_models_b200 as cdm
import unet_oracle as UO

mode = sys.argv[1] if len(sys.argv) > 1 else 'small'
train = 'train' in sys.argv[2:]
if mode == 'small':
    dim, mults, B, S = 32, (1, 2), 2, 32
else:
    dim, mults, B, S = 64, (1, 2, 4, 8), 8, 128
sd = UO.make_unet_state_dict(dim, mults, 3, seed=0)
with contextlib.redirect_stdout(io.StringIO()):
    u = cdm.Unet(dim=dim, dim_mults=mults, channels=3)
u.load_state_dict(sd)
u = u.cuda()
g = torch.Generator().manual_seed(1)
x = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).cuda()
t = torch.randint(0, 3, (B,), generator=g).cuda()
tgt = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).cuda()


def run():
    if train:
        from cold_diffusion_models_b200.deblurring import _LossFn
        for p in u.parameters():
            p.grad = None
        if getattr(u.engine, 'flat_grad', None) is not None:
            u.engine.flat_grad.zero_()
        y = u(x, t)
        _LossFn.apply(tgt, y, 1).backward()
    else:
        with torch.no_grad():
            y = u(x, t)
    torch.cuda.synchronize()
    snap = {k: v.clone() for k, v in u.engine._bufs.items()}
    snap[('OUT', ())] = y.detach().clone()
    if train and getattr(u.engine, 'flat_grad', None) is not None:
        snap[('FLAT_GRAD', ())] = u.engine.flat_grad.clone()
    return snap


run()
ref = run()
N = 8
bad = {}
for i in range(N):
    s = run()
    for k, v in s.items():
        if k in ref and v.shape == ref[k].shape and not torch.equal(v, ref[k]):
            a, b = v.double(), ref[k].double()
            fin = torch.isfinite(a) & torch.isfinite(b)
            r = ((a - b)[fin].norm() / (b[fin].norm() + 1e-30)).item()
            nn = int((~fin).sum().item())
            c, w, n2 = bad.get(k, (0, 0.0, 0))
            bad[k] = (c + 1, max(w, r), max(n2, nn))
print('mode', mode, 'train' if train else 'infer', '-- buffers in first-use order; only differing ones are listed')
for k in ref:
    if k in bad:
        print('  %-40s runs differing %d/%d   worst rel-L2 %.3e   non-finite %d' % (str(k), bad[k][0], N, bad[k][1], bad[k][2]))
if not bad:
    print('  all %d buffers bit-identical over %d reruns' % (len(ref), N))
