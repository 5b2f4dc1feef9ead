"""Per-launch time of every tensor-core weight-gradient GEMM of one Unet backward (config 3 network), grouped by shape,
under the K-split policies of csrc/wgrad_tc.cu (cd_wgrad_tc_set_split)."""
import sys, io, contextlib, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cold_diffusion_models_b200 as cdm
from cold_diffusion_models_b200._lib import lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
with contextlib.redirect_stdout(io.StringIO()):
    u = cdm.Unet(dim=64, dim_mults=(1, 2, 4, 8), channels=3).cuda()
x = torch.rand(B, 3, 128, 128, device='cuda') * 2 - 1
t = torch.randint(0, 200, (B,), device='cuda')
target = torch.rand(B, 3, 128, 128, device='cuda')


def step():
    y = u(x, t)
    ((y - target) ** 2).mean().backward()


configs = [('2waves-up', 0, 0), ('1wave', 1, 0), ('2waves', 2, 0), ('model-12k', 3, 12000)]
if len(sys.argv) > 2:
    configs = [c for c in configs if c[0] in sys.argv[2].split(',')]
res = collections.OrderedDict()
for name, pol, over in configs:
    lib.cd_wgrad_tc_set_split(pol, over)
    for _ in range(2):
        step()
    acc = collections.OrderedDict()
    for rep in range(3):
        u.engine.profile_wgrads = []
        step()
        torch.cuda.synchronize()
        for a, b, shp in u.engine.profile_wgrads:
            e = acc.setdefault(shp, [0, 0.0])
            e[0] += 1; e[1] += a.elapsed_time(b)
    u.engine.profile_wgrads = None
    res[name] = acc
lib.cd_wgrad_tc_set_split(3, 12000)
names = [c[0] for c in configs]
print("%-34s %3s " % ("(B,H,W,Cout,Cin,taps,stride)", "n") + " ".join("%10s" % n for n in names) + "   best TF/s")
tot = {n: 0.0 for n in names}
for shp in res[names[0]]:
    n = res[names[0]][shp][0] // 3
    us = []
    for nm in names:
        ms = res[nm][shp][1] / 3
        tot[nm] += ms
        us.append(ms / n * 1e3)
    fl = 2.0 * shp[0] * shp[1] * shp[2] * shp[3] * shp[4] * shp[5]
    print("%-34s %3d " % (str(shp), n) + " ".join("%10.1f" % v for v in us) + "   %8.1f" % (fl / min(us) / 1e6))
print("%-34s     " % "total ms per backward" + " ".join("%10.3f" % tot[n] for n in names))
