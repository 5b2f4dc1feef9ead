"""Per-launch time of every tensor-core convolution of one Unet forward (config 3 network), grouped by GEMM
shape, with the 1-CTA kernel and the SM-pair (cta_group::2) kernel side by side."""
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
res = {}
with torch.no_grad():
    for mode in (0, 1):
        lib.cd_conv_tc_set_2cta(mode)
        for _ in range(2):
            u(x, t)
        acc = collections.OrderedDict()
        for rep in range(5):
            u.engine.profile_convs, u.engine.profile_shapes = [], []
            u(x, t)
            torch.cuda.synchronize()
            for (a, b, f), shp in zip(u.engine.profile_convs, u.engine.profile_shapes):
                e = acc.setdefault(shp, [0, 0.0, f])
                e[0] += 1; e[1] += a.elapsed_time(b)
        u.engine.profile_convs = u.engine.profile_shapes = None
        res[mode] = acc
lib.cd_conv_tc_set_2cta(0)
print("%-44s %5s %9s %9s %8s %8s" % ("(B,Hg,Wg,Cout,K,nsrc,per_batch)", "n", "1cta us", "2cta us", "TF/s 1", "TF/s 2"))
tot = [0.0, 0.0]
for shp, (n, ms, f) in res[0].items():
    n2, ms2, _ = res[1][shp]
    n //= 5
    a, b = ms / 5 / n * 1e3, ms2 / 5 / n * 1e3
    tot[0] += ms / 5; tot[1] += ms2 / 5
    print("%-44s %5d %9.1f %9.1f %8.1f %8.1f" % (str(shp), n, a, b, f / a / 1e6, f / b / 1e6))
print("total conv ms per forward: 1cta %.3f   2cta %.3f" % (tot[0], tot[1]))
