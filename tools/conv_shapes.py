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
    for mode in (0, 1, 2, 3):
        lib.cd_conv_tc_set_2cta(min(mode, 1))
        lib.cd_conv_tc_set_2cta_bn(192 if mode == 2 else (128 if mode == 3 else 0))      # mode 2: the pair kernel also for the 128- / 64-wide N tiles
        lib.cd_conv_tc_set_halo(1 if mode == 3 else 0)                                   # mode 3: halo-tile kernel for the 3x3 convolutions
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
lib.cd_conv_tc_set_2cta(1)
lib.cd_conv_tc_set_2cta_bn(128)      # library default
lib.cd_conv_tc_set_halo(0)
print("%-44s %5s %9s %9s %9s %9s %8s %8s %8s %8s" % ("(B,Hg,Wg,Cout,K,nsrc,per_batch)", "n", "1cta us", "2cta us", "2cta-n us", "halo us", "TF/s 1", "TF/s 2", "TF/s 2n", "TF/s halo"))
tot = [0.0, 0.0, 0.0, 0.0]
for shp, (n, ms, f) in res[0].items():
    n //= 5
    us = [res[m][shp][1] / 5 / n * 1e3 for m in (0, 1, 2, 3)]
    for m in range(4):
        tot[m] += res[m][shp][1] / 5
    print("%-44s %5d %9.1f %9.1f %9.1f %9.1f %8.1f %8.1f %8.1f %8.1f" % ((str(shp), n) + tuple(us) + tuple(f / u / 1e6 for u in us)))
print("total conv ms per forward: 1cta %.3f   2cta (256-wide tiles, cost model) %.3f   2cta also for 128/64-wide tiles %.3f   halo-tile kernel for 3x3 (+ 2cta 256/128) %.3f" % tuple(tot))
