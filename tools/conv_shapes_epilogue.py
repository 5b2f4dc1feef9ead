"""Per-launch time of every tensor-core convolution of one Unet forward (config 3 network), grouped by GEMM shape, with the row
epilogue (default) and the line-coalesced epilogue of csrc/conv_epilogue.cuh side by side: mode 1 = only launches with <= 16 K
chunks per tile (the store-bound 1x1 projections), mode 2 = every launch (3/4/6 mainloop stages instead of 4/6/8).
Usage: python tools/conv_shapes_epilogue.py [batch]"""
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
    for mode in (0, 1, 3, 2):
        lib.cd_conv_tc_set_staged_epilogue(mode)
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
lib.cd_conv_tc_set_staged_epilogue(0)
MODES = (0, 1, 3, 2)
print("%-44s %5s %9s %9s %9s %9s %8s %8s" % ("(B,Hg,Wg,Cout,K,nsrc,per_batch)", "n", "rows us", "<=16 us", "<=48 us", "all us", "TF/s", "TF/s all"))
tot = {m: 0.0 for m in MODES}
for shp, (n, ms, f) in res[0].items():
    n //= 5
    us = {m: res[m][shp][1] / 5 / n * 1e3 for m in MODES}
    for m in MODES:
        tot[m] += res[m][shp][1] / 5
    print("%-44s %5d %9.1f %9.1f %9.1f %9.1f %8.1f %8.1f" % (str(shp), n, us[0], us[1], us[3], us[2], f / us[0] / 1e6, f / us[2] / 1e6))
print("total conv ms per forward: rows %.3f   <=16 chunks %.3f   <=48 chunks %.3f   all %.3f" % (tot[0], tot[1], tot[3], tot[2]))
