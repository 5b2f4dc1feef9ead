"""CPU check of the HOST LOGIC of the ConvNeXt Unet engine (engine.py / engine_bwd.py: buffer planning, concat-free skip
connections, tap lists of every convolution and of its data / weight gradients incl. the four parity classes of the
transpose convolution, the LinearAttention folding into per-batch weights, the time-embedding backward, flat gradient
views) and of `GaussianDiffusion.p_losses`: the schedules run on CPU tensors against tests/abi_emulator.py and must reproduce
the reference's output, loss and every parameter gradient (tests/golden/unet_small.npz, sample_small.npz).  The CUDA kernels
themselves are checked by the `-m gpu` tests."""
import io
import contextlib
import os
import numpy as np
import pytest
import torch

G = os.path.join(os.path.dirname(__file__), 'golden')


def load(name):
    z = np.load(os.path.join(G, name + '.npz'))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


@pytest.fixture()
def emu(monkeypatch):
    import abi_emulator
    monkeypatch.setattr(torch.Tensor, 'is_cuda', property(lambda self: True))
    with abi_emulator.patched():
        yield


def small_unet(g):
    import cold_diffusion_models_b200 as cdm
    with contextlib.redirect_stdout(io.StringIO()):
        u = cdm.Unet(dim=32, dim_mults=(1, 2), channels=3)
    u.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith('sd:')})
    return u


def test_unet_forward_and_every_gradient_on_the_emulated_abi(emu):
    g = load('unet_small')
    u = small_unet(g)
    y = u(g['x'], g['t'])
    assert rel(y.detach(), g['y']) < 2e-6
    loss = (g['target'] - y).abs().mean()
    assert abs(loss.item() - g['loss'].item()) < 1e-6
    loss.backward()
    named = dict(u.named_parameters())
    n, worst = 0, (-1.0, '')
    for k, v in g.items():
        if k.startswith('grad:'):
            mine = named[k[5:]].grad
        elif k.startswith('gsub:'):
            gr = named[k[5:]].grad.reshape(-1)
            mine = gr[::gr.numel() // 2048]
            assert abs(gr.double().norm().item() / g['gnorm:' + k[5:]].item() - 1) < 1e-4, k
        else:
            continue
        worst = max(worst, (rel(mine, v), k))
        n += 1
    assert n == len(named)
    assert worst[0] < 1e-4, worst


def test_p_losses_through_the_public_class_on_the_emulated_abi(emu):
    import cold_diffusion_models_b200 as cdm
    g, gs = load('unet_small'), load('sample_small')
    u = small_unet(g)
    key = 'Exponential_reflect|7|0.15|4|x0_step_down|0'
    gd = cdm.GaussianDiffusion(u, image_size=32, device_of_kernel='cpu', channels=3, timesteps=4, kernel_std=0.15, kernel_size=7,
                               blur_routine='Exponential_reflect', sampling_routine='x0_step_down')
    with torch.no_grad():
        loss = gd.p_losses(gs['x'], torch.tensor([3, 0]))
    assert abs(loss.item() - gs['loss:' + key].item()) < 1e-5


def test_trainer_step_equals_torch_adam_and_ema_on_the_emulated_abi(emu, monkeypatch, tmp_path):
    """Trainer.train_step (2 micro-batches, loss/2 each, flat-buffer gradients, fused Adam + EMA lerp, DB:1188-1204) against the
    same two micro-batches through the oracle + torch.optim.Adam + the reference's EMA formula"""
    import cold_diffusion_models_b200 as cdm
    import unet_oracle as UO
    import deblur_oracle as DO
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    g = load('unet_small')
    sd = {k[3:]: v for k, v in g.items() if k.startswith('sd:')}
    u = small_unet(g)
    gd = cdm.GaussianDiffusion(u, image_size=32, device_of_kernel='cpu', channels=3, timesteps=4, kernel_std=0.15, kernel_size=7,
                               blur_routine='Exponential_reflect', sampling_routine='x0_step_down', loss_type='l2')
    with contextlib.redirect_stdout(io.StringIO()):
        tr = cdm.Trainer(gd, None, image_size=32, train_batch_size=2, train_lr=1e-3, gradient_accumulate_every=2,
                         results_folder=str(tmp_path), dataset='synthetic', step_start_ema=0, update_ema_every=1, ema_decay=0.9)
    gen = torch.Generator().manual_seed(5)
    xs = [torch.rand(2, 3, 32, 32, generator=gen) * 2 - 1 for _ in range(2)]
    ts = [torch.tensor([3, 0]), torch.tensor([1, 2])]
    ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    orc = DO.DeblurOracle(lambda a, b: UO.unet_forward(ref, a, b), image_size=32, channels=3, timesteps=4, kernel_std=0.15,
                          kernel_size=7, blur_routine='Exponential_reflect', loss_type='l2')
    opt = torch.optim.Adam(list(ref.values()), lr=1e-3)
    for x, t in zip(xs, ts):
        (orc.p_losses(x, t) / 2).backward()
    opt.step()
    for x, t in zip(xs, ts):
        (gd.p_losses(x, t) / 2).backward()
    tr.opt.step(ema_mode=2, ema_beta=0.9)
    tr.opt.zero_grad()
    new, ema = gd.denoise_fn.state_dict(), tr.ema_model.denoise_fn.state_dict()
    for k in sd:
        assert rel(new[k], ref[k].detach()) < 1e-5, k
        assert rel(ema[k], sd[k] * 0.9 + 0.1 * ref[k].detach()) < 1e-5, k
    assert float(tr.opt.engine.flat_grad.abs().max()) == 0.0


def test_individual_incremental_routine_through_the_public_class(emu):
    """q_sample / p_losses / sample of the seventh blur routine (per-step kernel sizes 1,3,5,7; `sample` starts from the single
    step-t kernel, DB:401-402) on the emulated ABI against the reference"""
    import cold_diffusion_models_b200 as cdm
    g, gi = load('unet_small'), load('individual_small')
    u = small_unet(g)
    x = gi['x']
    for samp in ('default', 'x0_step_down'):
        gd = cdm.GaussianDiffusion(u, image_size=32, device_of_kernel='cpu', channels=3, timesteps=4, kernel_std=0.1, kernel_size=3,
                                   blur_routine='Individual_Incremental', sampling_routine=samp)
        for i, kconv in enumerate(gd.gaussian_kernels):
            assert torch.equal(kconv.weight[0, 0], gi['w%d' % i])
        tt = torch.tensor([3, 1])
        assert torch.allclose(gd.q_sample(x, tt), gi['q'], atol=3e-6)
        with torch.no_grad():
            assert abs(gd.p_losses(x, tt).item() - gi['loss'].item()) < 1e-5
        xt, dr, img = gd.sample(batch_size=2, img=x)
        assert rel(xt, gi['xt:' + samp]) < 1e-5 and rel(dr, gi['dr:' + samp]) < 1e-5 and rel(img, gi['img:' + samp]) < 1e-4, samp


def test_deblurring_sampling_helpers_through_the_public_class(emu):
    """sample_from_blur (partial blur start .. t-1, DB:863-925), all_sample (DB:609-689) and gen_sample with noise (DB:526-593;
    same torch seed -> same noise as the reference) on the emulated ABI against the reference"""
    import cold_diffusion_models_b200 as cdm
    g, gf = load('unet_small'), load('fb_small')
    u = small_unet(g)
    x = gf['x']
    for key in sorted(k[4:] for k in gf if k.startswith('img:')):
        routine, ks, std, T, samp = key.split('|')
        gd = cdm.GaussianDiffusion(u, image_size=32, device_of_kernel='cpu', channels=3, timesteps=int(T), kernel_std=float(std),
                                   kernel_size=int(ks), blur_routine=routine, sampling_routine=samp)
        for start in (0, 1):
            xt, dr, img = gd.sample_from_blur(batch_size=2, img=x, start=start)
            pre = ':%d:' % start + key
            assert rel(xt, gf['sfb_xt' + pre]) < 1e-5 and rel(dr, gf['sfb_dr' + pre]) < 1e-5 and rel(img, gf['sfb_img' + pre]) < 1e-4, pre
        X0s, Xts = gd.all_sample(batch_size=2, img=x)
        assert rel(torch.stack(X0s), gf['all_X0:' + key]) < 1e-4 and rel(torch.stack(Xts), gf['all_Xt:' + key]) < 1e-4, key
        torch.manual_seed(17)
        xt, dr, img = gd.gen_sample(batch_size=2, img=x, noise_level=0.05)
        assert rel(xt, gf['gen_xt:' + key]) < 1e-5 and rel(img, gf['gen_img:' + key]) < 1e-4, key


@pytest.mark.parametrize('tag,kw', [('residual', dict(residual=True)), ('notime', dict(with_time_emb=False)), ('outdim', dict(out_dim=5))])
def test_unet_constructor_options(emu, tag, kw):
    """Unet(residual=True) / Unet(with_time_emb=False) (the drivers' --residual / --remove_time_embed flags) / out_dim (DB:192-200):
    forward, L2 loss and every parameter gradient against the reference"""
    import cold_diffusion_models_b200 as cdm
    g = load('unet_options_small')
    with contextlib.redirect_stdout(io.StringIO()):
        u = cdm.Unet(dim=32, dim_mults=(1, 2), channels=3, **kw)
    pre = tag + ':sd:'
    base = {k[3:]: v for k, v in load('unet_small').items() if k.startswith('sd:')}
    extra = {k[len(pre):]: v for k, v in g.items() if k.startswith(pre)}
    u.load_state_dict({k: extra.get(k, base.get(k)) for k in u.state_dict()})
    y = u(g['x'], g['t'])
    assert rel(y.detach(), g[tag + ':y']) < 2e-6
    target = g['tgt5'] if tag == 'outdim' else g['x'].flip(0)
    loss = ((target - y) ** 2).mean()
    assert abs(loss.item() - g[tag + ':loss'].item()) < 1e-6
    loss.backward()
    worst = (-1.0, '')
    for n, p_ in u.named_parameters():
        gr = p_.grad.reshape(-1)
        assert abs(gr.double().norm().item() - g[tag + ':gnorm:' + n].item()) <= 1e-4 * g[tag + ':gnorm:' + n].item() + 1e-9, n
        worst = max(worst, (rel(gr[::max(1, gr.numel() // 256)], g[tag + ':gsub:' + n]), n))
    assert worst[0] < 1e-4, worst
