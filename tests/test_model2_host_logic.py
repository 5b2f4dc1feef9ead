"""CPU check of the HOST LOGIC of the DDPM-style `Model` (cold_diffusion_models_b200/model2.py, model2_train.py): the forward
schedule and the whole backward schedule (gradient routing through the concat buffers, tap lists of the data gradients incl.
the asymmetric-pad stride-2 Downsample, packed-weight layouts, per-batch weight gradients of the AttnBlock, time-embedding
backward) run on CPU tensors against tests/abi_emulator.py and must reproduce the reference's output and every parameter
gradient (tests/golden/model2_small.npz, model2_grads_small.npz).  The CUDA kernels are not exercised here."""
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
def model_and_goldens(monkeypatch):
    monkeypatch.setenv('COLDDIFF_MODEL_TRAINING', '1')
    import cold_diffusion_models_b200 as cdm
    g, gg = load('model2_small'), load('model2_grads_small')
    m = cdm.Model(resolution=16, in_channels=3, out_ch=3, ch=32, ch_mult=(1, 2), num_res_blocks=2, attn_resolutions=(8,), dropout=0.1)
    m.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith('sd:')})
    return m, g, gg


def test_forward_schedule_on_the_emulated_abi(model_and_goldens, monkeypatch):
    import abi_emulator
    m, g, _ = model_and_goldens
    m.eval()
    # inference path: cd_time_mlp2_fwd is not emulated -> drive the training-mode forward (same schedule, every stage kept)
    from cold_diffusion_models_b200 import model2_train
    with abi_emulator.patched(), torch.no_grad():
        monkeypatch.setattr(torch.Tensor, 'is_cuda', property(lambda self: True))
        y = model2_train.forward_train(m, g['x'], g['t'], {})
    assert rel(y, g['y']) < 2e-6


def test_inference_forward_and_sampling_on_the_emulated_abi(model_and_goldens, monkeypatch):
    """the production inference path of `Model` (fused time-embedding kernel, shared workspaces) and the Special_6_routine
    x0_step_down sampling around it (BASELINE config 2 shape at reduced size) against the reference"""
    import abi_emulator
    import cold_diffusion_models_b200 as cdm
    m, g, _ = model_and_goldens
    m.eval()
    monkeypatch.setattr(torch.Tensor, 'is_cuda', property(lambda self: True))
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    with abi_emulator.patched(), torch.no_grad():
        y = m(g['x'], g['t'])
        assert rel(y, g['y']) < 2e-6
        gd = cdm.GaussianDiffusion(m, image_size=16, device_of_kernel='cpu', channels=3, timesteps=6, loss_type='l1', kernel_std=0.1,
                                   kernel_size=3, blur_routine='Special_6_routine', train_routine='Final', sampling_routine='x0_step_down')
        xt, dr, img = gd.sample(batch_size=3, img=g['x'])
    assert rel(xt, g['s_xt']) < 1e-5 and rel(dr, g['s_dr']) < 1e-5 and rel(img, g['s_img']) < 1e-4


def test_backward_schedule_reproduces_every_reference_gradient(model_and_goldens, monkeypatch):
    import abi_emulator
    m, g, gg = model_and_goldens
    m.eval()                                  # dropout inactive, as in the golden
    with abi_emulator.patched():
        monkeypatch.setattr(torch.Tensor, 'is_cuda', property(lambda self: True))
        y = m(g['x'], g['t'])                 # autograd enabled -> ModelFunction
        loss = (gg['target'] - y).abs().mean()
        assert abs(loss.item() - gg['loss'].item()) < 1e-6
        loss.backward()
    named = dict(m.named_parameters())
    n = 0
    worst = (-1.0, '')
    for k, v in gg.items():
        if k.startswith('grad:'):
            mine = named[k[5:]].grad
        elif k.startswith('gsub:'):
            gr = named[k[5:]].grad.reshape(-1)
            mine = gr[::gr.numel() // 2048]
            assert abs(gr.double().norm().item() / gg['gnorm:' + k[5:]].item() - 1) < 1e-4, k
        else:
            continue
        e = rel(mine, v)
        # gradients that are zero in exact arithmetic are round-off on both sides (a per-channel shift in front of a GroupNorm
        # whose groups hold one channel: conv1.bias / temb_proj at ch = 32; the key bias in front of a row softmax): absolute check
        if (mine.double() - v.double()).abs().max().item() < 1e-7:
            e = 0.0
        worst = max(worst, (e, k))
        n += 1
    assert n == len(named)
    assert worst[0] < 1e-4, worst


def test_dropout_mask_is_a_pure_function_of_seed_and_index():
    """forward and backward must see the same mask: the emulated kernel (same hash as model2_bwd.cu) is deterministic in
    (seed, element index), keeps ~1-p of the elements and rescales by 1/(1-p)"""
    import ctypes as C
    import abi_emulator as E
    x = torch.ones(64 * 32)
    y1, y2, y3 = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    for y, seed in ((y1, 1234), (y2, 1234), (y3, 99)):
        E.cd_dropout(C.c_void_p(x.data_ptr()), 32, 64, 32, C.c_float(0.1), C.c_uint64(seed), C.c_void_p(y.data_ptr()), 32, None)
    assert torch.equal(y1, y2) and not torch.equal(y1, y3)
    kept = (y1 != 0).float().mean().item()
    assert 0.85 < kept < 0.95 and abs(y1.max().item() - 1 / 0.9) < 1e-6


def test_trainer_step_with_model_on_the_emulated_abi(model_and_goldens, monkeypatch, tmp_path):
    """`Trainer` + fused Adam/EMA over the flat buffers of `Model` (ModelEngine): one optimizer step on two micro-batches equals
    torch.optim.Adam on the oracle's autograd gradients; dropout inactive (eval mode) so both sides see the same network"""
    import io, contextlib
    import abi_emulator
    import cold_diffusion_models_b200 as cdm
    import model2_oracle as MO
    import deblur_oracle as DO
    m, g, _ = model_and_goldens
    monkeypatch.setattr(torch.Tensor, 'is_cuda', property(lambda self: True))
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    sd = {k[3:]: v for k, v in g.items() if k.startswith('sd:')}
    kw = dict(image_size=16, channels=3, timesteps=6, kernel_std=0.1, kernel_size=3, blur_routine='Special_6_routine', loss_type='l2')
    with abi_emulator.patched(), contextlib.redirect_stdout(io.StringIO()):
        gd = cdm.GaussianDiffusion(m, device_of_kernel='cpu', sampling_routine='x0_step_down', **kw)
        tr = cdm.Trainer(gd, None, image_size=16, train_batch_size=3, train_lr=1e-3, gradient_accumulate_every=2,
                         results_folder=str(tmp_path), dataset='synthetic', step_start_ema=0, update_ema_every=1, ema_decay=0.9)
        m.eval(); tr.ema_model.denoise_fn.eval()
        gen = torch.Generator().manual_seed(5)
        xs = [torch.rand(3, 3, 16, 16, generator=gen) * 2 - 1 for _ in range(2)]
        ts = [torch.tensor([5, 0, 2]), torch.tensor([1, 3, 4])]
        ref = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
        orc = DO.DeblurOracle(lambda a, b: MO.model_forward(ref, a, b, ch=32, num_resolutions=2, num_res_blocks=2), **kw)
        opt = torch.optim.Adam([v for v in ref.values() if v.requires_grad], lr=1e-3)
        for x, t in zip(xs, ts):
            (orc.p_losses(x, t) / 2).backward()
        opt.step()
        for x, t in zip(xs, ts):
            (gd.p_losses(x, t) / 2).backward()
        tr.opt.step(ema_mode=2, ema_beta=0.9)
    new, ema = m.state_dict(), tr.ema_model.denoise_fn.state_dict()
    checked = 0
    for k, v in sd.items():
        if not v.dtype.is_floating_point:
            continue
        if ref[k].grad.abs().max() < 1e-6:
            continue      # gradient zero in exact arithmetic (see the gradient test): Adam turns the round-off into +-lr steps
        # Adam's first step is lr * g / (|g| + eps): elements whose gradient is round-off-sized move by +-lr either way
        assert rel(new[k], ref[k].detach()) < 2e-4, k
        assert rel(ema[k], v * 0.9 + 0.1 * ref[k].detach()) < 2e-4, k
        checked += 1
    assert checked > 150


def test_dropout_plumbing_of_the_training_path(model_and_goldens, monkeypatch):
    """dropout active (train mode): the same host seed reproduces the forward, and the backward uses the forward's masks -- the
    loss change along -grad matches the first-order prediction"""
    import abi_emulator
    m, g, _ = model_and_goldens
    monkeypatch.setattr(torch.Tensor, 'is_cuda', property(lambda self: True))
    for b in (b for _, b in m._resblocks()):
        b.dropout.p = 0.3
    m.train()
    x, t = g['x'], g['t']
    target = torch.zeros_like(x)
    with abi_emulator.patched():
        torch.manual_seed(5)
        loss = ((target - m(x, t)) ** 2).mean()
        loss.backward()
        torch.manual_seed(6)
        other = ((target - m(x, t)) ** 2).mean()
        grads = {n: p.grad.clone() for n, p in m.named_parameters()}
        gnorm2 = sum((v.double() ** 2).sum().item() for v in grads.values())
        eps = 1e-3 / gnorm2 ** 0.5
        with torch.no_grad():
            for n, p in m.named_parameters():
                p.add_(grads[n], alpha=-eps)
        m.engine.mark_weights_dirty()
        torch.manual_seed(5)
        loss2 = ((target - m(x, t)) ** 2).mean()
    assert abs(other.item() - loss.item()) > 1e-6                       # different seed -> different masks
    predicted = -eps * gnorm2
    assert abs((loss2.item() - loss.item()) - predicted) < 0.1 * abs(predicted), (loss.item(), loss2.item(), predicted)
