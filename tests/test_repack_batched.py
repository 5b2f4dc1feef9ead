"""One-launch weight repacks (csrc/repack.cu: cd_pack_weight_batched / cd_unpack_wgrad_batched, ops.RepackBatch, the
COLDDIFF_BATCHED_REPACK switch of engine.py / engine_bwd.py), checked on the CPU:
  * the kernel SOURCE executed by tests/simt_cpu (both thread orders) against the numpy statement of the single-weight entry
    points, on a job table that takes every branch (tiled / element-strided, Conv2d / ConvTranspose2d, forward / data-gradient
    operand, partial tap lists, accumulate on and off, source clearing);
  * the host logic: with the switch on, the Unet's output, every parameter gradient (also after two accumulated backward
    passes and after a backward that raised half-way) and one Trainer step are the same as with the switch off.
The `-m gpu` counterpart is tests/test_helpers_and_variants_gpu.py::test_batched_repack_matches_the_single_launches."""
import contextlib
import ctypes as C
import io
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'simt_cpu'))
import abi_emulator as E  # noqa: E402

G = os.path.join(os.path.dirname(__file__), 'golden')


def load(name):
    z = np.load(os.path.join(G, name + '.npz'))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


@pytest.fixture(scope='module', params=['ascending', 'descending'])
def cpulib(request):
    import build
    lib = C.CDLL(build.build_all())          # SIMT_CPU_ASAN=1 applies (tests/simt_cpu/build.py)
    lib.simt_set_reverse_order(int(request.param == 'descending'))
    yield lib
    lib.simt_set_reverse_order(0)


def _job_set(kind):
    """(weight shape, taps, mode, transposed_conv, round_tf32) covering every branch of the two kernels"""
    from cold_diffusion_models_b200 import ops
    T3, T1, T4, T3D = ops.taps_conv(3, 1), ops.taps_conv(1, 0), ops.taps_conv(4, 1), ops.taps_conv_dgrad(3, 1)
    par = ops.taps_convT4_parity(1, 0)
    jobs = [((40, 24, 3, 3), T3, 0, False, 0),            # tiled, 960 rows: 3 full tiles + a partial one
            ((7, 5, 3, 3), T3, 0, False, 1),              # tiled, one partial tile; TF32 rounding (pack)
            ((16, 8, 4, 4), T4, 0, False, 0),             # tiled, KH*KW = 16
            ((16, 8, 4, 4), par, 0, False, 0),            # tiled, 4 of 16 taps (the other 12 must stay untouched on unpack)
            ((12, 20, 1, 1), T1, 0, False, 0),            # tiled, 1x1
            ((8, 16, 4, 4), par, 0, True, 0),             # ConvTranspose2d parity class: element-strided
            ((300, 40, 3, 3), T3, 0, True, 0)]            # element-strided with more work than one pass of its blocks
    if kind == 'pack':
        jobs += [((40, 24, 3, 3), T3D, 1, False, 0),      # data-gradient operand: element-strided
                 ((8, 16, 4, 4), T4, 1, True, 1)]
    return jobs


def _make(kind, seed):
    from cold_diffusion_models_b200 import ops
    g = torch.Generator().manual_seed(seed)
    batch = ops.RepackBatch(kind)
    bufs = []
    for shape, taps, mode, tr, rnd in _job_set(kind):
        O, I = (shape[1], shape[0]) if tr else (shape[0], shape[1])
        w = torch.randn(shape, generator=g)
        n, k = (O, I) if mode == 0 else (I, O)
        packed = torch.randn(len(taps), n, k, generator=g)
        if kind == 'pack':
            batch.add(w, taps, packed, shape=shape, mode=mode, transposed_conv=tr, round_tf32=rnd)
        else:
            batch.add(packed, taps, w, shape=shape, transposed_conv=tr)
        bufs.append((w, packed))
    batch._build(torch.device('cpu'))
    return batch, bufs


@pytest.mark.parametrize('kind,accumulate,clear_src', [('pack', 0, 0), ('unpack', 1, 1), ('unpack', 0, 0), ('unpack', 1, 0)])
def test_batched_repack_kernel_sources_against_the_single_weight_statements(cpulib, kind, accumulate, clear_src):
    res = []
    for impl in ('source', 'numpy'):
        batch, bufs = _make(kind, 11)
        tp, n, tot = C.c_void_p(batch._table.data_ptr()), len(batch), batch._total
        if kind == 'pack':
            rc = (cpulib.cd_pack_weight_batched if impl == 'source' else E.cd_pack_weight_batched)(tp, n, tot, C.c_void_p(0))
        else:
            rc = (cpulib.cd_unpack_wgrad_batched if impl == 'source' else E.cd_unpack_wgrad_batched)(tp, n, tot, accumulate, clear_src,
                                                                                                    C.c_void_p(0))
        assert rc == 0
        if kind == 'pack' and impl == 'numpy':              # the numpy statement leaves the TF32 rounding of the operand out
            for (_, pk), job in zip(bufs, _job_set(kind)):
                if job[4]:
                    pk.copy_(torch.from_numpy(E._tf32(pk.numpy())).view_as(pk))
        res.append(bufs)
    for j, ((w0, p0), (w1, p1)) in enumerate(zip(*res)):
        assert torch.equal(w0, w1), (kind, j)              # pure data movement (+ one add / one rounding): bit-exact
        assert torch.equal(p0, p1), (kind, j)
    if kind == 'unpack' and clear_src:
        assert all(float(p.abs().max()) == 0.0 for _, p in res[0])


def test_strided_jobs_are_correct_for_any_block_count(cpulib):
    """the element-strided branch must not depend on the caller's nblocks (include/colddiff.h)"""
    from cold_diffusion_models_b200 import ops
    outs = []
    for cap in (1, 3, 296):
        old = ops.RepackBatch.MAX_STRIDED_BLOCKS
        ops.RepackBatch.MAX_STRIDED_BLOCKS = cap
        try:
            batch, bufs = _make('pack', 3)
        finally:
            ops.RepackBatch.MAX_STRIDED_BLOCKS = old
        assert cpulib.cd_pack_weight_batched(C.c_void_p(batch._table.data_ptr()), len(batch), batch._total, C.c_void_p(0)) == 0
        outs.append([p for _, p in bufs])
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    for a, b in zip(outs[0], outs[2]):
        assert torch.equal(a, b)


# ----------------------------------------------------------------------------------------------------------------------
# host logic
# ----------------------------------------------------------------------------------------------------------------------
@pytest.fixture()
def emu(monkeypatch):
    from cold_diffusion_models_b200 import engine
    monkeypatch.setattr(torch.Tensor, 'is_cuda', property(lambda self: True))
    with E.patched():
        yield
    engine.batched_repack(False)


def small_unet(g):
    import cold_diffusion_models_b200 as cdm
    with contextlib.redirect_stdout(io.StringIO()):
        u = cdm.Unet(dim=32, dim_mults=(1, 2), channels=3)
    u.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith('sd:')})
    return u


def _grads(batched, passes=2, fail_first=False):
    from cold_diffusion_models_b200 import engine, ops
    engine.batched_repack(batched)
    g = load('unet_small')
    u = small_unet(g)
    launches = []
    real_call = ops.call

    def counting_call(name, *a):
        launches.append(name)
        return real_call(name, *a)
    ops.call = counting_call
    try:
        if fail_first:
            # a backward that dies after some weight gradients were accumulated into the packed buffers but before the unpack
            y = u(g['x'], g['t'])
            eng = u.engine
            orig = eng._time_bwd
            eng._time_bwd = lambda save: (_ for _ in ()).throw(RuntimeError('injected'))
            with pytest.raises(RuntimeError, match='injected'):
                (g['target'] - y).abs().mean().backward()
            eng._time_bwd = orig
            for p in u.parameters():
                p.grad.zero_()
        ys = []
        for _ in range(passes):
            y = u(g['x'], g['t'])
            (g['target'] - y).abs().mean().backward()
            ys.append(y.detach().clone())
    finally:
        ops.call = real_call
    return ys, {k: p.grad.clone() for k, p in u.named_parameters()}, launches


def test_switch_on_reproduces_switch_off_and_batches_the_launches(emu):
    y0, g0, l0 = _grads(False)
    y1, g1, l1 = _grads(True)
    assert all(torch.equal(a, b) for a, b in zip(y0, y1))
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k                # the emulator is deterministic: same sums in the same order
    # dense Conv2d gradients are stored packed (written in place): only the transposed convolutions still unpack
    assert l0.count('cd_unpack_wgrad') >= 4 and l0.count('cd_pack_weight') > 10
    assert l1.count('cd_unpack_wgrad') == 0 and l1.count('cd_pack_weight') == 0
    assert l1.count('cd_unpack_wgrad_batched') == 2        # one per backward pass
    assert l1.count('cd_pack_weight_batched') == 2         # forward operands + data-gradient operands, packed once (weights unchanged)


def test_switch_on_matches_the_reference_gradients(emu):
    g = load('unet_small')
    _, grads, _ = _grads(True, passes=1)
    n = 0
    for k, v in g.items():
        if k.startswith('grad:'):
            r = ((grads[k[5:]].double() - v.double()).norm() / (v.double().norm() + 1e-30)).item()
            assert r < 1e-4, (k, r)
            n += 1
    assert n > 0


def test_a_backward_that_raised_does_not_leak_partial_sums_into_the_next_one(emu):
    _, g0, _ = _grads(False, passes=1)
    _, g1, _ = _grads(True, passes=1, fail_first=True)
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k


def test_trainer_step_with_the_switch_on(emu, monkeypatch, tmp_path):
    import cold_diffusion_models_b200 as cdm
    from cold_diffusion_models_b200 import engine
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    g = load('unet_small')
    out = []
    for batched in (False, True):
        engine.batched_repack(batched)
        u = small_unet(g)
        gd = cdm.GaussianDiffusion(u, image_size=32, device_of_kernel='cpu', channels=3, timesteps=4, kernel_std=0.15, kernel_size=7,
                                   blur_routine='Exponential_reflect', sampling_routine='x0_step_down', loss_type='l2')
        with contextlib.redirect_stdout(io.StringIO()):
            tr = cdm.Trainer(gd, None, image_size=32, train_batch_size=2, train_lr=1e-3, gradient_accumulate_every=2,
                             results_folder=str(tmp_path), dataset='synthetic', step_start_ema=0, update_ema_every=1, ema_decay=0.9)
        gen = torch.Generator().manual_seed(5)
        for step in range(2):                               # the second step repacks the weights the first one updated
            for _ in range(2):
                x = torch.rand(2, 3, 32, 32, generator=gen) * 2 - 1
                (gd.p_losses(x, torch.tensor([3, step])) / 2).backward()
            tr.opt.step(ema_mode=2, ema_beta=0.9)
            tr.opt.zero_grad()
        out.append({k: v.clone() for k, v in gd.denoise_fn.state_dict().items()})
    for k in out[0]:
        assert torch.equal(out[0][k], out[1][k]), k
