"""cold_diffusion_models_b200.tuning: the guarded start-up selection of the opt-in kernel variants.
CPU checks of (a) the parent: a child that cannot run (no CUDA here) leaves every switch at its default and reports why;
(b) the child's decision procedure (`run_candidates`) driven on the emulated C ABI: a candidate whose results differ from the
default kernels' is rejected even when it is faster, a correct and faster one is accepted, a correct but slower one is not,
a candidate that raises ends the search with what was accepted before it."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

G = os.path.join(os.path.dirname(__file__), 'golden')


def test_parent_falls_back_to_defaults_when_the_child_cannot_run():
    from cold_diffusion_models_b200 import tuning, engine
    r = tuning.autotune(dim=32, dim_mults=(1, 2), image_size=32, batch=2, timeout=120)
    assert r['accepted'] == {}
    assert 'error' in r['report'] and r['report']['seconds'] >= 0
    assert engine.batched_repack() is False


@pytest.fixture()
def emu(monkeypatch):
    import abi_emulator
    from cold_diffusion_models_b200 import engine
    monkeypatch.setattr(torch.Tensor, 'is_cuda', property(lambda self: True))
    with abi_emulator.patched():
        yield
    engine.batched_repack(False)
    from cold_diffusion_models_b200 import trainer
    trainer.merge_micro_batches(False)


def test_decision_procedure_on_the_emulated_abi(emu, monkeypatch):
    import cold_diffusion_models_b200 as cdm
    from cold_diffusion_models_b200 import tuning, engine, _lib
    from cold_diffusion_models_b200 import trainer as trainer_mod
    z = np.load(os.path.join(G, 'unet_small.npz'))
    g = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}
    with contextlib.redirect_stdout(io.StringIO()):
        u = cdm.Unet(dim=32, dim_mults=(1, 2), channels=3)
    u.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith('sd:')})

    state = {'conv_2cta': 1, 'conv_staged_epilogue': 0, 'linattn_staged': 0, 'wgrad_bias_fusion': 0, 'layernorm_multi': 0, 'conv_simt_preload': 0}

    class FakeLib:                                       # the emulator has no kernel variants: record the switches instead
        def cd_conv_tc_set_staged_epilogue(self, v):
            state['conv_staged_epilogue'] = v
            return 0

        def cd_linattn_set_staged(self, v):
            state['linattn_staged'] = v
            return 0

        def cd_conv_simt_set_preload(self, v):
            state['conv_simt_preload'] = v
            return 0

        def cd_layernorm_set_multi(self, v):
            state['layernorm_multi'] = v
            return 0

        def cd_conv_tc_set_2cta(self, v):
            state['conv_2cta'] = v
            return 0

        def cd_wgrad_tc_set_bias_fusion(self, v):
            state['wgrad_bias_fusion'] = v
            return 0
    monkeypatch.setattr(_lib, 'lib', FakeLib())

    # "kernel variants" of the emulation: conv_staged_epilogue == 2 corrupts the output (a broken kernel that is very fast),
    # linattn_staged is correct but slow, batched_repack is correct and fast, mode 1 is correct and fast
    real_forward = type(u.engine).forward

    def forward(self, x, time, save=None, out=None):
        y = real_forward(self, x, time, save=save, out=out)
        if state['conv_staged_epilogue'] == 2:
            y[0, 0, 0, 0] += 1.0
        return y
    monkeypatch.setattr(type(u.engine), 'forward', forward)

    def timer(fn, n):                                    # (the decision logic is under test, not the clock: fn is not run)
        ms = 100.0
        if state['conv_staged_epilogue'] == 1:
            ms -= 10
        if state['conv_staged_epilogue'] == 2:
            ms -= 50
        if state['linattn_staged']:
            ms += 5
        if engine.batched_repack():
            ms -= 3
        if trainer_mod.merge_micro_batches():
            ms -= 7
        return ms
    reports = []
    xs = [g['x'], g['x'].flip(0) * 0.5]
    tgs = [g['target'], g['target'].flip(0)]
    ts = [g['t'], g['t'].flip(0)]
    cands = [c for c in tuning.CANDIDATES if not c[0].startswith('conv_2cta') and c[0] not in ('conv_staged_epilogue_mid_k', 'layernorm_multi', 'conv_simt_preload', 'wgrad_bias_fusion')]   # (these add nothing to the logic)
    rep = tuning.run_candidates(u, xs, tgs, ts, lambda: None, timer, 1, reports.append, candidates=cands)
    assert rep['complete'] and len(reports) == len(cands) + 3 and rep['inference_forward']['err_output'] <= rep['inference_forward']['tolerance']
    rows = {r['name']: r for r in rep['candidates']}
    assert rows['conv_staged_epilogue_short_k'].get('accepted') and rows['conv_staged_epilogue_short_k']['err_grad'] == 0.0
    assert 'rejected' in rows['conv_staged_epilogue_all'] and 'ms' not in rows['conv_staged_epilogue_all']      # wrong result: never timed
    assert not rows['linattn_staged'].get('accepted') and rows['linattn_staged']['ms'] == 95.0               # right but slower
    assert rows['batched_repack'].get('accepted')
    assert rows['merge_micro_batches'].get('accepted') and rows['merge_micro_batches']['err_grad'] < 1e-5    # same gradient, other summation order
    assert rep['accepted'] == {'conv_staged_epilogue': 1, 'batched_repack': 1, 'merge_micro_batches': 1} and rep['best_ms'] == 80.0
    assert state == {'conv_2cta': 1, 'conv_staged_epilogue': 1, 'linattn_staged': 0, 'wgrad_bias_fusion': 0, 'layernorm_multi': 0, 'conv_simt_preload': 0} and engine.batched_repack() is True      # left applied

    # a candidate that raises ends the search; what was accepted before it stands
    def boom(v):
        raise RuntimeError('CUDA error: an illegal memory access was encountered')
    monkeypatch.setattr(_lib.lib, 'cd_linattn_set_staged', lambda v: boom(v) if v else 0, raising=False)
    reports2 = []
    rep2 = tuning.run_candidates(u, xs, tgs, ts, lambda: None, timer, 1, reports2.append, candidates=[c for c in cands if c[0] in ('conv_staged_epilogue_short_k', 'linattn_staged')])
    assert 'complete' not in rep2 and rep2['accepted'] == {'conv_staged_epilogue': 1}
    assert rep2['candidates'][-1]['name'] == 'linattn_staged' and 'raised' in rep2['candidates'][-1]['rejected']


def test_trainer_step_over_merged_micro_batches_equals_the_accumulated_one(emu, monkeypatch, tmp_path):
    """Trainer.train_step with merge_micro_batches(): one pass over the concatenated micro-batches gives the parameters (Adam +
    EMA) of the two accumulated passes (DB:1188-1204); on the CPU generator one randint(2B) equals two randint(B) draws"""
    import cold_diffusion_models_b200 as cdm
    from cold_diffusion_models_b200 import trainer as trainer_mod
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    z = np.load(os.path.join(G, 'unet_small.npz'))
    g = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}
    out, calls = [], []
    for merged in (False, True):
        trainer_mod.merge_micro_batches(merged)
        with contextlib.redirect_stdout(io.StringIO()):
            u = cdm.Unet(dim=32, dim_mults=(1, 2), channels=3)
        u.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith('sd:')})
        gd = cdm.GaussianDiffusion(u, image_size=32, device_of_kernel='cpu', channels=3, timesteps=4, kernel_std=0.15, kernel_size=7,
                                   blur_routine='Exponential_reflect', sampling_routine='x0_step_down', loss_type='l2')
        with contextlib.redirect_stdout(io.StringIO()):
            tr = cdm.Trainer(gd, None, image_size=32, train_batch_size=2, train_lr=1e-3, gradient_accumulate_every=2,
                             results_folder=str(tmp_path), dataset='synthetic', step_start_ema=0, update_ema_every=1, ema_decay=0.9)
        n = [0]
        real = type(u.engine).forward

        def counting(self, *a, _real=real, **k):
            n[0] += 1
            return _real(self, *a, **k)
        monkeypatch.setattr(type(u.engine), 'forward', counting)
        gen = torch.Generator().manual_seed(5)
        torch.manual_seed(11)
        losses = []
        for step in range(2):
            bs = [torch.rand(2, 3, 32, 32, generator=gen) * 2 - 1 for _ in range(2)]
            losses.append(float(tr.train_step(batches=bs)))
            tr.step += 1
        monkeypatch.setattr(type(u.engine), 'forward', real)
        calls.append(n[0])
        out.append(({k: v.clone() for k, v in gd.denoise_fn.state_dict().items()}, losses))
    assert calls == [4, 2]                                   # two optimizer steps: 2 x 2 forward passes against 2 x 1
    for a, b in zip(out[0][1], out[1][1]):
        assert abs(a - b) < 1e-6 * max(1.0, abs(a))
    for k in out[0][0]:
        r = ((out[0][0][k].double() - out[1][0][k].double()).norm() / (out[0][0][k].double().norm() + 1e-30)).item()
        assert r < 1e-5, (k, r)


def test_forward_only_variants_get_a_second_chance_on_the_inference_forward(emu, monkeypatch):
    """a variant that is in the noise of a training step but measurably faster on the no_grad forward is accepted for sampling
    only (`accepted_sampling`); the CUDA-graph check cannot run on the CPU and is recorded as rejected, nothing else is affected"""
    import cold_diffusion_models_b200 as cdm
    from cold_diffusion_models_b200 import tuning, _lib
    z = np.load(os.path.join(G, 'unet_small.npz'))
    g = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}
    with contextlib.redirect_stdout(io.StringIO()):
        u = cdm.Unet(dim=32, dim_mults=(1, 2), channels=3)
    u.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith('sd:')})
    state = {}

    class FakeLib:
        def __getattr__(self, name):
            def setter(v, _n=name):
                state[_n] = v
                return 0
            return setter
    monkeypatch.setattr(_lib, 'lib', FakeLib())

    def timer(fn, n):
        if torch.is_grad_enabled():
            return 100.0 - (10 if state.get('cd_conv_tc_set_staged_epilogue') == 1 else 0)      # layernorm_multi: no visible gain
        return 10.0 - (1 if state.get('cd_layernorm_set_multi') else 0) + (2 if state.get('cd_linattn_set_staged') else 0)
    cands = [c for c in tuning.CANDIDATES if c[0] in ('conv_staged_epilogue_short_k', 'linattn_staged', 'layernorm_multi')]
    rep = tuning.run_candidates(u, [g['x']], [g['target']], [g['t']], lambda: None, timer, 1, lambda r: None, candidates=cands,
                                sampling_graph=True)
    assert rep['complete'] and rep['accepted'] == {'conv_staged_epilogue': 1}
    assert rep['accepted_sampling'] == {'layernorm_multi': 4}
    names = [r['name'] for r in rep['sampling_candidates']]
    assert names == ['linattn_staged', 'layernorm_multi', 'revert_conv_staged_epilogue']      # the accepted one is not tried again, only its revert
    assert 'rejected' in rep['sampling'] and not rep.get('sampling_cuda_graph')
    assert state['cd_layernorm_set_multi'] == 0 and state['cd_conv_tc_set_staged_epilogue'] == 1       # training switches left applied

    # the other direction: a switch accepted on the training step that slows the bare forward is reverted for sampling only
    def timer2(fn, n):
        if torch.is_grad_enabled():
            return 100.0 - (10 if state.get('cd_conv_tc_set_staged_epilogue') == 1 else 0)
        return 10.0 + (0.5 if state.get('cd_conv_tc_set_staged_epilogue') == 1 else 0)
    rep = tuning.run_candidates(u, [g['x']], [g['target']], [g['t']], lambda: None, timer2, 1, lambda r: None, candidates=cands[:1],
                                sampling_graph=True)
    assert rep['accepted'] == {'conv_staged_epilogue': 1} and rep['accepted_sampling'] == {'conv_staged_epilogue': 0}
    assert state['cd_conv_tc_set_staged_epilogue'] == 1                                                # training configuration left applied
