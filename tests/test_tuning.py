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


def test_decision_procedure_on_the_emulated_abi(emu, monkeypatch):
    import cold_diffusion_models_b200 as cdm
    from cold_diffusion_models_b200 import tuning, engine, _lib
    z = np.load(os.path.join(G, 'unet_small.npz'))
    g = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}
    with contextlib.redirect_stdout(io.StringIO()):
        u = cdm.Unet(dim=32, dim_mults=(1, 2), channels=3)
    u.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith('sd:')})

    state = {'conv_staged_epilogue': 0, 'linattn_staged': 0, 'wgrad_bias_fusion': 0}

    class FakeLib:                                       # the emulator has no kernel variants: record the switches instead
        def cd_conv_tc_set_staged_epilogue(self, v):
            state['conv_staged_epilogue'] = v
            return 0

        def cd_linattn_set_staged(self, v):
            state['linattn_staged'] = v
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

    def timer(fn, n):
        fn()
        ms = 100.0
        if state['conv_staged_epilogue'] == 1:
            ms -= 10
        if state['conv_staged_epilogue'] == 2:
            ms -= 50
        if state['linattn_staged']:
            ms += 5
        if engine.batched_repack():
            ms -= 3
        return ms
    reports = []
    rep = tuning.run_candidates(u, g['x'], g['target'], g['t'], lambda: None, timer, 1, reports.append)
    assert rep['complete'] and len(reports) == len(tuning.CANDIDATES) + 2
    rows = {r['name']: r for r in rep['candidates']}
    assert rows['conv_staged_epilogue_short_k'].get('accepted') and rows['conv_staged_epilogue_short_k']['err_grad'] == 0.0
    assert 'rejected' in rows['conv_staged_epilogue_all'] and 'ms' not in rows['conv_staged_epilogue_all']      # wrong result: never timed
    assert not rows['linattn_staged'].get('accepted') and rows['linattn_staged']['ms'] == 95.0               # right but slower
    assert rows['batched_repack'].get('accepted')
    assert rep['accepted'] == {'conv_staged_epilogue': 1, 'batched_repack': 1} and rep['best_ms'] == 87.0
    assert not rows['wgrad_bias_fusion'].get('accepted') and rows['wgrad_bias_fusion']['ms'] == 87.0        # right, no gain
    assert state == {'conv_staged_epilogue': 1, 'linattn_staged': 0, 'wgrad_bias_fusion': 0} and engine.batched_repack() is True      # left applied

    # a candidate that raises ends the search; what was accepted before it stands
    def boom(v):
        raise RuntimeError('CUDA error: an illegal memory access was encountered')
    monkeypatch.setattr(_lib.lib, 'cd_linattn_set_staged', lambda v: boom(v) if v else 0, raising=False)
    reports2 = []
    rep2 = tuning.run_candidates(u, g['x'], g['target'], g['t'], lambda: None, timer, 1, reports2.append)
    assert 'complete' not in rep2 and rep2['accepted'] == {'conv_staged_epilogue': 1}
    assert rep2['candidates'][-1]['name'] == 'linattn_staged' and 'raised' in rep2['candidates'][-1]['rejected']
