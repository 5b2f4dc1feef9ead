"""Guarded start-up selection of the opt-in kernel variants (`autotune()`).

Some kernel variants of libcolddiff are written and checked from source on the CPU but ship OFF because they have not been
validated on a B200 (NOTES.md, "Code that exists but has not run on a B200 yet"): the line-coalesced epilogue of the tcgen05
convolution (`cd_conv_tc_set_staged_epilogue`), the shared-memory-staged LinearAttention kernels (`cd_linattn_set_staged`), the
one-launch weight repacks (`engine.batched_repack`) and the bias gradient folded into the tcgen05 weight gradient
(`cd_wgrad_tc_set_bias_fusion`).  `autotune()` decides about them the way a BLAS library picks kernels at
start-up, but without trusting them: a CHILD process (so that a faulting or spinning kernel can neither poison this process'
CUDA context nor hang it -- the child is killed after `timeout` seconds) builds the network the caller is about to run, executes
one training micro-step (p_losses forward + backward) per candidate, and a candidate is accepted only if

  * the network output and EVERY parameter gradient agree with the default kernels' to within a small multiple of the
    run-to-run noise of the default kernels themselves (float atomics in the LinearAttention context and the weight-gradient
    split-K make two default runs differ in the last bits), and
  * the step is measurably faster than the best accepted configuration so far.

The parent applies the accepted switches to its own library instance and returns the child's report; if the child fails, hangs,
or finds nothing faster, every switch stays at its default.  Nothing here touches the arithmetic: every candidate computes the
same sums in the same order as the default kernel it replaces (see the kernels' headers).
"""
import json
import os
import subprocess
import sys
import time

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)

# (name, {switch: value}) in the order they are tried; each is tried on top of what has been accepted so far
CANDIDATES = [
    ('conv_staged_epilogue_short_k', {'conv_staged_epilogue': 1}),
    ('conv_staged_epilogue_mid_k', {'conv_staged_epilogue': 3}),
    ('conv_staged_epilogue_all', {'conv_staged_epilogue': 2}),
    ('linattn_staged', {'linattn_staged': 1}),
    ('batched_repack', {'batched_repack': 1}),
    ('layernorm_multi', {'layernorm_multi': 4}),
    ('conv_simt_preload', {'conv_simt_preload': 1}),
    # the SM-pair convolution kernel is validated; its default (1) follows a tile cost model fitted to batch 32 -- try the extremes
    ('conv_2cta_everywhere', {'conv_2cta': 2}),
    ('conv_2cta_off', {'conv_2cta': 0}),
    # last, because it is tcgen05 code that has never run: if it takes the child down, everything above has been decided
    ('wgrad_bias_fusion', {'wgrad_bias_fusion': 1}),
    # evaluated on top of everything else and reported separately by bench.py (it changes the shape of the step, not a kernel)
    ('merge_micro_batches', {'merge_micro_batches': 1}),
]
# candidates that change forward kernels (tried again on the inference forward alone)
FORWARD_CANDIDATES = ('conv_staged_epilogue_short_k', 'conv_staged_epilogue_mid_k', 'conv_staged_epilogue_all', 'linattn_staged', 'layernorm_multi',
                      'conv_simt_preload', 'conv_2cta_everywhere', 'conv_2cta_off')
FORWARD_SWITCHES = ('conv_staged_epilogue', 'linattn_staged', 'layernorm_multi', 'conv_simt_preload', 'conv_2cta')
DEFAULTS = {'conv_2cta': 1, 'conv_staged_epilogue': 0, 'linattn_staged': 1, 'batched_repack': 0, 'layernorm_multi': 4, 'conv_simt_preload': 1, 'merge_micro_batches': 0, 'wgrad_bias_fusion': 0}


def apply(settings):
    """set the library / engine switches named in `settings` (missing keys: unchanged)"""
    from . import _lib, engine
    if 'conv_2cta' in settings:
        _lib.lib.cd_conv_tc_set_2cta(int(settings['conv_2cta']))
    if 'conv_staged_epilogue' in settings:
        _lib.lib.cd_conv_tc_set_staged_epilogue(int(settings['conv_staged_epilogue']))
    if 'linattn_staged' in settings:
        _lib.lib.cd_linattn_set_staged(int(settings['linattn_staged']))
    if 'conv_simt_preload' in settings:
        _lib.lib.cd_conv_simt_set_preload(int(settings['conv_simt_preload']))
    if 'layernorm_multi' in settings:
        _lib.lib.cd_layernorm_set_multi(int(settings['layernorm_multi']))
    if 'batched_repack' in settings:
        engine.batched_repack(bool(settings['batched_repack']))
    if 'merge_micro_batches' in settings:
        from . import trainer
        trainer.merge_micro_batches(bool(settings['merge_micro_batches']))
    if 'wgrad_bias_fusion' in settings:
        _lib.lib.cd_wgrad_tc_set_bias_fusion(int(settings['wgrad_bias_fusion']))


def autotune(dim=64, dim_mults=(1, 2, 4, 8), channels=3, image_size=128, batch=32, accum=2, device=0, timeout=300, steps=5, verbose=False):
    """-> {'accepted': {switch: value}, 'report': {...}}; the accepted switches are applied to this process.  See the module text."""
    cmd = [sys.executable, '-m', 'cold_diffusion_models_b200.tuning', '--child', json.dumps(dict(
        dim=dim, dim_mults=list(dim_mults), channels=channels, image_size=image_size, batch=batch, accum=accum, device=device, steps=steps))]
    env = dict(os.environ)
    env['PYTHONPATH'] = ROOT + os.pathsep + env.get('PYTHONPATH', '')
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'MASTER_ADDR', 'MASTER_PORT', 'GROUP_RANK', 'LOCAL_WORLD_SIZE', 'TORCHELASTIC_RUN_ID',
              'COLDDIFF_CONV_STAGED_EPILOGUE', 'COLDDIFF_LINATTN_STAGED', 'COLDDIFF_BATCHED_REPACK'):
        env.pop(k, None)                       # the child is a plain single-GPU process starting from the library defaults
    t0 = time.time()
    report = None
    def last_report(text):
        """the child prints a cumulative report after every candidate: the last complete one counts (what was accepted BEFORE a
        candidate that crashed or hung the child has been validated)"""
        rep = None
        if isinstance(text, bytes):
            text = text.decode(errors='replace')
        for line in (text or '').splitlines():
            if line.startswith('AUTOTUNE_REPORT '):
                try:
                    rep = json.loads(line[len('AUTOTUNE_REPORT '):])
                except ValueError:
                    pass
        return rep

    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
        report = last_report(r.stdout)
        if report is None:
            report = {'error': 'child exited %d without a report' % r.returncode, 'stderr_tail': r.stderr[-600:]}
        elif r.returncode != 0:
            report['child_exit'] = r.returncode
    except subprocess.TimeoutExpired as e:
        report = last_report(e.stdout) or {}
        report['error_after'] = 'child killed after %d s' % timeout
    except Exception as e:  # pragma: no cover - spawn failures
        report = {'error': repr(e)[:300]}
    report['seconds'] = round(time.time() - t0, 1)
    accepted = dict(report.get('accepted', {})) if 'error' not in report else {}
    if accepted:
        apply(accepted)
    if verbose:
        print('autotune:', json.dumps(report), file=sys.stderr)
    return {'accepted': accepted, 'report': report}


# ----------------------------------------------------------------------------------------------------------------------
# child process
# ----------------------------------------------------------------------------------------------------------------------
def _child(cfg):
    import contextlib
    import io
    import torch
    import cold_diffusion_models_b200 as cdm
    dev = torch.device('cuda', int(cfg['device']))
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        unet = cdm.Unet(dim=cfg['dim'], dim_mults=tuple(cfg['dim_mults']), channels=cfg['channels']).to(dev)
    B, S, Cc = cfg['batch'], cfg['image_size'], cfg['channels']
    g = torch.Generator().manual_seed(7)
    A = int(cfg.get('accum', 2))
    x = [(torch.rand(B, Cc, S, S, generator=g) * 2 - 1).to(dev) for _ in range(A)]
    target = [(torch.rand(B, Cc, S, S, generator=g) * 2 - 1).to(dev) for _ in range(A)]
    t = [torch.randint(0, 200, (B,), generator=g).to(dev) for _ in range(A)]

    def timer(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    run_candidates(unet, x, target, t, torch.cuda.synchronize, timer, int(cfg.get('steps', 3)),
                   emit=lambda rep: print('AUTOTUNE_REPORT ' + json.dumps(rep), flush=True), sampling_graph=True)


def run_candidates(unet, x, target, t, sync, timer, steps, emit, candidates=None, min_gain=0.005, sampling_graph=False):
    """the child's decision procedure (device-agnostic so that tests can drive it on the emulated C ABI): tries CANDIDATES in
    order on top of what has been accepted, calls emit(cumulative report) after every decision; returns the final report"""
    import torch
    from cold_diffusion_models_b200.deblurring import _LossFn
    candidates = CANDIDATES if candidates is None else candidates

    from cold_diffusion_models_b200 import trainer as _trainer
    if torch.is_tensor(x):                              # one micro-batch given: a list of one
        x, target, t = [x], [target], [t]
    A = len(x)
    xc, tc, tgc = (torch.cat(x), torch.cat(t), torch.cat(target)) if A > 1 else (x[0], t[0], target[0])

    def step():
        """the gradient of one optimizer step, the way Trainer.train_step accumulates it: A micro-batches with loss / A each, or
        (merge_micro_batches) one pass over their concatenation"""
        if A > 1 and _trainer.merge_micro_batches():
            y = unet(xc, tc)
            _LossFn.apply(tgc, y, 1).backward()
            return y
        ys = []
        for xi, ti, gi in zip(x, t, target):
            y = unet(xi, ti)
            (_LossFn.apply(gi, y, 1) / A).backward()
            ys.append(y)
        return torch.cat(ys) if A > 1 else ys[0]

    def run():
        """one micro-step with the smooth L2 loss (an L1 loss' sign() gradient would turn 1e-4-level output noise into
        O(1) gradient differences) -> (output, flat gradient)"""
        for p in unet.parameters():
            p.grad = None
        if getattr(unet.engine, 'flat_grad', None) is not None:
            unet.engine.flat_grad.zero_()
        y = step()
        sync()
        return y.detach().clone(), unet.engine.flat_grad.detach().clone()

    def rel(a, b):
        return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()

    def block_errors(ga, gb):
        """worst relative error over the per-parameter slices of the flat gradient (a wrong small tensor must not hide
        behind the norm of the big ones)"""
        worst = 0.0
        for n, (off, k) in unet.engine._offsets.items():
            a, b = ga[off:off + k], gb[off:off + k]
            nb = b.double().norm().item()
            if nb > 0:
                worst = max(worst, (a.double() - b.double()).norm().item() / nb)
        return worst

    apply(DEFAULTS)
    run()                                               # warm-up: buffers, packed weights, function attributes
    y0, g0 = run()
    y1, g1 = run()
    noise_y, noise_g = rel(y1, y0), block_errors(g1, g0)
    tol_y, tol_g = max(1e-4, 8 * noise_y), max(2e-3, 8 * noise_g)
    timer(step, 1)
    base_ms = timer(step, steps)
    report = {'noise': {'output': noise_y, 'grad': noise_g}, 'tolerance': {'output': tol_y, 'grad': tol_g}, 'default_ms': base_ms,
              'candidates': [], 'accepted': {}, 'best_ms': base_ms}
    emit(report)
    accepted, best_ms = {}, base_ms
    for name, sw in candidates:
        trial = dict(accepted)
        trial.update(sw)
        row = {'name': name, 'switches': sw}
        try:
            apply(dict(DEFAULTS, **trial))
            run()                                       # first use of the variant: attributes, job tables
            y, gr = run()
            ey, eg = rel(y, y0), block_errors(gr, g0)
            row.update(err_output=ey, err_grad=eg, finite=bool(torch.isfinite(y).all().item() and torch.isfinite(gr).all().item()))
            if row['finite'] and ey <= tol_y and eg <= tol_g:
                timer(step, 1)
                ms = timer(step, steps)
                # A/B: the configuration accepted so far is timed again right after the candidate (clocks drift over the run)
                apply(dict(DEFAULTS, **accepted))
                timer(step, 1)
                ref = timer(step, steps)
                row['ms'], row['ms_reference'] = ms, ref
                if ms < (1.0 - min_gain) * ref:
                    accepted, best_ms = trial, ms
                    row['accepted'] = True
            else:
                row['rejected'] = 'results differ from the default kernels'
        except Exception as e:
            row['rejected'] = 'raised: ' + repr(e)[:200]
            report['candidates'].append(row)
            report['accepted'], report['best_ms'] = accepted, best_ms
            emit(report)
            return report                               # a CUDA error is sticky: nothing after it can be trusted
        report['candidates'].append(row)
        report['accepted'], report['best_ms'] = accepted, best_ms
        emit(report)
    # ---- inference forward (sampling loops run it under no_grad: no saved tensors, other epilogue variants): the accepted
    # switches must reproduce the default kernels there too, or nothing is accepted at all
    if accepted:
        try:
            with torch.no_grad():
                apply(DEFAULTS)
                yd = unet(x[0], t[0]).clone()
                yd2 = unet(x[0], t[0]).clone()
                apply(dict(DEFAULTS, **accepted))
                ya = unet(x[0], t[0]).clone()
                sync()
            err, tol = rel(ya, yd), max(1e-4, 8 * rel(yd2, yd))
            report['inference_forward'] = {'err_output': err, 'tolerance': tol}
            if not (err <= tol and bool(torch.isfinite(ya).all().item())):
                report['inference_forward']['rejected'] = 'accepted switches change the no_grad forward: all switches dropped'
                accepted, best_ms = {}, base_ms
        except Exception as e:
            report['inference_forward'] = {'rejected': 'raised: ' + repr(e)[:200]}
            accepted, best_ms = {}, base_ms
        report['accepted'], report['best_ms'] = accepted, best_ms
        emit(report)
    apply(dict(DEFAULTS, **accepted))
    # ---- forward-only decisions for the sampling loops: a variant that only touches forward kernels can be worth 3 % of a reverse
    # step and still disappear in the noise of a training step; candidates not accepted above are tried again on the no_grad forward
    samp = {}
    if sampling_graph:
        try:
            with torch.no_grad():
                fwd = lambda: unet(x[0], t[0])
                apply(DEFAULTS)
                yd = fwd().clone()
                yd2 = fwd().clone()
                sync()
                tol = max(1e-4, 8 * rel(yd2, yd))
                rows = []
                for name, sw in candidates:
                    if name not in FORWARD_CANDIDATES or all(accepted.get(k) == v for k, v in sw.items()):
                        continue
                    row = {'name': name}
                    trial = dict(accepted, **samp)
                    trial.update(sw)
                    apply(dict(DEFAULTS, **trial))
                    ya = fwd().clone()
                    ya = fwd().clone()
                    sync()
                    row['err_output'] = rel(ya, yd)
                    if row['err_output'] <= tol and bool(torch.isfinite(ya).all().item()):
                        timer(fwd, 1)
                        ms = timer(fwd, 2 * steps)
                        apply(dict(DEFAULTS, **dict(accepted, **samp)))
                        timer(fwd, 1)
                        ref = timer(fwd, 2 * steps)
                        row['ms'], row['ms_reference'] = ms, ref
                        if ms < (1.0 - min_gain) * ref:
                            samp.update(sw)
                            row['accepted'] = True
                    else:
                        row['rejected'] = 'results differ from the default kernels'
                    rows.append(row)
                # ... and the other way round: a switch accepted on the training step that slows the bare forward down is put back
                # to its default for sampling
                for key in sorted(k for k in accepted if k in FORWARD_SWITCHES and k not in samp):
                    row = {'name': 'revert_' + key}
                    cur = dict(accepted, **samp)
                    apply(dict(DEFAULTS, **dict(cur, **{key: DEFAULTS[key]})))
                    timer(fwd, 1)
                    ms = timer(fwd, 2 * steps)
                    apply(dict(DEFAULTS, **cur))
                    timer(fwd, 1)
                    ref = timer(fwd, 2 * steps)
                    row['ms'], row['ms_reference'] = ms, ref
                    if ms < (1.0 - min_gain) * ref:
                        samp[key] = DEFAULTS[key]
                        row['accepted'] = True
                    rows.append(row)
                report['sampling_candidates'] = rows
        except Exception as e:
            report['sampling_candidates_error'] = repr(e)[:200]
            samp = {}
        report['accepted_sampling'] = samp
        emit(report)
    apply(dict(DEFAULTS, **dict(accepted, **samp)))
    # ---- CUDA-graph replay of that forward (engine.enable_cuda_graph) against the eager launches -----
    if sampling_graph:
        row = {'name': 'sampling_cuda_graph'}
        try:
            with torch.no_grad():
                fwd = lambda: unet(x[0], t[0])
                ye = fwd().clone()
                ye2 = fwd().clone()
                sync()
                tol = max(1e-4, 8 * rel(ye2, ye))
                timer(fwd, 1)
                ms_eager = timer(fwd, steps)
                unet.engine.enable_cuda_graph(True)
                try:
                    yg = fwd().clone()                 # captures
                    yg = fwd().clone()                 # replays
                    sync()
                    row.update(err_output=rel(yg, ye), ms_eager=ms_eager)
                    if row['err_output'] <= tol and bool(torch.isfinite(yg).all().item()):
                        timer(fwd, 1)
                        row['ms'] = timer(fwd, steps)
                        if row['ms'] < (1.0 - min_gain) * ms_eager:
                            row['accepted'] = True
                            report['sampling_cuda_graph'] = True
                    else:
                        row['rejected'] = 'results differ from the eager launches'
                finally:
                    unet.engine.enable_cuda_graph(False)
        except Exception as e:
            row['rejected'] = 'raised: ' + repr(e)[:200]
        report['sampling'] = row
    apply(dict(DEFAULTS, **accepted))
    report['complete'] = True
    emit(report)
    return report


if __name__ == '__main__':
    if len(sys.argv) >= 3 and sys.argv[1] == '--child':
        _child(json.loads(sys.argv[2]))
    else:
        print(json.dumps(autotune(verbose=True)))
