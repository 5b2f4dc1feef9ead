"""CPU checks of the host-side tables the product path builds (no kernels run): they must equal what the unmodified
reference builds (tests/golden/*.npz), independently of the oracle."""
import os
import numpy as np
import torch

G = os.path.join(os.path.dirname(__file__), 'golden')


def load(name):
    z = np.load(os.path.join(G, name + '.npz'))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def test_defading_generation_schedule_tables_match_reference():
    from cold_diffusion_models_b200.defading_generation import get_kernels_with_schedule, get_reverse_kernels_with_schedule
    g = load('defading_gen_small')
    a = get_kernels_with_schedule(4, 32, 0.6, 3)
    assert tuple(a.shape) == tuple(g['rev0:alphas'].shape) == (4, 1, 32, 32)
    assert torch.allclose(a, g['rev0:alphas'], atol=1e-6) and torch.allclose(1. - a, g['rev0:one_minus_alphas'], atol=1e-6)
    om = get_reverse_kernels_with_schedule(4, 32, 0.6, 3)
    assert torch.allclose(om, g['rev1:one_minus_alphas'], atol=1e-6) and torch.allclose(1. - om, g['rev1:alphas'], atol=1e-6)
    # the last reverse entry is the un-faded image (all ones), the first the product of the first T-1 kernels
    assert torch.equal(om[-1], torch.ones(1, 32, 32))


def test_cosine_schedule_of_the_demixing_package_matches_reference():
    from cold_diffusion_models_b200.denoising import cosine_beta_schedule
    g = load('denoise_small')
    ac = torch.cumprod(1. - cosine_beta_schedule(5), axis=0)
    assert torch.allclose(torch.sqrt(ac), g['sqrt_ac'], atol=1e-7) and torch.allclose(torch.sqrt(1. - ac), g['sqrt_1mac'], atol=1e-7)


def test_package_exports_mirror_the_reference_inits():
    import cold_diffusion_models_b200 as root
    from cold_diffusion_models_b200 import (deblurring_diffusion_pytorch, denoising_diffusion_pytorch, resolution_diffusion_pytorch,
                                            defading_diffusion_pytorch, defading_generation_diffusion_pytorch,
                                            demixing_diffusion_pytorch, snowification_diffusion)
    for pkg in (denoising_diffusion_pytorch, defading_generation_diffusion_pytorch, demixing_diffusion_pytorch):
        assert set(pkg.__all__) == {'GaussianDiffusion', 'Unet', 'Trainer'}
        for n in pkg.__all__:
            assert hasattr(pkg, n)
    for n in ('GaussianDiffusion', 'Unet', 'Trainer', 'Model'):
        assert hasattr(deblurring_diffusion_pytorch, n) and hasattr(root, n)
    # constructor keywords of the two-input packages (DM:311-322, DFGEN:348-361)
    import inspect
    kw = inspect.signature(demixing_diffusion_pytorch.GaussianDiffusion.__init__).parameters
    assert list(kw)[1:] == ['denoise_fn', 'image_size', 'channels', 'timesteps', 'loss_type', 'train_routine', 'sampling_routine', 'discrete']
    kw = inspect.signature(defading_generation_diffusion_pytorch.GaussianDiffusion.__init__).parameters
    assert list(kw)[1:] == ['denoise_fn', 'image_size', 'channels', 'timesteps', 'loss_type', 'train_routine', 'sampling_routine',
                            'reverse', 'kernel_std', 'initial_mask']
    assert kw['kernel_std'].default == 0.15 and kw['initial_mask'].default == 11 and kw['reverse'].default is False
    kw = inspect.signature(demixing_diffusion_pytorch.GaussianDiffusion.gen_sample).parameters
    assert list(kw)[1:] == ['batch_size', 'img', 'noise_level', 't']


def test_state_dict_wire_format_of_every_package_matches_the_reference():
    """checkpoints are {'step','model','ema'} of GaussianDiffusion.state_dict() (DB:1140-1149): every key and shape the reference
    writes must exist with the same shape here (tests/golden/state_dict_format.json, recorded from the unmodified reference)"""
    import io, json, contextlib
    import cold_diffusion_models_b200 as cdm
    from cold_diffusion_models_b200 import (resolution_diffusion_pytorch as rs, defading_diffusion_pytorch as df,
                                            denoising_diffusion_pytorch as dn, demixing_diffusion_pytorch as dm,
                                            defading_generation_diffusion_pytorch as dg, snowification_diffusion as sn)
    fmt = json.load(open(os.path.join(G, 'state_dict_format.json')))
    with contextlib.redirect_stdout(io.StringIO()):
        mk = lambda: cdm.Unet(dim=32, dim_mults=(1, 2), channels=3)
        built = {
            'deblurring': cdm.GaussianDiffusion(mk(), image_size=32, device_of_kernel='cpu', channels=3, timesteps=4, kernel_std=0.15,
                                                kernel_size=7, blur_routine='Exponential_reflect'),
            'deblurring_model': cdm.GaussianDiffusion(cdm.Model(resolution=16, in_channels=3, out_ch=3, ch=32, ch_mult=(1, 2), num_res_blocks=2,
                                                                attn_resolutions=(8,), dropout=0.1), image_size=16, device_of_kernel='cpu',
                                                      channels=3, timesteps=3, kernel_std=0.1, kernel_size=3, blur_routine='Special_6_routine'),
            'resolution': rs.GaussianDiffusion(mk(), image_size=32, device_of_kernel='cpu', channels=3, timesteps=4,
                                               resolution_routine='Incremental_factor_2'),
            'defading': df.GaussianDiffusion(mk(), image_size=32, device_of_kernel='cpu', channels=3, timesteps=4, kernel_std=0.6,
                                             initial_mask=3, fade_routine='Incremental'),
            'denoising': dn.GaussianDiffusion(mk(), image_size=32, channels=3, timesteps=5),
            'demixing': dm.GaussianDiffusion(mk(), image_size=32, channels=3, timesteps=5),
            'defading_generation': dg.GaussianDiffusion(mk(), image_size=32, channels=3, timesteps=4, kernel_std=0.6, initial_mask=3),
            'decolor': sn.GaussianDiffusion(mk(), image_size=32, device_of_kernel='cpu', channels=3, timesteps=4,
                                            forward_process_type='Decolorization', decolor_routine='Linear'),
        }
    for tag, ref in fmt.items():
        mine = {k: list(v.shape) for k, v in built[tag].state_dict().items()}
        missing = [k for k in ref if k not in mine]
        extra = [k for k in mine if k not in ref]
        assert not missing and not extra, (tag, missing[:5], extra[:5])
        bad = [k for k in ref if mine[k] != ref[k]]
        assert not bad, (tag, bad[:5])


def test_api_surface_matches_the_reference():
    """every constructor keyword (name, kind, default) of the exported classes and every public method of GaussianDiffusion with
    its parameter names, as recorded from the unmodified reference (tests/golden/api_surface.json).  Extra parameters on our side
    are allowed when they have defaults (private ones carry a leading underscore)."""
    import inspect, json
    import cold_diffusion_models_b200 as cdm
    from cold_diffusion_models_b200 import (deblurring_diffusion_pytorch as db, resolution_diffusion_pytorch as rs,
                                            defading_diffusion_pytorch as df, denoising_diffusion_pytorch as dn,
                                            demixing_diffusion_pytorch as dm, defading_generation_diffusion_pytorch as dg,
                                            snowification_diffusion as sn)
    mods = dict(deblurring=db, resolution=rs, defading=df, denoising=dn, demixing=dm, defading_generation=dg, snowification=sn)
    api = json.load(open(os.path.join(G, 'api_surface.json')))
    problems = []
    for key, entry in api.items():
        tag, cn = key.split('.')
        cls = getattr(mods[tag], cn)
        for mn, ref in entry.items():
            fn = getattr(cls, mn, None)
            if fn is None:
                problems.append((key, mn, 'missing'))
                continue
            mine = [(n, str(p.kind), None if p.default is inspect._empty else repr(p.default))
                    for n, p in inspect.signature(fn).parameters.items() if n != 'self' and not n.startswith('_')]
            ref = [tuple(r) for r in ref]
            ref_names = {r[0] for r in ref}
            extras = [m for m in mine if m[0] not in ref_names]
            if any(m[2] is None for m in extras):
                problems.append((key, mn, 'extra parameter without a default', extras))
            mine = [m for m in mine if m[0] in ref_names]           # keyword supersets (e.g. `shuffle`) are fine
            if [m[0] for m in mine] != [r[0] for r in ref]:
                problems.append((key, mn, 'names', [m[0] for m in mine], [r[0] for r in ref]))
            elif mn == '__init__' and mine != ref:
                problems.append((key, mn, 'kinds/defaults', [m for m, r in zip(mine, ref) if m != r][:4]))
    assert not problems, problems[:12]


def test_snowification_get_model_and_dataset_exports():
    """`diffusion.model.get_model.get_model` / `diffusion.get_dataset` of the snowification package (get_model.py:4-38)"""
    import io, contextlib
    from cold_diffusion_models_b200.snowification_diffusion import get_model, get_dataset
    from cold_diffusion_models_b200.snowification_diffusion.model.get_model import get_model as gm
    from cold_diffusion_models_b200 import Unet, Model

    class Args:
        pass
    a = Args(); a.model = 'UnetResNet'; a.dataset = 'cifar10_train'
    m = get_model(a)
    assert gm is get_model and isinstance(m, Model) and m.resolution == 32 and m.ch == 128
    a.dataset = 'celebA_train'
    assert get_model(a).resolution == 128
    a.model = 'UnetConvNext'
    with contextlib.redirect_stdout(io.StringIO()):
        u = get_model(a, with_time_emb=False)
    assert isinstance(u, Unet) and u.time_mlp is None
    assert get_dataset('unknown', '/tmp', (32, 32)) is None
