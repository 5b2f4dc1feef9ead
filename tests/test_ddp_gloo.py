"""N>1 host logic on CPU (gloo, world_size 2): the flat gradient layout + single all-reduce + 1/N scale used by
Trainer.train_step reproduce single-process training on the concatenated batch (the reference's DataParallel
semantics: loss = mean of equal-size per-replica means, DB:1192)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import unet_oracle as UO
    from cold_diffusion_models_b200.engine_bwd import flat_offsets, allreduce_mean_
    torch.set_num_threads(2)
    sd = {k: v.clone().requires_grad_(True) for k, v in UO.make_unet_state_dict(32, (1,), 3, seed=1).items()}
    g = torch.Generator().manual_seed(7)
    x = torch.rand(4, 3, 16, 16, generator=g) * 2 - 1
    t = torch.tensor([0, 3, 1, 2])
    tgt = torch.rand(4, 3, 16, 16, generator=g) * 2 - 1
    lo, hi = rank * 2, rank * 2 + 2                      # per-rank shard of the global batch
    (tgt[lo:hi] - UO.unet_forward(sd, x[lo:hi], t[lo:hi])).abs().mean().backward()
    table, total = flat_offsets([(n, p.numel()) for n, p in sd.items()])
    flat = torch.zeros(total)
    for n, p in sd.items():
        off, k = table[n]
        assert off % 4 == 0
        flat[off:off + k] = p.grad.reshape(-1)
    scale = allreduce_mean_(flat, world)
    flat *= scale
    if rank == 0:
        torch.save(dict(flat=flat, table=table), out)
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_full_batch(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import unet_oracle as UO
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    sd = {k: v.clone().requires_grad_(True) for k, v in UO.make_unet_state_dict(32, (1,), 3, seed=1).items()}
    g = torch.Generator().manual_seed(7)
    x = torch.rand(4, 3, 16, 16, generator=g) * 2 - 1
    t = torch.tensor([0, 3, 1, 2])
    tgt = torch.rand(4, 3, 16, 16, generator=g) * 2 - 1
    (tgt - UO.unet_forward(sd, x, t)).abs().mean().backward()
    for n, p in sd.items():
        off, k = got['table'][n]
        a, b = got['flat'][off:off + k], p.grad.reshape(-1)
        assert torch.allclose(a, b, atol=1e-6, rtol=1e-4), n


# ---------------------------------------------------------------------------------------------------------------------------
# the real Trainer.train_step on two gloo ranks (CPU tensors, C ABI emulated by tests/abi_emulator.py): flat-buffer gradients,
# ONE all-reduce per optimizer step, 1/N folded into the fused Adam -- both ranks must end with the same weights, equal to a
# single-process step on the concatenated batch
# ---------------------------------------------------------------------------------------------------------------------------
def _small_gd(seed_sd=3):
    import io, contextlib
    import unet_oracle as UO
    import cold_diffusion_models_b200 as cdm
    sd = UO.make_unet_state_dict(32, (1, 2), 3, seed=seed_sd)
    with contextlib.redirect_stdout(io.StringIO()):
        u = cdm.Unet(dim=32, dim_mults=(1, 2), channels=3)
    u.load_state_dict(sd)
    return cdm.GaussianDiffusion(u, image_size=32, device_of_kernel='cpu', channels=3, timesteps=4, kernel_std=0.15, kernel_size=7,
                                 blur_routine='Exponential_reflect', sampling_routine='x0_step_down', loss_type='l2')


def _batches():
    g = torch.Generator().manual_seed(11)
    return [torch.rand(2, 3, 32, 32, generator=g) * 2 - 1 for _ in range(2)]


def _trainer_worker(rank, world, port, out):
    import io, contextlib
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    import abi_emulator
    import cold_diffusion_models_b200 as cdm
    torch.Tensor.is_cuda = property(lambda self: True)
    torch.Tensor.cuda = lambda self, *a, **k: self
    with abi_emulator.patched(), contextlib.redirect_stdout(io.StringIO()):
        gd = _small_gd(seed_sd=3 + rank)                    # ranks start from DIFFERENT weights ...
        tr = cdm.Trainer(gd, None, image_size=32, train_batch_size=2, train_lr=1e-3, gradient_accumulate_every=1,
                         results_folder=os.path.dirname(out), dataset='synthetic')
        assert tr._world == world and tr._rank == rank
        assert tr.ds.seed == 1234 + 1000003 * rank          # ... read different images ...
        import unet_oracle as UO                            # ... and the Trainer broadcasts rank 0's parameters and EMA weights
        want = UO.make_unet_state_dict(32, (1, 2), 3, seed=3)
        for net in (gd.denoise_fn, tr.ema_model.denoise_fn):
            have = net.state_dict()
            assert all(torch.equal(have[k], want[k]) for k in want)
        torch.manual_seed(100 + rank)                       # forward() draws t ~ randint(0, T, (B,)) from the global generator
        tr.train_step(batches=[_batches()[rank]])
        # the all-reduce of the finished suffix of the flat gradient (ups, mid, final: issued while the down path is still being
        # differentiated) + the rest at the end cover the buffer exactly once
        eng = gd.denoise_fn.engine
        rng = tr.overlapped_ranges
        assert len(rng) >= 1 and rng[0][1] == eng._grad_total and all(a[0] == b[1] for a, b in zip(rng, rng[1:])), rng
        assert 0 < rng[-1][0] < eng._grad_total // 2 and rng[-1][0] == eng._group_start['ups.0']
        late = [n for n, _ in gd.denoise_fn.named_parameters() if n.startswith('time_mlp.') or '.mlp.1.' in n]
        assert late and all(eng._offsets[n][0] < eng._group_start['downs.0'] for n in late)     # end-of-backward gradients sit first
    if rank == 0:
        torch.save({k: v.clone() for k, v in gd.denoise_fn.state_dict().items()}, out)
    else:
        torch.save({k: v.clone() for k, v in gd.denoise_fn.state_dict().items()}, out + '.r1')
    dist.destroy_process_group()


def test_trainer_step_on_two_gloo_ranks_equals_single_process_on_the_concatenated_batch(tmp_path, monkeypatch):
    import io, contextlib
    sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_trainer_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r0, r1 = torch.load(out), torch.load(out + '.r1')
    for k in r0:
        assert torch.equal(r0[k], r1[k]), k                 # replicas stay bit-identical
    # single process: the same four images as one batch, the same t values
    import abi_emulator
    import cold_diffusion_models_b200 as cdm
    monkeypatch.setattr(torch.Tensor, 'is_cuda', property(lambda self: True))
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    ts = []
    for rank in range(2):
        torch.manual_seed(100 + rank)
        ts.append(torch.randint(0, 4, (2,)).long())
    with abi_emulator.patched(), contextlib.redirect_stdout(io.StringIO()):
        gd = _small_gd()
        tr = cdm.Trainer(gd, None, image_size=32, train_batch_size=4, train_lr=1e-3, gradient_accumulate_every=1,
                         results_folder=str(tmp_path), dataset='synthetic')
        gd.p_losses(torch.cat(_batches()), torch.cat(ts)).backward()
        tr.opt.step(ema_mode=1)
    one = gd.denoise_fn.state_dict()
    for k in r0:
        d = (r0[k].double() - one[k].double()).norm() / (one[k].double().norm() + 1e-30)
        assert d < 1e-5, (k, d.item())
