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
