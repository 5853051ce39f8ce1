"""CPU, world_size 2 over gloo: host-side logic of the data-parallel path (sharding, flat gradient bucket,
loss reduction).  The kernels themselves need a GPU and are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from re2nn_seq_b200 import dist as rd
        torch.manual_seed(0)
        model = torch.nn.Module()
        model.a = torch.nn.Parameter(torch.zeros(3, 4))
        model.b = torch.nn.Parameter(torch.zeros(5), requires_grad=False)      # frozen: not in the bucket
        model.c = torch.nn.Parameter(torch.zeros(2))
        model.a.grad = torch.full((3, 4), float(rank + 1))
        model.c.grad = None                                                      # missing grad counts as zero
        bucket = rd.GradBucket(model)
        nbytes = bucket.all_reduce()
        assert nbytes == (12 + 2) * 4
        assert torch.equal(model.a.grad, torch.full((3, 4), 3.0))               # 1 + 2
        assert torch.equal(model.c.grad, torch.zeros(2))
        # second step: gradients are views of the bucket -> backward accumulates into it, no pack / unpack copies
        bucket.zero_grad()
        assert model.a.grad.data_ptr() == bucket.flat.data_ptr()
        ((model.a.sum() + 2.0 * model.c.sum()) * float(rank + 1)).backward()
        bucket.all_reduce()
        assert model.a.grad.data_ptr() == bucket.flat.data_ptr()
        assert torch.equal(model.a.grad, torch.full((3, 4), 3.0))
        assert torch.equal(model.c.grad, torch.full((2,), 6.0))
        assert torch.equal(bucket.flat, torch.cat([torch.full((12,), 3.0), torch.full((2,), 6.0)]))
        # sharding covers the batch exactly once, in order
        lens = torch.arange(1, 8)
        lo, hi = rd.shard_bounds(7, world, rank)
        (mine,) = rd.shard_batch([lens], world, rank)
        assert torch.equal(mine, lens[lo:hi])
        assert rd.global_token_count(mine) == int(lens.sum())
        # CE semantics: per-rank sum / global token count, summed over ranks == global mean
        per_tok = torch.arange(28, dtype=torch.float64)
        offs = np.concatenate([[0], np.cumsum(lens.numpy())])
        local = per_tok[offs[lo]:offs[hi]].sum() / 28.0
        tot = rd.all_reduce_loss(local)
        assert abs(tot.item() - per_tok.mean().item()) < 1e-12
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_two_rank_bucket_and_sharding():
    world = 2
    port = _free_port()
    ctx = mp.get_context('spawn')
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(out.keys()) == [0, 1]


def test_shard_bounds_cover():
    from re2nn_seq_b200.dist import shard_bounds
    for n in (1, 7, 8, 4096, 65537):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
