"""GPU, 2 ranks over NCCL: sharded training step + one gradient all-reduce == single-GPU full batch."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(crf):
    import re2nn_seq_b200 as r
    from re2nn_seq_b200 import synth
    args = synth.make_args(farnn=2, use_crf=crf, update_nonlinear='tanh', beta=0.1)
    f = synth.make_decompose_factors(3, 300, 64, 48, 9, 16, dtype=np.float32)
    x, lens, lab = synth.make_batch(4, 37, 14, 300, 9)
    torch.manual_seed(0)
    m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f)
    with torch.no_grad():
        for n in ('Wss1', 'Wrs1', 'Wss2', 'Wrs2'):
            getattr(m, n).mul_(0.1)
        m.bs1.fill_(0.2)
        m.bs2.fill_(-0.1)
        if crf:
            m.crf.transitions.copy_(torch.from_numpy(synth.crf_transitions(0, m.C)))
    return m, x, lens, lab


def _worker(rank, world, port, crf, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from re2nn_seq_b200 import dist as rd
        m, x, lens, lab = _build(crf)
        m = m.cuda()
        xs, ls, ys = rd.shard_batch([torch.from_numpy(a).cuda() for a in (x, lens, lab)], world, rank)
        m.global_tokens = rd.global_token_count(ls)
        loss, pred, _ = m.forward_local(xs, ys, ls, train=True)
        loss.backward()
        rd.GradBucket(m).all_reduce()
        tot = rd.all_reduce_loss(loss)
        out[rank] = (tot.item(), {k: v.grad.cpu().numpy() for k, v in m.named_parameters() if v.requires_grad},
                     pred.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('crf', [1, 0])
def test_two_gpu_training_step_matches_single(crf):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, crf, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    m, x, lens, lab = _build(crf)
    m = m.cuda()
    loss, pred, _ = m.forward_local(*(torch.from_numpy(a).cuda() for a in (x, lab, lens)), train=True)
    loss.backward()
    l0, g0, p0 = out[0]
    l1, g1, p1 = out[1]
    assert abs(l0 - loss.item()) <= 1e-5 * abs(loss.item())
    np.testing.assert_array_equal(np.concatenate([p0, p1]), pred.cpu().numpy())     # shards decode the same tags
    for k, v in m.named_parameters():
        if v.requires_grad:
            ref = v.grad.cpu().numpy()
            np.testing.assert_array_equal(g0[k], g1[k])                              # ranks agree bit for bit
            scale = max(np.abs(ref).max(), 1e-30)
            assert np.abs(g0[k] - ref).max() / scale < 2e-5, k
