"""GPU: ragged / degenerate shapes through every arithmetic mode (tile tails, single tokens, odd dimensions)."""
import numpy as np
import pytest
import torch

from helpers import oracle_params, rel_err
from oracle import re2nn_oracle as orc

pytestmark = pytest.mark.gpu


class _Z(dict):
    files = property(lambda self: list(self.keys()))


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


CASES = [
    # V, S, R, C, D, B, Lmax, lengths override
    (11, 5, 3, 2, 4, 1, 1, None),                 # one sequence, one token
    (11, 5, 3, 2, 4, 3, 6, [1, 6, 1]),            # single-token sequences next to a full one
    (40, 17, 9, 3, 7, 130, 5, None),              # batch tile tail (128 + 2), odd S / R / C
    (40, 33, 20, 6, 5, 257, 9, None),             # two full tiles + 1 row
    (40, 64, 40, 4, 8, 64, 3, [3] * 64),          # all sequences the same length
]


@pytest.mark.parametrize('prec', ['fp32', 'tf32x3', 'fp16x3', 'bf16'])
@pytest.mark.parametrize('case', range(len(CASES)))
@pytest.mark.parametrize('farnn,crf', [(0, 1), (2, 0)])
def test_edge_shapes(case, prec, farnn, crf):
    import re2nn_seq_b200 as r
    from re2nn_seq_b200 import ops, synth
    if prec != 'fp32' and not ops.has_tcgen05():
        pytest.skip('no tcgen05')
    V, S, R, C, D, B, Lmax, lens_override = CASES[case]
    args = synth.make_args(farnn=farnn, use_crf=crf, update_nonlinear='tanh', beta=0.2)
    f = synth.make_decompose_factors(case, V, S, R, C, D, lang_frac=0.5, dtype=np.float32)
    x, lens, lab = synth.make_batch(case + 1, B, Lmax, V, C)
    if lens_override is not None:
        lens = np.asarray(lens_override, np.int64)
        pad = np.arange(Lmax)[None, :] >= lens[:, None]
        x[pad] = V
    torch.manual_seed(case)
    m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f)
    with torch.no_grad():
        if farnn:
            for n in ('Wss1', 'Wrs1', 'Wss2', 'Wrs2'):
                getattr(m, n).mul_(0.2)
            m.bs1.fill_(0.1)
            m.bs2.fill_(-0.1)
        if crf:
            m.crf.transitions.copy_(torch.from_numpy(synth.crf_transitions(case, m.C)))
    m = m.cuda()
    m.precision = prec
    with torch.no_grad():
        sc = m.forward_scores(_t(x), _t(lens)).cpu().numpy()
        loss, pred, true = m.forward_local(_t(x), _t(lab), _t(lens), train=True)
        _, pred_g, _ = m.forward_local(_t(x), _t(lab), _t(lens), train=False)          # CUDA-graph path
    z = _Z({'p.' + k: v.detach().cpu().numpy() for k, v in m.state_dict().items()})
    L = int(lens.max())
    truth, _, _ = orc.decompose_scores(oracle_params(z, np.float64), x, lens, args)
    mask = orc.length_mask(lens, L)
    tol = 3e-2 if prec == 'bf16' else 1e-5
    assert rel_err(sc[mask], truth[mask]) < tol
    assert (sc[~mask] == 0).all()                                   # rows past the length are written as zeros
    assert torch.equal(pred, pred_g)
    assert pred.shape[0] == int(lens.sum()) == true.shape[0]
    if prec != 'bf16':
        o_loss, o_pred, o_true, _ = orc.decompose_forward_local(oracle_params(z, np.float32), x, lab, lens, args, 0, True)
        np.testing.assert_array_equal(pred.cpu().numpy(), o_pred)
        np.testing.assert_array_equal(true.cpu().numpy(), o_true)
        assert rel_err(loss.item(), o_loss) < 1e-5


def test_argument_errors_are_reported_not_fatal():
    """Bad arguments come back as RuntimeError with a message (the reference raises Python exceptions too)."""
    import re2nn_seq_b200 as r
    from re2nn_seq_b200 import synth
    args = synth.make_args(farnn=0, use_crf=1, update_nonlinear='tanh')
    f = synth.make_decompose_factors(0, 11, 5, 3, 2, 4, dtype=np.float32)
    m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f).cuda()
    x, lens, lab = synth.make_batch(1, 2, 4, 11, 2)
    with pytest.raises(RuntimeError, match='float32|int64|contiguous|expected'):
        with torch.no_grad():
            m.forward_scores(_t(x).int(), _t(lens))                  # wrong index dtype
    with pytest.raises(AssertionError):
        m.crf.neg_log_likelihood_loss(torch.zeros(2, 4, 3, device='cuda'), None, _t(lab), lengths=_t(lens))   # crf.py:58
    m.precision = 'no-such-mode'
    with pytest.raises(KeyError):
        with torch.no_grad():
            m.forward_scores(_t(x), _t(lens))


def test_crf_empty_and_overlong_lengths():
    """A zero-length sequence contributes nothing to the CRF loss, gets a zero feature gradient and decodes to nothing;
    lengths past the row are clamped; neighbouring sequences are untouched (the kernels used to index
    position n - 1 = -1)."""
    import re2nn_seq_b200 as r
    from re2nn_seq_b200 import synth
    rs = np.random.RandomState(3)
    B, L, ntag = 9, 6, 7
    T = ntag + 2
    feats = rs.randn(B, L, T).astype(np.float32)
    tags = rs.randint(0, ntag, size=(B, L)).astype(np.int64)
    lens = rs.randint(1, L + 1, size=B).astype(np.int64)
    crf = r.CRF(ntag, True).cuda()
    with torch.no_grad():
        crf.transitions.copy_(_t(synth.crf_transitions(1, T)))

    def run(fe, tg, lengths):
        f = _t(fe).requires_grad_(True)
        crf.transitions.grad = None
        loss = crf.neg_log_likelihood_loss(f, None, _t(tg), lengths=_t(lengths))
        loss.backward()
        return loss.item(), f.grad.cpu().numpy(), crf.transitions.grad.cpu().numpy().copy()

    keep = np.arange(B) != 4
    l2, g2, t2 = run(feats[keep], tags[keep], lens[keep])          # the same batch without sequence 4
    lens0 = lens.copy()
    lens0[4] = 0
    l0, g0, t0 = run(feats, tags, lens0)
    assert abs(l0 - l2) <= 1e-5 * abs(l2)
    assert np.abs(g0[4]).max() == 0.0
    np.testing.assert_allclose(g0[keep], g2, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(t0, t2, rtol=1e-4, atol=1e-5)
    lens_big = lens.copy()
    lens_big[2] = L + 5                                    # clamped to the row
    lens_clamped = lens.copy()
    lens_clamped[2] = L
    la, ga, _ = run(feats, tags, lens_big)
    lb, gb, _ = run(feats, tags, lens_clamped)
    assert la == lb and np.array_equal(ga, gb)


@pytest.mark.parametrize('ntag', [3, 30, 72, 94])
def test_crf_backward_three_warps_per_sequence_matches_one(ntag):
    """CRF backward with three warps per sequence (T <= 96: one transition row per lane, named barriers) against the
    one-warp-per-sequence sweep: same feature gradients bit for bit, transition gradients up to the order of the final
    atomic adds; ragged lengths incl. one-token and empty sequences, hard-constrained (-1e4) transitions."""
    import re2nn_seq_b200 as r
    from re2nn_seq_b200 import _lib
    rs = np.random.RandomState(ntag)
    B, L, T = 61, 11, ntag + 2
    feats = (rs.randn(B, L, T) * 3.0).astype(np.float32)
    lens = rs.randint(0, L + 1, size=B).astype(np.int64)
    lens[:4] = [L, 1, 0, 2]
    tags = rs.randint(0, ntag, size=(B, L)).astype(np.int64)
    trans = rs.randn(T, T).astype(np.float32)
    trans[rs.rand(T, T) < 0.2] = -10000.0
    trans[:, T - 2] = -10000.0
    trans[T - 1, :] = -10000.0
    mask = (np.arange(L)[None, :] < lens[:, None])
    out = {}
    try:
        for split in (1, 0):
            _lib.check(_lib.fn['re2nn_debug_set_crf_backward_split'](split), 'crf split')
            crf = r.CRF(ntag, True).cuda()
            with torch.no_grad():
                crf.transitions.copy_(torch.from_numpy(trans).cuda())
            f = torch.from_numpy(feats).cuda().requires_grad_(True)
            loss = crf.neg_log_likelihood_loss(f, torch.from_numpy(mask).cuda(), torch.from_numpy(tags).cuda())
            loss.backward()
            out[split] = (loss.item(), f.grad.clone(), crf.transitions.grad.clone())
    finally:
        _lib.check(_lib.fn['re2nn_debug_set_crf_backward_split'](1), 'crf split')
    assert out[1][0] == out[0][0]
    assert torch.equal(out[1][1], out[0][1])
    scale = out[0][2].abs().max().item() + 1e-30
    assert (out[1][2] - out[0][2]).abs().max().item() / scale < 2e-6


@pytest.mark.parametrize('B,L', [(129, 7), (1024, 35), (4096, 35), (5000, 128), (16384, 255), (300, 0)])
def test_length_order_kernel_matches_torch_sort(B, L):
    """One-launch longest-first schedule (counting sort + offsets + gathers) == torch.sort(stable, descending) +
    cumsum + index_select, incl. ties everywhere, empty sequences, batches that are no multiple of the 1024-row tile."""
    from re2nn_seq_b200 import ops
    rs = np.random.RandomState(B + L)
    lens = torch.from_numpy(rs.randint(0, L + 1, size=B).astype(np.int64)).cuda()
    got = ops.length_order(lens, L)
    assert got is not None
    order, ls, offs, offs_sorted = got
    _, want_order = torch.sort(lens.to(torch.int16), descending=True, stable=True)
    want_offs = torch.cumsum(lens, 0) - lens
    assert torch.equal(order, want_order)
    assert torch.equal(offs, want_offs)
    assert torch.equal(ls, lens.index_select(0, want_order))
    assert torch.equal(offs_sorted, want_offs.index_select(0, want_order))


def test_length_order_kernel_range():
    from re2nn_seq_b200 import ops
    lens = torch.ones(20000, dtype=torch.int64, device='cuda')
    assert ops.length_order(lens, 5) is None                      # beyond the single-CTA batch limit: torch path
    assert ops.length_order(lens[:100], 300) is None              # more length bins than the kernel keeps


def test_training_decode_on_side_stream_is_the_same_decode():
    """forward_local(train=True) decodes on a side stream next to the CRF loss: same loss, predictions and gradients as
    the single-stream order."""
    from test_gpu_parity import _random_decompose
    m, args, x, lens, lab = _random_decompose(41, 300, 64, 40, 12, 30, 200, 9, farnn=0, use_crf=1, update_nonlinear='tanh',
                                              beta=0.1)
    xt, lt, yt = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (x, lens, lab))
    out = {}
    for on in (True, False):
        m.overlap_decode = on
        m.zero_grad(set_to_none=True)
        loss, pred, true = m.forward_local(xt, yt, lt, train=True)
        loss.backward()
        torch.cuda.synchronize()
        out[on] = (loss.item(), pred.clone(), true.clone(), m.S1.grad.clone())
    assert out[True][0] == out[False][0]
    for i in (1, 2, 3):
        assert torch.equal(out[True][i], out[False][i])
