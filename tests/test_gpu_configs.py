"""GPU: the remaining BASELINE.json configurations as parity cases (shapes of cfg1, cfg4, cfg5 at batch sizes
the oracle finishes in seconds) plus size-independent properties at the full cfg2 batch."""
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


def _decompose(seed, V, S, R, C, D, B, Lmax, fixed_len=False, **flags):
    import re2nn_seq_b200 as r
    from re2nn_seq_b200 import synth
    args = synth.make_args(**flags)
    f = synth.make_decompose_factors(seed, V, S, R, C, D, dtype=np.float32)
    x, lens, lab = synth.make_batch(seed + 1, B, Lmax, V, C, fixed_len=fixed_len)
    torch.manual_seed(seed)
    m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f)
    with torch.no_grad():
        if args.use_crf:
            m.crf.transitions.copy_(torch.from_numpy(synth.crf_transitions(seed, m.C)))
    return m.cuda(), args, x, lens, lab


def _oracle64(m, args, x, lens):
    z = _Z({'p.' + k: v.detach().cpu().numpy() for k, v in m.state_dict().items()})
    sc, _, _ = orc.decompose_scores(oracle_params(z, np.float64), x, lens, args)
    return sc


@pytest.mark.parametrize('prec', ['fp32', 'tf32x3', 'fp16x3'])
def test_cfg4_shapes_long_recurrence(prec):
    """ATIS-ZH-shaped: S=512, R=256, len<=128 (long-recurrence stress), small batch."""
    from re2nn_seq_b200 import ops
    if prec != 'fp32' and not ops.has_tcgen05():
        pytest.skip('no tcgen05')
    m, args, x, lens, lab = _decompose(11, 2000, 512, 256, 127, 100, 40, 128, farnn=0, use_crf=1,
                                       update_nonlinear='tanh', beta=0.1)
    m.precision = prec
    with torch.no_grad():
        sc = m.forward_scores(_t(x), _t(lens)).cpu().numpy()
        _, pred, _ = m.forward_local(_t(x), _t(lab), _t(lens), train=False)
    truth = _oracle64(m, args, x, lens)
    mask = orc.length_mask(lens, 128)
    assert rel_err(sc[mask], truth[mask]) < 1e-5
    z = _Z({'p.' + k: v.detach().cpu().numpy() for k, v in m.state_dict().items()})
    p32 = oracle_params(z, np.float32)
    o_pred = orc.decode(p32, truth.astype(np.float32), lens, args, m.C, 0, True)
    assert (pred.cpu().numpy() != o_pred).sum() == 0


@pytest.mark.parametrize('prec', ['fp32', 'tf32x3', 'fp16x3', 'bf16'])
def test_cfg5_shapes(prec):
    """Scale-sweep shapes S=1024, R=512, C=128, len=64 (fixed), small batch; exercises the 256-wide n-tiles."""
    from re2nn_seq_b200 import ops
    if prec != 'fp32' and not ops.has_tcgen05():
        pytest.skip('no tcgen05')
    m, args, x, lens, lab = _decompose(12, 900, 1024, 512, 128, 100, 300, 64, fixed_len=True, farnn=0, use_crf=1,
                                       update_nonlinear='tanh', beta=0.1)
    m.precision = prec
    with torch.no_grad():
        sc = m.forward_scores(_t(x), _t(lens)).cpu().numpy()
    truth = _oracle64(m, args, x[:24], lens[:24])           # oracle on a slice: rows are independent
    err = rel_err(sc[:24], truth)
    # tensor-core fp32 accumulation truncates: at K = 1024..1536 the 3xTF32 path is good to 3e-5, not 1e-5
    assert err < {'bf16': 3e-2, 'tf32x3': 3e-5, 'fp16x3': 3e-5, 'fp32': 1e-5}[prec], err


def test_cfg1_onehot_shapes_exact():
    """ATIS-BIO-shaped exact automaton (V=900, S=300, C=127+1, len<=46, B=32): integer path counts, bit-exact."""
    import re2nn_seq_b200 as r
    from re2nn_seq_b200 import synth
    args = synth.make_args(method='onehot', rand_constant=0.0)
    a = synth.make_onehot_automaton(21, 900, 300, 127, dtype=np.float32)
    x, lens, lab = synth.make_batch(22, 32, 46, 900, 127)
    m = r.FARNN_S_O_I_S(a['language_tensor'], a['output_mat'], a['wildcard_mat'], a['output_wildcard_vector'],
                        a['final_vector'], a['start_vector'], None, args, 0, False)
    with torch.no_grad():
        loss, pred, true = m.forward_local(torch.from_numpy(x), torch.from_numpy(lab), torch.from_numpy(lens), train=False)
        sc = m.forward_score(torch.from_numpy(x), None, torch.from_numpy(lens)).numpy()
    p = {k: v for k, v in a.items() if k != 'language_rows'}
    p.update(h0=a['start_vector'], hT=a['final_vector'])
    o_sc = orc.onehot_scores(p, x, lens, args)
    assert np.isfinite(sc).all() and sc.max() < 2 ** 24
    np.testing.assert_array_equal(sc, o_sc)
    _, o_pred, o_true, _ = orc.onehot_forward_local(p, x, lab, lens, args, 0, train=False)
    np.testing.assert_array_equal(pred.numpy(), o_pred)


def test_full_cfg2_batch_properties():
    """Full B=4096 batch: (1) rows are independent -> any permutation of the batch permutes the outputs;
    (2) padding is never read -> garbage in the pad columns changes nothing; (3) a slice agrees with the oracle."""
    m, args, x, lens, lab = _decompose(13, 12000, 300, 200, 72, 100, 4096, 35, farnn=0, use_crf=1,
                                       update_nonlinear='tanh', beta=0.1)
    xt, lt, yt = _t(x), _t(lens), _t(lab)
    offs = np.concatenate([[0], np.cumsum(lens)])
    with torch.no_grad():
        _, pred, _ = m.forward_local(xt, yt, lt, train=False)
        perm = np.random.RandomState(0).permutation(4096)
        _, pred_p, _ = m.forward_local(_t(x[perm]), _t(lab[perm]), _t(lens[perm]), train=False)
        x2 = x.copy()
        pad = np.arange(35)[None, :] >= lens[:, None]
        x2[pad] = np.random.RandomState(1).randint(0, 12000, size=int(pad.sum()))
        _, pred_g, _ = m.forward_local(_t(x2), yt, lt, train=False)
    pred = pred.cpu().numpy()
    pred_p = pred_p.cpu().numpy()
    offs_p = np.concatenate([[0], np.cumsum(lens[perm])])
    for j in (0, 1, 17, 4095):
        b = perm[j]
        np.testing.assert_array_equal(pred_p[offs_p[j]:offs_p[j + 1]], pred[offs[b]:offs[b + 1]])
    np.testing.assert_array_equal(pred_g.cpu().numpy(), pred)
    z = _Z({'p.' + k: v.detach().cpu().numpy() for k, v in m.state_dict().items()})
    _, o_pred, _, _ = orc.decompose_forward_local(oracle_params(z, np.float32), x[:64], lab[:64], lens[:64], args, 0, False)
    np.testing.assert_array_equal(pred[:offs[64]], o_pred)


def test_longest_first_order_is_invisible():
    """Sorting by length (tile skipping) changes neither predictions, loss nor gradients (up to fp32 summation order)."""
    m, args, x, lens, lab = _decompose(14, 800, 96, 64, 11, 32, 300, 17, farnn=2, use_crf=1, update_nonlinear='tanh',
                                       beta=0.1)
    with torch.no_grad():
        for n in ('Wss1', 'Wrs1', 'Wss2', 'Wrs2'):
            getattr(m, n).mul_(0.1)
    res = []
    for flag in (True, False):
        m.sort_by_length = flag
        for q in m.parameters():
            q.grad = None
        loss, pred, true = m.forward_local(_t(x), _t(lab), _t(lens), train=True)
        loss.backward()
        res.append((loss.item(), pred.cpu().numpy(), true.cpu().numpy(),
                    {k: v.grad.cpu().numpy() for k, v in m.named_parameters() if v.requires_grad}))
    (l1, p1, t1, g1), (l0, p0, t0, g0) = res
    np.testing.assert_array_equal(p1, p0)
    np.testing.assert_array_equal(t1, t0)
    assert abs(l1 - l0) <= 1e-5 * abs(l0)
    for k in g0:
        assert rel_err(g1[k], g0[k]) < 2e-5, k


# ---- the exact benchmarked paths ------------------------------------------------------------------
def _reference_tags(cfg_flags, f, x, lab, lens, crf_seed):
    """Decoded tags of the UNMODIFIED reference (oracle/_ref, torch CPU fp32) for the given rows."""
    from oracle import make_ref
    why = make_ref.verify()
    if why:
        pytest.skip(why)
    make_ref.import_reference()
    from src_seq.farnn.model_decompose_single import FARNN_S_D_W_I_S as RefModel
    from re2nn_seq_b200 import synth
    torch.manual_seed(crf_seed)
    ref = RefModel(V=f['V'], S1=f['S1'], S2=f['S2'], C_output_mat=f['C_output_mat'], wildcard_mat=f['wildcard_mat'],
                   wildcard_output_vector=f['wildcard_output_vector'], final_vector=f['final_vector'],
                   start_vector=f['start_vector'], pretrained_word_embed=f['pretrained_word_embed'], priority_mat=None,
                   args=synth.make_args(**cfg_flags), o_idx=0, is_cuda=False)
    with torch.no_grad():
        ref.crf.transitions.copy_(torch.from_numpy(synth.crf_transitions(crf_seed, ref.C)))
    ref.eval()
    Lm = int(lens.max())
    with torch.no_grad():
        _, pred, _ = ref.forward_local(torch.from_numpy(x[:, :Lm].copy()), torch.from_numpy(lab[:, :Lm].copy()),
                                       torch.from_numpy(lens.copy()), train=False)
    return pred.numpy()


def test_cfg2_benchmarked_path_vs_reference():
    """cfg2 exactly as bench.py times it: B=4096, precision='auto' (fp16x3), CUDA graph replay, resident kernel,
    four chunk streams.  Decoded tags of a 128-row slice of EVERY chunk of the length-sorted batch against the
    unmodified reference classes, bit-exact."""
    import re2nn_seq_b200 as r
    from re2nn_seq_b200 import ops, synth
    if not ops.has_tcgen05():
        pytest.skip('no tcgen05')
    flags = dict(farnn=0, use_crf=1, update_nonlinear='tanh', beta=0.1)
    args = synth.make_args(**flags)
    f = synth.make_decompose_factors(0, 12000, 300, 200, 72, 100, dtype=np.float32)
    x, lens, lab = synth.make_batch(1000, 4096, 35, 12000, 72)
    torch.manual_seed(13)
    m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f)
    with torch.no_grad():
        m.crf.transitions.copy_(torch.from_numpy(synth.crf_transitions(13, m.C)))
    m = m.cuda().eval()
    m.precision = 'auto'
    m.infer_chunks = 4
    xt, lt, yt = _t(x), _t(lens), _t(lab)
    with torch.no_grad():
        for _ in range(3):                      # second sighting captures the graph, third replays it
            _, pred, _ = m.forward_local(xt, yt, lt, train=False)
        assert m._resolved_precision() == 'fp16x3' and ops.recurrence_is_resident(300, 200, 0, 'fp16x3')
        chunks = m._infer_chunks(4096)
    assert getattr(m, '_graphs', None), 'the CUDA-graph path was not taken'
    pred = pred.cpu().numpy()
    offs = np.concatenate([[0], np.cumsum(lens)])
    order = np.argsort(-lens, kind='stable')        # the module's longest-first processing order
    assert chunks is not None and len(chunks) == 4
    for b0, b1 in chunks:
        rows = order[b0:b0 + 128]
        want = _reference_tags(flags, f, x[rows], lab[rows], lens[rows], 13)
        got = np.concatenate([pred[offs[b]:offs[b + 1]] for b in rows])
        mism = int((got != want).sum())
        assert mism == 0, 'chunk [%d,%d): %d of %d tags differ from the reference' % (b0, b1, mism, len(want))


@pytest.mark.parametrize('prec', ['fp16x3', 'tf32x3'])
def test_cfg5_shapes_decoded_tags_vs_reference(prec):
    """S=1024, R=512, C=128, len=64: decoded tags (Viterbi over T=131) of the parity-grade modes against the unmodified
    reference on a 48-sequence slice, bit-exact; scores against the fp64 oracle."""
    from re2nn_seq_b200 import ops, synth
    if not ops.has_tcgen05():
        pytest.skip('no tcgen05')
    flags = dict(farnn=0, use_crf=1, update_nonlinear='tanh', beta=0.1)
    m, args, x, lens, lab = _decompose(12, 900, 1024, 512, 128, 100, 300, 64, fixed_len=True, **flags)
    m.precision = prec
    with torch.no_grad():
        _, pred, _ = m.forward_local(_t(x), _t(lab), _t(lens), train=False)
    f = synth.make_decompose_factors(12, 900, 1024, 512, 128, 100, dtype=np.float32)
    want = _reference_tags(flags, f, x[:48], lab[:48], lens[:48], 12)
    got = pred.cpu().numpy()[:48 * 64]
    assert int((got != want).sum()) == 0


def test_onehot_s1024_exact():
    """Onehot at the scale-sweep state count (S=1024, small vocabulary): integer path counts, bit-exact scores and tags."""
    import re2nn_seq_b200 as r
    from re2nn_seq_b200 import synth
    args = synth.make_args(method='onehot', rand_constant=0.0)
    a = synth.make_onehot_automaton(23, 12, 1024, 128, lang_frac=0.5, dtype=np.float32)
    x, lens, lab = synth.make_batch(24, 6, 12, 12, 128)
    m = r.FARNN_S_O_I_S(a['language_tensor'], a['output_mat'], a['wildcard_mat'], a['output_wildcard_vector'],
                        a['final_vector'], a['start_vector'], None, args, 0, False)
    with torch.no_grad():
        loss, pred, true = m.forward_local(torch.from_numpy(x), torch.from_numpy(lab), torch.from_numpy(lens), train=False)
        sc = m.forward_score(torch.from_numpy(x), None, torch.from_numpy(lens)).numpy()
    p = {k: v for k, v in a.items() if k != 'language_rows'}
    p.update(h0=a['start_vector'], hT=a['final_vector'])
    o_sc = orc.onehot_scores(p, x, lens, args)
    assert np.isfinite(sc).all() and sc.max() < 2 ** 24
    np.testing.assert_array_equal(sc, o_sc)
    _, o_pred, _, _ = orc.onehot_forward_local(p, x, lab, lens, args, 0, train=False)
    np.testing.assert_array_equal(pred.numpy(), o_pred)


def test_cfg3_full_batch_gradients_vs_fp64_oracle(capsys):
    """cfg3 as bench.py trains it: B=1024, forward GEMMs in the `auto` training precision (fp16x3), BPTT GEMMs 3xTF32.
    Every parameter gradient against the float64 autograd oracle over the SAME 1024 sequences; the measured
    per-parameter errors are printed (they are the evidence for the bound written in DESIGN.md section 2)."""
    import re2nn_seq_b200 as r
    from re2nn_seq_b200 import ops, synth
    from oracle import re2nn_oracle_torch as ot
    if not ops.has_tcgen05():
        pytest.skip('no tcgen05')
    flags = dict(farnn=0, use_crf=1, update_nonlinear='tanh', beta=0.1)
    m, args, x, lens, lab = _decompose(7, 12000, 300, 200, 72, 100, 1024, 35, **flags)
    m.train()
    m.train_precision = 'auto'
    loss, pred, _ = m.forward_local(_t(x), _t(lab), _t(lens), train=True)
    loss.backward()
    rename = {'embedding.weight': 'embedding', 'crf.transitions': 'crf_transitions',
              'priority_layer.priority_mat': 'priority_mat', 'priority_layer.priority_bias': 'priority_bias'}
    p = {rename.get(k, k): v.detach().cpu().numpy().astype(np.float64) for k, v in m.state_dict().items()}
    names = [rename.get(k, k) for k, v in m.named_parameters() if v.requires_grad]
    o_loss, o_grads = ot.grads(p, x, lab, lens, args, names=names)
    assert abs(loss.item() - o_loss) <= 1e-5 * abs(o_loss)
    errs = {}
    for k, v in m.named_parameters():
        if v.requires_grad:
            ref = o_grads[rename.get(k, k)]
            errs[k] = rel_err(v.grad.cpu().numpy(), ref)
    with capsys.disabled():
        print('\ncfg3 B=1024 gradient errors vs fp64 oracle (max-norm relative): ' +
              ', '.join('%s %.2e' % kv for kv in sorted(errs.items())))
    for k, e in errs.items():
        assert e < 2e-5, '%s: %.3e' % (k, e)         # the bound DESIGN.md section 2 states


@pytest.mark.parametrize('prec', ['bf16', 'fp16x3', 'tf32x3'])
def test_fused_label_score_operand_is_bit_identical(prec):
    """Per-step tensor-core path (S=1024: not resident): the (alpha * beta) operand written by the forward direction's
    epilogue (re2nn_decompose_recurrence ab_out, backward direction first) gives exactly the scores of the unfused
    path on every valid position, ragged lengths included."""
    from re2nn_seq_b200 import ops
    if not ops.has_tcgen05():
        pytest.skip('no tcgen05')
    m, args, x, lens, lab = _decompose(31, 900, 1024, 512, 128, 100, 200, 24, farnn=0, use_crf=1, update_nonlinear='tanh',
                                       beta=0.1)
    m.precision = prec
    with torch.no_grad():
        assert ops.recurrence_fuses(1024, 512, 0, prec)
        l0 = ops.launches()
        fused = m.forward_scores(_t(x), _t(lens), fuse=True).cpu().numpy()
        n_fused = ops.launches() - l0
        plain = m.forward_scores(_t(x), _t(lens)).cpu().numpy()
    mask = orc.length_mask(lens, int(lens.max()))
    np.testing.assert_array_equal(fused[mask], plain[mask])
    assert n_fused > 4 * int(lens.max())          # single-direction launches: 4 per step
