"""CPU: pin oracle/re2nn_oracle.py against every golden fixture produced by the reference."""
import numpy as np
import pytest

from conftest import golden_files
from helpers import args_of, load_golden, oracle_params, rel_err
from oracle import re2nn_oracle as orc

TOL = 1e-5   # north_star: scores/losses within 1e-5 relative in fp32


@pytest.mark.parametrize('name', golden_files('dec_') + golden_files('sf_'))
def test_decompose_forward_local(name):
    z, meta = load_golden(name)
    p, args = oracle_params(z), args_of(meta)
    dense_v = z['dense_v'] if meta['kind'] == 'sf' else None
    loss, pred, true, scores = orc.decompose_forward_local(p, z['x'], z['labels'], z['lengths'], args,
                                                           o_idx=meta['o_idx'], train=True, dense_v=dense_v)
    L = int(z['lengths'].max())
    mask = orc.length_mask(z['lengths'], L)
    assert rel_err(scores[mask], z['all_scores'][mask]) < TOL
    if meta['flags'].get('marryup_type', 'none') == 'none':      # KD / PR mixing is out-of-scope torch code
        assert rel_err(scores, z['all_scores']) < TOL            # pad positions too: the oracle computes them
        assert rel_err(loss, z['loss']) < TOL
    np.testing.assert_array_equal(pred, z['pred'])       # decoded tags: bit-exact
    np.testing.assert_array_equal(true, z['true'])


@pytest.mark.parametrize('name', golden_files('one_'))
def test_onehot(name):
    z, meta = load_golden(name)
    p, args = oracle_params(z), args_of(meta)
    loss, pred, true, scores = orc.onehot_forward_local(p, z['x'], z['labels'], z['lengths'], args,
                                                        o_idx=meta['o_idx'], train=True)
    if meta['flags']['rand_constant'] == 0:
        np.testing.assert_array_equal(scores, z['all_scores'])   # 0/1 automaton: exact path counts
    assert rel_err(scores, z['all_scores']) < TOL
    assert rel_err(loss, z['loss']) < TOL
    np.testing.assert_array_equal(pred, z['pred'])
    np.testing.assert_array_equal(true, z['true'])
    re_pred, re_scores = orc.onehot_forward_RE(p, z['x'], z['lengths'], args, o_idx=meta['o_idx'])
    np.testing.assert_array_equal(re_pred, z['re_pred'])
    assert rel_err(re_scores, z['re_scores']) < TOL


@pytest.mark.parametrize('name', golden_files('crf_'))
def test_crf(name):
    z, _ = load_golden(name)
    mask = orc.length_mask(z['lengths'], z['feats'].shape[1])
    loss = orc.crf_nll(z['feats'], mask, z['tags'], z['transitions'])
    assert rel_err(loss, z['loss']) < TOL
    path = orc.crf_viterbi(z['feats'], mask, z['transitions'])
    np.testing.assert_array_equal(path, z['path'])       # includes the reference's pad-column quirks


def test_helpers_ragged():
    a = np.arange(12).reshape(3, 4)
    lens = np.array([4, 1, 2])
    r = orc.reverse_rows(a, lens)
    np.testing.assert_array_equal(r, [[3, 2, 1, 0], [4, 5, 6, 7], [9, 8, 10, 11]])
    np.testing.assert_array_equal(orc.flatten_rows(a, lens), [0, 1, 2, 3, 4, 8, 9])
    np.testing.assert_array_equal(orc.length_mask(lens), a % 4 < lens[:, None])


@pytest.mark.parametrize('name', [n for n in golden_files('dec_') + golden_files('sf_')
                                  if 'max' not in n and '_kd' not in n and '_pr' not in n])
def test_decompose_gradients(name):
    """The autograd restatement reproduces the reference's own parameter gradients."""
    from helpers import golden_grads
    from oracle import re2nn_oracle_torch as ot
    z, meta = load_golden(name)
    p, args = oracle_params(z, np.float64), args_of(meta)
    gold = golden_grads(z)
    rename = {'embedding.weight': 'embedding', 'crf.transitions': 'crf_transitions'}
    names = [rename.get(k, k) for k in gold]
    dense_v = z['dense_v'] if meta['kind'] == 'sf' else None
    loss, g = ot.grads(p, z['x'], z['labels'], z['lengths'], args, dense_v, names=names)
    assert rel_err(loss, z['loss']) < 1e-5
    for k, ref in gold.items():
        mine = g[rename.get(k, k)]
        if np.abs(ref).max() == 0:
            assert mine is None or np.abs(mine).max() < 1e-7, k
        else:
            assert rel_err(mine, ref) < 2e-4, (k, rel_err(mine, ref))    # reference grads are fp32
