"""GPU: the FST (non-independent) model family against fixtures generated from the unmodified reference classes
(tests/golden/make_golden_fst.py): scores / loss 1e-5 relative, decoded tags bit-exact, every parameter gradient."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import golden_files
from helpers import GOLDEN, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5
GRAD_TOL = 2e-4


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    return z, json.loads(str(z['meta']))


def build_fst_module(z, meta, load_state=True):
    """Construct the product module exactly as make_golden_fst.py constructed the reference one."""
    import re2nn_seq_b200 as r
    from re2nn_seq_b200 import synth
    ins = {k[3:]: z[k] for k in z.files if k.startswith('in.')}
    kind = meta['kind']
    flags = dict(meta['flags'])
    indep = {'fst_dw': 0, 'fst_dwi': 1, 'fst_o': 0}.get(kind, flags.pop('independent', None))
    if kind == 'fst_oi':
        indep = 2 if 'ind2' in meta.get('name', '') else None
    torch.manual_seed(meta['seed'])
    if kind == 'fst_dw':
        m = r.FARNN_S_D_W(priority_mat=None, args=synth.make_args(independent=0, **flags), o_idx=meta['o_idx'], **ins)
    elif kind == 'fst_dwi':
        m = r.FARNN_S_D_W_I(priority_mat=None, args=synth.make_args(independent=1, **flags), o_idx=meta['o_idx'], **ins)
    elif kind == 'fst_o':
        pm = ins.pop('priority_mat', None)
        m = r.FARNN_S_O(priority_mat=pm, args=synth.make_args(independent=0, **flags), o_idx=meta['o_idx'], is_cuda=False, **ins)
    else:
        pm = ins.pop('priority_mat', None)
        m = r.FARNN_S_O_I(priority_mat=pm, args=synth.make_args(independent=meta['independent'], **flags),
                          o_idx=meta['o_idx'], is_cuda=False, **ins)
    if load_state:
        m.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('p.')}, strict=True)
    return m


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize('name', golden_files('fst_'))
def test_fst_golden(name):
    z, meta = _load(name)
    meta['name'] = name
    if meta['kind'] == 'fst_oi':
        meta['independent'] = 2 if 'ind2' in name else 1
    m = build_fst_module(z, meta).cuda()
    x, lab, lens = _t(z['x']), _t(z['labels']), _t(z['lengths'])
    has_grads = any(k.startswith('g.') for k in z.files)
    L = int(z['lengths'].max())
    if meta['kind'] in ('fst_o', 'fst_oi'):
        with torch.no_grad():
            sc = m.forward_score(x, lab, lens).cpu().numpy()
            re_pred, re_scores = m.forward_RE(x, lab, lens)
        np.testing.assert_array_equal(re_pred.cpu().numpy(), z['re_pred'])
        assert rel_err(re_scores.cpu().numpy(), z['re_scores']) < TOL
    else:
        with torch.no_grad():
            sc = m.forward_scores(x, lens).cpu().numpy()
    want = z['all_scores']
    valid = np.arange(want.shape[1])[None, :] < z['lengths'][:, None]
    if meta['kind'] in ('fst_o', 'fst_oi'):
        assert rel_err(sc, want) < TOL                                   # the onehot classes score every position
    else:
        assert rel_err(sc[:, :L][valid[:, :L]], want[valid]) < TOL
    if has_grads:
        loss, pred, true = m.forward_local(x, lab, lens, train=True)
        loss.backward()
    else:
        with torch.no_grad():
            loss, pred, true = m.forward_local(x, lab, lens, train=True)
    assert rel_err(loss.item(), float(z['loss'])) < TOL
    np.testing.assert_array_equal(pred.cpu().numpy(), z['pred'])
    np.testing.assert_array_equal(true.cpu().numpy(), z['true'])
    if has_grads:
        params = dict(m.named_parameters())
        for k in z.files:
            if k.startswith('g.'):
                g = params[k[2:]].grad
                assert g is not None, k
                assert rel_err(g.cpu().numpy(), z[k]) < GRAD_TOL, k
