"""Golden fixtures of the FST (non-independent) model family from the UNMODIFIED reference classes.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_fst.py
Writes tests/golden/fst_*.npz.  Each fixture holds the constructor inputs (``in.<name>``, so a test can rebuild the
product module with the same torch seed), the reference module's post-construction state_dict (``p.<name>``), the
batch, and the reference's own outputs: all_scores, loss, flat predictions and every parameter gradient
(``g.<name>``).  Classes: FARNN_S_D_W (farnn/model_decompose.py), FARNN_S_D_W_I (model_decompose_independent.py),
FARNN_S_O and FARNN_S_O_I (model_onehot.py).
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
warnings.filterwarnings('ignore')

from re2nn_seq_b200 import synth  # noqa: E402
from src_seq.farnn.model_decompose import FARNN_S_D_W  # noqa: E402
from src_seq.farnn.model_decompose_independent import FARNN_S_D_W_I  # noqa: E402
from src_seq.farnn.model_onehot import FARNN_S_O, FARNN_S_O_I  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def _save(name, model, ins, batch, outs, meta):
    out = {'in.' + k: v for k, v in ins.items() if v is not None}
    out.update({'p.' + k: _np(v) for k, v in model.state_dict().items()})
    for k, v in model.named_parameters():
        if v.requires_grad and v.grad is not None:
            out['g.' + k] = _np(v.grad)
    out.update(batch)
    out.update(outs)
    out['meta'] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'loss', float(outs['loss']), 'pred[:8]', outs['pred'][:8], 'grads', sum(k.startswith('g.') for k in out))


def _shrink_gates(model, args):
    if args.farnn >= 1:
        with torch.no_grad():
            for n in ['Wss1', 'Wrs1'] + (['Wss2', 'Wrs2'] if args.farnn == 2 else []):
                getattr(model, n).mul_(0.3)
            model.bs1.fill_(0.2)
            if args.farnn == 2:
                model.bs2.fill_(-0.1)


def _spy_decode(model, captured):
    orig = model.decode

    def spy(all_scores, flat, mask, lens):
        captured['all_scores'] = all_scores
        return orig(all_scores, flat, mask, lens)
    model.decode = spy


def decompose_fst_inputs(seed, V, S, R, RW, C, D):
    rs = np.random.RandomState(seed)
    sc = 1.0 / np.sqrt(R)
    ins = dict(V=rs.randn(V + 1, R) * sc, C=rs.rand(C, R) * 0.6 + 0.1, S1=rs.randn(S, R) * sc, S2=rs.randn(S, R) * sc,
               C_wildcard=rs.rand(C, RW) * 0.5, S1_wildcard=rs.randn(S, RW) * 0.4, S2_wildcard=rs.randn(S, RW) * 0.4,
               wildcard_wildcard=(rs.rand(S, S) < 0.15).astype(np.float64) * 0.5,
               final_vector=(rs.rand(S) < 0.4).astype(np.float64), start_vector=(rs.rand(S) < 0.4).astype(np.float64),
               pretrained_word_embed=rs.randn(V + 1, D))
    ins['V'][V] = 0.0
    ins['pretrained_word_embed'][V] = 0.0
    ins['start_vector'][0] = 1.0
    ins['final_vector'][1] = 1.0
    return ins


def case_dw(name, seed, dims, flags, backward=True):
    V, S, R, RW, C, D, B, Lmax = dims
    args = synth.make_args(independent=0, **flags)
    ins = decompose_fst_inputs(seed, V, S, R, RW, C, D)
    x, lengths, labels = synth.make_batch(seed + 1, B, Lmax, V, C)
    torch.manual_seed(seed)
    m = FARNN_S_D_W(priority_mat=None, args=args, o_idx=1, **ins)
    m.is_cuda = False
    if m.use_crf:
        m.crf.gpu = False
    _shrink_gates(m, args)
    if m.use_crf:
        with torch.no_grad():
            m.crf.transitions.add_(torch.from_numpy(synth.crf_transitions(seed, m.C, noise=0.3)) * 0 +
                                   torch.randn(m.C, m.C, generator=torch.Generator().manual_seed(seed)) * 0.3)
    cap = {}
    _spy_decode(m, cap)
    loss, pred, true = m.forward_local(torch.from_numpy(x), torch.from_numpy(labels), torch.from_numpy(lengths), train=True)
    if backward:
        loss.backward()
    _save(name, m, ins, dict(x=x, lengths=lengths, labels=labels),
          dict(all_scores=_np(cap['all_scores']), loss=_np(loss), pred=_np(pred), true=_np(true)),
          dict(kind='fst_dw', flags=flags, o_idx=1, dims=dict(V=V, S=S, R=R, RW=RW, C=C, D=D, B=B, Lmax=Lmax), seed=seed))


def case_dwi(name, seed, dims, flags, backward=True):
    V, S, R, RO, C, D, B, Lmax = dims
    args = synth.make_args(independent=1, **flags)
    rs = np.random.RandomState(seed)
    sc = 1.0 / np.sqrt(R)
    ins = dict(V=rs.randn(V + 1, R) * sc, S1=rs.randn(S, R) * sc, S2=rs.randn(S, R) * sc, C_output=rs.rand(C, RO) * 0.5 + 0.1,
               S1_output=rs.rand(S, RO) * 0.8, S2_output=rs.rand(S, RO) * 0.8,
               wildcard_mat=(rs.rand(S, S) < 0.15).astype(np.float64) * 0.5, wildcard_output=rs.rand(S, S) * 0.2,
               final_vector=(rs.rand(S) < 0.4).astype(np.float64), start_vector=(rs.rand(S) < 0.4).astype(np.float64),
               pretrained_word_embed=rs.randn(V + 1, D))
    if flags.get('train_mode') == 'max':       # keep the max-product states away from the all-zero fixed point
        for k in ('V', 'S1', 'S2'):
            ins[k] = np.abs(ins[k]) * 1.5
    ins['V'][V] = 0.0
    ins['pretrained_word_embed'][V] = 0.0
    ins['start_vector'][0] = 1.0
    ins['final_vector'][1] = 1.0
    x, lengths, labels = synth.make_batch(seed + 1, B, Lmax, V, C)
    torch.manual_seed(seed)
    m = FARNN_S_D_W_I(priority_mat=None, args=args, o_idx=1, **ins)
    m.is_cuda = False
    if m.use_crf:
        m.crf.gpu = False
        with torch.no_grad():
            m.crf.transitions.add_(torch.randn(m.C, m.C, generator=torch.Generator().manual_seed(seed)) * 0.3)
    _shrink_gates(m, args)
    cap = {}
    _spy_decode(m, cap)
    loss, pred, true = m.forward_local(torch.from_numpy(x), torch.from_numpy(labels), torch.from_numpy(lengths), train=True)
    if backward:
        loss.backward()
    _save(name, m, ins, dict(x=x, lengths=lengths, labels=labels),
          dict(all_scores=_np(cap['all_scores']), loss=_np(loss), pred=_np(pred), true=_np(true)),
          dict(kind='fst_dwi', flags=flags, o_idx=1, dims=dict(V=V, S=S, R=R, RO=RO, C=C, D=D, B=B, Lmax=Lmax), seed=seed))


def case_o(name, seed, dims, flags, noise, independent=None, priority=False):
    V, S, C, B, Lmax = dims
    flags = dict(flags, rand_constant=noise, method='onehot')
    rs = np.random.RandomState(seed)
    x, lengths, labels = synth.make_batch(seed + 1, B, Lmax, V, C)
    pm = np.eye(C + 1) + 0.1 * rs.randn(C + 1, C + 1) if priority else None
    start = (rs.rand(S) < 0.4).astype(np.float64)
    final = (rs.rand(S) < 0.4).astype(np.float64)
    start[0] = 1.0
    final[1] = 1.0
    if independent is None:
        args = synth.make_args(independent=0, **flags)
        ins = dict(language_tensor=(rs.rand(V + 1, C + 1, S, S) < 0.12).astype(np.float64),
                   wildcard_tensor=(rs.rand(C + 1, S, S) < 0.08).astype(np.float64),
                   wildcard_wildcard_mat=(rs.rand(S, S) < 0.1).astype(np.float64), final_vector=final, start_vector=start)
        ins['language_tensor'][V] = 0.0
        torch.manual_seed(seed)
        m = FARNN_S_O(priority_mat=pm, args=args, o_idx=1, is_cuda=False, **ins)
        kind = 'fst_o'
    else:
        args = synth.make_args(independent=independent, **flags)
        ins = dict(language_tensor=(rs.rand(V + 1, S, S) < 0.3).astype(np.float64),
                   output_tensor=(rs.rand(C + 1, S, S) < 0.3).astype(np.float64),
                   wildcard_mat=(rs.rand(S, S) < 0.15).astype(np.float64),
                   output_wildcard_mat=(rs.rand(S, S) < 0.2).astype(np.float64), final_vector=final, start_vector=start)
        ins['language_tensor'][V] = 0.0
        torch.manual_seed(seed)
        m = FARNN_S_O_I(priority_mat=pm, args=args, o_idx=1, is_cuda=False, **ins)
        kind = 'fst_oi'
    xt, lt, yt = torch.from_numpy(x), torch.from_numpy(lengths), torch.from_numpy(labels)
    all_scores = m.forward_score(xt, yt, lt, train=True)
    loss, pred, true = m.forward_local(xt, yt, lt, train=True)
    loss.backward()
    with torch.no_grad():
        re_pred, re_scores = m.forward_RE(xt, yt, lt, train=False)
    ins['priority_mat'] = pm
    _save(name, m, ins, dict(x=x, lengths=lengths, labels=labels),
          dict(all_scores=_np(all_scores), loss=_np(loss), pred=_np(pred), true=_np(true), re_pred=_np(re_pred),
               re_scores=_np(re_scores)),
          dict(kind=kind, flags=flags, o_idx=1, dims=dict(V=V, S=S, C=C, B=B, Lmax=Lmax), seed=seed, priority=bool(priority)))


def main():
    tr_all = dict(train_beta=1, train_h0=1, train_hT=1, train_V_embed=1, train_wildcard=1, train_wildcard_wildcard=1,
                  train_word_embed=1)
    d = (20, 9, 6, 4, 4, 5, 6, 7)      # V, S, R, R_W / R_O, C, D, B, Lmax
    case_dw('fst_dw_f0_tanh_crf', 50, d, dict(farnn=0, update_nonlinear='tanh', use_crf=1, beta=0.3, **tr_all))
    case_dw('fst_dw_f2_tanh_ce', 51, d, dict(farnn=2, update_nonlinear='tanh', use_crf=0, beta=0.5, additional_nonlinear='tanh',
                                             **tr_all))
    case_dw('fst_dw_f1_relu_crf_add', 52, d, dict(farnn=1, update_nonlinear='relu', use_crf=1, beta=0.2, additional_states=2,
                                                  rand_constant=1e-2))
    case_dw('fst_dw_f0_tanh_max', 53, d, dict(farnn=0, update_nonlinear='tanh', use_crf=0, beta=0.4, train_mode='max'),
            backward=False)
    case_dwi('fst_dwi_f0_tanh_crf', 60, d, dict(farnn=0, update_nonlinear='tanh', use_crf=1, beta=0.3, **tr_all))
    case_dwi('fst_dwi_f2_tanh_ce', 61, d, dict(farnn=2, update_nonlinear='tanh', use_crf=0, beta=0.5, local_loss_func='CE',
                                               **tr_all))
    case_dwi('fst_dwi_f1_relu_max', 62, d, dict(farnn=1, update_nonlinear='relu', use_crf=0, beta=0.2, train_mode='max'))
    od = (12, 7, 3, 5, 6)              # V, S, C, B, Lmax
    case_o('fst_o_sum', 70, od, dict(), 1e-2)
    case_o('fst_o_max_ce_prio', 71, od, dict(train_mode='max', local_loss_func='CE', use_priority=1, train_wildcard=1,
                                             train_wildcard_wildcard=1), 1e-2, priority=True)
    case_o('fst_oi_ind2', 72, od, dict(), 1e-2, independent=2)
    case_o('fst_oi_ind1_ce_max', 73, od, dict(train_mode='max', local_loss_func='CE'), 1e-2, independent=1)
    case_o('fst_oi_ind2_exact', 74, od, dict(), 0.0, independent=2)


if __name__ == '__main__':
    main()
