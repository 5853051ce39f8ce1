"""Generate golden fixtures from the UNMODIFIED reference classes.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz.  Each fixture holds, for one small seeded case: the reference
module's post-construction state_dict (``p.<name>``), the inputs, and the reference's own
outputs (all_scores, loss, flat predictions, Viterbi paths, parameter gradients ``g.<name>``).
The fixtures pin oracle/re2nn_oracle.py (tests/test_oracle_golden.py) and the CUDA path
(tests/test_gpu_*.py); /root/reference is never needed at test time.
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
from src_seq.baselines.crf import CRF  # noqa: E402
from src_seq.farnn.model_decompose_single import FARNN_S_D_W_I_S, FARNN_S_SF  # noqa: E402
from src_seq.farnn.model_onehot import FARNN_S_O_I_S  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def _state(model):
    return {'p.' + k: _np(v) for k, v in model.state_dict().items()}


def _grads(model):
    out = {}
    for k, v in model.named_parameters():
        if v.requires_grad and v.grad is not None:
            out['g.' + k] = _np(v.grad)
    return out


def _randomise_trainables(model, seed, names):
    """Give gates / CRF transitions non-trivial values so parity is not checked at a fixed point."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, scale in names:
            t = dict(model.named_parameters())[n]
            t.add_(torch.randn(t.shape, generator=g) * scale)


def decompose_case(name, seed, dims, flags, sf=False, priority=False, marry=False):
    V, S, R, C, D, B, Lmax = dims
    args = synth.make_args(**flags)
    f = synth.make_decompose_factors(seed, V, S, R, C, D, lang_frac=0.5)
    x, lengths, labels = synth.make_batch(seed + 1, B, Lmax, V, C)
    pm = None
    if priority:
        rs = np.random.RandomState(seed + 2)
        pm = np.eye(C + 1) + 0.1 * rs.randn(C + 1, C + 1)
    torch.manual_seed(seed)
    if sf:
        model = FARNN_S_SF(S1=f['S1'], S2=f['S2'], C_output_mat=f['C_output_mat'],
                           wildcard_mat=f['wildcard_mat'], wildcard_output_vector=f['wildcard_output_vector'],
                           final_vector=f['final_vector'], start_vector=f['start_vector'],
                           priority_mat=pm, args=args, o_idx=1, is_cuda=False)
    else:
        model = FARNN_S_D_W_I_S(V=f['V'], S1=f['S1'], S2=f['S2'], C_output_mat=f['C_output_mat'],
                                wildcard_mat=f['wildcard_mat'], wildcard_output_vector=f['wildcard_output_vector'],
                                final_vector=f['final_vector'], start_vector=f['start_vector'],
                                pretrained_word_embed=f['pretrained_word_embed'], priority_mat=pm,
                                args=args, o_idx=1, is_cuda=False)
    rnd = []
    if args.farnn >= 1:
        # reference init is randn (std 1) which saturates the gates; shrink to keep them informative
        with torch.no_grad():
            for n in ['Wss1', 'Wrs1'] + (['Wss2', 'Wrs2'] if args.farnn == 2 else []):
                getattr(model, n).mul_(0.3)
            model.bs1.fill_(0.2)
            if args.farnn == 2:
                model.bs2.fill_(-0.1)
    if args.use_crf:
        rnd.append(('crf.transitions', 0.3))
    if flags.get('train_beta'):
        rnd.append(('beta_vec', 0.2))
    _randomise_trainables(model, seed + 3, rnd)
    # the wildcard_output_vector is zero in synthetic data; make the CE branch see it
    if args.local_loss_func != 'CE1':
        with torch.no_grad():
            model.wildcard_output_vector.add_(0.25)

    xt, lt, yt = torch.from_numpy(x), torch.from_numpy(lengths), torch.from_numpy(labels)
    captured = {}
    orig_decode = model.decode

    def spy(all_scores, flat, mask, lens):
        captured['all_scores'] = all_scores
        return orig_decode(all_scores, flat, mask, lens)
    model.decode = spy
    out = {}
    if sf:
        rs = np.random.RandomState(seed + 4)
        vecs = (rs.randn(B, Lmax, R) / np.sqrt(R)).astype(np.float32)
        out['dense_v'] = vecs
        loss, pred, true = model(torch.from_numpy(vecs), yt, lt, True)
    elif marry:
        rs = np.random.RandomState(seed + 5)
        re_tags = rs.rand(B, Lmax, C).astype(np.float32)              # teacher scores, B x L x C (RE.py output)
        out['re_tags'] = re_tags
        loss, pred, true = model.forward_local(xt, yt, lt, train=True, re_tags=torch.from_numpy(re_tags))
    else:
        loss, pred, true = model.forward_local(xt, yt, lt, train=True)
    loss.backward()
    out.update(_state(model))
    out.update(_grads(model))
    out.update(x=x, lengths=lengths, labels=labels, all_scores=_np(captured['all_scores']),
               loss=_np(loss), pred=_np(pred), true=_np(true))
    out['meta'] = np.array(json.dumps(dict(kind='sf' if sf else 'decompose', flags=flags, o_idx=1,
                                           dims=dict(V=V, S=S, R=R, C=C, D=D, B=B, Lmax=Lmax))))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'loss', float(loss), 'pred[:8]', _np(pred)[:8])


def onehot_case(name, seed, dims, flags, noise=0.0, priority=False):
    V, S, C, B, Lmax = dims
    flags = dict(flags, rand_constant=noise, method='onehot')
    args = synth.make_args(**flags)
    a = synth.make_onehot_automaton(seed, V, S, C, lang_frac=0.6)
    x, lengths, labels = synth.make_batch(seed + 1, B, Lmax, V, C)
    pm = None
    if priority:
        rs = np.random.RandomState(seed + 2)
        pm = np.eye(C + 1) + 0.1 * rs.randn(C + 1, C + 1)
    torch.manual_seed(seed)
    model = FARNN_S_O_I_S(a['language_tensor'], a['output_mat'], a['wildcard_mat'], a['output_wildcard_vector'],
                          a['final_vector'], a['start_vector'], pm, args, 1, False)
    if args.local_loss_func != 'CE1':
        with torch.no_grad():
            model.output_wildcard_vector.add_(0.25)
    xt, lt, yt = torch.from_numpy(x), torch.from_numpy(lengths), torch.from_numpy(labels)
    all_scores = model.forward_score(xt, yt, lt, train=True)
    loss, pred, true = model.forward_local(xt, yt, lt, train=True)
    loss.backward()
    with torch.no_grad():
        re_pred, re_scores = model.forward_RE(xt, yt, lt, train=False)
    out = {}
    out.update(_state(model))
    out.update(_grads(model))
    out.update(x=x, lengths=lengths, labels=labels, all_scores=_np(all_scores), loss=_np(loss),
               pred=_np(pred), true=_np(true), re_pred=_np(re_pred), re_scores=_np(re_scores))
    out['meta'] = np.array(json.dumps(dict(kind='onehot', flags=flags, o_idx=1,
                                           dims=dict(V=V, S=S, C=C, B=B, Lmax=Lmax))))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'loss', float(loss), 'pred[:8]', _np(pred)[:8])


def crf_case(name, seed, tagset, B, L, integer_feats=False):
    rs = np.random.RandomState(seed)
    T = tagset + 2
    crf = CRF(tagset, False)
    with torch.no_grad():
        crf.transitions.copy_(torch.from_numpy(synth.crf_transitions(seed, T, noise=0.0 if integer_feats else 0.5)))
    if integer_feats:
        feats = rs.randint(0, 3, size=(B, L, T)).astype(np.float32)   # many exact ties
    else:
        feats = rs.randn(B, L, T).astype(np.float32)
    lengths = rs.randint(1, L + 1, size=(B,)).astype(np.int64)
    lengths[0] = L
    if B > 1:
        lengths[1] = 1
    tags = rs.randint(0, tagset, size=(B, L)).astype(np.int64)
    mask = np.arange(L)[None, :] < lengths[:, None]
    ft = torch.from_numpy(feats).requires_grad_(True)
    loss = crf.neg_log_likelihood_loss(ft, torch.from_numpy(mask), torch.from_numpy(tags))
    loss.backward()
    with torch.no_grad():
        _, path = crf._viterbi_decode(torch.from_numpy(feats), torch.from_numpy(mask))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), feats=feats, lengths=lengths, tags=tags,
                        transitions=_np(crf.transitions), loss=_np(loss), path=_np(path),
                        g_feats=_np(ft.grad), g_transitions=_np(crf.transitions.grad),
                        meta=np.array(json.dumps(dict(kind='crf', tagset=tagset))))
    print(name, 'loss', float(loss))


def main():
    dims = (30, 12, 8, 5, 6, 7, 9)     # V, S, R, C, D, B, Lmax
    tr_all = dict(train_beta=1, train_h0=1, train_hT=1, train_V_embed=1, train_wildcard=1,
                  train_wildcard_wildcard=1, train_word_embed=1)
    decompose_case('dec_f0_tanh_crf', 10, dims, dict(farnn=0, update_nonlinear='tanh', use_crf=1, beta=0.1))
    decompose_case('dec_f0_none_ce', 11, dims, dict(farnn=0, update_nonlinear='none', use_crf=0, beta=0.3))
    decompose_case('dec_f1_relu_ce', 12, dims, dict(farnn=1, update_nonlinear='relu', use_crf=0, beta=0.5))
    decompose_case('dec_f2_tanh_crf_add', 13, dims, dict(farnn=2, update_nonlinear='tanh', use_crf=1, beta=0.1,
                                                       additional_states=3, rand_constant=1e-2))
    decompose_case('dec_f2_relutanh_ce_all', 14, dims, dict(farnn=2, update_nonlinear='relutanh', use_crf=0, beta=0.4,
                                                          additional_nonlinear='tanh', **tr_all))
    decompose_case('dec_f0_tanh_max', 15, dims, dict(farnn=0, update_nonlinear='tanh', use_crf=0, beta=0.2,
                                                   train_mode='max'))
    decompose_case('dec_f0_prio_crf_sig', 16, dims, dict(farnn=0, update_nonlinear='tanh', use_crf=1, beta=0.1,
                                                       use_priority=1, additional_nonlinear='sigmoid'),
                   priority=True)
    decompose_case('dec_f1_tanh_ce_plain', 17, dims, dict(farnn=1, update_nonlinear='tanh', use_crf=0, beta=0.6,
                                                        local_loss_func='CE', additional_nonlinear='relu', **tr_all))
    decompose_case('dec_f2_none_crf_relutanh', 18, dims, dict(farnn=2, update_nonlinear='none', use_crf=1, beta=0.1,
                                                            additional_nonlinear='relutanh', train_c_output=1))
    decompose_case('dec_f0_tanh_ce_kd', 21, dims, dict(farnn=0, update_nonlinear='tanh', use_crf=0, beta=0.3,
                                                     marryup_type='kd', c1_kdpr=2.0, c2_kdpr=0.4), marry=True)
    decompose_case('dec_f2_tanh_crf_pr', 22, dims, dict(farnn=2, update_nonlinear='tanh', use_crf=1, beta=0.2,
                                                      marryup_type='pr', c1_kdpr=1.5, c2_kdpr=0.3, c3_pr=0.9), marry=True)
    decompose_case('dec_f0_tanh_ml', 23, dims, dict(farnn=0, update_nonlinear='tanh', use_crf=0, beta=0.3,
                                                  local_loss_func='ML', margin=0.3))
    decompose_case('sf_f2_tanh_crf', 19, dims, dict(farnn=2, update_nonlinear='tanh', use_crf=1), sf=True)
    decompose_case('sf_f0_relu_ce', 20, dims, dict(farnn=0, update_nonlinear='relu', use_crf=0), sf=True)
    odims = (25, 11, 4, 6, 8)          # V, S, C, B, Lmax
    onehot_case('one_sum_none', 30, odims, dict(update_nonlinear='none'))
    onehot_case('one_sum_tanh_noise', 31, odims, dict(update_nonlinear='tanh'), noise=1e-2)
    onehot_case('one_max_relu', 32, odims, dict(update_nonlinear='relu', train_mode='max'), noise=1e-2)
    onehot_case('one_sum_relutanh_ce_prio', 33, odims, dict(update_nonlinear='relutanh', local_loss_func='CE',
                                                           use_priority=1), noise=1e-2, priority=True)
    onehot_case('one_sum_relu_ml', 34, odims, dict(update_nonlinear='relu', local_loss_func='ML', margin=0.5), noise=1e-2)
    crf_case('crf_rand', 40, 6, 9, 11)
    crf_case('crf_ties', 41, 5, 8, 7, integer_feats=True)
    crf_case('crf_single', 42, 3, 1, 1)
    crf_case('crf_wide', 43, 75, 5, 12)


if __name__ == '__main__':
    main()
