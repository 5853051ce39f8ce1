"""CPU oracle for the RE2NN-SEQ transducer hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's PyTorch algorithm for the path named in
BASELINE.json (the FA-RNN i-FST recurrences, per-position label scoring, CRF forward /
Viterbi and their gradients).  It is the *checker*: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  The product
(re2nn_seq_b200/) never imports anything under oracle/ and has no CPU fallback.

Parity pinning: the reference has no tests or golden vectors of its own.
The oracle is pinned against outputs of the reference's own classes, imported unmodified
from /root/reference in the build container by tests/golden/make_golden.py; the resulting
fixtures are committed under tests/golden/*.npz and tests/test_oracle_golden.py checks the
oracle against every one of them.

Each function cites the reference file:line it restates (paths relative to
/root/reference/src_seq/).  All arithmetic is done in the dtype of the parameter arrays
(float32 to mirror the reference; float64 can be passed in to get a higher-precision truth).
"""
import numpy as np

START_TAG = -2   # baselines/crf.py:11
STOP_TAG = -1    # baselines/crf.py:12


# --------------------------------------------------------------------------------------
# small helpers: utils.py:133-199
# --------------------------------------------------------------------------------------
def length_mask(lengths, max_len=None):
    """utils.py:133-144 get_length_mask."""
    max_len = int(max_len or lengths.max())
    return np.arange(max_len)[None, :] < lengths[:, None]


def reverse_rows(a, lengths):
    """utils.py:183-189 reverse: flip the first lengths[b] entries of every row."""
    out = a.copy()
    for b in range(a.shape[0]):
        n = int(lengths[b])
        out[b, :n] = a[b, :n][::-1]
    return out


def flatten_rows(a, lengths):
    """utils.py:153-164 flatten: concatenate the valid prefixes, batch-major."""
    return np.concatenate([a[b, :int(lengths[b])] for b in range(a.shape[0])], axis=0)


def semiring_sum(h, tr):
    """utils.py:198-199 _matmul: einsum('bs,bsj->bj')."""
    return np.einsum('bs,bsj->bj', h, tr)


def semiring_max(h, tr):
    """utils.py:192-195 _maxmul: max over s of h[b,s]*tr[b,s,j]."""
    return (h[:, :, None] * tr).max(axis=1)


def nonlinear(x, kind):
    """update_nonlinear / additional_nonlinear switch (model_decompose_single.py:184-191,
    model_decompose.py:228-237)."""
    if kind == 'relu':
        return np.maximum(x, 0)
    if kind == 'tanh':
        return np.tanh(x)
    if kind == 'relutanh':
        return np.tanh(np.maximum(x, 0))
    if kind == 'sigmoid':
        return _sigmoid(x)
    return x


def _sigmoid(x):
    with np.errstate(over='ignore'):
        return (1.0 / (1.0 + np.exp(-x))).astype(x.dtype)


def priority(scores, p):
    """farnn/priority.py:20-30: scores @ priority_mat + priority_bias."""
    return scores @ p['priority_mat'] + p['priority_bias']


# --------------------------------------------------------------------------------------
# decompose i-FST: farnn/model_decompose_single.py (FARNN_S_D_W_I_S, FARNN_S_SF)
# --------------------------------------------------------------------------------------
def token_factor(p, ids, args):
    """model_decompose.py:222-241 get_generalized_v_embed_vec applied to V_embed[ids]
    (model_decompose_single.py:140-141)."""
    v_vec = p['V_embed'][ids]
    gen = p['embedding'][ids] @ p['embed_r_generalized']
    gen = nonlinear(gen, args.additional_nonlinear)
    return v_vec * p['beta_vec'] + gen * (1 - p['beta_vec'])


def output_vector_sum(p, args):
    """model_decompose_single.py:231-234."""
    o = p['C_output_mat'].sum(0)
    if args.local_loss_func != 'CE1':
        o = o + p['wildcard_output_vector']
    return o


def decompose_step(p, h, v, h_init, o, args, is_forward):
    """model_decompose_single.py:138-200 get_forward_score (also FARNN_S_SF :417-481)."""
    k = args.sigmoid_exponent
    if args.farnn == 0:
        hbar = h
    elif args.farnn == 1:
        hbar = h
        zt = _sigmoid((h @ p['Wss1'] + v @ p['Wrs1'] + p['bs1']) * k)
    elif args.farnn == 2:
        zt = _sigmoid((h @ p['Wss1'] + v @ p['Wrs1'] + p['bs1']) * k)
        rt = _sigmoid((h @ p['Wss2'] + v @ p['Wrs2'] + p['bs2']) * k)
        hbar = (1 - rt) * h_init + rt * h
    else:
        raise NotImplementedError()
    if not is_forward:
        hbar = hbar * o
    if args.train_mode == 'max':
        temp = np.einsum('br,sr->bsr', v, p['S1'])
        tr = np.einsum('sr,bjr->bjs', p['S2'], temp) + p['wildcard_mat']
        if is_forward:
            nxt = semiring_max(hbar, tr)
        else:
            nxt = semiring_max(hbar, tr.transpose(0, 2, 1))
    else:
        if is_forward:
            nxt = ((hbar @ p['S1']) * v) @ p['S2'].T + hbar @ p['wildcard_mat']
        else:
            nxt = ((hbar @ p['S2']) * v) @ p['S1'].T + hbar @ p['wildcard_mat'].T
    if is_forward:
        nxt = nxt * o
    nxt = nonlinear(nxt, args.update_nonlinear)
    if args.farnn == 0:
        return nxt
    return (1 - zt) * h + zt * nxt


def decompose_scores(p, x, lengths, args, dense_v=None):
    """model_decompose_single.py:220-272 (token input) / :494-545 (FARNN_S_SF, dense_v B x L x R).

    Returns (all_scores B x L x C', alpha B x (L+1) x S, beta B x (L+1) x S) where alpha/beta are
    h0_forward_score and reversed_backward_score_x of the reference (:257-261)."""
    B = x.shape[0] if dense_v is None else dense_v.shape[0]
    L = int(lengths.max())
    dt = p['S1'].dtype
    S = p['S1'].shape[0]
    if dense_v is None:
        bx = reverse_rows(x, lengths)
    else:
        bv = reverse_rows(dense_v, lengths)
    h0 = np.repeat(p['h0'][None, :], B, 0)
    hT = np.repeat(p['hT'][None, :], B, 0)
    hf, hb = h0.copy(), hT.copy()
    fwd = np.zeros((B, L, S), dt)
    bwd = np.zeros((B, L, S), dt)
    o = output_vector_sum(p, args)
    for i in range(L):
        vf = token_factor(p, x[:, i], args) if dense_v is None else dense_v[:, i]
        hf = decompose_step(p, hf, vf, h0, o, args, True)
        fwd[:, i] = hf
        vb = token_factor(p, bx[:, i], args) if dense_v is None else bv[:, i]
        hb = decompose_step(p, hb, vb, hT, o, args, False)
        bwd[:, i] = hb
    alpha = np.concatenate([h0[:, None], fwd], 1)
    beta = reverse_rows(np.concatenate([hT[:, None], bwd], 1), lengths + 1)
    ab = alpha[:, 1:L + 1] * beta[:, 1:L + 1]
    all_scores = np.einsum('bls,cs->blc', ab, p['C_output_mat'])
    if args.use_priority:
        all_scores = priority(all_scores, p)
    return all_scores.astype(dt), alpha, beta


def ce_loss(flat_scores, flat_labels):
    """nn.CrossEntropyLoss() default = mean over tokens (model_decompose.py:80)."""
    m = flat_scores.max(1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(flat_scores - m).sum(1))
    picked = flat_scores[np.arange(len(flat_labels)), flat_labels]
    return (lse - picked).mean(dtype=flat_scores.dtype)


def ml_loss(flat_scores, flat_labels, margin):
    """nn.MultiMarginLoss(margin) (p = 1, mean): sum_{j != y} max(0, margin - x_y + x_j) / C per token
    (local_loss_func == 'ML': model_decompose.py:84-85, model_onehot.py:61-62)."""
    n, C = flat_scores.shape
    picked = flat_scores[np.arange(n), flat_labels][:, None]
    h = np.maximum(0, flat_scores.dtype.type(margin) - picked + flat_scores)
    h[np.arange(n), flat_labels] = 0
    return (h.sum(1, dtype=flat_scores.dtype) / flat_scores.dtype.type(C)).mean(dtype=flat_scores.dtype)


def local_loss(flat_scores, flat_labels, args):
    if args.local_loss_func == 'ML':
        return ml_loss(flat_scores, flat_labels, args.margin)
    return ce_loss(flat_scores, flat_labels)


def decode(p, all_scores, lengths, args, C, o_idx, use_crf):
    """model_decompose.py:339-371 decode.  C already includes the +2 CRF tags when use_crf."""
    if use_crf:
        sc = all_scores.copy()
        if args.local_loss_func == 'CE1':
            sc[:, :, C - 3] = np.minimum(sc[:, :, C - 3], np.asarray(args.threshold, sc.dtype))
        path = crf_viterbi(sc, length_mask(lengths), p['crf_transitions'])
        pred = flatten_rows(path, lengths)
        if args.local_loss_func == 'CE1':
            pred[pred == C - 3] = o_idx
        return pred
    flat = flatten_rows(all_scores, lengths).copy()
    if args.local_loss_func == 'CE1':
        flat[:, C - 1] = np.minimum(flat[:, C - 1], np.asarray(args.threshold, flat.dtype))
    pred = flat.argmax(1).astype(np.int64)   # numpy argmax = first maximal index, like torch.max
    if args.local_loss_func == 'CE1':
        pred[pred == C - 1] = o_idx
    return pred


def decompose_forward_local(p, x, label, lengths, args, o_idx=0, train=True, dense_v=None):
    """model_decompose_single.py:207-304 forward_local / :483-580 FARNN_S_SF.forward
    (marryup_type 'none').  Returns (loss or None, flat_pred, flat_true, all_scores)."""
    all_scores, _, _ = decompose_scores(p, x, lengths, args, dense_v)
    C = p['C_output_mat'].shape[0]
    use_crf = bool(args.use_crf)
    flat_true = flatten_rows(label, lengths)
    loss = None
    if train:
        if use_crf:
            loss = crf_nll(all_scores, length_mask(lengths), label, p['crf_transitions'])
        else:
            loss = local_loss(flatten_rows(all_scores, lengths), flat_true, args)
    pred = decode(p, all_scores, lengths, args, C, o_idx, use_crf)
    return loss, pred, flat_true, all_scores


# --------------------------------------------------------------------------------------
# onehot i-FST: farnn/model_onehot.py (FARNN_S_O_I_S)
# --------------------------------------------------------------------------------------
def onehot_scores(p, x, lengths, args):
    """model_onehot.py:351-428 FARNN_S_O_I_S.forward_score.  Runs all x.shape[1] steps."""
    B, L = x.shape
    dt = p['language_tensor'].dtype
    S = p['output_mat'].shape[1]
    semiring = semiring_max if args.train_mode == 'max' else semiring_sum
    bx = reverse_rows(x, lengths)
    h0 = np.repeat(p['h0'][None, :], B, 0)
    hT = np.repeat(p['hT'][None, :], B, 0)
    hf, hb = h0.copy(), hT.copy()
    fwd = np.zeros((B, L, S), dt)
    bwd = np.zeros((B, L, S), dt)
    sum_tensor = p['language_tensor'] + p['wildcard_mat']
    o = p['output_mat'].sum(0)
    if args.local_loss_func != 'CE1':
        o = o + p['output_wildcard_vector']
    for i in range(L):
        hf = semiring(hf, sum_tensor[x[:, i]]) * o
        hf = nonlinear(hf, args.update_nonlinear)
        fwd[:, i] = hf
        hb = semiring(hb * o, sum_tensor[bx[:, i]].transpose(0, 2, 1))
        hb = nonlinear(hb, args.update_nonlinear)
        bwd[:, i] = hb
    alpha = np.concatenate([h0[:, None], fwd], 1)
    beta = reverse_rows(np.concatenate([hT[:, None], bwd], 1), lengths + 1)
    ab = alpha[:, 1:] * beta[:, 1:]
    all_scores = np.einsum('cs,bls->blc', p['output_mat'], ab)
    if args.use_priority:
        all_scores = priority(all_scores, p)
    return all_scores.astype(dt)


def onehot_local_decode(flat_scores, args, C, o_idx):
    """model_onehot.py:162-180 local_decode."""
    sc = flat_scores.copy()
    if args.local_loss_func == 'CE1':
        sc[:, C - 1] = np.minimum(sc[:, C - 1], np.asarray(args.threshold, sc.dtype))
    pred = sc.argmax(1).astype(np.int64)
    if args.local_loss_func == 'CE1':
        pred[pred == C - 1] = o_idx
    return pred


def onehot_forward_local(p, x, label, lengths, args, o_idx=0, train=True):
    """model_onehot.py:131-146 forward_local."""
    all_scores = onehot_scores(p, x, lengths, args)
    flat_true = flatten_rows(label, lengths)
    flat_scores = flatten_rows(all_scores, lengths)
    loss = local_loss(flat_scores, flat_true, args) if train else None
    pred = onehot_local_decode(flat_scores, args, p['output_mat'].shape[0], o_idx)
    return loss, pred, flat_true, all_scores


def onehot_forward_RE(p, x, lengths, args, o_idx=0):
    """model_onehot.py:148-160 forward_RE: un-flattened decode, B x L.  Under CE1 the reference
    returns the *clamped* clone of the scores (:153-154,160), so the oracle does too."""
    all_scores = onehot_scores(p, x, lengths, args)
    C = p['output_mat'].shape[0]
    sc = all_scores.copy()
    if args.local_loss_func == 'CE1':
        sc[:, :, C - 1] = np.minimum(sc[:, :, C - 1], np.asarray(args.threshold, sc.dtype))
    pred = sc.argmax(2).astype(np.int64)
    if args.local_loss_func == 'CE1':
        pred[pred == C - 1] = o_idx
    return pred, sc


# --------------------------------------------------------------------------------------
# CRF: baselines/crf.py
# --------------------------------------------------------------------------------------
def log_sum_exp(vec):
    """crf.py:16-27: max-shifted logsumexp over axis 1 of (B, T_from, T_to)."""
    m = vec.max(1, keepdims=True)
    return m[:, 0] + np.log(np.exp(vec - m).sum(1))


def crf_log_partition(feats, mask, trans):
    """crf.py:48-99 _calculate_PZ.  Returns per-sequence log Z (the reference sums them)."""
    B, L, T = feats.shape
    part = feats[:, 0, :] + trans[START_TAG][None, :]
    for t in range(1, L):
        cur = feats[:, t, None, :] + trans[None] + part[:, :, None]
        new = log_sum_exp(cur)
        part = np.where(mask[:, t, None], new, part)
    fin = log_sum_exp(trans[None] + part[:, :, None])
    return fin[:, STOP_TAG]


def crf_gold_score(feats, mask, tags, trans):
    """crf.py:202-251 _score_sentence.  Returns per-sequence gold path scores."""
    B, L, T = feats.shape
    prev = np.concatenate([np.full((B, 1), T - 2, tags.dtype), tags[:, :-1]], 1)
    energy = feats[np.arange(B)[:, None], np.arange(L)[None, :], tags] + trans[prev, tags]
    energy = np.where(mask, energy, 0).sum(1, dtype=feats.dtype)
    lens = mask.sum(1)
    last = tags[np.arange(B), lens - 1]
    return energy + trans[last, STOP_TAG]


def crf_nll(feats, mask, tags, trans):
    """crf.py:253-260 neg_log_likelihood_loss: sum over the batch of (log Z - gold)."""
    z = crf_log_partition(feats, mask, trans).sum(dtype=feats.dtype)
    g = crf_gold_score(feats, mask, tags, trans).sum(dtype=feats.dtype)
    return z - g


def crf_viterbi(feats, mask, trans):
    """crf.py:102-195 _viterbi_decode.  Returns decode_idx B x L int64 exactly as the reference
    leaves it (pad positions hold 0 except the last column which holds the final pointer)."""
    B, L, T = feats.shape
    lens = mask.sum(1).astype(np.int64)
    part = feats[:, 0, :] + trans[START_TAG][None, :]
    hist = [part]
    bps = []
    for t in range(1, L):
        cur = feats[:, t, None, :] + trans[None] + part[:, :, None]
        bp = cur.argmax(1)
        part = cur.max(1)
        hist.append(part)
        bp[~mask[:, t]] = 0
        bps.append(bp)
    bps.append(np.zeros((B, T), np.int64))
    hist = np.stack(hist, 1)                      # B x L x T
    bps = np.stack(bps, 1).astype(np.int64)       # B x L x T
    last_part = hist[np.arange(B), lens - 1]      # B x T
    last_vals = last_part[:, :, None] + trans[None]
    pointer = last_vals.argmax(1)[:, STOP_TAG].astype(np.int64)
    bps[np.arange(B), lens - 1, :] = pointer[:, None]
    out = np.zeros((B, L), np.int64)
    out[:, L - 1] = pointer
    for t in range(L - 2, -1, -1):
        pointer = bps[np.arange(B), t, pointer]
        out[:, t] = pointer
    return out
