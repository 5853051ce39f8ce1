"""Gradient oracle.  TEST INFRASTRUCTURE ONLY (same rules as re2nn_oracle.py).

The reference obtains every gradient from torch.autograd over its own forward
(train_decompose.py:192 `loss.backward()`).  This file restates that forward with differentiable
torch CPU ops (float64 by default) so that tests can ask for d loss / d parameter at sizes where no
golden fixture exists.  It is pinned against the reference's own gradients stored in
tests/golden/*.npz (tests/test_oracle_golden.py::test_decompose_gradients).
File:line citations as in re2nn_oracle.py (paths relative to /root/reference/src_seq/).
"""
import torch


def _nl(x, kind):
    if kind == 'relu':
        return torch.relu(x)
    if kind == 'tanh':
        return torch.tanh(x)
    if kind == 'relutanh':
        return torch.tanh(torch.relu(x))
    if kind == 'sigmoid':
        return torch.sigmoid(x)
    return x


def _reverse(a, lengths):
    out = a.clone()
    for b in range(a.shape[0]):
        n = int(lengths[b])
        out[b, :n] = a[b, :n].flip([0])
    return out


def _step(p, h, v, h_init, o, args, fwd):
    """farnn/model_decompose_single.py:138-200 (sum semiring)."""
    k = args.sigmoid_exponent
    if args.farnn >= 1:
        zt = torch.sigmoid((h @ p['Wss1'] + v @ p['Wrs1'] + p['bs1']) * k)
    hbar = h
    if args.farnn == 2:
        rt = torch.sigmoid((h @ p['Wss2'] + v @ p['Wrs2'] + p['bs2']) * k)
        hbar = (1 - rt) * h_init + rt * h
    if not fwd:
        hbar = hbar * o
    if fwd:
        nxt = ((hbar @ p['S1']) * v) @ p['S2'].T + hbar @ p['wildcard_mat']
        nxt = nxt * o
    else:
        nxt = ((hbar @ p['S2']) * v) @ p['S1'].T + hbar @ p['wildcard_mat'].T
    nxt = _nl(nxt, args.update_nonlinear)
    if args.farnn == 0:
        return nxt
    return (1 - zt) * h + zt * nxt


def decompose_scores(p, x, lengths, args, dense_v=None):
    """model_decompose_single.py:220-272 / :494-545.  p: dict of torch tensors (requires_grad as wanted)."""
    L = int(lengths.max())
    B = len(lengths)
    if dense_v is None:
        vt = p['V_embed'] * p['beta_vec'] + _nl(p['embedding'] @ p['embed_r_generalized'], args.additional_nonlinear) \
            * (1 - p['beta_vec'])                                    # model_decompose.py:222-241, per token id
        vf = vt[x[:, :L]]
    else:
        vf = dense_v[:, :L]
    vb = _reverse(vf, lengths)
    o = p['C_output_mat'].sum(0)
    if args.local_loss_func != 'CE1':
        o = o + p['wildcard_output_vector']
    h0 = p['h0'].unsqueeze(0).repeat(B, 1)
    hT = p['hT'].unsqueeze(0).repeat(B, 1)
    hf, hb = h0, hT
    fs, bs = [], []
    for i in range(L):
        hf = _step(p, hf, vf[:, i], h0, o, args, True)
        fs.append(hf)
        hb = _step(p, hb, vb[:, i], hT, o, args, False)
        bs.append(hb)
    alpha = torch.stack([h0] + fs, 1)
    beta = _reverse(torch.stack([hT] + bs, 1), lengths + 1)
    scores = torch.einsum('bls,cs->blc', alpha[:, 1:] * beta[:, 1:], p['C_output_mat'])
    if args.use_priority:
        scores = scores @ p['priority_mat'] + p['priority_bias']
    return scores


def crf_nll(feats, lengths, tags, trans):
    """baselines/crf.py:48-99,202-260."""
    B, L, T = feats.shape
    total = feats.new_zeros(())
    for b in range(B):
        n = int(lengths[b])
        part = feats[b, 0] + trans[T - 2]
        gold = feats[b, 0, tags[b, 0]] + trans[T - 2, tags[b, 0]]
        for t in range(1, n):
            part = torch.logsumexp(part[:, None] + trans + feats[b, t][None, :], 0)
            gold = gold + feats[b, t, tags[b, t]] + trans[tags[b, t - 1], tags[b, t]]
        z = torch.logsumexp(part + trans[:, T - 1], 0)
        gold = gold + trans[tags[b, n - 1], T - 1]
        total = total + (z - gold)
    return total


def decompose_loss(p, x, labels, lengths, args, dense_v=None):
    scores = decompose_scores(p, x, lengths, args, dense_v)
    L = int(lengths.max())
    if args.use_crf:
        return crf_nll(scores, lengths, labels, p['crf_transitions']), scores
    mask = torch.arange(L)[None, :] < lengths[:, None]
    flat = scores[mask]
    if args.local_loss_func == 'ML':      # nn.MultiMarginLoss(margin)  (model_decompose.py:84-85)
        return torch.nn.functional.multi_margin_loss(flat, labels[:, :L][mask], margin=float(args.margin)), scores
    return torch.nn.functional.cross_entropy(flat, labels[:, :L][mask]), scores


def grads(p_np, x, labels, lengths, args, dense_v=None, dtype=torch.float64, names=None):
    """numpy params -> (loss, {name: grad ndarray}) for the requested parameter names (default: all float params)."""
    p = {k: torch.tensor(v, dtype=dtype) for k, v in p_np.items()}
    names = names or [k for k in p if k not in ('priority_mat', 'priority_bias')]
    for k in names:
        p[k].requires_grad_(True)
    dv = None if dense_v is None else torch.tensor(dense_v, dtype=dtype)
    loss, _ = decompose_loss(p, torch.as_tensor(x), torch.as_tensor(labels), torch.as_tensor(lengths), args, dv)
    loss.backward()
    return float(loss), {k: (p[k].grad.numpy() if p[k].grad is not None else None) for k in names}
