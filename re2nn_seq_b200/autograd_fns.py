"""torch.autograd.Function wrappers: forward and backward both run in libre2nn_b200.so."""
import torch

from . import ops
from ._lib import V_DENSE, V_TOKEN


def _key(tensors):
    return tuple((t.data_ptr(), t._version) for t in tensors)


def _prepare(consts, p, dense_v, cache):
    """Token table / gate table / output-vector sum.  Cached per parameter version when no grad is needed."""
    farnn = consts['farnn']
    if dense_v is None:
        deps = [p['V_embed'], p['embedding'], p['embed_r_generalized'], p['beta_vec']]
    else:
        deps = []
    gdeps = [p[n] for n in ('Wrs1', 'bs1', 'Wrs2', 'bs2') if n in p]
    odeps = [p['C_output_mat'], p['wildcard_output_vector']]
    key = (_key(deps), _key(gdeps), _key(odeps), consts['additional_nonlinear'], consts['ce1'])
    if cache is not None and dense_v is None and cache.get('key') == key:
        return cache['vtab'], cache['gtab'], cache['o']
    if dense_v is None:
        vtab = ops.token_table(p['V_embed'], p['embedding'], p['embed_r_generalized'], p['beta_vec'],
                               consts['additional_nonlinear'])
    else:
        vtab = dense_v.reshape(-1, dense_v.shape[-1])
    gtab = None
    if farnn >= 1:
        gtab = ops.gate_table(vtab, p['Wrs1'], p['bs1'], p.get('Wrs2'), p.get('bs2'), farnn)
    o = ops.output_vector_sum(p['C_output_mat'], None if consts['ce1'] else p['wildcard_output_vector'])
    if cache is not None and dense_v is None:
        cache.update(key=key, vtab=vtab, gtab=gtab, o=o)
    return vtab, gtab, o


def _weight_prep(consts, p, cache):
    """Operand-format weight copies, cached per parameter version (inference on the tensor-core precisions)."""
    prec = consts['precision']
    if cache is None or prec == 'fp32' or consts.get('max_semiring'):
        return None
    deps = [p[n] for n in ('S1', 'S2', 'wildcard_mat', 'Wss1', 'Wss2') if n in p]
    key = (_key(deps), prec, consts['farnn'])
    if cache.get('wkey') == key:
        return cache['wprep']
    buf = ops.weight_prep(p['S1'], p['S2'], p['wildcard_mat'], p.get('Wss1'), p.get('Wss2'), consts['farnn'], prec)
    cache.update(wkey=key, wprep=buf)
    return buf


class _DecomposeScores(torch.autograd.Function):
    @staticmethod
    def forward(ctx, consts, names, pr, x, dense_v, lengths, L, cache, *tensors):
        p = {n: t.detach().contiguous() for n, t in zip(names, tensors)}
        need_grad = consts['grad_on'] and (any(ctx.needs_input_grad[8:]) or
                                           (dense_v is not None and ctx.needs_input_grad[4]))
        mx = consts.get('max_semiring', False)
        ctx.max_semiring = mx
        if mx:
            need_grad = False        # inference-only semiring; backward raises
        vtab, gtab, o = _prepare(consts, p, dense_v, None if need_grad else cache)
        wprep = None if need_grad else _weight_prep(consts, p, cache)
        Lpad = x.shape[1] if dense_v is None else dense_v.shape[1]
        pm, pb = (pr if consts['use_priority'] else (None, None))
        if consts.get('fuse_scores') and not need_grad and not mx and consts['farnn'] == 0 and not consts['full_pad'] \
                and consts['precision'] != 'fp32':
            # decode-only inference: (alpha * beta) comes straight out of the forward direction's epilogue
            fused = ops.decompose_recurrence_fused(x, lengths, L, vtab, p['S1'], p['S2'], p['wildcard_mat'], o, p['h0'],
                                                   p['hT'], consts['update_nonlinear'], consts['precision'],
                                                   v_mode=V_TOKEN if dense_v is None else V_DENSE, Lpad=Lpad, wprep=wprep)
            if fused is not None:
                B, S = lengths.shape[0], p['S1'].shape[0]
                return ops.label_scores_ab(fused[0], B, L, S, p['C_output_mat'], pm, pb, consts['precision'])
        alpha, beta, saves = ops.decompose_recurrence(
            x, lengths, L, vtab, gtab, p['S1'], p['S2'], p['wildcard_mat'], o, p['h0'], p['hT'],
            p.get('Wss1'), p.get('Wss2'), consts['farnn'], consts['update_nonlinear'], consts['sigmoid_exponent'],
            precision=consts['precision'], v_mode=V_TOKEN if dense_v is None else V_DENSE,
            full_pad=consts['full_pad'], save_for_backward=need_grad, Lpad=Lpad, max_semiring=mx, wprep=wprep)
        scores = ops.label_scores(alpha, beta, lengths, p['C_output_mat'], pm, pb, full_pad=consts['full_pad'],
                                  precision='fp32' if mx else consts['precision'])
        if need_grad:
            ctx.consts, ctx.names, ctx.pr, ctx.L = consts, names, pr, L
            ctx.saved = (p, x, dense_v, lengths, vtab, gtab, o, alpha, beta, saves)
        return scores

    @staticmethod
    def backward(ctx, dscores):
        from . import backward as bw
        if ctx.max_semiring:
            raise NotImplementedError("re2nn_b200: train_mode='max' (max-product semiring) is inference-only")
        grads = bw.decompose_backward(ctx, dscores.contiguous())
        dense_g = grads.pop('__dense_v__', None)
        out = [None, None, None, None, dense_g, None, None, None]
        for i, n in enumerate(ctx.names):
            out.append(grads.get(n) if ctx.needs_input_grad[8 + i] else None)
        return tuple(out)


def decompose_scores(consts, names, tensors, pr, x, dense_v, lengths, L, cache=None):
    consts = dict(consts, grad_on=torch.is_grad_enabled())     # Function.forward itself always runs under no_grad
    return _DecomposeScores.apply(consts, names, pr, x, dense_v, lengths, L, cache, *tensors)


class _CeLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, lengths, labels, n_total):
        loss = ops.ce_loss(scores, lengths, labels, n_total)
        ctx.save_for_backward(scores, lengths, labels)
        ctx.n_total = n_total
        return loss

    @staticmethod
    def backward(ctx, g):
        scores, lengths, labels = ctx.saved_tensors
        return ops.ce_loss_backward(scores, lengths, labels, ctx.n_total, g.contiguous().float()), None, None, None


def ce_loss(scores, lengths, labels, n_total):
    """Mean cross entropy over the n_total valid tokens (nn.CrossEntropyLoss default, model_decompose.py:80)."""
    return _CeLoss.apply(scores.contiguous(), lengths, labels, n_total)


class _OnehotScores(torch.autograd.Function):
    @staticmethod
    def forward(ctx, consts, pr, x, lengths, L, h0, hT, language, W, output_mat, out_wild):
        need_grad = consts['grad_on'] and any(ctx.needs_input_grad[5:])
        o = ops.output_vector_sum(output_mat, None if consts['ce1'] else out_wild)
        # language + W once per parameter version (the reference redoes this V*S*S add every forward)
        cache = consts.get('cache')
        key = (language.data_ptr(), language._version, W.data_ptr(), W._version)
        if cache is not None and cache.get('key') == key:
            summed = cache['sum']
        else:
            summed = ops.onehot_sum_tensor(language, W)
            if cache is not None:
                cache.clear()
                cache.update(key=key, sum=summed)
        alpha, beta = ops.onehot_recurrence(x, lengths, L, summed, None, o, h0, hT, consts['update_nonlinear'],
                                            consts['max_semiring'], consts['full_pad'], presummed=True)
        pm, pb = (pr if consts['use_priority'] else (None, None))
        scores = ops.label_scores(alpha, beta, lengths, output_mat, pm, pb, full_pad=consts['full_pad'])
        if need_grad:
            ctx.consts, ctx.pr, ctx.L = consts, pr, L
            ctx.saved = (x, lengths, h0, hT, summed, W, output_mat, o, alpha, beta)
        return scores

    @staticmethod
    def backward(ctx, dscores):
        from . import backward as bw
        dlang = bw.onehot_backward(ctx, dscores.contiguous())
        return (None, None, None, None, None, None, None, dlang, None, None, None)


def onehot_scores(consts, tensors, pr, x, lengths, L):
    t = [v.detach() if not v.requires_grad else v for v in tensors]
    consts = dict(consts, grad_on=torch.is_grad_enabled())
    return _OnehotScores.apply(consts, pr, x, lengths, L, *t)
