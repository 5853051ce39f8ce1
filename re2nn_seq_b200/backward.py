"""Gradient assembly for the autograd wrappers: calls the CUDA backward entry points and maps their
outputs onto the reference's parameter names (no arithmetic besides two broadcast adds)."""
import torch

from . import ops


def decompose_backward(ctx, dscores):
    consts, names = ctx.consts, ctx.names
    p, x, dense_v, lengths, vtab, gtab, o, alpha, beta, saves = ctx.saved
    want = {n for i, n in enumerate(names) if ctx.needs_input_grad[8 + i]}
    pr_mat = ctx.pr[0].detach() if consts['use_priority'] else None
    g = ops.decompose_backward(consts, p, x, dense_v, lengths, ctx.L, vtab, o, alpha, beta, saves, dscores, pr_mat, want)
    out = {k: v for k, v in g.items() if k in want}
    # o = sum_c C_output_mat[c, :] (+ wildcard_output_vector): every row of C_output_mat receives d_o
    if 'C_output_mat' in want:
        out['C_output_mat'] = g['C_output_mat'] + g['o'].unsqueeze(0)
    if 'wildcard_output_vector' in want:
        out['wildcard_output_vector'] = torch.zeros_like(p['wildcard_output_vector']) if consts['ce1'] else g['o']
    if dense_v is None:
        tw = want & {'V_embed', 'beta_vec', 'embed_r_generalized', 'embedding'}
        if tw:
            out.update(ops.token_table_backward(g['vtab'], p['V_embed'], p['embedding'], p['embed_r_generalized'],
                                                p['beta_vec'], consts['additional_nonlinear'], tw))
    else:
        out['__dense_v__'] = g['vtab'].view_as(dense_v)
    for n in ('bs1', 'bs2'):
        if n in out:
            out[n] = out[n].view_as(p[n])
    return out


def onehot_backward(ctx, dscores):
    consts = ctx.consts
    if consts['max_semiring']:
        raise NotImplementedError("re2nn_b200: the max-product onehot backward is not built (train_mode='max')")
    x, lengths, h0, hT, language, W, output_mat, o, alpha, beta = ctx.saved
    pr_mat = ctx.pr[0].detach() if consts['use_priority'] else None
    return ops.onehot_backward(x, lengths, ctx.L, language, W, o, h0, hT, alpha, beta, dscores, output_mat, pr_mat,
                               consts['update_nonlinear'], consts['full_pad'], presummed=True)   # `language` = language + W
