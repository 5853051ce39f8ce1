"""Drop-ins for the FST (non-independent) model family -- SURVEY.md section 8 rows a16 / f4:

    FARNN_S_D_W     farnn/model_decompose.py:10-456               4-order CP factors, wildcard factors
    FARNN_S_D_W_I   farnn/model_decompose_independent.py:12-300   language x output factorisation (independent = 1)
    FARNN_S_O       farnn/model_onehot.py:8-180                   exact FST, label-indexed transitions V x C x S x S
    FARNN_S_O_I     farnn/model_onehot.py:183-306                 exact, language x output tensors

Same class names, constructor signatures, parameter names / requires_grad flags / state_dict keys, RNG draw order and
forward_local / forward_score / forward_RE / local_decode contracts as the reference classes.

How they map onto the kernels:

* FARNN_S_D_W keeps the rank-R structure (`((h S1) * _R) S2^T + h Wsum`, model_decompose.py:278-291), so it runs on the
  i-FST recurrence kernels unchanged (tensor cores, resident kernel, gates, BPTT): token table scaled by
  sum_c C_embed[c], wildcard matrix = S1w diag(sum_c C_wildcard[c]) S2w^T + wildcard_wildcard, no output mask.  Only the
  score differs (model_decompose.py:309-324: two rank contractions against [C_embed | C_wildcard]); it is three calls
  of the library GEMM with the element-wise products between them, and its gradient reaches the recurrence through
  `re2nn_decompose_backward`'s dalpha_in / dbeta_in.
* The other three multiply a label / output mask into the transition matrix ELEMENT-WISE
  (model_decompose_independent.py:172-176, model_onehot.py:271-284), which destroys the rank structure: every
  (sequence, step) owns a dense S x S matrix.  They are built from two library primitives -- `re2nn_gemm_nt` (the
  step-GEMM mainloops) and `re2nn_batched_vecmat` (sum / max semiring, argmax routing for the max backward) -- with
  the gathers and element-wise products between them left to torch (device-side glue; autograd differentiates it).
  No shipped configuration uses these classes; they are complete and parity-tested, not tuned.
* train_mode = 'max' trains through `re2nn_batched_vecmat`'s saved argmax (first maximal source state, torch.max's
  tie-break) in the three dense classes; `MaxProductTrainer` below gives FARNN_S_D_W_I_S the same path.
"""
import numpy as np
import torch
from torch import nn

from . import autograd_fns, ops
from ._lib import V_TOKEN
from .bert_embeddings import _Aggregate
from .crf import CRF
from .model_decompose_single import _ADD_NL, _UPDATE_NL, _DecomposeBase, _nl_name
from .priority import PriorityLayer
from .utils import exclusive_offsets, flatten


# ---- differentiable wrappers of the library primitives ----------------------------------------------------------------
class _MatmulNT(torch.autograd.Function):
    """C = A @ B^T (A: M x K, B: N x K) on the library GEMM; dA = dC @ B, dB = dC^T @ A on the same kernel."""

    @staticmethod
    def forward(ctx, A, B, precision):
        A, B = A.contiguous(), B.contiguous()
        ctx.save_for_backward(A, B)
        ctx.precision = precision
        return ops.gemm_nt(A, B, precision)

    @staticmethod
    def backward(ctx, dC):
        A, B = ctx.saved_tensors
        dC = dC.contiguous()
        dA = ops.gemm_nt(dC, B.t().contiguous(), 'fp32') if ctx.needs_input_grad[0] else None
        dB = ops.gemm_nt(dC.t().contiguous(), A.t().contiguous(), 'fp32') if ctx.needs_input_grad[1] else None
        return dA, dB, None


def matmul_nt(A, B, precision='fp32'):
    return _MatmulNT.apply(A, B, precision)


class _VecMat(torch.autograd.Function):
    """out[b,s] = (+|max)_j h[b,j] T[b,j,s]  (transposed: T[b,s,j]) -- utils.py:192-199 on one dense matrix per row."""

    @staticmethod
    def forward(ctx, h, T, transposed, maxp):
        h, T = h.contiguous(), T.contiguous()
        out, idx = ops.batched_vecmat(h, T, transposed, maxp)
        ctx.save_for_backward(h, T, idx)
        ctx.transposed, ctx.maxp = transposed, maxp
        return out

    @staticmethod
    def backward(ctx, dout):
        h, T, idx = ctx.saved_tensors
        dout = dout.contiguous()
        tr = ctx.transposed
        if not ctx.maxp:
            dh, _ = ops.batched_vecmat(dout, T, not tr, False)
            dT = (dout.unsqueeze(2) * h.unsqueeze(1)) if tr else (h.unsqueeze(2) * dout.unsqueeze(1))
            return dh, dT, None, None
        ix = idx.long()                                       # B x S: the source state that won target s
        hsel = h.gather(1, ix)
        if tr:                                                # T[b, s, ix[b,s]]
            tsel = T.gather(2, ix.unsqueeze(2)).squeeze(2)
            dT = torch.zeros_like(T).scatter_(2, ix.unsqueeze(2), (dout * hsel).unsqueeze(2))
        else:                                                 # T[b, ix[b,s], s]
            tsel = T.gather(1, ix.unsqueeze(1)).squeeze(1)
            dT = torch.zeros_like(T).scatter_(1, ix.unsqueeze(1), (dout * hsel).unsqueeze(1))
        dh = torch.zeros_like(h).scatter_add_(1, ix, dout * tsel)
        return dh, dT, None, None


def vecmat(h, T, transposed=False, maxp=False):
    return _VecMat.apply(h, T, transposed, maxp)


def _apply_nl(x, name):
    if name == 'relu':
        return torch.relu(x)
    if name == 'tanh':
        return torch.tanh(x)
    if name == 'relutanh':
        return torch.tanh(torch.relu(x))
    if name == 'sigmoid':
        return torch.sigmoid(x)
    return x


def _reverse_tokens(x, lengths):
    """utils.py:183-189 `reverse` for a B x L index matrix without the Python loop over the batch."""
    L = x.shape[1]
    pos = torch.arange(L, device=x.device).unsqueeze(0)
    n = lengths.unsqueeze(1)
    return x.gather(1, torch.where(pos < n, n - 1 - pos, pos))


def _beta_rows(hT, backward_scores, lengths):
    """reversed_backward_score_x[:, i + 1] for i = 0..L-1 (e.g. model_decompose.py:420-431): row n-1-i of
    cat([hT, backward_scores]) while i < n, row i + 1 afterwards."""
    B, L, S = backward_scores.shape
    full = torch.cat([hT.view(1, 1, S).expand(B, 1, S), backward_scores], dim=1)          # B x (L+1) x S
    pos = torch.arange(L, device=full.device).unsqueeze(0)
    n = lengths.unsqueeze(1)
    idx = torch.where(pos < n, n - 1 - pos, pos + 1)
    return full.gather(1, idx.unsqueeze(2).expand(B, L, S))


class _FstRecurrence(torch.autograd.Function):
    """alpha, beta of the rank-R recurrence with no output mask, differentiable w.r.t. the (already scaled) token
    table, S1, S2, the wildcard matrix, h0 / hT and the gate parameters."""

    @staticmethod
    def forward(ctx, consts, x, lengths, L, vtab, S1, S2, W, h0, hT, Wss1, Wrs1, bs1, Wss2, Wrs2, bs2):
        t = [q.detach().contiguous() if q is not None else None for q in (vtab, S1, S2, W, h0, hT, Wss1, Wrs1, bs1, Wss2, Wrs2, bs2)]
        vtab, S1, S2, W, h0, hT, Wss1, Wrs1, bs1, Wss2, Wrs2, bs2 = t
        farnn = consts['farnn']
        need = consts['grad_on'] and any(ctx.needs_input_grad[4:])
        gtab = ops.gate_table(vtab, Wrs1, bs1, Wrs2, bs2, farnn) if farnn >= 1 else None
        o = torch.ones((S1.shape[0],), dtype=torch.float32, device=S1.device)
        mx = consts['max_semiring']
        if mx and need:
            raise NotImplementedError("re2nn_b200 FARNN_S_D_W: train_mode='max' is inference-only for this class")
        alpha, beta, saves = ops.decompose_recurrence(
            x, lengths, L, vtab, gtab, S1, S2, W, o, h0, hT, Wss1, Wss2, farnn, consts['update_nonlinear'],
            consts['sigmoid_exponent'], precision='fp32' if mx else consts['precision'], v_mode=V_TOKEN, full_pad=False,
            save_for_backward=need, Lpad=x.shape[1], max_semiring=mx, zero_fill=True)
        if need:
            ctx.consts, ctx.L = consts, L
            ctx.saved = (x, lengths, vtab, o, alpha, beta, saves,
                         dict(S1=S1, S2=S2, wildcard_mat=W, h0=h0, hT=hT, Wss1=Wss1, Wrs1=Wrs1, bs1=bs1, Wss2=Wss2, Wrs2=Wrs2,
                              bs2=bs2))
        return alpha, beta

    @staticmethod
    def backward(ctx, dalpha, dbeta):
        x, lengths, vtab, o, alpha, beta, saves, p = ctx.saved
        p = {k: v for k, v in p.items() if v is not None}
        names = ['vtab', 'S1', 'S2', 'wildcard_mat', 'h0', 'hT', 'Wss1', 'Wrs1', 'bs1', 'Wss2', 'Wrs2', 'bs2']
        want = {n for i, n in enumerate(names) if ctx.needs_input_grad[4 + i] and (n == 'vtab' or n in p)}
        consts = dict(ctx.consts, full_pad=False, ce1=True)
        g = ops.decompose_backward(consts, p, x, None, lengths, ctx.L, vtab, o, alpha, beta, saves, None, None, want,
                                   dalpha_in=dalpha.contiguous(), dbeta_in=dbeta.contiguous())
        out = []
        for n in names:
            v = g.get(n) if n in want else None
            if v is not None and n in ('bs1', 'bs2'):
                v = v.view_as(p[n])
            out.append(v)
        return (None, None, None, None) + tuple(out)


# ---- max-product TRAINING of the i-FST classes ---------------------------------------------------
# train_mode = 'max' materialises one dense transition per sequence and step in the reference too
# (model_decompose_single.py:159-166, model_onehot.py:374-403 with utils.py:192-195); the gradient follows the argmax
# saved by re2nn_batched_vecmat.  Inference keeps the dedicated kernels (re2nn_decompose_max_recurrence, the MAXP
# onehot kernel); these differentiable forms run only when a gradient is needed.
def ifst_decompose_max_scores(m, x, dense_v, lengths, L):
    """all_scores B x L x C of FARNN_S_D_W_I_S / FARNN_S_SF under the max-product semiring, differentiable."""
    a = m.args
    S = m._S_full
    B = lengths.shape[0]
    k = float(a.sigmoid_exponent)
    nl = _nl_name(a.update_nonlinear, _UPDATE_NL)
    if dense_v is None:
        vt = _Aggregate.apply(m.V_embed, m.embedding.weight, m.embed_r_generalized, m.beta_vec,
                              _nl_name(a.additional_nonlinear, _ADD_NL))
        x = x[:, :L]
        xr = _reverse_tokens(x, lengths)
        v_at = lambda i, rev: vt[(xr if rev else x)[:, i]]
    else:
        v = dense_v[:, :L]
        pos = torch.arange(L, device=v.device).unsqueeze(0)
        n = lengths.unsqueeze(1)
        ridx = torch.where(pos < n, n - 1 - pos, pos)
        vr = v.gather(1, ridx.unsqueeze(2).expand(B, L, v.shape[2]))
        v_at = lambda i, rev: (vr if rev else v)[:, i]
    o = m.C_output_mat.sum(0)
    if a.local_loss_func != 'CE1':
        o = o + m.wildcard_output_vector
    h0 = m.h0.view(1, S).expand(B, S)
    hT = m.hT.view(1, S).expand(B, S)

    def step(h, V_vec, h_init, is_forward):
        h_bar = h
        if a.farnn >= 1:
            zt = torch.sigmoid(k * (matmul_nt(h, m.Wss1.t().contiguous()) + matmul_nt(V_vec, m.Wrs1.t().contiguous()) + m.bs1))
        if a.farnn == 2:
            rt = torch.sigmoid(k * (matmul_nt(h, m.Wss2.t().contiguous()) + matmul_nt(V_vec, m.Wrs2.t().contiguous()) + m.bs2))
            h_bar = (1 - rt) * h_init + rt * h
        if not is_forward:
            h_bar = h_bar * o
        X = (m.S1.unsqueeze(0) * V_vec.unsqueeze(1)).reshape(B * S, m.S1.shape[1])
        Tr = matmul_nt(X, m.S2).view(B, S, S) + m.wildcard_mat
        hn = vecmat(h_bar, Tr, transposed=not is_forward, maxp=True)
        if is_forward:
            hn = hn * o
        hn = _apply_nl(hn, nl)
        if a.farnn >= 1:
            hn = (1 - zt) * h + zt * hn
        return hn

    hf, hb, fw, bw = h0, hT, [], []
    for i in range(L):
        hf = step(hf, v_at(i, False), h0, True)
        fw.append(hf)
        hb = step(hb, v_at(i, True), hT, False)
        bw.append(hb)
    alpha = torch.stack(fw, dim=1)                                          # h0_forward_score[:, i + 1]
    beta = _beta_rows(m.hT, torch.stack(bw, dim=1), lengths)                # reversed_backward_score_x[:, i + 1]
    scores = matmul_nt((alpha * beta).reshape(B * L, S), m.C_output_mat).view(B, L, m.C)
    if a.use_priority:
        scores = (matmul_nt(scores.reshape(B * L, m.C), m.priority_layer.priority_mat.t().contiguous()) +
                  m.priority_layer.priority_bias).view(B, L, m.C)
    return scores


def ifst_onehot_max_scores(m, x, lengths):
    """all_scores B x L x C of FARNN_S_O_I_S under the max-product semiring, differentiable (model_onehot.py:366-426)."""
    a = m.args
    B, L = x.shape
    S = m.S
    nl = a.update_nonlinear if a.update_nonlinear in _UPDATE_NL else 'none'
    T = m.language_tensor + m.wildcard_mat
    o = m.output_mat.sum(0)
    if a.local_loss_func != 'CE1':
        o = o + m.output_wildcard_vector
    xr = _reverse_tokens(x, lengths)
    hf = m.h0.view(1, S).expand(B, S)
    hb = m.hT.view(1, S).expand(B, S)
    fw, bw = [], []
    for i in range(L):
        hf = _apply_nl(vecmat(hf, T[x[:, i]], False, True) * o, nl)
        fw.append(hf)
        hb = _apply_nl(vecmat(hb * o, T[xr[:, i]], True, True), nl)
        bw.append(hb)
    alpha = torch.stack(fw, dim=1)
    beta = _beta_rows(m.hT, torch.stack(bw, dim=1), lengths)
    Cn = m.output_mat.shape[0]
    scores = matmul_nt((alpha * beta).reshape(B * L, S), m.output_mat).view(B, L, Cn)
    if a.use_priority:
        scores = (matmul_nt(scores.reshape(B * L, Cn), m.priority_layer.priority_mat.t().contiguous()) +
                  m.priority_layer.priority_bias).view(B, L, Cn)
    return scores


# ---- FARNN_S_D_W ------------------------------------------------------------------------------------------------------------
class FARNN_S_D_W(_DecomposeBase):
    def __init__(self, V=None, C=None, S1=None, S2=None, C_wildcard=None, S1_wildcard=None, S2_wildcard=None,
                 wildcard_wildcard=None, final_vector=None, start_vector=None, pretrained_word_embed=None,
                 priority_mat=None, args=None, o_idx=0):
        super().__init__()
        self.is_cuda = torch.cuda.is_available()
        self.additional_states = args.additional_states
        self.args = args
        self.embedding = nn.Embedding.from_pretrained(torch.from_numpy(pretrained_word_embed).float(),
                                                      freeze=(not args.train_word_embed))
        self.C, self.R_W = C_wildcard.shape
        self.S, _ = S1_wildcard.shape
        _, self.R = C.shape
        self.t = 1
        self.use_crf = bool(args.use_crf)
        if self.use_crf:
            self.crf = CRF(self.C, self.is_cuda)
            self.C += 2
        self.priority_layer = PriorityLayer(self.C, priority_mat)
        self.random = bool(args.random)
        self.h0 = nn.Parameter(self.pad_additional_states(torch.from_numpy(start_vector).float()),
                               requires_grad=bool(args.train_h0))
        self.hT = nn.Parameter(self.pad_additional_states(torch.from_numpy(final_vector).float()),
                               requires_grad=bool(args.train_hT))
        self.init_forward_parameters(S1, S2, C, V, S1_wildcard, S2_wildcard, C_wildcard, wildcard_wildcard)
        self.beta = args.beta
        self.beta_vec = nn.Parameter(torch.tensor([self.beta] * self.R).float(), requires_grad=bool(args.train_beta))
        self.o_idx = o_idx
        self.not_o_idxs = [i for i in range(self.C) if i != self.o_idx]
        self.initialize()

    def init_forward_parameters(self, S1, S2, C, V, S1_w, S2_w, C_w, W):
        a = self.args
        self.S1 = nn.Parameter(self.pad_additional_states(torch.from_numpy(S1).float()), requires_grad=True)
        self.S2 = nn.Parameter(self.pad_additional_states(torch.from_numpy(S2).float()), requires_grad=True)
        self.V_embed = nn.Parameter(torch.from_numpy(V).float(), requires_grad=bool(a.train_V_embed))
        G = torch.matmul(self.embedding.weight.data.pinverse(), self.V_embed.data)
        self.embed_r_generalized = nn.Parameter(G, requires_grad=True)
        if a.use_crf == 1:
            C = np.concatenate((C, self.get_random((2, self.R)).numpy() * a.rand_constant), axis=0)
            C_w = np.concatenate((C_w, self.get_random((2, self.R_W)).numpy() * a.rand_constant), axis=0)
        self.C_wildcard = nn.Parameter(self.pad_additional_states(torch.from_numpy(C_w).float()),
                                       requires_grad=bool(a.train_wildcard))
        self.C_embed = nn.Parameter(torch.from_numpy(C).float(), requires_grad=True)
        self.S1_wildcard = nn.Parameter(self.pad_additional_states(torch.from_numpy(S1_w).float()),
                                        requires_grad=bool(a.train_wildcard))
        self.S2_wildcard = nn.Parameter(self.pad_additional_states(torch.from_numpy(S2_w).float()),
                                        requires_grad=bool(a.train_wildcard))
        self.wildcard_wildcard = nn.Parameter(self.pad_additional_states(torch.from_numpy(W).float()),
                                              requires_grad=bool(a.train_wildcard_wildcard))
        self._fst_gate_params(self.S + self.additional_states)
        if self.random:
            for w in (self.S1, self.S2, self.C_embed, self.V_embed, self.S1_wildcard, self.S2_wildcard, self.C_wildcard,
                      self.embed_r_generalized, self.wildcard_wildcard):
                nn.init.xavier_normal_(w)
            nn.init.normal_(self.h0)
            nn.init.normal_(self.hT)

    def _fst_gate_params(self, S_full):
        """model_decompose.py:137-171 (the farnn == 1 branch also re-initialises bs1 under xavier)."""
        a = self.args
        if a.farnn == 1:
            self.Wss1 = nn.Parameter(torch.randn((S_full, S_full)).float(), requires_grad=True)
            self.Wrs1 = nn.Parameter(torch.randn((self.R, S_full)).float(), requires_grad=True)
            self.bs1 = nn.Parameter(torch.ones((1, S_full)).float() * a.bias_init, requires_grad=True)
            if a.xavier:
                nn.init.xavier_normal_(self.Wss1)
                nn.init.xavier_normal_(self.Wrs1)
                nn.init.xavier_normal_(self.bs1)
        if a.farnn == 2:
            self.Wss1 = nn.Parameter(torch.randn((S_full, S_full)).float(), requires_grad=True)
            self.Wrs1 = nn.Parameter(torch.randn((self.R, S_full)).float(), requires_grad=True)
            self.bs1 = nn.Parameter(torch.ones((1, S_full)).float() * a.bias_init, requires_grad=True)
            self.Wss2 = nn.Parameter(torch.randn((S_full, S_full)).float(), requires_grad=True)
            self.Wrs2 = nn.Parameter(torch.randn((self.R, S_full)).float(), requires_grad=True)
            self.bs2 = nn.Parameter(torch.ones((1, S_full)).float() * a.bias_init, requires_grad=True)
            if a.xavier:
                for w in (self.Wss1, self.Wrs1, self.Wss2, self.Wrs2):
                    nn.init.xavier_normal_(w)

    # -- shared by FARNN_S_D_W / FARNN_S_D_W_I -----------------------------------------------------------------
    def _token_table(self):
        """get_generalized_v_embed_vec evaluated per token id (model_decompose.py:222-241): (V+1) x R."""
        return _Aggregate.apply(self.V_embed, self.embedding.weight, self.embed_r_generalized, self.beta_vec,
                                _nl_name(self.args.additional_nonlinear, _ADD_NL))

    def _gates(self):
        g = [getattr(self, n, None) for n in ('Wss1', 'Wrs1', 'bs1', 'Wss2', 'Wrs2', 'bs2')]
        return g

    def _apply_priority(self, scores):
        if not self.args.use_priority:
            return scores
        B, L, Cn = scores.shape
        out = matmul_nt(scores.reshape(B * L, Cn), self.priority_layer.priority_mat.t().contiguous())
        return (out + self.priority_layer.priority_bias).view(B, L, Cn)

    def forward_scores(self, input, lengths, shape=None):
        dev = self._device()
        x = input.to(dev).contiguous()
        lengths = lengths.to(dev).contiguous()
        L, _ = shape or self._host_shape(lengths)
        B = x.shape[0]
        S = self._S_full
        consts = dict(self._recurrence_consts(), grad_on=torch.is_grad_enabled())
        vt = self._token_table()                                           # V_vec per token
        vtab_rec = vt * self.C_embed.sum(0)                                # _R = V_vec * C_vec_sum   (:255-256)
        Wsum = matmul_nt(self.S1_wildcard * self.C_wildcard.sum(0), self.S2_wildcard) + self.wildcard_wildcard   # (:326-331)
        alpha, beta = _FstRecurrence.apply(consts, x, lengths, L, vtab_rec, self.S1, self.S2, Wsum, self.h0, self.hT,
                                           *self._gates())
        # get_final_score (:309-324) at alpha index i, beta index i + 1 (:426-427)
        a_prev = torch.cat([self.h0.view(1, 1, S).expand(B, 1, S), alpha[:, :L - 1]], dim=1).reshape(B * L, S)
        aS = matmul_nt(a_prev, torch.cat([self.S1, self.S1_wildcard], dim=1).t().contiguous())
        bS = matmul_nt(beta.reshape(B * L, S), torch.cat([self.S2, self.S2_wildcard], dim=1).t().contiguous())
        prod = aS * bS
        v = vt[x[:, :L]].reshape(B * L, self.R)
        prod = torch.cat([prod[:, :self.R] * v, prod[:, self.R:]], dim=1)
        scores = matmul_nt(prod, torch.cat([self.C_embed, self.C_wildcard], dim=1)).view(B, L, self.C)
        return self._apply_priority(scores)

    def forward_local(self, input, label, lengths, train=True):
        dev = self._device()
        lengths = lengths.to(dev).contiguous()
        shape = self._host_shape(lengths)
        all_scores = self.forward_scores(input, lengths, shape)
        return self._finish(all_scores, label, lengths, train, None, shape)


# ---- FARNN_S_D_W_I -------------------------------------------------------------------------------------------------------------
class FARNN_S_D_W_I(FARNN_S_D_W):
    def __init__(self, V=None, S1=None, S2=None, C_output=None, S1_output=None, S2_output=None, wildcard_mat=None,
                 wildcard_output=None, final_vector=None, start_vector=None, pretrained_word_embed=None, priority_mat=None,
                 args=None, o_idx=0):
        nn.Module.__init__(self)
        self.is_cuda = torch.cuda.is_available()
        self.additional_states = args.additional_states
        self.args = args
        self.embedding = nn.Embedding.from_pretrained(torch.from_numpy(pretrained_word_embed).float(),
                                                      freeze=(not args.train_word_embed))
        self.C, self.R_O = C_output.shape
        self.S, self.R = S1.shape
        self.t = 1
        self.use_crf = bool(args.use_crf)
        if self.use_crf:
            self.crf = CRF(self.C, self.is_cuda)
            self.C += 2
        self.priority_layer = PriorityLayer(self.C, priority_mat)
        self.random = bool(args.random)
        self.h0 = nn.Parameter(self.pad_additional_states(torch.from_numpy(start_vector).float()),
                               requires_grad=bool(args.train_h0))
        self.hT = nn.Parameter(self.pad_additional_states(torch.from_numpy(final_vector).float()),
                               requires_grad=bool(args.train_hT))
        self.init_forward_parameters(S1, S2, V, S1_output, S2_output, C_output, wildcard_mat, wildcard_output)
        self.beta = args.beta
        self.beta_vec = nn.Parameter(torch.tensor([self.beta] * self.R).float(), requires_grad=bool(args.train_beta))
        self.o_idx = o_idx
        self.not_o_idxs = [i for i in range(self.C) if i != self.o_idx]
        self.initialize()

    def init_forward_parameters(self, S1, S2, V, S1_o, S2_o, C_o, W, W_o):
        a = self.args
        self.S1 = nn.Parameter(self.pad_additional_states(torch.from_numpy(S1).float()), requires_grad=True)
        self.S2 = nn.Parameter(self.pad_additional_states(torch.from_numpy(S2).float()), requires_grad=True)
        self.V_embed = nn.Parameter(torch.from_numpy(V).float(), requires_grad=bool(a.train_V_embed))
        G = torch.matmul(self.embedding.weight.data.pinverse(), self.V_embed.data)
        self.embed_r_generalized = nn.Parameter(G, requires_grad=True)
        if a.use_crf:
            C_o = np.concatenate((C_o, self.get_random((2, self.R_O)).numpy() * a.rand_constant), axis=0)
        self.C_output = nn.Parameter(self.pad_additional_states(torch.from_numpy(C_o).float()),
                                     requires_grad=bool(a.train_wildcard))
        self.S1_output = nn.Parameter(self.pad_additional_states(torch.from_numpy(S1_o).float()),
                                      requires_grad=bool(a.train_wildcard))
        self.S2_output = nn.Parameter(self.pad_additional_states(torch.from_numpy(S2_o).float()),
                                      requires_grad=bool(a.train_wildcard))
        self.wildcard_mat = nn.Parameter(self.pad_additional_states(torch.from_numpy(W).float()),
                                         requires_grad=bool(a.train_wildcard_wildcard))
        self.wildcard_output = nn.Parameter(self.pad_additional_states(torch.from_numpy(W_o).float()),
                                            requires_grad=bool(a.train_wildcard_wildcard)) if W_o is not None else None
        self._fst_gate_params(self.S + self.additional_states)
        if self.random:
            for w in (self.S1, self.S2, self.V_embed, self.S1_output, self.S2_output, self.C_output,
                      self.embed_r_generalized, self.wildcard_mat):
                nn.init.xavier_normal_(w)
            nn.init.normal_(self.h0)
            nn.init.normal_(self.hT)

    def _device(self):
        ops.require_cuda()
        dev = self.S1.device
        if dev.type != 'cuda':
            raise RuntimeError("re2nn_b200: %s must live on a CUDA device (call .cuda()); there is no CPU path"
                               % type(self).__name__)
        return dev

    def _dense_transition(self, V_vec):
        """Tr[b,j,s] = sum_r S1[j,r] V_vec[b,r] S2[s,r] + wildcard_mat[j,s]   (model_decompose_independent.py:172-175)"""
        B = V_vec.shape[0]
        S = self._S_full
        X = (self.S1.unsqueeze(0) * V_vec.unsqueeze(1)).reshape(B * S, self.R)
        return matmul_nt(X, self.S2).view(B, S, S) + self.wildcard_mat

    def _step(self, h, tok, h_init, vt, O, is_forward, maxp):
        a = self.args
        k = float(a.sigmoid_exponent)
        V_vec = vt[tok]
        h_bar = h
        if a.farnn >= 1:
            zt = torch.sigmoid(k * (matmul_nt(h, self.Wss1.t().contiguous()) + matmul_nt(V_vec, self.Wrs1.t().contiguous()) + self.bs1))
        if a.farnn == 2:
            rt = torch.sigmoid(k * (matmul_nt(h, self.Wss2.t().contiguous()) + matmul_nt(V_vec, self.Wrs2.t().contiguous()) + self.bs2))
            h_bar = (1 - rt) * h_init + rt * h
        Tr = self._dense_transition(V_vec) * O
        hn = _apply_nl(vecmat(h_bar, Tr, transposed=not is_forward, maxp=maxp), _nl_name(a.update_nonlinear, _UPDATE_NL))
        if a.farnn >= 1:
            hn = (1 - zt) * h + zt * hn
        return hn

    def forward_scores(self, input, lengths, shape=None):
        dev = self._device()
        x = input.to(dev).contiguous()
        lengths = lengths.to(dev).contiguous()
        L, _ = shape or self._host_shape(lengths)
        x = x[:, :L]
        B, S = x.shape[0], self._S_full
        maxp = self.args.train_mode == 'max'
        vt = self._token_table()
        O = matmul_nt(self.S1_output * self.C_output.sum(0), self.S2_output)               # get_output_tensor_sum (:208-215)
        if self.args.local_loss_func != 'CE1':
            O = O + self.wildcard_output
        xr = _reverse_tokens(x, lengths)
        h0 = self.h0.view(1, S).expand(B, S)
        hT = self.hT.view(1, S).expand(B, S)
        hf, hb, fw, bw = h0, hT, [], []
        for i in range(L):
            hf = self._step(hf, x[:, i], h0, vt, O, True, maxp)
            fw.append(hf)
            hb = self._step(hb, xr[:, i], hT, vt, O, False, maxp)
            bw.append(hb)
        fw, bw = torch.stack(fw, dim=1), torch.stack(bw, dim=1)
        alpha = torch.cat([h0.unsqueeze(1), fw[:, :L - 1]], dim=1)          # h0_forward_score[:, i]
        beta = _beta_rows(self.hT, bw, lengths)                             # reversed_backward_score_x[:, i + 1]
        s1s2_out = (self.S1_output.t().unsqueeze(2) * self.S2_output.t().unsqueeze(1)).reshape(self.R_O, S * S)
        scores = []
        for i in range(L):                                                  # get_final_score (:198-206)
            bss = self._dense_transition(vt[x[:, i]])
            M = alpha[:, i].unsqueeze(2) * beta[:, i].unsqueeze(1) * bss
            br = matmul_nt(M.reshape(B, S * S), s1s2_out)
            scores.append(matmul_nt(br, self.C_output))
        return self._apply_priority(torch.stack(scores, dim=1))


# ---- exact (onehot) FST classes -----------------------------------------------------------------------------------------
class FARNN_S_O(nn.Module):
    def __init__(self, language_tensor=None, wildcard_tensor=None, wildcard_wildcard_mat=None, final_vector=None,
                 start_vector=None, priority_mat=None, args=None, o_idx=0, is_cuda=False):
        super().__init__()
        self.is_cuda = torch.cuda.is_available() and is_cuda
        self.args = args
        C, S, S = wildcard_tensor.shape
        self.S, self.C = S, C
        self.amp = args.rand_constant
        noisy = lambda a: torch.from_numpy(a).float() + torch.rand_like(torch.from_numpy(a).float()) * self.amp   # utils.py:273-274
        self.h0 = nn.Parameter(noisy(start_vector), requires_grad=False)
        self.hT = nn.Parameter(noisy(final_vector), requires_grad=False)
        self.language_tensor = nn.Parameter(noisy(language_tensor), requires_grad=True)                    # V x C x S x S
        self.wildcard_tensor = nn.Parameter(noisy(wildcard_tensor), requires_grad=bool(args.train_wildcard))   # C x S x S
        self.wildcard_wildcard_mat = nn.Parameter(torch.from_numpy(wildcard_wildcard_mat).float(),
                                                  requires_grad=bool(args.train_wildcard_wildcard))
        self.priority_layer = PriorityLayer(self.C, priority_mat)
        self.o_idx = o_idx
        self.initialize()
        if torch.cuda.is_available():
            self.cuda()

    def initialize(self):
        a = self.args
        self.t = 1
        if a.local_loss_func not in ('CE', 'CE1', 'ML'):
            raise NotImplementedError()

    def _device(self):
        ops.require_cuda()
        if not self.language_tensor.is_cuda:
            self.cuda()
        return self.language_tensor.device

    def _recurrence(self, x, lengths, sum_tensor, mask=None):
        """h <- relu(semiring(h, sum_tensor[x_t] [* mask])) in both directions (model_onehot.py:88-103, 271-290);
        -> (alpha rows i, beta rows i + 1) as the score loops index them."""
        B, L = x.shape
        S = self.S
        maxp = self.args.train_mode == 'max'
        xr = _reverse_tokens(x, lengths)
        hf = self.h0.view(1, S).expand(B, S)
        hb = self.hT.view(1, S).expand(B, S)
        fw, bw = [], []
        for i in range(L):
            Tr = sum_tensor[x[:, i]]
            Trb = sum_tensor[xr[:, i]]
            if mask is not None:
                Tr, Trb = Tr * mask, Trb * mask
            hf = torch.relu(vecmat(hf, Tr, False, maxp))
            fw.append(hf)
            hb = torch.relu(vecmat(hb, Trb, True, maxp))
            bw.append(hb)
        fw, bw = torch.stack(fw, dim=1), torch.stack(bw, dim=1)
        alpha = torch.cat([self.h0.view(1, 1, S).expand(B, 1, S), fw[:, :L - 1]], dim=1)
        return alpha, _beta_rows(self.hT, bw, lengths)

    def _apply_priority(self, score):
        if not self.args.use_priority:
            return score
        return matmul_nt(score, self.priority_layer.priority_mat.t().contiguous()) + self.priority_layer.priority_bias

    def _scores(self, x, lengths):
        B, L = x.shape
        lang, wild = self.language_tensor, self.wildcard_tensor
        sum_tensor = lang.sum(1) + wild.sum(0)                                     # V x S x S   (:81-85)
        if self.args.local_loss_func != 'CE1':
            sum_tensor = sum_tensor + self.wildcard_wildcard_mat
        all_tensor = lang + wild                                                   # V x C x S x S
        alpha, beta = self._recurrence(x, lengths, sum_tensor)
        out = []
        for i in range(L):                                                         # (:119-133)
            Tr = all_tensor[x[:, i]]                                               # B x C x S x S
            sc = torch.relu(Tr * alpha[:, i].view(B, 1, self.S, 1) * beta[:, i].view(B, 1, 1, self.S)).sum(dim=(2, 3))
            out.append(self._apply_priority(sc))
        return torch.stack(out, dim=1)

    def forward_score(self, input, label, lengths, train=True):
        out_dev = input.device
        dev = self._device()
        return self._scores(input.to(dev).contiguous(), lengths.to(dev).contiguous()).to(out_dev)

    def forward_local(self, input, label, lengths, train=True):
        out_dev = input.device
        dev = self._device()
        dl = lengths.to(dev).contiguous()
        scores = self._scores(input.to(dev).contiguous(), dl).contiguous()
        N = int(lengths.sum())
        label = label.to(dev)
        flattened_true_labels = flatten(label, dl)
        loss = None
        if train and self.args.local_loss_func == 'ML':
            from .kd import ml_loss
            loss = ml_loss(scores, dl, label, self.args.margin)
        elif train:
            loss = autograd_fns.ce_loss(scores, dl, label.contiguous(), N)
        with torch.no_grad():
            ce1 = self.args.local_loss_func == 'CE1'
            pred, _ = ops.argmax_decode(scores.detach(), dl, exclusive_offsets(dl), N, clamp_col=self.C - 1 if ce1 else -1,
                                        threshold=self.args.threshold, o_idx=self.o_idx)
        if loss is not None:
            loss = loss.to(out_dev)
        return loss, pred.to(out_dev), flattened_true_labels.to(out_dev)

    def forward_RE(self, input, label, lengths, train=False):
        out_dev = input.device
        dev = self._device()
        with torch.no_grad():
            dl = lengths.to(dev).contiguous()
            scores = self._scores(input.to(dev).contiguous(), dl).contiguous()
            ce1 = self.args.local_loss_func == 'CE1'
            _, pred = ops.argmax_decode(scores, dl, None, 0, clamp_col=self.C - 1 if ce1 else -1,
                                        threshold=self.args.threshold, o_idx=self.o_idx, want_flat=False, want_padded=True)
            if ce1:
                scores = scores.clone()
                scores[:, :, self.C - 1].clamp_(max=float(self.args.threshold))
        return pred.to(out_dev), scores.to(out_dev)

    def local_decode(self, all_scores=None):
        assert torch.is_tensor(all_scores)
        out_dev = all_scores.device
        dev = self._device()
        with torch.no_grad():
            sc = all_scores.detach().to(dev).float().contiguous().unsqueeze(1)
            N = sc.shape[0]
            ones = torch.ones((N,), dtype=torch.int64, device=dev)
            ce1 = self.args.local_loss_func == 'CE1'
            pred, _ = ops.argmax_decode(sc, ones, torch.arange(N, dtype=torch.int64, device=dev), N,
                                        clamp_col=self.C - 1 if ce1 else -1, threshold=self.args.threshold, o_idx=self.o_idx)
        return pred.to(out_dev)


class FARNN_S_O_I(FARNN_S_O):
    def __init__(self, language_tensor=None, output_tensor=None, wildcard_mat=None, output_wildcard_mat=None,
                 final_vector=None, start_vector=None, priority_mat=None, args=None, o_idx=0, is_cuda=False):
        nn.Module.__init__(self)
        self.is_cuda = torch.cuda.is_available() and is_cuda
        self.args = args
        C, S, S = output_tensor.shape
        self.S, self.C = S, C
        self.amp = args.rand_constant
        noisy = lambda a: torch.from_numpy(a).float() + torch.rand_like(torch.from_numpy(a).float()) * self.amp
        self.h0 = nn.Parameter(noisy(start_vector), requires_grad=False)
        self.hT = nn.Parameter(noisy(final_vector), requires_grad=False)
        self.language_tensor = nn.Parameter(noisy(language_tensor), requires_grad=True)          # V x S x S
        self.wildcard_mat = nn.Parameter(noisy(wildcard_mat), requires_grad=False)               # S x S
        self.output_tensor = nn.Parameter(torch.from_numpy(output_tensor).float(), requires_grad=False)   # C x S x S
        self.output_wildcard_mat = nn.Parameter(torch.from_numpy(output_wildcard_mat).float(), requires_grad=False) \
            if output_wildcard_mat is not None else None
        self.priority_layer = PriorityLayer(self.C, priority_mat)
        self.o_idx = o_idx
        self.initialize()
        if torch.cuda.is_available():
            self.cuda()

    def _scores(self, x, lengths):
        B, L = x.shape
        S = self.S
        sum_tensor = self.language_tensor + self.wildcard_mat                      # V x S x S   (:259)
        sum_output = self.output_tensor.sum(0)
        if self.args.local_loss_func != 'CE1':
            sum_output = sum_output + self.output_wildcard_mat
        mask = sum_output if self.args.independent == 2 else None
        alpha, beta = self._recurrence(x, lengths, sum_tensor, mask)
        out_flat = self.output_tensor.reshape(self.C, S * S)
        out = []
        for i in range(L):                                                         # get_final_score (:226-230)
            M = alpha[:, i].unsqueeze(2) * beta[:, i].unsqueeze(1) * sum_tensor[x[:, i]]
            out.append(self._apply_priority(matmul_nt(M.reshape(B, S * S), out_flat)))
        return torch.stack(out, dim=1)
