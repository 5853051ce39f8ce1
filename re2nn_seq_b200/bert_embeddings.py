"""Drop-ins for the caller one step upstream of FARNN_S_SF (reference: src_seq/farnn/bert_embeddings.py:31-128,
src_seq/farnn/model_decompose_single_with_bert.py:16-68): the rank factors fed to the recurrence

    v[b,t] = V_embed[x[b,t]] * beta_vec + phi_add(emb[b,t] @ embed_r_generalized) * (1 - beta_vec)

* static word embeddings (`WordEmbedding`): `emb` is a row gather, so v is a row gather of the hoisted token table
  (`re2nn_token_table`, one GEMM per parameter version) -- exactly what FARNN_S_D_W_I_S does internally.
* contextual embeddings (BERT): `emb` is B x L x D produced by the encoder.  The encoder itself stays the host's
  business (transformers, outside SURVEY section 8); everything after it -- the (B*L) x D x R GEMM, the additional
  nonlinearity and the blend -- is ONE fused launch of `re2nn_token_table` over B*L rows, with the hand-written
  backward (`re2nn_token_table_backward`) for V_embed, the embeddings, embed_r_generalized and beta_vec.
"""
import numpy as np
import torch
from torch import nn

from . import ops
from .model_decompose_single import FARNN_S_SF


class WordEmbedding(nn.Module):
    """bert_embeddings.py:31-43"""

    def __init__(self, args, word_embed):
        super().__init__()
        self.args = args
        self.embedding = nn.Embedding.from_pretrained(torch.from_numpy(np.asarray(word_embed)).float(),
                                                      freeze=(not args.train_word_embed))
        self.static_embed = torch.from_numpy(np.asarray(word_embed)).float()

    def forward(self, inp, lengths=None):
        return self.embedding(inp)


class _Aggregate(torch.autograd.Function):
    """rows x R = v_rows * beta + phi(emb_rows @ G) * (1 - beta), forward and backward through the C-ABI."""

    @staticmethod
    def forward(ctx, v_rows, emb_rows, G, beta_vec, nl):
        ctx.nl = nl
        ctx.save_for_backward(v_rows, emb_rows, G, beta_vec)
        return ops.token_table(v_rows, emb_rows, G, beta_vec, nl)

    @staticmethod
    def backward(ctx, dout):
        v_rows, emb_rows, G, beta_vec = ctx.saved_tensors
        names = ('V_embed', 'embedding', 'embed_r_generalized', 'beta_vec')
        want = [n for n, need in zip(names, ctx.needs_input_grad[:4]) if need]
        g = ops.token_table_backward(dout.contiguous(), v_rows, emb_rows, G, beta_vec, ctx.nl, want) if want else {}
        return g.get('V_embed'), g.get('embedding'), g.get('embed_r_generalized'), g.get('beta_vec'), None


class EmbedAggregator(nn.Module):
    """bert_embeddings.py:46-128.  `embed`: optional encoder module for `args.use_bert` (e.g. the reference's
    BertEmbedding); it must expose `static_embed` (V x D) and return B x L x D from its forward."""

    def __init__(self, args, V, word_embed, embed=None):
        super().__init__()
        self.args = args
        self.V_embed = nn.Parameter(torch.from_numpy(np.asarray(V)).float(), requires_grad=bool(args.train_V_embed))
        self.V, self.R = self.V_embed.size()
        if bool(getattr(args, 'use_bert', 0)):
            if embed is None:
                raise ValueError('EmbedAggregator: args.use_bert needs the contextual encoder passed as embed= '
                                 '(the BERT encoder is host-side code, not part of this library)')
            self.embed = embed
        else:
            self.embed = WordEmbedding(args, word_embed)
        embed_init_weight = self.embed.static_embed                          # V x D
        self.embed_r_generalized = nn.Parameter(torch.matmul(embed_init_weight.pinverse(), self.V_embed.data),
                                                requires_grad=True)        # D x R
        self.beta = args.beta
        self.beta_vec = nn.Parameter(torch.tensor([self.beta] * self.R).float(), requires_grad=bool(args.train_beta))
        if args.random:
            nn.init.xavier_normal_(self.V_embed)
            nn.init.xavier_normal_(self.embed_r_generalized)

    def get_generalized_v_embed_vec(self, v_batch_vec, emb_batch_vec):
        """B x L x R, B x L x D -> B x L x R (bert_embeddings.py:82-97), one fused launch."""
        ops.require_cuda()
        B, L, R = v_batch_vec.shape
        dev = self.V_embed.device
        out = _Aggregate.apply(v_batch_vec.to(dev).reshape(B * L, R).contiguous().float(),
                               emb_batch_vec.to(dev).reshape(B * L, -1).contiguous().float(),
                               self.embed_r_generalized, self.beta_vec, self.args.additional_nonlinear)
        return out.view(B, L, R)

    def _clip(self, inp, lengths):
        L = int(lengths.max())
        return inp[:, :L] if L < inp.shape[1] else inp

    def forward(self, inp, lengths):
        inp = self._clip(inp, lengths).to(self.V_embed.device)
        return self.get_generalized_v_embed_vec(self.V_embed[inp], self.embed(inp, lengths))

    def forward_bert(self, inp, bert_input, bert_attend_mask, bert_valid_mask, lengths):
        inp = self._clip(inp, lengths).to(self.V_embed.device)
        emb = self.embed(bert_input, bert_attend_mask, bert_valid_mask, lengths)         # B x L x D
        return self.get_generalized_v_embed_vec(self.V_embed[inp], emb)


class FARNN_S_bert(nn.Module):
    """model_decompose_single_with_bert.py:16-68: EmbedAggregator -> FARNN_S_SF."""

    def __init__(self, V=None, S1=None, S2=None, C_output_mat=None, wildcard_mat=None, wildcard_output_vector=None,
                 final_vector=None, start_vector=None, static_embed=None, priority_mat=None, args=None, o_idx=0,
                 is_cuda=True, embed=None):
        super().__init__()
        self.embed = EmbedAggregator(args, V, static_embed, embed=embed)
        self.slot_filler = FARNN_S_SF(S1=S1, S2=S2, C_output_mat=C_output_mat, wildcard_mat=wildcard_mat,
                                      wildcard_output_vector=wildcard_output_vector, final_vector=final_vector,
                                      start_vector=start_vector, priority_mat=priority_mat, args=args, o_idx=o_idx,
                                      is_cuda=is_cuda)

    def forward(self, input, bert_input, bert_attend_mask, bert_valid_mask, lengths, label, train=True, re_tags=None):
        vecs = self.embed.forward_bert(input, bert_input, bert_attend_mask, bert_valid_mask, lengths)
        return self.slot_filler(vecs, label, lengths, train=train, re_tags=re_tags)

    def forward_static(self, input, lengths, label, train=True, re_tags=None):
        """Same pipeline on the static word embeddings (args.use_bert = 0)."""
        vecs = self.embed(input, lengths)
        return self.slot_filler(vecs, label, lengths, train=train, re_tags=re_tags)
