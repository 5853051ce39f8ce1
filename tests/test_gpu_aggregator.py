"""GPU: EmbedAggregator / FARNN_S_bert (the caller one step upstream of FARNN_S_SF) against the
reference formula in plain torch fp32 (bert_embeddings.py:82-97) and its autograd gradients."""
import numpy as np
import pytest
import torch
from torch import nn

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _setup(nl, seed=0, V=300, S=64, R=40, C=12, D=50):
    import re2nn_seq_b200 as r
    from re2nn_seq_b200 import synth
    args = synth.make_args(farnn=0, use_crf=0, update_nonlinear='tanh', beta=0.3, additional_nonlinear=nl,
                           train_V_embed=1, train_beta=1, train_word_embed=1)
    f = synth.make_decompose_factors(seed, V, S, R, C, D, dtype=np.float32)
    return r, args, f


def _formula(agg, inp, emb, nl):
    g = torch.einsum('bld,dr->blr', emb, agg.embed_r_generalized)
    g = {'none': lambda t: t, 'relu': torch.relu, 'tanh': torch.tanh, 'sigmoid': torch.sigmoid,
         'relutanh': lambda t: torch.tanh(torch.relu(t))}[nl](g)
    return agg.V_embed[inp] * agg.beta_vec + g * (1 - agg.beta_vec)


@pytest.mark.parametrize('nl', ['none', 'tanh', 'relutanh', 'sigmoid'])
def test_aggregator_matches_formula_and_autograd(nl):
    r, args, f = _setup(nl)
    torch.manual_seed(1)
    agg = r.EmbedAggregator(args, f['V'], f['pretrained_word_embed']).cuda()
    rs = np.random.RandomState(2)
    B, L = 37, 11
    inp = torch.from_numpy(rs.randint(0, 300, size=(B, L + 3))).cuda()
    lens = torch.from_numpy(rs.randint(2, L + 1, size=B)).cuda()
    lens[0] = L
    w = torch.randn(B, L, 40, device='cuda')
    out = agg(inp, lens)
    (out * w).sum().backward()
    got = {k: v.grad.clone() for k, v in agg.named_parameters() if v.grad is not None}
    agg.zero_grad()
    want = _formula(agg, inp[:, :L], agg.embed.embedding(inp[:, :L]), nl)
    (want * w).sum().backward()
    assert rel_err(out.detach().cpu().numpy(), want.detach().cpu().numpy()) < 1e-5
    assert set(got) == {k for k, v in agg.named_parameters() if v.grad is not None}
    for k, v in agg.named_parameters():
        assert rel_err(got[k].cpu().numpy(), v.grad.cpu().numpy()) < 2e-5, k


class _Encoder(nn.Module):
    """Stands in for the BERT encoder: any module with static_embed and a B x L x D forward."""

    def __init__(self, table):
        super().__init__()
        self.static_embed = torch.from_numpy(table).float()
        self.proj = nn.Embedding.from_pretrained(torch.from_numpy(table).float(), freeze=False)

    def forward(self, bert_input, attend_mask, valid_mask, lengths):
        return torch.tanh(self.proj(bert_input)) * 0.5


def test_farnn_s_bert_pipeline_and_gradients_to_the_encoder():
    r, args, f = _setup('tanh', seed=3)
    args.use_bert = 1
    torch.manual_seed(0)
    enc = _Encoder(f['pretrained_word_embed'])
    keep = ('S1', 'S2', 'C_output_mat', 'wildcard_mat', 'wildcard_output_vector', 'final_vector', 'start_vector')
    m = r.FARNN_S_bert(V=f['V'], static_embed=f['pretrained_word_embed'], priority_mat=None, args=args, o_idx=0,
                       embed=enc, **{k: f[k] for k in keep}).cuda()
    rs = np.random.RandomState(4)
    B, L = 48, 9
    x = torch.from_numpy(rs.randint(0, 300, size=(B, L))).cuda()
    lens = torch.from_numpy(rs.randint(2, L + 1, size=B).astype(np.int64)).cuda()
    lens[0] = L
    lab = torch.from_numpy(rs.randint(0, 12, size=(B, L))).cuda()
    loss, pred, true = m(x, x, None, None, lens, lab, train=True)
    loss.backward()
    assert torch.isfinite(loss) and pred.shape == true.shape == (int(lens.sum()),)
    assert m.embed.embed.proj.weight.grad is not None and float(m.embed.embed.proj.weight.grad.abs().sum()) > 0
    assert float(m.embed.embed_r_generalized.grad.abs().sum()) > 0
    # the same factors fed directly to FARNN_S_SF give the same loss
    with torch.no_grad():
        vecs = m.embed.forward_bert(x, x, None, None, lens)
    loss2, pred2, _ = m.slot_filler(vecs, lab, lens, train=True)
    assert rel_err(loss.item(), loss2.item()) < 1e-6
    assert torch.equal(pred, pred2)
