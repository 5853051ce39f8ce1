"""CRF with the reference's interface (src_seq/baselines/crf.py:29-260), computed by the
warp-per-sequence CUDA kernels in csrc/crf.cu."""
import torch
from torch import nn

from . import ops

START_TAG = -2
STOP_TAG = -1


class _CrfNll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, transitions, lengths, tags, grad_on):
        need = grad_on and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        loss, _, part = ops.crf_nll(feats, transitions, lengths, tags, save=need)
        if need:
            ctx.save_for_backward(feats, transitions, lengths, tags, part)
        return loss

    @staticmethod
    def backward(ctx, g):
        feats, transitions, lengths, tags, part = ctx.saved_tensors
        dfeats, dtrans = ops.crf_nll_backward(feats, transitions, lengths, tags, part, g.contiguous().float())
        return dfeats, dtrans, None, None, None


class CRF(nn.Module):
    def __init__(self, tagset_size, gpu):
        super().__init__()
        self.gpu = gpu
        self.tagset_size = tagset_size
        T = tagset_size + 2
        init = torch.zeros(T, T)
        init[:, START_TAG] = -10000.0      # nothing transitions into START   (crf.py:40)
        init[STOP_TAG, :] = -10000.0       # nothing leaves STOP              (crf.py:41)
        self.transitions = nn.Parameter(init, requires_grad=True)

    @staticmethod
    def _lengths(mask):
        return mask.long().sum(1)

    def _prep(self, feats):
        ops.require_cuda()
        dev = self.transitions.device
        if not self.transitions.is_cuda:
            raise RuntimeError("re2nn_b200 CRF: module must live on a CUDA device (call .cuda())")
        return feats.to(dev).float().contiguous()

    def neg_log_likelihood_loss(self, feats, mask, tags, lengths=None):
        assert feats.size(2) == self.tagset_size + 2          # crf.py:58
        feats = self._prep(feats)
        dev = feats.device
        lengths = (self._lengths(mask) if lengths is None else lengths).to(dev).contiguous()
        tags = tags.to(dev).contiguous()
        return _CrfNll.apply(feats, self.transitions, lengths, tags, torch.is_grad_enabled())

    def _viterbi_decode(self, feats, mask, lengths=None):
        assert feats.size(2) == self.tagset_size + 2          # crf.py:114
        with torch.no_grad():
            feats = self._prep(feats.detach())
            lengths = (self._lengths(mask) if lengths is None else lengths).to(feats.device).contiguous()
            _, padded = ops.crf_viterbi(feats, self.transitions.detach(), lengths, want_padded=True)
        return None, padded

    def forward(self, feats, mask):
        return self._viterbi_decode(feats, mask)
