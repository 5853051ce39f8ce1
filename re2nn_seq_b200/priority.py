"""PriorityLayer with the reference's parameters (src_seq/farnn/priority.py:8-30).

Inside the FARNN modules the product is fused into re2nn_label_scores (second GEMM pass) and this module only
owns the two non-trainable parameters so that state_dict keys match (priority_layer.priority_mat /
priority_layer.priority_bias).  Called on its own, forward() computes scores @ priority_mat + priority_bias
(priority.py:20-30) with the library's GEMM."""
import torch
from torch import nn


class PriorityLayer(nn.Module):
    def __init__(self, C, priority_mat=None, priority_bias=None):
        super().__init__()
        base = torch.eye(C).float()
        if priority_mat is not None:
            given = torch.from_numpy(priority_mat).float()
            k = given.shape[0]
            base[:k, :k] = given
        self.priority_mat = nn.Parameter(base, requires_grad=False)
        self.priority_bias = nn.Parameter(torch.zeros(C).float(), requires_grad=False)

    def forward(self, scores):
        """scores (... x C) -> scores @ priority_mat + priority_bias, same shape (priority.py:20-30)."""
        from . import ops
        ops.require_cuda()
        dev = self.priority_mat.device
        if dev.type != 'cuda':
            raise RuntimeError("re2nn_b200 PriorityLayer: module must live on a CUDA device (call .cuda())")
        x = scores.to(dev).float().contiguous()
        out = ops.affine(x.reshape(-1, x.shape[-1]), self.priority_mat.detach(), self.priority_bias.detach())
        return out.reshape(x.shape).to(scores.device)
