"""PriorityLayer with the reference's parameters (src_seq/farnn/priority.py:8-30).

The matrix product itself is fused into re2nn_label_scores (second GEMM pass); this module only
owns the two non-trainable parameters so that state_dict keys match
(priority_layer.priority_mat / priority_layer.priority_bias)."""
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
        raise RuntimeError("PriorityLayer is applied inside the fused label-score kernel; "
                           "call the owning FARNN module instead")
