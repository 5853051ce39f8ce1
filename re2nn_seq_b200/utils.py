"""Batch-shaping helpers with the reference's names and semantics (src_seq/utils.py:133-189),
written without per-row Python loops.  Pure index plumbing on integer/bool tensors."""
import torch


def get_length_mask(length, max_len=None):
    """B -> B x max_len bool, True where position < length (utils.py:133-144)."""
    assert len(length.shape) == 1
    max_len = int(max_len or length.max().item())
    return torch.arange(max_len, device=length.device, dtype=length.dtype).expand(len(length), max_len) \
        < length.unsqueeze(1)


def flatten(input, length):
    """Concatenate the valid prefixes batch-major (utils.py:153-164): B x L x ? -> N x ?."""
    L = input.shape[1]
    mask = get_length_mask(length.to(input.device), L)
    return input[mask]


def reverse(input, lengths):
    """Flip the first lengths[b] entries of every row, leave the pad tail in place (utils.py:183-189)."""
    B, L = input.shape[0], input.shape[1]
    pos = torch.arange(L, device=input.device).unsqueeze(0).expand(B, L)
    n = lengths.to(input.device).unsqueeze(1)
    src = torch.where(pos < n, n - 1 - pos, pos)
    idx = src.reshape(B, L, *([1] * (input.dim() - 2))).expand_as(input)
    return torch.gather(input, 1, idx)


def exclusive_offsets(lengths):
    """Start of every sequence in the flattened (batch-major, valid-only) layout."""
    return torch.cumsum(lengths, 0) - lengths
