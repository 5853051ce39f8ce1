"""Shadows src_seq/farnn/priority.py."""
from re2nn_seq_b200.priority import PriorityLayer  # noqa: F401
