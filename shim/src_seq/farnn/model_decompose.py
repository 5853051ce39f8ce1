"""Shadows src_seq/farnn/model_decompose.py (imported at train_decompose.py:10, test.py:6)."""
from re2nn_seq_b200.model_fst import FARNN_S_D_W  # noqa: F401
