"""Shadows src_seq/farnn/model_decompose_independent.py (imported at train_decompose.py:11, test.py:7)."""
from re2nn_seq_b200.model_fst import FARNN_S_D_W_I  # noqa: F401
