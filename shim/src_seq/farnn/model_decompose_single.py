"""Shadows src_seq/farnn/model_decompose_single.py (imported at train_decompose.py:12, test.py:8,
model_decompose_single_with_bert.py:13)."""
from re2nn_seq_b200.model_decompose_single import FARNN_S_D_W_I_S, FARNN_S_SF  # noqa: F401
