"""Shadows src_seq/farnn/model_onehot.py (imported at train_onehot.py:11, RE.py:6)."""
from re2nn_seq_b200.model_fst import FARNN_S_O, FARNN_S_O_I  # noqa: F401
from re2nn_seq_b200.model_onehot import FARNN_S_O_I_S  # noqa: F401
