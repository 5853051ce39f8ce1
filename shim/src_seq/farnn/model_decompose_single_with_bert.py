"""Shadows src_seq/farnn/model_decompose_single_with_bert.py (imported at train_decompose_ptm.py:11)."""
from re2nn_seq_b200.bert_embeddings import FARNN_S_bert  # noqa: F401
