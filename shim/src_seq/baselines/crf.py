"""Shadows src_seq/baselines/crf.py: CRF on the warp-per-sequence CUDA kernels (re2nn_seq_b200/crf.py)."""
from re2nn_seq_b200.crf import CRF, START_TAG, STOP_TAG  # noqa: F401
