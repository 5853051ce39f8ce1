"""re2nn_seq_b200 — B200-native (sm_100a) implementation of the RE2NN-SEQ transducer hot path.

Drop-in module classes (same names / signatures as the reference's src_seq.farnn / src_seq.baselines):
    FARNN_S_O_I_S, FARNN_S_D_W_I_S, FARNN_S_SF, CRF, PriorityLayer, EmbedAggregator, FARNN_S_bert,
    FARNN_S_D_W, FARNN_S_D_W_I, FARNN_S_O, FARNN_S_O_I (the FST family, model_fst.py)
The kernels live in libre2nn_b200.so (C-ABI: include/re2nn_b200.h), built by re2nn_seq_b200.build.
"""
from . import _lib  # noqa: F401  (fails loudly if the shared object is missing)
from .crf import CRF
from .model_decompose_single import FARNN_S_D_W_I_S, FARNN_S_SF
from .model_onehot import FARNN_S_O_I_S
from .priority import PriorityLayer
from .bert_embeddings import EmbedAggregator, FARNN_S_bert, WordEmbedding
from .model_fst import FARNN_S_D_W, FARNN_S_D_W_I, FARNN_S_O, FARNN_S_O_I

__all__ = ['FARNN_S_O_I_S', 'FARNN_S_D_W_I_S', 'FARNN_S_SF', 'CRF', 'PriorityLayer', 'EmbedAggregator', 'FARNN_S_bert',
           'WordEmbedding', 'FARNN_S_D_W', 'FARNN_S_D_W_I', 'FARNN_S_O', 'FARNN_S_O_I']
