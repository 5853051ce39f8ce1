"""Shadows src_seq/farnn/bert_embeddings.py: the aggregator upstream of FARNN_S_SF is one fused launch; the BERT
encoder classes themselves stay the reference's (loaded from its own file)."""
import importlib.util
import os

from re2nn_seq_b200.bert_embeddings import EmbedAggregator, WordEmbedding  # noqa: F401
from src_seq import reference_root

_ref = os.path.join(reference_root(), 'src_seq', 'farnn', 'bert_embeddings.py')
if os.path.exists(_ref):
    try:
        _spec = importlib.util.spec_from_file_location('src_seq.farnn._reference_bert_embeddings', _ref)
        _mod = importlib.util.module_from_spec(_spec)
        _spec.loader.exec_module(_mod)
        for _n in dir(_mod):
            if not _n.startswith('_') and _n not in ('EmbedAggregator', 'WordEmbedding'):
                globals()[_n] = getattr(_mod, _n)
    except Exception:      # transformers / checkpoints absent: the BERT encoders are out of scope, the aggregator is not
        pass
