"""`src_seq` shadow package: run the reference's own drivers UNCHANGED on the B200 kernels.

    PYTHONPATH=/path/to/re2nn_seq_b200_repo/shim:/path/to/re2nn_seq_b200_repo  python main.py --method decompose ...

Only the hot-path modules are shadowed (each re-exports the drop-in classes of re2nn_seq_b200 under the names the
drivers import -- train_decompose.py:10-12, train_onehot.py:11, RE.py:6, test.py:6-8, train_decompose_ptm.py:11,
farnn/model_decompose_single_with_bert.py:13, baselines/neural_softmax.py:5):

    src_seq.farnn.model_onehot                     FARNN_S_O, FARNN_S_O_I, FARNN_S_O_I_S
    src_seq.farnn.model_decompose                  FARNN_S_D_W
    src_seq.farnn.model_decompose_independent      FARNN_S_D_W_I
    src_seq.farnn.model_decompose_single           FARNN_S_D_W_I_S, FARNN_S_SF
    src_seq.farnn.model_decompose_single_with_bert FARNN_S_bert
    src_seq.farnn.bert_embeddings                  EmbedAggregator, WordEmbedding
    src_seq.farnn.priority                         PriorityLayer
    src_seq.baselines.crf                          CRF

Every other module (main, data, init_params, train_*, val, RE, utils, metrics, rule_utils, wfa, tools, ...) resolves
to the reference's own file: this package's __path__ continues into the reference tree named by
$RE2NN_REFERENCE_ROOT (default /root/reference, else the vendored oracle/_ref of this repo).
RE2NN_SHIM=off turns the shadowing off (every import falls through to the reference).
"""
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(os.path.dirname(_HERE))


def reference_root():
    cand = [os.environ.get('RE2NN_REFERENCE_ROOT'), '/root/reference', os.path.join(_REPO, 'oracle', '_ref')]
    for c in cand:
        if c and os.path.isdir(os.path.join(c, 'src_seq')):
            return c
    raise ImportError("src_seq shim: no reference tree found (set RE2NN_REFERENCE_ROOT to the RE2NN-SEQ checkout)")


def chain(path_list, *sub):
    """Continue a (sub)package into the reference tree; with RE2NN_SHIM=off the reference comes FIRST."""
    ref = os.path.join(reference_root(), 'src_seq', *sub)
    if os.path.isdir(ref) and ref not in path_list:
        if os.environ.get('RE2NN_SHIM', 'on').lower() in ('off', '0', 'false'):
            path_list.insert(0, ref)
        else:
            path_list.append(ref)


chain(__path__)
