"""Seeded synthetic automata, factors and token batches.

Everything here is host-side numpy; the same generators feed the golden-vector
script (tests/golden/make_golden.py), the parity tests and bench.py so that the
reference, the oracle and the CUDA path always see identical inputs.

Shapes follow the reference's constructor contracts:
  onehot    FARNN_S_O_I_S(language_tensor (V+1)xSxS, output_mat (C+1)xS, wildcard_mat SxS,
            output_wildcard_vector S, final_vector S, start_vector S, ...)
            (/root/reference/src_seq/farnn/model_onehot.py:311-344)
  decompose FARNN_S_D_W_I_S(V (V+1)xR, S1 SxR, S2 SxR, C_output_mat (C+1)xS, wildcard_mat SxS,
            wildcard_output_vector S, final_vector S, start_vector S, pretrained_word_embed (V+1)xD, ...)
            (/root/reference/src_seq/farnn/model_decompose_single.py:13-67)
"""
import argparse

import numpy as np

# Flags of /root/reference/src_seq/main.py:14-100 that the hot path reads,
# with main.py's defaults unless the BASELINE configs fix another value.
_ARG_DEFAULTS = dict(
    method='decompose', independent=2, train_mode='sum', local_loss_func='CE1',
    update_nonlinear='none', additional_nonlinear='none', farnn=0, sigmoid_exponent=5,
    bias_init=5.0, xavier=0, use_crf=0, use_priority=0, threshold=0.5, margin=0.3,
    additional_states=0, rand_constant=1e-5, random_pad_func='uniform', beta=1.0,
    train_beta=0, train_h0=0, train_hT=0, train_V_embed=0, train_wildcard=0,
    train_wildcard_wildcard=0, train_c_output=1, train_word_embed=0, random=0,
    marryup_type='none', c1_kdpr=1.0, c2_kdpr=1.0, c3_pr=1.0, seq_max_len=30, bz=500,
    rank=150, embed_dim=100,
)


def make_args(**overrides):
    """argparse.Namespace carrying every flag the farnn modules read from ``self.args``."""
    d = dict(_ARG_DEFAULTS)
    for k in overrides:
        if k not in d:
            raise KeyError("unknown hot-path flag %r" % k)
    d.update(overrides)
    return argparse.Namespace(**d)


def make_batch(seed, B, Lmax, V, C, fixed_len=False):
    """tokens x ~ U{0..V-1} (pads = V), lengths ~ U{ceil(Lmax/3)..Lmax} with lengths[0]=Lmax, labels ~ U{0..C-1}."""
    rs = np.random.RandomState(seed)
    if fixed_len:
        lengths = np.full((B,), Lmax, dtype=np.int64)
    else:
        lo = max(1, -(-Lmax // 3))
        lengths = rs.randint(lo, Lmax + 1, size=(B,)).astype(np.int64)
        lengths[0] = Lmax
    x = rs.randint(0, V, size=(B, Lmax)).astype(np.int64)
    labels = rs.randint(0, C, size=(B, Lmax)).astype(np.int64)
    pos = np.arange(Lmax)[None, :]
    pad = pos >= lengths[:, None]
    x[pad] = V
    labels[pad] = 0
    return x, lengths, labels


def _language_rows(rs, V, frac):
    n = max(1, int(round(V * frac)))
    return np.sort(rs.choice(V, size=n, replace=False))


def make_onehot_automaton(seed, V, S, C, lang_frac=0.05, dtype=np.float64):
    """Random 0/1 rule automaton in i-FST form: every state carries exactly one label."""
    rs = np.random.RandomState(seed)
    lang = np.zeros((V + 1, S, S), dtype=dtype)
    rows = _language_rows(rs, V, lang_frac)
    for r in rows:
        lang[r] = (rs.rand(S, S) < (2.0 / S)).astype(dtype)
    start_states = rs.choice(S, size=min(3, S), replace=False)
    wildcard = (rs.rand(S, S) < (1.0 / S)).astype(dtype)
    wildcard[start_states, start_states] = 1.0
    state_label = rs.randint(0, C + 1, size=(S,))
    output_mat = np.zeros((C + 1, S), dtype=dtype)
    output_mat[state_label, np.arange(S)] = 1.0
    start_vector = np.zeros((S,), dtype=dtype)
    start_vector[start_states] = 1.0
    final_vector = (rs.rand(S) < 0.1).astype(dtype)
    final_vector[rs.randint(0, S)] = 1.0
    output_wildcard_vector = np.zeros((S,), dtype=dtype)
    return dict(language_tensor=lang, output_mat=output_mat, wildcard_mat=wildcard,
                output_wildcard_vector=output_wildcard_vector, final_vector=final_vector,
                start_vector=start_vector, language_rows=rows)


def make_decompose_factors(seed, V, S, R, C, D, lang_frac=0.05, dtype=np.float64, dense_vocab=False):
    """Rank-R i-FST factors with O(1) state norms under tanh."""
    rs = np.random.RandomState(seed)
    sc = 1.0 / np.sqrt(R)
    S1 = (rs.randn(S, R) * sc).astype(dtype)
    S2 = (rs.randn(S, R) * sc).astype(dtype)
    V_embed = np.zeros((V + 1, R), dtype=dtype)
    if dense_vocab:
        V_embed[:V] = rs.randn(V, R) * sc
    else:
        rows = _language_rows(rs, V, lang_frac)
        V_embed[rows] = rs.randn(len(rows), R) * sc
    E = rs.randn(V + 1, D).astype(dtype)
    E[V] = 0.0
    start_states = rs.choice(S, size=min(3, S), replace=False)
    wildcard = (rs.rand(S, S) < (1.0 / S)).astype(dtype)
    wildcard[start_states, start_states] = 1.0
    state_label = rs.randint(0, C + 1, size=(S,))
    C_output_mat = np.zeros((C + 1, S), dtype=dtype)
    C_output_mat[state_label, np.arange(S)] = 1.0
    start_vector = np.zeros((S,), dtype=dtype)
    start_vector[start_states] = 1.0
    final_vector = (rs.rand(S) < 0.1).astype(dtype)
    final_vector[rs.randint(0, S)] = 1.0
    wildcard_output_vector = np.zeros((S,), dtype=dtype)
    return dict(V=V_embed, S1=S1, S2=S2, C_output_mat=C_output_mat, wildcard_mat=wildcard,
                wildcard_output_vector=wildcard_output_vector, final_vector=final_vector,
                start_vector=start_vector, pretrained_word_embed=E)


def crf_transitions(seed, T, noise=0.1):
    """Reference CRF init (crf.py:39-41) plus N(0, noise) so Viterbi is non-trivial."""
    rs = np.random.RandomState(seed)
    tr = np.zeros((T, T), dtype=np.float32)
    tr[:, T - 2] = -10000.0
    tr[T - 1, :] = -10000.0
    tr += (rs.randn(T, T) * noise).astype(np.float32)
    return tr


# BASELINE.json configs (SURVEY.md §8d "config table").
CONFIGS = {
    'cfg1': dict(kind='onehot', V=900, C=127, S=300, Lmax=46, B=32),
    'cfg2': dict(kind='decompose', V=12000, C=72, S=300, R=200, D=100, Lmax=35, B=4096,
                 use_crf=1, update_nonlinear='tanh', beta=0.1),
    'cfg3': dict(kind='decompose', V=12000, C=72, S=300, R=200, D=100, Lmax=35, B=1024,
                 use_crf=1, update_nonlinear='tanh', beta=0.1, train=True),
    'cfg4': dict(kind='decompose', V=2000, C=127, S=512, R=256, D=100, Lmax=128, B=4096,
                 use_crf=1, update_nonlinear='tanh', beta=0.1),
    'cfg5': dict(kind='decompose', V=900, C=128, S=1024, R=512, D=100, Lmax=64, B=65536,
                 use_crf=1, update_nonlinear='tanh', beta=0.1, fixed_len=True),
}
