"""Shared helpers for the parity tests (fixtures -> oracle parameter dicts)."""
import json
import os

import numpy as np

from re2nn_seq_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# reference state_dict key -> oracle key
_RENAME = {'embedding.weight': 'embedding', 'crf.transitions': 'crf_transitions',
           'priority_layer.priority_mat': 'priority_mat', 'priority_layer.priority_bias': 'priority_bias'}


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    meta = json.loads(str(z['meta']))
    return z, meta


def oracle_params(z, dtype=np.float32):
    p = {}
    for k in z.files:
        if k.startswith('p.'):
            n = k[2:]
            p[_RENAME.get(n, n)] = z[k].astype(dtype)
    return p


def golden_grads(z):
    return {k[2:]: z[k] for k in z.files if k.startswith('g.')}


def args_of(meta):
    return synth.make_args(**meta['flags'])


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
