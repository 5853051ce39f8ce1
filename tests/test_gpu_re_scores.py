"""GPU: teacher-score sweep + `.re.score` cache (RE.py:15-52,66-75,190) with the drop-in onehot model."""
import pickle

import numpy as np
import pytest
import torch

from helpers import oracle_params
from oracle import re2nn_oracle as orc

pytestmark = pytest.mark.gpu


def test_predict_by_re_sweep_and_cache(tmp_path):
    import re2nn_seq_b200 as r
    from re2nn_seq_b200 import re_scores, synth
    V, S, C = 60, 40, 9
    args = synth.make_args(method='onehot', rand_constant=0.0, threshold=0.99, use_crf=0, beta=1)
    a = synth.make_onehot_automaton(1, V, S, C, dtype=np.float64)
    m = r.FARNN_S_O_I_S(a['language_tensor'], a['output_mat'], a['wildcard_mat'], a['output_wildcard_vector'],
                        a['final_vector'], a['start_vector'], None, args, 2, False)
    splits = {}
    for i, (name, n) in enumerate((('train', 150), ('dev', 37), ('test', 64))):
        x, lens, lab = synth.make_batch(20 + i, n, 12, V, C)
        splits[name] = tuple(torch.from_numpy(t) for t in (x, lab, lens))
    path = str(tmp_path / 'automata.pkl')
    got = re_scores.predict_by_RE(m, splits, path, bz=32)
    assert len(got) == 6 and all(t.device.type == 'cpu' for t in got)
    # same bytes come back from the cache, in the reference's format (a plain pickle of the 6-tuple)
    with open(path + '.re.score', 'rb') as f:
        raw = pickle.load(f)
    again = re_scores.predict_by_RE(m, splits, path, bz=32)
    for x, y, z in zip(got, raw, again):
        assert torch.equal(x, y) and torch.equal(x, z)
    # every split against the numpy restatement of forward_RE (0/1 automaton: integer path counts, bit-exact)
    p = {k: np.asarray(v, dtype=np.float32) for k, v in a.items() if k not in ('priority_mat', 'language_rows')}
    p.update(h0=p['start_vector'], hT=p['final_vector'])
    for j, name in enumerate(('train', 'dev', 'test')):
        x, lab, lens = (t.numpy() for t in splits[name])
        pred, sc = orc.onehot_forward_RE(p, x, lens, args, 2)
        sc = sc.copy()
        sc[sc == np.float32(0.99)] = 1.0
        np.testing.assert_array_equal(got[j].numpy(), pred)
        np.testing.assert_array_equal(got[3 + j].numpy(), sc)
