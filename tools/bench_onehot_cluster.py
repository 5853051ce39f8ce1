"""cfg1-shaped onehot recurrence: kernel time for cluster sizes 1 / 2 / 4 at several batch sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
import re2nn_seq_b200 as r
from re2nn_seq_b200 import synth, _lib

V, S, C, L = 900, 300, 127, 46
args = synth.make_args(method='onehot', rand_constant=0.0)
rs = np.random.RandomState(0)
a8 = synth.make_onehot_automaton(0, 8, S, C, dtype=np.float32)
lang = np.zeros((V + 1, S, S), dtype=np.float32)
for i, rr in enumerate(rs.choice(V, size=V // 20, replace=False)):
    lang[rr] = a8['language_tensor'][i % 8]
m = r.FARNN_S_O_I_S(lang, a8['output_mat'], a8['wildcard_mat'], a8['output_wildcard_vector'], a8['final_vector'],
                    a8['start_vector'], None, args, 0, False)
for B in (32, 64, 128, 1024):
    x, lens, lab = synth.make_batch(1, B, L, V, C)
    xt, lt = torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda()
    ntok = int(lens.sum())
    ref = None
    for stream, nc in ((4, 1), (8, 1), (4, 2), (8, 2), (4, 4), (8, 4)):      # stream column = rows in flight
        _lib.check(_lib.fn['re2nn_debug_set_onehot_cluster'](nc + (8 if stream == 4 else 0)), 'nc')
        ts = [m.time_recurrence(xt, lt) for _ in range(6)][2:]
        with torch.no_grad():
            sc = m.forward_score(xt, None, lt)
        ref = sc if ref is None else ref
        ms = float(np.mean(ts))
        print('B=%d rows_in_flight=%d nc=%d: %.3f ms  %.1f GB/s algorithmic  same=%s' % (B, stream, nc, ms, ntok * 2.0 * S * S * 4 / ms / 1e6,
                                                                      bool(torch.equal(sc, ref))))
_lib.check(_lib.fn['re2nn_debug_set_onehot_cluster'](0), 'nc')
