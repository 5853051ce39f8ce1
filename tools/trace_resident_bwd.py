"""Debug helper: where one steady-state step of the resident BPTT sweep spends its cycles (cfg3 shapes)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import re2nn_seq_b200 as r
from re2nn_seq_b200 import synth, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
c = synth.CONFIGS['cfg3']
args = synth.make_args(farnn=0, use_crf=1, update_nonlinear='tanh', beta=0.1)
f = synth.make_decompose_factors(0, c['V'], c['S'], c['R'], c['C'], c['D'], dtype=np.float32)
x, lens, lab = synth.make_batch(1000, B, c['Lmax'], c['V'], c['C'], fixed_len=True)
m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f).cuda().train()
m.train_precision = 'auto'
xt, lt, yt = (torch.from_numpy(a).cuda() for a in (x, lens, lab))
buf = torch.zeros(32 * 65536, dtype=torch.int64, device='cuda')
for it in range(3):
    m.zero_grad()
    loss, _, _ = m.forward_local(xt, yt, lt, train=True)
    torch.cuda.synchronize()
    if it == 2:
        _lib.check(_lib.fn['re2nn_debug_set_tc_trace'](C.c_void_p(buf.data_ptr())), 'trace')
    loss.backward()
    torch.cuda.synchronize()
_lib.check(_lib.fn['re2nn_debug_set_tc_trace'](None), 'trace')
t = buf.cpu().numpy().reshape(-1, 32)
t = t[t[:, 0] != 0]
t = t[t[:, 8] != 0]
print('%d CTAs traced; kernel lifetime median %.0f cycles' % (len(t), np.median(t[:, 7] - t[:, 0])))
names = {8: 'producer: G1 B tiles issued, waits for rows (step k)', 9: 'producer: rows of step k-1 published (ready)', 10: 'producer: all G1 loads issued',
         14: 'mma: first G1 k-block landed', 15: 'mma: G1 issued + committed', 18: 'epilogue: G1 tile stored',
         11: 'producer: G2 weight tiles issued, waits for Q', 12: 'producer: Q published (ready)', 13: 'producer: all G2 loads issued',
         16: 'mma: first G2 k-block landed', 17: 'mma: G2 issued + committed', 19: 'epilogue: G2 tile stored',
         26: 'producer: next step waits for rows', 30: 'epi warp0 G2: published',
         4: 'epi warp0 G2: first prefetch batch issued', 5: 'epi warp0 G2: accumulator ready', 20: 'epi warp0 G2 chunk0: before tmem ld', 21: 'epi warp0 G2 chunk0: transposed', 22: 'epi warp0 G2 chunk0: rows 0-15 stored', 23: 'epi warp0 G2 chunk0: rows 16-31 stored', 24: 'epi warp0 G2 chunk1: before tmem ld', 25: 'epi warp0 G2 chunk1: transposed'}
ref = t[:, 9:10]
order = sorted(names, key=lambda s: np.median(t[:, s] - ref[:, 0]))
for s in order:
    d = t[:, s] - ref[:, 0]
    print('   %-52s p50 %7.0f  min %7.0f  max %7.0f' % (names[s], np.median(d), d.min(), d.max()))
