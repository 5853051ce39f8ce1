"""Debug helper: where one steady-state step (k = 2) of the resident recurrence kernel spends its cycles (cfg2)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
import re2nn_seq_b200 as r
from re2nn_seq_b200 import synth, _lib, ops

prec = sys.argv[1] if len(sys.argv) > 1 else 'fp16x3'
cfg = sys.argv[2] if len(sys.argv) > 2 else 'cfg2'
c = synth.CONFIGS[cfg]
args = synth.make_args(farnn=0, use_crf=1, update_nonlinear='tanh', beta=0.1)
f = synth.make_decompose_factors(0, c['V'], c['S'], c['R'], c['C'], c['D'], dtype=np.float32)
x, lens, lab = synth.make_batch(1000, min(c['B'], 4096), c['Lmax'], c['V'], c['C'], fixed_len=True)
m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f).cuda().eval()
m.precision = prec
m.use_cuda_graph = False
xt, lt, yt = (torch.from_numpy(a).cuda() for a in (x, lens, lab))
buf = torch.zeros(32 * 65536, dtype=torch.int64, device='cuda')
orig = ops.label_scores
def patched(*a, **k):
    _lib.check(_lib.fn['re2nn_debug_set_tc_trace'](None), 'trace')      # stop tracing before the score GEMM
    return orig(*a, **k)
ops.label_scores = patched
with torch.no_grad():
    for _ in range(2):
        m.forward_scores(xt, lt)
    torch.cuda.synchronize()
    _lib.check(_lib.fn['re2nn_debug_set_tc_trace'](C.c_void_p(buf.data_ptr())), 'trace')
    m.forward_scores(xt, lt)
    torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(-1, 32)
t = t[t[:, 0] != 0]
t = t[t[:, 8] != 0]
print('%s %s: %d CTAs traced; kernel lifetime median %.0f cycles' % (prec, cfg, len(t), np.median(t[:, 7] - t[:, 0])))
names = {8: 'producer: G1 B tiles issued, waits for rows (step k)', 9: 'producer: rows of step k-1 published (ready)', 10: 'producer: all G1 loads issued',
         14: 'mma: first G1 k-block landed', 15: 'mma: G1 issued + committed', 18: 'epilogue: G1 tile stored',
         11: 'producer: G2 weight tiles issued, waits for Q', 12: 'producer: Q published (ready)', 13: 'producer: all G2 loads issued',
         16: 'mma: first G2 k-block landed', 17: 'mma: G2 issued + committed', 19: 'epilogue: G2 tile stored',
         26: 'producer: next step waits for rows', 27: 'producer: rows of step k published (ready)',
         4: 'epi warp0 G2: waits for accumulator', 5: 'epi warp0 G2: accumulator ready', 20: 'epi warp0 G2 chunk0: start',
         23: 'epi warp0 G2 chunk1: start', 24: 'epi warp0 G2 chunk1: transposed', 25: 'epi warp0 G2 chunk1: stored',
         30: 'epi warp0 G2: published'}
names[21] = 'epi warp0 G2 chunk0: transposed'; names[22] = 'epi warp0 G2 chunk0: stored'
ref = t[:, 9:10]
order = sorted(names, key=lambda s: np.median(t[:, s] - ref[:, 0]))
for s in order:
    d = t[:, s] - ref[:, 0]
    print('   %-52s p50 %7.0f  min %7.0f  max %7.0f' % (names[s], np.median(d), d.min(), d.max()))
