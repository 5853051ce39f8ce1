"""Debug helper: per-CTA phase breakdown of the LAST step-GEMM2 launch of a real cfg2 recurrence (fixed lengths)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
import re2nn_seq_b200 as r
from re2nn_seq_b200 import synth, _lib, ops

prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
if len(sys.argv) > 2:
    _lib.check(_lib.fn['re2nn_debug_set_tc_cta_group'](int(sys.argv[2])), 'cg')
c = synth.CONFIGS['cfg2']
args = synth.make_args(farnn=0, use_crf=1, update_nonlinear='tanh', beta=0.1)
f = synth.make_decompose_factors(0, c['V'], c['S'], c['R'], c['C'], c['D'], dtype=np.float32)
x, lens, lab = synth.make_batch(1000, c['B'], c['Lmax'], c['V'], c['C'], fixed_len=True)
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
import re2nn_seq_b200.autograd_fns as af
with torch.no_grad():
    for _ in range(2):
        m.forward_scores(xt, lt)
    torch.cuda.synchronize()
    _lib.check(_lib.fn['re2nn_debug_set_tc_trace'](C.c_void_p(buf.data_ptr())), 'trace')
    m.forward_scores(xt, lt)
    torch.cuda.synchronize()
allt = buf.cpu().numpy().reshape(256, 256, 32)
which = int(sys.argv[3]) if len(sys.argv) > 3 else 41          # launch index: odd = GEMM2 of step (index // 2)
for which in (which - 1, which):
    t = allt[which]
    t = t[t[:, 0] != 0]
    t = t[t[:, 5] != 0]          # CTAs that ran an epilogue
    lead = t[t[:, 3] != 0]       # ... and issued the MMAs (the peer of a pair does not)
    print('%s: launch %d (%s), %d CTAs with a tile (%d issuing MMAs); cycles relative to kernel entry, median over CTAs'
          % (prec, which, 'GEMM2' if which & 1 else 'GEMM1', len(t), len(lead)))
    for name, col in (('setup done', t[:, 2] - t[:, 0]), ('previous grid complete', lead[:, 1] - lead[:, 0]),
                      ('first epilogue prefetch issued', t[:, 4] - t[:, 0]), ('all MMAs issued', lead[:, 3] - lead[:, 0]),
                      ('accumulator ready', t[:, 5] - t[:, 0]), ('epilogue done (warp 2)', t[:, 6] - t[:, 0]),
                      ('exit', t[:, 7] - t[:, 0])):
        print('   %-32s p50 %7.0f   min %7.0f   max %7.0f' % (name, np.median(col), col.min(), col.max()))
    kb = lead[:, 8:20] - lead[:, 1:2]
    print('   epilogue warp 2 after accumulator ready, per 32-col chunk [start, tmem->smem done, rows 0-15 stored, rows 16-31 stored]:',
          [int(v) for v in np.median(t[:, 20:32] - t[:, 5:6], 0)])
    print('   k-block arrival after previous grid complete:', ' '.join('%d' % v for v in np.median(kb, 0) if v > 0))
