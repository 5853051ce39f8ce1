"""Debug helper: per-CTA phase breakdown of the LAST step-GEMM2 launch of a real cfg2 recurrence (fixed lengths)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
import re2nn_seq_b200 as r
from re2nn_seq_b200 import synth, _lib, ops

prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
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
t = buf.cpu().numpy().reshape(-1, 32)
t = t[t[:, 0] != 0]
names = ['alive', 'setup', 'mma_issue(from setup)', 'prefetch(from setup)', 'acc_ready(from setup)', 'epilogue(warp2)', 'total']
d = np.stack([t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 2], t[:, 5] - t[:, 2], t[:, 6] - t[:, 5], t[:, 7] - t[:, 0]], 1)
print('%s: last traced launch, %d CTAs' % (prec, len(t)))
for n, col in zip(names, d.T):
    print('   %-28s mean %8.0f  p50 %8.0f  max %8.0f cycles' % (n, col.mean(), np.median(col), col.max()))
