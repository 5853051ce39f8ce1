"""Debug helper: in-situ timeline of the step-GEMM launches of one cfg2 forward (direct launches and graph replay)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
import re2nn_seq_b200 as r
from re2nn_seq_b200 import synth, _lib

prec = sys.argv[1] if len(sys.argv) > 1 else 'fp16x3'
c = synth.CONFIGS['cfg2']
args = synth.make_args(farnn=0, use_crf=1, update_nonlinear='tanh', beta=0.1)
f = synth.make_decompose_factors(0, c['V'], c['S'], c['R'], c['C'], c['D'], dtype=np.float32)
x, lens, lab = synth.make_batch(1000, c['B'], c['Lmax'], c['V'], c['C'])
m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f).cuda().eval()
m.precision = prec
xt, lt, yt = (torch.from_numpy(a).cuda() for a in (x, lens, lab))
buf = torch.zeros(2 * 4096, dtype=torch.int64, device='cuda')
for graph in (False, True):
    m.use_cuda_graph = graph
    with torch.no_grad():
        for _ in range(3):
            m.forward_local(xt, yt, lt, train=False)
        torch.cuda.synchronize()
        buf.zero_()
        _lib.check(_lib.fn['re2nn_debug_set_tc_timeline'](C.c_void_p(buf.data_ptr())), 'tl')
        m.forward_local(xt, yt, lt, train=False)
        torch.cuda.synchronize()
        _lib.check(_lib.fn['re2nn_debug_set_tc_timeline'](None), 'tl')
    t = buf.cpu().numpy().reshape(-1, 2)
    t = t[t[:, 0] > 0]
    dur = (t[:, 1] - t[:, 0]) / 1e3
    gap = (t[1:, 0] - t[:-1, 1]) / 1e3
    print('%s graph=%s: %d launches; CTA(0,0,0) lifetime us: median %.1f (GEMM1 %.1f / GEMM2 %.1f); gap to next kernel us: median %.1f; span %.1f us'
          % (prec, graph, len(t), np.median(dur), np.median(dur[0:70:2]), np.median(dur[1:70:2]), np.median(gap[:69]),
             (t[69, 1] - t[0, 0]) / 1e3 if len(t) >= 70 else -1))
    print('   first 8 (start_rel, dur):', [(round((a - t[0, 0]) / 1e3, 1), round((b - a) / 1e3, 1)) for a, b in t[:8]])
