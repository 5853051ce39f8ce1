"""Per-launch durations of the fused-scoring schedule at cfg5: backward-direction launches, then forward (fused epilogue)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
import re2nn_seq_b200 as r
from re2nn_seq_b200 import synth, ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
prec = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
c = dict(synth.CONFIGS['cfg5'])
args = synth.make_args(farnn=0, use_crf=1, update_nonlinear='tanh', beta=0.1)
f = synth.make_decompose_factors(0, c['V'], c['S'], c['R'], c['C'], c['D'], dtype=np.float32)
x, lens, lab = synth.make_batch(1, B, c['Lmax'], c['V'], c['C'], fixed_len=True)
m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f).cuda().eval()
m.precision = prec
xt, lt = torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda()
with torch.no_grad():
    for fuse in (True, False):
        m.forward_scores(xt, lt, fuse=fuse)
        torch.cuda.synchronize()
        ops.profile_enable(True); ops.profile_read()
        m.forward_scores(xt, lt, fuse=fuse)
        torch.cuda.synchronize()
        for cls in (1, 2):
            iv = ops.profile_intervals(cls)
            d = np.array([e - s for s, e in iv])
            n = len(d)
            print('fuse=%s class %d: %d launches, total %.2f ms; first half mean %.4f, second half mean %.4f' % (
                fuse, cls, n, d.sum(), d[:n // 2].mean(), d[n // 2:].mean()))
        ops.profile_read(); ops.profile_enable(False)
