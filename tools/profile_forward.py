"""Run N cfg2 inference forwards (direct launches, no graph) inside a cudaProfilerStart/Stop range.
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
      python tools/profile_forward.py fp16x3 3
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
import re2nn_seq_b200 as r
from re2nn_seq_b200 import synth

prec = sys.argv[1] if len(sys.argv) > 1 else 'auto'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = synth.CONFIGS['cfg2']
farnn = int(sys.argv[3]) if len(sys.argv) > 3 else 0
args = synth.make_args(farnn=farnn, use_crf=1, update_nonlinear='tanh', beta=0.1)
f = synth.make_decompose_factors(0, c['V'], c['S'], c['R'], c['C'], c['D'], dtype=np.float32)
x, lens, lab = synth.make_batch(1000, c['B'], c['Lmax'], c['V'], c['C'])
torch.manual_seed(0)
m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f)
with torch.no_grad():
    m.crf.transitions.copy_(torch.from_numpy(synth.crf_transitions(0, m.C)))
m = m.cuda().eval()
m.precision = prec
m.use_cuda_graph = False
xt, lt, yt = (torch.from_numpy(a).cuda() for a in (x, lens, lab))
with torch.no_grad():
    for _ in range(2):
        m.forward_local(xt, yt, lt, train=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(n):
        m.forward_local(xt, yt, lt, train=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print('profiled %d forwards, precision %s' % (n, m._resolved_precision()))
