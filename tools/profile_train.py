"""N cfg3 training steps (fwd + CRF loss + bwd) inside a cudaProfilerStart/Stop range (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
import re2nn_seq_b200 as r
from re2nn_seq_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
farnn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
c = synth.CONFIGS['cfg3']
args = synth.make_args(farnn=farnn, use_crf=1, update_nonlinear='tanh', beta=0.1)
f = synth.make_decompose_factors(0, c['V'], c['S'], c['R'], c['C'], c['D'], dtype=np.float32)
x, lens, lab = synth.make_batch(1000, c['B'], c['Lmax'], c['V'], c['C'])
torch.manual_seed(0)
m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f)
with torch.no_grad():
    m.crf.transitions.copy_(torch.from_numpy(synth.crf_transitions(0, m.C)))
m = m.cuda().train()
m.train_precision = sys.argv[3] if len(sys.argv) > 3 else 'auto'
xt, lt, yt = (torch.from_numpy(a).cuda() for a in (x, lens, lab))
def step():
    for q in m.parameters():
        q.grad = None
    loss, _, _ = m.forward_local(xt, yt, lt, train=True)
    loss.backward()
for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(n):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled', n, 'training steps')
