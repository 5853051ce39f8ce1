"""How much of a cfg3 training step is GPU-busy?  Sums kernel durations (CUPTI via torch.profiler) and compares with
the wall time of the step: a large gap means the step is bound by launch overhead, not by the kernels."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
import re2nn_seq_b200 as r
from re2nn_seq_b200 import synth

c = synth.CONFIGS['cfg3']
farnn = int(sys.argv[1]) if len(sys.argv) > 1 else 0
args = synth.make_args(farnn=farnn, use_crf=1, update_nonlinear='tanh', beta=0.1, sigmoid_exponent=5, bias_init=5.0)
f = synth.make_decompose_factors(0, c['V'], c['S'], c['R'], c['C'], c['D'], dtype=np.float32)
x, lens, lab = synth.make_batch(1000, c['B'], c['Lmax'], c['V'], c['C'])
torch.manual_seed(0)
m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f)
with torch.no_grad():
    m.crf.transitions.copy_(torch.from_numpy(synth.crf_transitions(0, m.C)))
m = m.cuda().train()
m.train_precision = "auto"
xt, lt, yt = (torch.from_numpy(a).cuda() for a in (x, lens, lab))
def step():
    for q in m.parameters():
        q.grad = None
    loss, _, _ = m.forward_local(xt, yt, lt, train=True)
    loss.backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 10 * 1e3
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
busy = sum(e.device_time for e in ev) / 5 / 1e3 if hasattr(ev[0], 'device_time') else sum(e.cuda_time for e in ev) / 5 / 1e3
print('wall %.3f ms/step; GPU-busy (sum of kernel durations) %.3f ms/step; %d kernels/step' % (wall, busy, len(ev) / 5))
agg = {}
for e in ev:
    d = e.device_time if hasattr(e, 'device_time') else e.cuda_time
    a = agg.setdefault(e.name[:70], [0, 0.0]); a[0] += 1; a[1] += d
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print('   %-72s %5.1f / step  %8.1f us / step' % (k, v[0] / 5, v[1] / 5))
