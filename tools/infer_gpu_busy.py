"""Per-kernel GPU time of one cfg2 inference forward as replayed from the CUDA graph (CUPTI via torch.profiler)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
import re2nn_seq_b200 as r
from re2nn_seq_b200 import synth

prec = sys.argv[1] if len(sys.argv) > 1 else 'auto'
farnn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
c = synth.CONFIGS['cfg2']
args = synth.make_args(farnn=farnn, use_crf=1, update_nonlinear='tanh', beta=0.1)
f = synth.make_decompose_factors(0, c['V'], c['S'], c['R'], c['C'], c['D'], dtype=np.float32)
x, lens, lab = synth.make_batch(1000, c['B'], c['Lmax'], c['V'], c['C'])
torch.manual_seed(0)
m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f)
with torch.no_grad():
    m.crf.transitions.copy_(torch.from_numpy(synth.crf_transitions(0, m.C)))
m = m.cuda().eval()
m.precision = prec
xt, lt, yt = (torch.from_numpy(a).cuda() for a in (x, lens, lab))
with torch.no_grad():
    for _ in range(5):
        m.forward_local(xt, yt, lt, train=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        m.forward_local(xt, yt, lt, train=False)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 20 * 1e3
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            m.forward_local(xt, yt, lt, train=False)
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
dur = lambda e: e.device_time if hasattr(e, 'device_time') else e.cuda_time
print('%s farnn=%d: wall %.3f ms/forward (no L2 flush); GPU-busy %.3f ms/forward' % (m._resolved_precision(), farnn, wall, sum(dur(e) for e in ev) / 5 / 1e3))
agg = {}
for e in ev:
    a = agg.setdefault(e.name[:80], [0, 0.0]); a[0] += 1; a[1] += dur(e)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print('   %-82s %5.1f / fwd  %8.1f us / fwd' % (k, v[0] / 5, v[1] / 5))
