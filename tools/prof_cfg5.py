"""cfg5-shaped decompose inference: whole-batch time and per-class step-GEMM time (library CUDA events).
Usage: python tools/prof_cfg5.py [B] [prec ...]      (default B=65536, bf16)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
import re2nn_seq_b200 as r
from re2nn_seq_b200 import synth, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
precs = sys.argv[2:] or ['bf16']
cfg = os.environ.get('CFG', 'cfg5')
reps = int(os.environ.get('REPS', '2'))
c = dict(synth.CONFIGS[cfg])
peaks = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))
args = synth.make_args(farnn=0, use_crf=1, update_nonlinear='tanh', beta=0.1)
f = synth.make_decompose_factors(0, c['V'], c['S'], c['R'], c['C'], c['D'], dtype=np.float32)
x, lens, lab = synth.make_batch(1, B, c['Lmax'], c['V'], c['C'], fixed_len=c.get('fixed_len', False))
torch.manual_seed(0)
m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f)
with torch.no_grad():
    m.crf.transitions.copy_(torch.from_numpy(synth.crf_transitions(0, m.C)))
m = m.cuda().eval()
m.use_cuda_graph = False
xt, lt, yt = (torch.from_numpy(t).cuda() for t in (x, lens, lab))
ntok = int(lens.sum())
S, R, D, Cp = c['S'], c['R'], c['D'], m.C
fl = 8 * S * R + 4 * S * S + 2 * D * R + 2 * S * Cp
for prec in precs:
    m.precision = prec
    with torch.no_grad():
        m.forward_local(xt, yt, lt, train=False)
        torch.cuda.synchronize()
        ops.profile_enable(True)
        ops.profile_read()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            m.forward_local(xt, yt, lt, train=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        pm, pn = ops.profile_read()
        ops.profile_enable(False)
    rec_fl = 2.0 * ntok * (2.0 * S * R + 2.0 * (R + S) * S)
    print('%s %s B=%d: %.3f ms/batch %.3g tok/s whole-step %.1f TFLOP/s = %.1f%% of %.0f; peak mem %.1f GB' % (
        cfg, prec, B, ms, ntok / ms * 1e3, ntok * fl / ms / 1e9, 100 * ntok * fl / ms / 1e9 / peaks['bf16_tflops_sustained'],
        peaks['bf16_tflops_sustained'], torch.cuda.max_memory_allocated() / 2**30))
    tot = sum(pm) / reps
    print('   classes ms/batch: gate %.3f g1 %.3f g2 %.3f resident %.3f (launches %s) -> recurrence %.1f TFLOP/s = %.1f%%' % (
        pm[0] / reps, pm[1] / reps, pm[2] / reps, pm[3] / reps, [n // reps for n in pn], rec_fl / tot / 1e9,
        100 * rec_fl / tot / 1e9 / peaks['bf16_tflops_sustained']))
