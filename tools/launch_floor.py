"""Debug helper: per-launch floor of the step-GEMM chain (tiny shapes => kernels are trivial, what remains is
launch + dependency latency) next to the real cfg2 shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
import re2nn_seq_b200 as r
from re2nn_seq_b200 import synth, ops

def run(B, S, R, L, prec, reps=20):
    args = synth.make_args(farnn=0, use_crf=0, update_nonlinear='tanh', beta=0.1)
    f = synth.make_decompose_factors(0, 500, S, R, 8, 16, dtype=np.float32)
    x, lens, lab = synth.make_batch(1, B, L, 500, 8, fixed_len=True)
    m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f).cuda()
    m.precision = prec
    xt, lt = torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda()
    with torch.no_grad():
        for _ in range(3):
            m.forward_scores(xt, lt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            m.forward_scores(xt, lt)
        e1.record()
        t_enq = time.perf_counter() - t0
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print('B=%5d S=%4d R=%4d L=%3d %-7s: %.3f ms per forward (GPU), %.1f us per step-GEMM launch; host enqueue %.3f ms per forward'
          % (B, S, R, L, prec, ms, 1e3 * ms / (2 * L), 1e3 * t_enq / reps))

for prec in ('bf16', 'fp16x3', 'fp32'):
    run(128, 32, 16, 35, prec)
for prec in ('bf16', 'fp16x3', 'tf32x3'):
    run(4096, 300, 200, 35, prec)
