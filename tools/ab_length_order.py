"""A/B on one box: one-launch schedule (ops.length_order) vs the torch.sort / cumsum / gather pipeline, cfg2 with and
without gates, CUDA-graph replay.  python tools/ab_length_order.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import re2nn_seq_b200 as r
from re2nn_seq_b200 import ops, synth

c = dict(synth.CONFIGS['cfg2'])
x, lens, lab = synth.make_batch(1000, c['B'], c['Lmax'], c['V'], c['C'])
xd, ld, yd = torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(lab).cuda()
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')
fast = ops.length_order
for farnn in (0, 2):
    args = synth.make_args(farnn=farnn, use_crf=1, update_nonlinear='tanh', beta=0.1, sigmoid_exponent=5, bias_init=5.0)
    f = synth.make_decompose_factors(0, c['V'], c['S'], c['R'], c['C'], c['D'], dtype=np.float32)
    res = {}
    for mode in ('fast', 'torch', 'fast', 'torch'):
        ops.length_order = fast if mode == 'fast' else (lambda lengths, L: None)
        torch.manual_seed(0)
        m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f).cuda().eval()
        with torch.no_grad():
            for _ in range(4):
                m.forward_local(xd, yd, ld, train=False)
            torch.cuda.synchronize()
            ts = []
            for _ in range(20):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); m.forward_local(xd, yd, ld, train=False); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
        res.setdefault(mode, []).append(float(np.median(ts)))
    print('farnn', farnn, {k: ['%.3f' % v for v in vs] for k, vs in res.items()}, flush=True)
