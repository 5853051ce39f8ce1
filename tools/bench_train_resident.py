"""A/B of the training step with resident launches (forward with save slabs + BPTT sweep) vs per-step launches, over
the per-GPU batch size.  python tools/bench_train_resident.py [B ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import re2nn_seq_b200 as r
from re2nn_seq_b200 import _lib, dist as rd, synth


def main():
    batches = [int(v) for v in sys.argv[1:]] or [1024, 2048, 4096, 8192]
    c = dict(synth.CONFIGS['cfg3'])
    args = synth.make_args(farnn=0, use_crf=1, update_nonlinear='tanh', beta=0.1, sigmoid_exponent=5, bias_init=5.0)
    f = synth.make_decompose_factors(0, c['V'], c['S'], c['R'], c['C'], c['D'], dtype=np.float32)
    torch.manual_seed(0)
    m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f)
    with torch.no_grad():
        m.crf.transitions.copy_(torch.from_numpy(synth.crf_transitions(0, m.C)))
    m = m.cuda().train()
    m.train_precision = 'auto'
    bucket = rd.GradBucket(m)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')
    for B in batches:
        x, lens, lab = synth.make_batch(1000, B, c['Lmax'], c['V'], c['C'])
        xd, ld, yd = torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(lab).cuda()

        def step():
            bucket.zero_grad()
            loss, _, _ = m.forward_local(xd, yd, ld, train=True)
            loss.backward()
            return loss

        res = {}
        for mode in (3, 1, 2, 0, 3, 1, 2, 0):
            _lib.check(_lib.fn['re2nn_debug_set_resident_train'](mode), 'resident_train')
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                flush.fill_(1.0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                step()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            res.setdefault(mode, []).append(float(np.median(ts)))
        print('B=%5d  ' % B + '   '.join('%s %s' % (n, ['%.2f' % v for v in res[k]]) for k, n in ((3, 'fwd+bwd resident'), (1, 'fwd only'), (2, 'bwd only'), (0, 'per-step'))), flush=True)
    _lib.check(_lib.fn['re2nn_debug_set_resident_train'](0), 'resident_train')


if __name__ == '__main__':
    main()
