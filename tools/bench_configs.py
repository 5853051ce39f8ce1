"""Timings of the other BASELINE.json configurations (not the bench line): cfg1 onehot, cfg4, cfg5 (decompose + onehot).
Usage: python tools/bench_configs.py [cfg1|cfg4|cfg5|cfg5_onehot] ..."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
import re2nn_seq_b200 as r
from re2nn_seq_b200 import synth, ops

peaks = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json'))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0, 'bf16_tflops_sustained': 1400.0}


def timeit(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def onehot(V, S, C, B, L, fixed, reps, tag):
    args = synth.make_args(method='onehot', rand_constant=0.0)
    rs = np.random.RandomState(0)
    a = synth.make_onehot_automaton(0, 8, S, C, dtype=np.float32)          # small generator, then tile the vocabulary
    lang = torch.zeros((V + 1, S, S), dtype=torch.float32)
    rows = rs.choice(V, size=max(1, V // 20), replace=False)
    for i, rr in enumerate(rows):
        lang[rr] = torch.from_numpy(a['language_tensor'][i % 8])
    x, lens, lab = synth.make_batch(1, B, L, V, C, fixed_len=fixed)
    m = r.FARNN_S_O_I_S(lang.numpy(), a['output_mat'], a['wildcard_mat'], a['output_wildcard_vector'], a['final_vector'],
                        a['start_vector'], None, args, 0, False)
    xt, lt, yt = (torch.from_numpy(t).cuda() for t in (x, lens, lab))
    with torch.no_grad():
        ms = timeit(lambda: m.forward_local(xt, yt, lt, train=False), reps)
    ntok = int(lens.sum())
    by = ntok * (2.0 * S * S * 4 + 2 * S * 4 + (C + 1) * 4)
    print('%s onehot V=%d S=%d B=%d L=%d: %.3f ms/batch, %.3g tok/s, algorithmic %.1f GB/s = %.1f %% of measured HBM peak %.0f GB/s'
          % (tag, V, S, B, L, ms, ntok / ms * 1e3, by / ms / 1e6, 100 * by / ms / 1e6 / peaks['hbm_gbs'], peaks['hbm_gbs']))


def decompose(tag, cfgname, B, reps, precs):
    c = dict(synth.CONFIGS[cfgname])
    c['B'] = B
    args = synth.make_args(farnn=0, use_crf=1, update_nonlinear='tanh', beta=0.1)
    f = synth.make_decompose_factors(0, c['V'], c['S'], c['R'], c['C'], c['D'], dtype=np.float32)
    x, lens, lab = synth.make_batch(1, c['B'], c['Lmax'], c['V'], c['C'], fixed_len=c.get('fixed_len', False))
    torch.manual_seed(0)
    m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f)
    with torch.no_grad():
        m.crf.transitions.copy_(torch.from_numpy(synth.crf_transitions(0, m.C)))
    m = m.cuda().eval()
    xt, lt, yt = (torch.from_numpy(t).cuda() for t in (x, lens, lab))
    ntok = int(lens.sum())
    S, R, D, Cp = c['S'], c['R'], c['D'], m.C
    fl = 8 * S * R + 4 * S * S + 2 * D * R + 2 * S * Cp
    for prec in precs:
        m.precision = prec
        with torch.no_grad():
            ms = timeit(lambda: m.forward_local(xt, yt, lt, train=False), reps)
        print('%s decompose %s S=%d R=%d B=%d L<=%d %s: %.3f ms/batch, %.3g tok/s, %.1f TFLOP/s algorithmic = %.1f %% of measured bf16 peak'
              % (tag, cfgname, S, R, c['B'], c['Lmax'], prec, ms, ntok / ms * 1e3, ntok * fl / ms / 1e9,
                 100 * ntok * fl / ms / 1e9 / peaks['bf16_tflops_sustained']))
    del m
    torch.cuda.empty_cache()


which = sys.argv[1:] or ['cfg1', 'cfg4']
if 'cfg1' in which:
    onehot(900, 300, 127, 32, 46, False, 10, 'cfg1')
    onehot(900, 300, 127, 1024, 46, False, 5, 'cfg1xB1024')
if 'cfg4' in which:
    decompose('cfg4', 'cfg4', 4096, 5, ['fp16x3', 'bf16'])
if 'cfg5' in which:
    decompose('cfg5/8', 'cfg5', 8192, 3, ['bf16', 'fp16x3'])
if 'cfg5_full' in which:
    decompose('cfg5', 'cfg5', 65536, 2, ['bf16'])
if 'cfg5_onehot' in which:
    onehot(900, 1024, 128, 256, 64, True, 3, 'cfg5-onehot(B=256)')
