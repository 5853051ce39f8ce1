#!/usr/bin/env python
"""bench.py — headline benchmark of the RE2NN-SEQ transducer hot path on B200.

Metric (BASELINE.json): valid token positions / second of decompose i-FST inference + Viterbi decode.
Workload at N=1: BASELINE.json configs[1] ("cfg2": V=12000, C=72, S=300, R=200, D=100, len<=35, B=4096,
tanh update, CRF, beta=0.1), synthetic automaton factors and token batches (re2nn_seq_b200/synth.py).
One "step" = one forward_local(train=False) over one batch = scores + Viterbi for every sequence.
N>1: one process per GPU (torchrun), every rank runs its own batch of the same shape (weak scaling,
no data-path collective: sequences are independent).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16|tf32x3]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'token positions/sec (decompose i-FST inference + Viterbi)'
UNIT = 'tokens/s'
METRIC_TRAIN = 'token positions/sec (decompose i-FST training step: fwd + CRF loss + bwd + grad all-reduce)'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='cfg2')
    ap.add_argument('--precision', default=os.environ.get('RE2NN_PRECISION', 'auto'))
    ap.add_argument('--farnn', type=int, default=0)
    ap.add_argument('--mode', default='infer', choices=['infer', 'train'],
                    help="infer = BASELINE configs[1] (headline); train = configs[2]: fwd+bwd+grad all-reduce, B=1024/GPU")
    ap.add_argument('--train-precision', default='auto', help='forward GEMMs of --mode train: auto|fp32|tf32x3|fp16x3')
    ap.add_argument('--cta-group', type=int, default=0, help='debug: force the tcgen05 CTA-group size (0 = cost model)')
    ap.add_argument('--resident', type=int, default=1, help='debug: 0 = one launch per step GEMM instead of the resident recurrence kernel')
    ap.add_argument('--backward-tc', type=int, default=1, help='debug: 0 = BPTT step GEMMs on fp32 CUDA cores')
    ap.add_argument('--infer-chunks', type=int, default=4, help='streams the graphed inference body forks into (1 = single stream)')
    ap.add_argument('--cpu-sample', type=int, default=256, help='sequences in the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


def flops_per_position(S, R, D, Cp, farnn):
    """F = 8SR + 4S^2 + 2DR + 2SC' (+ 8S^2 + 8SR for farnn=2)."""
    f = 8 * S * R + 4 * S * S + 2 * D * R + 2 * S * Cp
    if farnn == 2:
        f += 8 * S * S + 8 * S * R
    elif farnn == 1:
        f += 4 * S * S + 4 * S * R
    return f


def build_workload(cfg, seed, farnn):
    from re2nn_seq_b200 import synth
    c = dict(synth.CONFIGS[cfg])
    args = synth.make_args(farnn=farnn, use_crf=c.get('use_crf', 1), update_nonlinear=c.get('update_nonlinear', 'tanh'),
                           beta=c.get('beta', 0.1), sigmoid_exponent=5, bias_init=5.0)
    f = synth.make_decompose_factors(0, c['V'], c['S'], c['R'], c['C'], c['D'], dtype=np.float32)
    x, lens, lab = synth.make_batch(1000 + seed, c['B'], c['Lmax'], c['V'], c['C'], fixed_len=c.get('fixed_len', False))
    return c, args, f, x, lens, lab


def oracle_params_from_module(m):
    rename = {'embedding.weight': 'embedding', 'crf.transitions': 'crf_transitions',
              'priority_layer.priority_mat': 'priority_mat', 'priority_layer.priority_bias': 'priority_bias'}
    return {rename.get(k, k): v.detach().cpu().numpy() for k, v in m.state_dict().items()}


def cpu_leg(p, args, x, lens, lab, sample, repeats=1):
    """Time the oracle port (numpy restatement of the reference's own algorithm) on the host cores."""
    from oracle import re2nn_oracle as orc
    xs, ls, ys = x[:sample], lens[:sample], lab[:sample]
    Lm = int(ls.max())
    xs, ys = xs[:, :Lm], ys[:, :Lm]
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.decompose_forward_local(p, xs, ys, ls, args, 0, train=False)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return float(ls.sum()) / best, best


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML every 10 ms while the timed region runs."""
    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}

    def __init__(self, index):
        self.index, self.sm, self.bits, self.stop_flag, self.th, self.max = index, [], 0, False, None, None
        self.power = []

    def _loop(self):
        import pynvml as nv
        try:
            nv.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(',')[self.index])
                except Exception:
                    pass
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            while not self.stop_flag:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    try:
                        self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    except Exception:
                        pass
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                time.sleep(0.01)
        except Exception as e:   # noqa
            self.err = str(e)

    def start(self):
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=2)
        reasons = sorted(n for b, n in self.REASONS.items() if self.bits & b)
        return {'sm_mhz': float(np.median(self.sm)) if self.sm else None, 'sm_max_mhz': self.max,
                'samples': len(self.sm), 'power_w_max': max(self.power) if self.power else None, 'reasons': reasons}


def run_reference(a):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the Python
    reference itself cannot travel to the GPU box).  Rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    try:
        import torch
        torch.set_num_threads(os.cpu_count())
    except Exception:
        pass
    c, args, f, x, lens, lab = build_workload(a.config, 0, a.farnn)
    import torch
    # parameters exactly as the product module initialises them (host-only construction)
    import re2nn_seq_b200 as r
    from re2nn_seq_b200 import synth
    torch.manual_seed(0)
    m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f)
    with torch.no_grad():
        m.crf.transitions.copy_(torch.from_numpy(synth.crf_transitions(0, m.C)))
    p = oracle_params_from_module(m)
    sample = min(a.cpu_sample, c['B'])
    for _ in range(min(a.warmup, 1)):
        cpu_leg(p, args, x, lens, lab, sample)
    times, toks = [], float(lens[:sample].sum())
    for _ in range(a.steps):
        _, dt = cpu_leg(p, args, x, lens, lab, sample)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = toks / (ms / 1e3)
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
            'warmup': a.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'cfg2 decompose i-FST inference + Viterbi (V=12000,C=72,S=300,R=200,len<=35)',
                       'sample': 'first %d of B=%d sequences per step' % (sample, c['B'])},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
                             'sample': 'oracle/re2nn_oracle.py on the first %d sequences of the batch' % sample},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    a = parse()
    if a.impl == 'reference':
        return run_reference(a)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    import re2nn_seq_b200 as r
    from re2nn_seq_b200 import ops, synth
    if not a.resident:
        from re2nn_seq_b200 import _lib
        _lib.check(_lib.fn['re2nn_debug_set_resident'](0), 'resident')
    if not a.backward_tc:
        from re2nn_seq_b200 import _lib
        _lib.check(_lib.fn['re2nn_debug_set_backward_tc'](0), 'backward_tc')
    if a.cta_group:
        from re2nn_seq_b200 import _lib
        _lib.check(_lib.fn['re2nn_debug_set_tc_cta_group'](a.cta_group), 'cta_group')

    if a.mode == 'train' and a.config == 'cfg2':
        a.config = 'cfg3'
    c, args, f, x, lens, lab = build_workload(a.config, rank, a.farnn)
    torch.manual_seed(0)
    m = r.FARNN_S_D_W_I_S(args=args, o_idx=0, **f)
    with torch.no_grad():
        m.crf.transitions.copy_(torch.from_numpy(synth.crf_transitions(0, m.C)))
    m.infer_chunks = a.infer_chunks
    m.precision = a.precision      # 'auto': parity-grade tensor-core mode (split fp16 / 3xTF32), fp32 without tcgen05
    with torch.no_grad():
        prec = m._resolved_precision()
    m.precision = prec
    p_oracle = oracle_params_from_module(m) if rank == 0 else None
    m = m.cuda().eval()
    bucket = None
    if a.mode == 'train':
        from re2nn_seq_b200 import dist as rd
        m.train()
        m.train_precision = a.train_precision
        prec = m._resolved_precision()          # grad mode is on here: resolves the training precision
        bucket = rd.GradBucket(m)

    n_tok = int(lens.sum())
    xd, ld, yd = torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(lab).cuda()
    xh, lh, yh = torch.from_numpy(x).pin_memory(), torch.from_numpy(lens).pin_memory(), torch.from_numpy(lab).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')   # > 126 MB L2

    def train_step(xg, yg, lg):
        for q in bucket.params:
            q.grad = None
        loss, pred, _ = m.forward_local(xg, yg, lg, train=True)
        loss.backward()
        bucket.all_reduce()                # one all-reduce(SUM) of the flat gradient bucket (NCCL when N>1)
        return loss, pred, None

    def step_device():
        if a.mode == 'train':
            return train_step(xd, yd, ld)
        with torch.no_grad():
            return m.forward_local(xd, yd, ld, train=False)

    pred_host = torch.empty((n_tok,), dtype=torch.int64).pin_memory()

    def step_e2e():
        xg = xh.cuda(non_blocking=True)
        lg = lh.cuda(non_blocking=True)
        yg = yh.cuda(non_blocking=True)
        if a.mode == 'train':
            _, pred, _ = train_step(xg, yg, lg)
        else:
            with torch.no_grad():
                _, pred, _ = m.forward_local(xg, yg, lg, train=False)
        pred_host.copy_(pred, non_blocking=True)

    for _ in range(max(a.warmup, 3)):
        step_device()
    torch.cuda.synchronize()

    # parity self-check on the bench batch: decoded tags of the timed mode vs the fp32 CUDA-core path
    mismatch = None
    if a.mode == 'infer' and prec != 'fp32':
        with torch.no_grad():
            _, p_fast, _ = m.forward_local(xd, yd, ld, train=False)
            m.precision = 'fp32'
            _, p_ref, _ = m.forward_local(xd, yd, ld, train=False)
            m.precision = prec
        mismatch = int((p_fast != p_ref).sum().item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region 1: inputs resident in HBM -------------------------------------------------------
    sampler = ClockSampler(local)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    l0 = ops.launches()
    barrier()
    sampler.start()
    for i in range(a.steps):
        flush.zero_()                      # L2 flush between timed iterations (outside the event pair)
        ev[i][0].record()
        step_device()
        ev[i][1].record()
    barrier()
    launches_total = ops.launches() - l0
    launches = launches_total // a.steps
    total_ms = sum(s.elapsed_time(e) for s, e in ev)

    # ---- roofline pass: the same K steps again, every step-GEMM launch bracketed by CUDA events recorded
    # inside the library on the launching stream (kept out of region 1 so the events do not perturb `value`)
    m.use_cuda_graph = False               # events are recorded around direct launches, not inside a graph replay
    ops.profile_enable(True)
    ops.profile_read()
    barrier()
    evp = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for i in range(a.steps):
        flush.zero_()
        evp[i][0].record()
        step_device()
        evp[i][1].record()
    barrier()
    clocks = sampler.stop()
    prof_ms, prof_n = ops.profile_read()
    ops.profile_enable(False)
    m.use_cuda_graph = True
    prof_step_ms = sum(s.elapsed_time(e) for s, e in evp) / a.steps

    # ---- timed region 2: end to end from pinned host buffers --------------------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for i in range(a.steps):
        flush.zero_()
        ev2[i][0].record()
        step_e2e()
        ev2[i][1].record()
    barrier()
    total_ms2 = sum(s.elapsed_time(e) for s, e in ev2)

    t = torch.tensor([total_ms, total_ms2], dtype=torch.float64, device='cuda')
    tok = torch.tensor([float(n_tok)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tok, op=dist.ReduceOp.SUM)
    total_ms, total_ms2 = t.tolist()
    all_tok = tok.item()

    if rank == 0:
        ms_step = total_ms / a.steps
        value = all_tok / (ms_step / 1e3)
        e2e = all_tok / (total_ms2 / a.steps / 1e3)
        S, R, D, Cp = c['S'], c['R'], c['D'], m.C
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak_tf = peaks.get('bf16_tflops_sustained') or 1400.0
        peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside the step)' if peaks else \
            'fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)'
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'r01_traffic.json'))).get(prec)
        except Exception:
            pass
        if prof_n[3] > 0:
            # dominant kernel: the resident recurrence kernel (ALL steps of both directions in one launch).
            # Algorithmic flops per launch: every valid position, both directions, G1 (2*S*R) + G2 (2*(R+S)*S)
            # [+ gate GEMM 2*S*S*farnn]
            dom_name = 'resident recurrence: all steps, G1 + G2 + fused epilogues (%s)' % prec
            dom_cls = 3
            dom_flops = 2.0 * n_tok * (2.0 * S * R + 2.0 * (R + S) * S + 2.0 * S * S * a.farnn)
            try:
                traffic = json.load(open(os.path.join(ROOT, 'profiles', 'r01_traffic.json'))).get(prec + '_resident')
            except Exception:
                traffic = None
        else:
            # dominant kernel: GEMM2 (+ state epilogue), both directions per launch: 2 dirs * 2*B*S*(R+S) flops
            dom_name = 'step GEMM2 + state epilogue (%s)' % prec
            dom_cls = 2
            dom_flops = 2 * 2.0 * c['B'] * S * (R + S)
        g2_ms = prof_ms[dom_cls] / max(prof_n[dom_cls], 1)
        achieved = dom_flops / (g2_ms * 1e-3) / 1e12 if g2_ms > 0 else 0.0
        line = {
            'metric': METRIC if a.mode == 'infer' else METRIC_TRAIN, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': max(a.warmup, 3),
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'fp32': 'f32', 'bf16': 'bf16', 'tf32x3': 'tf32x3', 'fp16x3': 'fp16x3'}[prec], 'data': 'synthetic',
            'config': {'workload': '%s decompose i-FST %s (V=%d,C=%d,S=%d,R=%d,D=%d,len<=%d,B=%d per GPU)'
                                   % (a.config, 'training step: fwd + CRF loss + bwd + grad all-reduce + Viterbi'
                                      if a.mode == 'train' else 'inference + Viterbi', c['V'], c['C'], S, R, D,
                                      c['Lmax'], c['B']),
                       'mode': a.mode,
                       'farnn': a.farnn, 'precision': prec, 'tokens_per_step_per_gpu': n_tok,
                       'tag_mismatches_vs_fp32_path': mismatch,
                       'l2': 'flushed between timed iterations (256 MB write)',
                       'cuda_graph': a.mode == 'infer',
                       'whole_step_tflops': value * flops_per_position(S, R, D, Cp, a.farnn) / 1e12},
            'roofline': {'bound': 'tensor', 'kernel': dom_name, 'achieved': achieved,
                         'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf, 'traffic': traffic,
                         'peak_source': peak_src, 'launches_timed': prof_n[dom_cls], 'avg_launch_ms': g2_ms,
                         'kernel_share_of_step': (prof_ms[dom_cls] / a.steps) / prof_step_ms if prof_step_ms > 0 else None,
                         'profiled_ms_per_step': prof_step_ms,
                         'class_ms_per_step': {'gate': prof_ms[0] / a.steps, 'gemm1': prof_ms[1] / a.steps,
                                               'gemm2': prof_ms[2] / a.steps, 'resident': prof_ms[3] / a.steps}},
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': int(xh.numel() * 8 + lh.numel() * 8 + yh.numel() * 8),
                    'd2h_bytes_per_step': int(n_tok * 8)},
            'gpu_launches': int(launches_total), 'gpu_launches_per_step': int(launches), 'clocks': clocks,
        }
        if not a.no_cpu_baseline:
            try:
                torch.set_num_threads(os.cpu_count())
            except Exception:
                pass
            sample = min(a.cpu_sample, c['B'])
            v, dt = cpu_leg(p_oracle, args, x, lens, lab, sample)
            line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
                                    'sample': 'oracle/re2nn_oracle.py (numpy port of the reference algorithm) on the '
                                              'first %d of %d sequences, %.2f s' % (sample, c['B'], dt)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
