"""Multicast (2 x 2 cluster) step GEMM: correctness against torch and A/B timing against the CTA-pair kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from re2nn_seq_b200 import ops, _lib

torch.manual_seed(0)
for (M, N, K) in ((16384, 1024, 512), (20000, 512, 1536), (65536, 1024, 1536)):
    A = torch.randn(M, K, device='cuda') / K ** 0.5
    B = torch.randn(N, K, device='cuda')
    ref = (A.bfloat16().float() @ B.bfloat16().float().t())
    for mc in (1, 0):
        _lib.check(_lib.fn['re2nn_debug_set_tc_multicast'](mc), 'mc')
        C = ops.gemm_nt(A, B, 'bf16')
        torch.cuda.synchronize()
        err = ((C - ref).abs().max() / ref.abs().max()).item()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.gemm_nt(A, B, 'bf16')
        e1.record()
        torch.cuda.synchronize()
        print('M=%d N=%d K=%d multicast=%d: rel err %.2e, %.3f ms per call (incl. operand conversion)' % (
            M, N, K, mc, err, e0.elapsed_time(e1) / 5))
_lib.check(_lib.fn['re2nn_debug_set_tc_multicast'](0), 'mc')
