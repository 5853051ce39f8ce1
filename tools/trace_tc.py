"""Debug helper: per-phase cycle breakdown of the tcgen05 step GEMMs (run on the GPU box)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
from re2nn_seq_b200 import ops, _lib

names = ['alive', 'setup', 'mma_issue(from setup)', 'prefetch(from setup)', 'acc_ready(from setup)', 'epilogue', 'exit']
for (M, N, K, prec) in [(8192, 300, 500, 'bf16'), (8192, 200, 304, 'bf16'), (8192, 300, 500, 'tf32x3'), (131072, 512, 1024, 'bf16')]:
    A = torch.randn(M, K, device='cuda'); B = torch.randn(N, K, device='cuda')
    buf = torch.zeros(32 * 65536, dtype=torch.int64, device='cuda')
    ops.gemm_nt(A, B, prec)
    _lib.check(_lib.fn['re2nn_debug_set_tc_trace'](C.c_void_p(buf.data_ptr())), 'trace')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm_nt(A, B, prec); e1.record(); torch.cuda.synchronize()
    _lib.check(_lib.fn['re2nn_debug_set_tc_trace'](None), 'trace')
    t = buf.cpu().numpy().reshape(-1, 32)
    t = t[t[:, 0] != 0]
    d = np.stack([t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 2], t[:, 5] - t[:, 2],
                  t[:, 6] - t[:, 5], t[:, 7] - t[:, 0]], 1)
    print('M=%d N=%d K=%d %s: %d CTAs, wall %.1f us (incl. 2 convert kernels)' % (M, N, K, prec, len(t), e0.elapsed_time(e1) * 1e3))
    kb = t[:, 8:32] - t[:, 2:3]
    kb = np.where(t[:, 8:32] > 0, kb, 0)
    print('   k-block arrival (cycles after setup), median over CTAs:', [int(np.median(kb[:, i])) for i in range(24) if kb[:, i].max() > 0])
    for n, col in zip(names, d.T):
        print('   %-28s mean %8.0f  p50 %8.0f  max %8.0f cycles' % (n, col.mean(), np.median(col), col.max()))
