"""Small forward + backward through every kernel family, for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import re2nn_seq_b200 as r
from test_gpu_parity import _random_decompose

def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()

for farnn in (0, 2):
    m, args, x, lens, lab = _random_decompose(3, 200, 80, 48, 10, 30, 300, 7, farnn=farnn, use_crf=1,
                                              update_nonlinear='tanh', beta=0.1)
    m.use_cuda_graph = False
    for prec in ('fp16x3', 'bf16', 'fp32'):
        m.precision = prec
        with torch.no_grad():
            m.forward_local(t(x), t(lab), t(lens), train=False)
    m.train_precision = 'auto'
    loss, _, _ = m.forward_local(t(x), t(lab), t(lens), train=True)
    loss.backward()
    torch.cuda.synchronize()
    print('farnn', farnn, 'ok', float(loss))
