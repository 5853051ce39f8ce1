"""TEST / BASELINE INFRASTRUCTURE -- not product code.

Runs the UNMODIFIED reference classes (vendored by oracle/make_ref.py into oracle/_ref, sha256-pinned) on the host
cores: the CPU baseline and `bench.py --impl reference` arm, and a second parity checker next to the numpy
restatement.  Nothing here imports re2nn_seq_b200: the synthetic generators are loaded from their source file
(synth.py only needs numpy), so the product's shared object is never mapped into a reference process.

Reference entry points used (cited from /root/reference/src_seq):
  farnn/model_decompose_single.py:13-136  FARNN_S_D_W_I_S.__init__      :207-304 forward_local
  farnn/model_onehot.py:311-344           FARNN_S_O_I_S.__init__         :131-146 forward_local
  baselines/crf.py:102-195                CRF._viterbi_decode (through decode)
"""
import importlib.util
import os
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def load_synth():
    """re2nn_seq_b200/synth.py as a stand-alone module (no package import, no shared object)."""
    spec = importlib.util.spec_from_file_location('re2nn_synth_standalone', os.path.join(ROOT, 're2nn_seq_b200', 'synth.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _reference():
    if __package__:
        from . import make_ref
    else:
        import make_ref
    return make_ref.import_reference()


def build_decompose(synth, cfg, farnn=0, seed=0, module_seed=0):
    """The reference FARNN_S_D_W_I_S on the CPU with the bench's synthetic factors; constructed under the same
    torch seed as the product module, so both start from bit-identical parameters (tests/test_host_modules.py)."""
    import torch
    _reference()
    from src_seq.farnn.model_decompose_single import FARNN_S_D_W_I_S
    c = dict(synth.CONFIGS[cfg])
    args = synth.make_args(farnn=farnn, use_crf=c.get('use_crf', 1), update_nonlinear=c.get('update_nonlinear', 'tanh'),
                           beta=c.get('beta', 0.1), sigmoid_exponent=5, bias_init=5.0)
    f = synth.make_decompose_factors(seed, c['V'], c['S'], c['R'], c['C'], c['D'], dtype=np.float32)
    torch.manual_seed(module_seed)
    m = FARNN_S_D_W_I_S(V=f['V'], S1=f['S1'], S2=f['S2'], C_output_mat=f['C_output_mat'], wildcard_mat=f['wildcard_mat'],
                        wildcard_output_vector=f['wildcard_output_vector'], final_vector=f['final_vector'],
                        start_vector=f['start_vector'], pretrained_word_embed=f['pretrained_word_embed'],
                        priority_mat=None, args=args, o_idx=0, is_cuda=False)
    with torch.no_grad():
        m.crf.transitions.copy_(torch.from_numpy(synth.crf_transitions(0, m.C)))
    m.eval()
    return m, args, c


def build_onehot(synth, V, S, C, seed=0):
    """The reference FARNN_S_O_I_S on the CPU with a random 0/1 rule automaton (rand_constant = 0: exact counts)."""
    import torch
    _reference()
    from src_seq.farnn.model_onehot import FARNN_S_O_I_S
    args = synth.make_args(method='onehot', rand_constant=0.0)
    a = synth.make_onehot_automaton(seed, V, S, C, dtype=np.float32)
    torch.manual_seed(0)
    m = FARNN_S_O_I_S(a['language_tensor'], a['output_mat'], a['wildcard_mat'], a['output_wildcard_vector'],
                      a['final_vector'], a['start_vector'], None, args, 0, False)
    m.eval()
    return m, args, a


def run_forward_local(m, x, lab, lens, threads=None, repeats=1, train=False):
    """-> (pred int64[N], seconds of the best of `repeats`, per-repeat seconds).  Inference under no_grad, as
    val.py:11-28 calls it; inputs trimmed to the longest sequence of the sample like the reference's collate."""
    import torch
    if threads:
        torch.set_num_threads(int(threads))
    Lm = int(lens.max())
    xs = torch.from_numpy(np.ascontiguousarray(x[:, :Lm]))
    ys = torch.from_numpy(np.ascontiguousarray(lab[:, :Lm]))
    ls = torch.from_numpy(np.ascontiguousarray(lens))
    times, pred = [], None
    for _ in range(repeats):
        t0 = time.perf_counter()
        with torch.no_grad():
            _, pred, _ = m.forward_local(xs, ys, ls, train=train)
        times.append(time.perf_counter() - t0)
    return pred.cpu().numpy().astype(np.int64), min(times), times
