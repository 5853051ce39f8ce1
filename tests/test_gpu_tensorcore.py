"""GPU: the tcgen05 mainloop, first in isolation (C = A @ B^T against a float64 product), then inside
the decompose recurrence (bf16 with a stated bound, tf32x3 at the fp32 parity tolerance)."""
import numpy as np
import pytest
import torch

from helpers import oracle_params, rel_err
from oracle import re2nn_oracle as orc

pytestmark = pytest.mark.gpu

SHAPES = [(128, 64, 64), (128, 64, 128), (200, 300, 300), (256, 32, 64), (700, 48, 96), (333, 77, 130), (4096, 200, 300), (1000, 520, 1030),
          (8192, 304, 500), (40000, 512, 264)]


def _need_tc():
    from re2nn_seq_b200 import ops
    if not ops.has_tcgen05():
        pytest.skip('no tcgen05 device')


@pytest.fixture(params=[1, 2], ids=['cta1', 'ctapair'])
def cta_group(request):
    """Force single-CTA tiles (128 x bn) or CTA pairs (cta_group::2, 256 x bn) for the tcgen05 GEMMs of the test."""
    from re2nn_seq_b200 import _lib
    _lib.check(_lib.fn['re2nn_debug_set_tc_cta_group'](request.param), 'cta_group')
    yield request.param
    _lib.check(_lib.fn['re2nn_debug_set_tc_cta_group'](0), 'cta_group')


@pytest.mark.parametrize('prec,tol', [('fp32', 2e-6), ('tf32x3', 2e-5), ('fp16x3', 2e-5), ('bf16', 1.5e-2)])
@pytest.mark.parametrize('M,N,K', SHAPES)
def test_gemm_nt(prec, tol, M, N, K, cta_group):
    from re2nn_seq_b200 import ops
    if prec != 'fp32':
        _need_tc()
    elif cta_group == 2:
        pytest.skip('fp32 runs the SIMT mainloop')
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((M, K), generator=g).cuda()
    B = torch.randn((N, K), generator=g).cuda()
    C = ops.gemm_nt(A, B, prec)
    ref = A.double() @ B.double().t()
    err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
    assert err < tol, 'rel err %.3e' % err


def test_gemm_nt_exact_integers(cta_group):
    """Small-integer operands are exact in bf16 and tf32: any layout / descriptor error shows up as a wrong integer."""
    from re2nn_seq_b200 import ops
    _need_tc()
    g = torch.Generator().manual_seed(1)
    A = torch.randint(-3, 4, (300, 200), generator=g).float().cuda()
    B = torch.randint(-3, 4, (150, 200), generator=g).float().cuda()
    ref = (A.double() @ B.double().t()).float()
    for prec in ('bf16', 'tf32x3', 'fp16x3'):
        C = ops.gemm_nt(A, B, prec)
        assert torch.equal(C, ref), prec


def _model(farnn, crf, seed=5, B=160, S=300, R=200):
    from test_gpu_parity import _random_decompose
    return _random_decompose(seed, 2000, S, R, 72, 100, B, 35, farnn=farnn, use_crf=crf, update_nonlinear='tanh',
                             beta=0.1)


def _truth(m, args, x, lens):
    class _Z(dict):
        files = property(lambda self: list(self.keys()))
    z = _Z({'p.' + k: v.detach().cpu().numpy() for k, v in m.state_dict().items()})
    p64 = oracle_params(z, np.float64)
    sc, _, _ = orc.decompose_scores(p64, x, lens, args)
    return sc, z


@pytest.mark.parametrize('mode', ['tf32x3', 'fp16x3'])
@pytest.mark.parametrize('farnn', [0, 2])
def test_recurrence_tf32x3_matches_fp32_tolerance(farnn, mode, cta_group):
    _need_tc()
    m, args, x, lens, lab = _model(farnn, 1)
    truth, z = _truth(m, args, x, lens)
    mask = orc.length_mask(lens, 35)
    xt, lt, yt = (torch.from_numpy(a).cuda() for a in (x, lens, lab))
    with torch.no_grad():
        m.precision = 'fp32'
        s32 = m.forward_scores(xt, lt).cpu().numpy()
        _, p32, _ = m.forward_local(xt, yt, lt, train=False)
        m.precision = mode
        s3 = m.forward_scores(xt, lt).cpu().numpy()
        _, p3, _ = m.forward_local(xt, yt, lt, train=False)
    assert rel_err(s32[mask], truth[mask]) < 1e-5
    assert rel_err(s3[mask], truth[mask]) < 1e-5                # tensor cores at the fp32 parity bar
    assert torch.equal(p32, p3)                                  # decoded tags identical to the fp32 path


@pytest.mark.parametrize('farnn', [0, 2])
def test_recurrence_bf16_stated_bound(farnn):
    """bf16 factors: stated bound 3e-2 relative on scores (max-norm); tag agreement with the fp32 path reported and >= 97 %."""
    _need_tc()
    m, args, x, lens, lab = _model(farnn, 1)
    truth, z = _truth(m, args, x, lens)
    mask = orc.length_mask(lens, 35)
    xt, lt, yt = (torch.from_numpy(a).cuda() for a in (x, lens, lab))
    with torch.no_grad():
        m.precision = 'fp32'
        _, p32, _ = m.forward_local(xt, yt, lt, train=False)
        m.precision = 'bf16'
        sb = m.forward_scores(xt, lt).cpu().numpy()
        _, pb, _ = m.forward_local(xt, yt, lt, train=False)
    err = rel_err(sb[mask], truth[mask])
    agree = (p32 == pb).float().mean().item()
    print('bf16 rel err %.3e, tag agreement %.5f' % (err, agree))
    assert err < 3e-2
    assert agree >= 0.97
