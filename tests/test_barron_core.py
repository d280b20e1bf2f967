"""CPU check of the arithmetic shared with the (experimental) CUDA Barron-loss kernels: the per-element functions
of neural-light-transport_b200/csrc/nlt_barron_core.h are compiled for the host (tests/barron_host_check.cpp, g++)
and compared with the reference-pinned oracle (oracle/barron_oracle.py, float64 + autograd): 1-D analysis and its
exact adjoint on every length 1..40, the whole loss + gradient on odd sizes, tiny sizes (multiple reflections),
with and without alpha weights."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import barron_oracle as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'neural-light-transport_b200', 'csrc')
LOG_Z = math.log(2.0 * math.e * 0.6019072301972346)        # log(2 e K_1(1)); the reference's spline agrees to 1e-10


@pytest.fixture(scope='module')
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp('barron') / 'barron_host_check.so')
    subprocess.check_call(['g++', '-O2', '-shared', '-fPIC', '-std=c++17', '-I', CSRC, '-o', so,
                           os.path.join(ROOT, 'tests', 'barron_host_check.cpp')])
    lib = C.CDLL(so)
    fp = C.POINTER(C.c_float)
    lib.barron_host.restype = C.c_int
    lib.barron_host.argtypes = [fp, fp, fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float,
                                fp, fp]
    lib.barron_host_analysis.argtypes = [fp, C.c_int32, fp, fp]
    lib.barron_host_adjoint.argtypes = [fp, fp, C.c_int32, fp]
    return lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def test_log_z_constant():
    assert abs(LOG_Z - B.log_partition_alpha1_closed_form()) <= 1e-12


@pytest.mark.parametrize('n', list(range(1, 41)))
def test_analysis_and_exact_adjoint_1d(host, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32)
    nl, nh = (n + 1) // 2, n // 2
    lo, hi = np.zeros(max(nl, 1), np.float32), np.zeros(max(nh, 1), np.float32)
    host.barron_host_analysis(_p(x), n, _p(lo), _p(hi))
    f_lo, f_hi = B.cdf97_analysis_filters()
    xt = torch.from_numpy(x.astype(np.float64)).reshape(1, 1, n).requires_grad_(True)
    want_lo = B._filter_decimate(xt, f_lo, 1, 0).reshape(-1)
    want_hi = B._filter_decimate(xt, f_hi, 1, 1).reshape(-1) if nh > 0 else torch.zeros(0, dtype=torch.float64)
    np.testing.assert_allclose(lo[:nl], want_lo.detach().numpy(), rtol=0, atol=2e-6)
    np.testing.assert_allclose(hi[:nh], want_hi.detach().numpy(), rtol=0, atol=2e-6)
    # adjoint: gradient of <g_lo, lo(x)> + <g_hi, hi(x)> with respect to x
    g_lo = rng.standard_normal(max(nl, 1)).astype(np.float32)
    g_hi = rng.standard_normal(max(nh, 1)).astype(np.float32)
    dx = np.zeros(n, np.float32)
    host.barron_host_adjoint(_p(g_lo), _p(g_hi), n, _p(dx))
    obj = (want_lo * torch.from_numpy(g_lo[:nl].astype(np.float64))).sum()
    if nh > 0:
        obj = obj + (want_hi * torch.from_numpy(g_hi[:nh].astype(np.float64))).sum()
    obj.backward()
    np.testing.assert_allclose(dx, xt.grad.reshape(-1).numpy(), rtol=0, atol=5e-6)


@pytest.mark.parametrize('Bn,H,W,levels,use_alpha', [(2, 32, 48, 5, False), (1, 83, 71, 5, True), (2, 33, 17, 5, True),
                                                     (1, 17, 40, 5, False), (3, 5, 7, 3, False), (1, 2, 2, 1, True),
                                                     (1, 64, 64, 5, False)])
def test_loss_and_gradient_match_the_pinned_oracle(host, Bn, H, W, levels, use_alpha):
    g = torch.Generator().manual_seed(H * 100 + W)
    gt = torch.rand(Bn, H, W, 3, generator=g, dtype=torch.float64)
    pred = (gt + 0.03 * torch.randn(Bn, H, W, 3, generator=g, dtype=torch.float64)).requires_grad_(True)
    alpha = torch.rand(Bn, H, W, 1, generator=g, dtype=torch.float64) if use_alpha else None
    loss_scale = 0.25
    # oracle with the same number of levels
    res = (gt - pred) * (alpha if use_alpha else 1.0)
    x = B.rgb_to_syuv(res).permute(0, 3, 1, 2).reshape(Bn * 3, H, W)
    flat = B.wavelet_flatten(B.wavelet_construct(x, levels)).reshape(Bn, 3, H, W)
    nll = B.charbonnier(flat, 0.01) + math.log(0.01) + LOG_Z
    per = nll.mean(dim=(1, 2, 3))
    (per.sum() * loss_scale).backward()
    f32 = lambda t: np.ascontiguousarray(t.detach().numpy().astype(np.float32))
    p32, g32, a32 = f32(pred), f32(gt), (f32(alpha) if use_alpha else None)
    loss = np.zeros(Bn, np.float32)
    d_pred = np.zeros_like(p32)
    rc = host.barron_host(_p(p32), _p(g32), _p(a32), Bn, H, W, levels, 0.01, LOG_Z, loss_scale, _p(loss), _p(d_pred))
    assert rc == 0
    np.testing.assert_allclose(loss, per.detach().numpy(), rtol=2e-5, atol=2e-5)
    want = pred.grad.numpy()
    err = np.linalg.norm(d_pred.astype(np.float64) - want) / np.linalg.norm(want)
    assert err <= 2e-5, err
