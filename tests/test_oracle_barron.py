"""Pins oracle/barron_oracle.py (the "barron" training-loss term) against the
reference's OWN fixtures -- tests/golden/barron_reference_fixtures.npz, built by tests/golden/make_barron_golden.py
from third_party/robust_loss/data/{wavelet_golden.mat, partition_spline.npz} -- and against the golden index
vectors of third_party/robust_loss/wavelet_test.py:89-121.  CPU only."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import barron_oracle as B

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'barron_reference_fixtures.npz')


@pytest.fixture(scope='module')
def gold():
    return np.load(GOLD)


def test_reflecting_pad_golden_vectors():
    # wavelet_test.py:89-105  (n = 8, 17 below, 13 above)
    n = 8
    want = np.concatenate((np.arange(3, 0, -1), np.arange(n), np.arange(n - 2, 0, -1), np.arange(n),
                           np.arange(n - 2, 0, -1), np.arange(7)))
    np.testing.assert_array_equal(B.reflect_indices(n, 17, 13).numpy(), want)
    # wavelet_test.py:107-121 (n = 11, 15 below, 7 above)
    n = 11
    want = np.concatenate((np.arange(5, n), np.arange(n - 2, 0, -1), np.arange(n), np.arange(n - 2, 2, -1)))
    np.testing.assert_array_equal(B.reflect_indices(n, 15, 7).numpy(), want)
    # reflect([A, B, C, D], 2, 2) = [C, B, A, B, C, D, C, B]  (docstring example, wavelet.py:108-110)
    np.testing.assert_array_equal(B.reflect_indices(4, 2, 2).numpy(), [2, 1, 0, 1, 2, 3, 2, 1])
    np.testing.assert_array_equal(B.reflect_indices(1, 3, 2).numpy(), [0] * 6)


def test_cdf97_pyramid_matches_the_reference_golden(gold):
    """wavelet_test.py:168-172 (testConstructMatchesGoldenData), same tolerance 1e-5 -- in float64 the match to
    the independently produced pyramid is limited by the 12-digit filter taps."""
    im = torch.from_numpy(gold['image'])
    levels = int(gold['num_levels'])
    assert levels == 5 and tuple(im.shape) == (3, 83, 71)
    pyr = B.wavelet_construct(im, levels)
    assert len(pyr) == levels + 1
    worst = 0.0
    for d in range(levels):
        for k in range(3):
            want = gold['band_%d_%d' % (d, k)]
            assert tuple(pyr[d][k].shape) == want.shape, (d, k)
            worst = max(worst, float(np.abs(pyr[d][k].numpy() - want).max()))
    worst = max(worst, float(np.abs(pyr[-1].numpy() - gold['residual']).max()))
    assert worst <= 1e-5, worst
    # and in float32, the precision the reference runs it in
    pyr32 = B.wavelet_construct(im.float(), levels)
    err32 = max(float(np.abs(pyr32[d][k].double().numpy() - gold['band_%d_%d' % (d, k)]).max())
                for d in range(levels) for k in range(3))
    assert err32 <= 1e-5, err32


def test_filters_and_volume_preservation():
    lo, hi = B.cdf97_analysis_filters()
    assert len(lo) == 9 and len(hi) == 7
    # wavelet_test.py:123-130: the separable lowpass doubles the magnitude
    assert abs(float((lo[:, None] * lo[None, :]).sum()) - 2.0) <= 1e-10
    # wavelet_test.py:132-146: flatten(construct(.)) has a unit Jacobian determinant on a power-of-two size
    fun = lambda z: B.wavelet_flatten(B.wavelet_construct(z.reshape(1, 4, 4), 2)).reshape(-1)
    jac = torch.autograd.functional.jacobian(fun, torch.rand(16, dtype=torch.float64))
    assert abs(float(torch.linalg.det(jac)) - 1.0) <= 1e-5
    # util_test.py:132-139: the scaled YUV transform is volume preserving
    m = torch.tensor(B._YUV_FROM_RGB, dtype=torch.float64) * B.SYUV_SCALE
    assert abs(float(torch.linalg.det(m)) - 1.0) <= 1e-5
    x = torch.rand(5, 3, dtype=torch.float64)
    torch.testing.assert_close(B.rgb_to_syuv(x), x @ m)


def test_flatten_layout_on_a_tiny_pyramid():
    r = torch.full((1, 1, 1), 0.0)
    b = [torch.full((1, 1, 1), float(v)) for v in (1, 2, 3)]       # level 1: band0, band1, band2
    c = [torch.full((1, 2, 2), float(v)) for v in (4, 5, 6)]       # level 0
    flat = B.wavelet_flatten(((c[0], c[1], c[2]), (b[0], b[1], b[2]), r))
    want = torch.tensor([[[0, 2, 5, 5], [3, 1, 5, 5], [6, 6, 4, 4], [6, 6, 4, 4]]], dtype=torch.float32)
    torch.testing.assert_close(flat, want)


def test_log_partition_at_alpha_one(gold):
    """The spline of partition_spline.npz is documented as accurate to 1e-6 (distribution.py:150-179)."""
    log_z = B.log_partition(1.0, int(gold['spline_x_scale']), int(gold['spline_knot_lo']), gold['spline_values'],
                            gold['spline_tangents'])
    assert abs(B.partition_curve(1.0) - 1.2) <= 1e-12            # docstring pair (1, ~1.2), distribution.py:97-98
    assert abs(log_z - B.log_partition_alpha1_closed_form()) <= 1e-6
    assert abs(log_z - 1.1855) <= 1e-3


def test_charbonnier_is_the_general_loss_at_alpha_one():
    x = torch.linspace(-0.2, 0.2, 41, dtype=torch.float64)
    scale, alpha = 0.01, 1.0
    beta = abs(alpha - 2.0)
    general = (beta / alpha) * (((x / scale) ** 2 / beta + 1.0) ** (0.5 * alpha) - 1.0)     # general.py:110-113
    torch.testing.assert_close(B.charbonnier(x, scale), general)
    assert float(B.charbonnier(torch.zeros(1), scale)) == 0.0


def test_barron_loss_shapes_reduction_and_gradient(gold):
    log_z = B.log_partition(1.0, int(gold['spline_x_scale']), int(gold['spline_knot_lo']), gold['spline_values'],
                            gold['spline_tangents'])
    g = torch.Generator().manual_seed(0)
    gt = torch.rand(2, 32, 48, 3, generator=g, dtype=torch.float64)
    pred = (gt + 0.05 * torch.randn(2, 32, 48, 3, generator=g, dtype=torch.float64)).requires_grad_(True)
    per = B.barron_loss(gt, pred, log_z, keep_batch=True)
    assert per.shape == (2,)
    torch.testing.assert_close(B.barron_loss(gt, pred, log_z), per.mean())
    # identical images: every coefficient is 0 -> the loss is the constant log(scale) + log Z(1)
    same = B.barron_loss(gt, gt, log_z)
    assert abs(float(same) - (math.log(0.01) + log_z)) <= 1e-12
    # weights = alpha blend of both images (nlt/losses.py:109-111)
    w = torch.rand(2, 32, 48, 1, generator=g, dtype=torch.float64)
    torch.testing.assert_close(B.barron_loss(gt, pred, log_z, weights=w),
                               B.barron_loss(gt * w, pred * w, log_z))
    per.sum().backward()
    assert pred.grad is not None and torch.isfinite(pred.grad).all() and float(pred.grad.abs().max()) > 0
    with pytest.raises(ValueError):
        B.image_nll(torch.zeros(1, 8, 8, 3, dtype=torch.float64), log_z)     # 8x8 cannot hold 5 levels
