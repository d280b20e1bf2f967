"""Pins the oracle (oracle/nlt_oracle.py): hand-computed known answers, an
independent direct-loop numpy restatement, adjoint identities, and the golden
vectors under tests/golden/.  CPU only."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import nlt_oracle as O
from oracle import np_ref as R

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
KS = [(1, 1), (2, 1), (2, 2), (3, 1), (3, 2), (4, 2), (1, 2), (3, 3)]


def t64(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64)


def test_gen_feat_n_docstring_and_golden():
    # the one example the reference carries (nlt/util/net.py:23)
    assert O.gen_feat_n(8, 64) == [8, 16, 32, 64, 64, 32, 16, 8, 4, 3]
    assert O.gen_feat_n(16, 256) == [16, 32, 64, 128, 256, 256, 128, 64, 32, 16, 8, 4, 3]
    gold = np.load(os.path.join(GOLD, 'gen_feat_n_reference.npz'))
    for key in gold.files:
        a, b, f = (int(v) for v in key.split('_'))
        assert O.gen_feat_n(a, b, f) == gold[key].tolist(), key


def test_same_pad_table():
    # k=2,s=1 pads (0,1); k=2,s=2 none; k=3,s=2 on even n pads (0,1)
    assert O.same_pad(8, 2, 1) == (0, 1)
    assert O.same_pad(8, 2, 2) == (0, 0)
    assert O.same_pad(8, 3, 2) == (0, 1)
    assert O.same_pad(8, 3, 1) == (1, 1)
    assert O.same_pad(7, 3, 2) == (1, 1)
    assert O.same_pad(8, 4, 2) == (1, 1)


def test_deconv_known_answer_1d():
    # hand tap check: k2 s1 deconv of [1,2,3,4] with w=[2,3] -> [2,7,12,17]
    x = t64([1, 2, 3, 4]).view(1, 1, 4, 1)
    w = t64([2, 3]).view(1, 2, 1, 1)
    y = O.conv2d_transpose_same(x, w, None, 1)
    assert y.flatten().tolist() == [2, 7, 12, 17]
    # k2 s2: pure depth-to-space: out[2i+d] = w[d] x[i]
    y = O.conv2d_transpose_same(x, w, None, 2)   # H: 1 -> 2 (kh=1 < s: second row is zero)
    assert y[0, 0, :, 0].tolist() == [2, 3, 4, 6, 6, 9, 8, 12]
    assert y[0, 1, :, 0].abs().sum() == 0
    # forward k2 s1 SAME pads at the END: y[i] = 2 x[i] + 3 x[i+1]
    y = O.conv2d_same(x, w, None, 1)
    assert y.flatten().tolist() == [8, 13, 18, 8]
    y = O.conv2d_same(x, w, None, 2)
    assert y.flatten().tolist() == [8, 18]


@pytest.mark.parametrize('k,s', KS)
@pytest.mark.parametrize('hw', [(8, 8), (7, 5), (4, 6)])
def test_conv_matches_direct_loops(k, s, hw):
    rng = np.random.default_rng(k * 10 + s)
    x = rng.standard_normal((2, hw[0], hw[1], 3))
    w = rng.standard_normal((k, k, 3, 4))
    b = rng.standard_normal(4)
    got = O.conv2d_same(t64(x), t64(w), t64(b), s).numpy()
    np.testing.assert_allclose(got, R.conv2d_same(x, w, b, s), rtol=1e-12, atol=1e-12)
    wt = rng.standard_normal((k, k, 4, 3))
    got = O.conv2d_transpose_same(t64(x), t64(wt), t64(b), s).numpy()
    np.testing.assert_allclose(got, R.conv2d_transpose_same(x, wt, b, s), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('k,s', KS)
def test_adjoint_identity(k, s):
    # <conv(x), y> == <x, deconv(y)> when deconv shares the conv's kernel buffer
    rng = np.random.default_rng(100 + k * 10 + s)
    n = 4 * s
    x = t64(rng.standard_normal((1, n, n, 3)))
    w = t64(rng.standard_normal((k, k, 3, 5)))
    cx = O.conv2d_same(x, w, None, s)
    y = t64(rng.standard_normal(tuple(cx.shape)))
    dy = O.conv2d_transpose_same(y, w, None, s)   # (kh,kw,Co=3,Ci=5) read as transpose kernel
    assert dy.shape == x.shape
    assert abs(float((cx * y).sum() - (x * dy).sum())) < 1e-9


def test_resampler_known_answers():
    d = np.arange(12, dtype=np.float64).reshape(1, 3, 4, 1)   # d[y,x] = 4y + x
    pts = [(0, 0), (1, 2), (0.5, 0.5), (2.25, 1.5), (-0.5, 0), (3.5, 2), (3, 2), (-1, 0), (4, 1), (0, 3), (1.5, -0.25)]
    warp = np.array(pts, dtype=np.float64).reshape(1, 1, len(pts), 2)
    got = O.resampler(t64(d), t64(warp)).numpy().reshape(-1)
    want = [0, 9, 2.5, 8.25, 0, 5.5, 11, 0, 0, 0, 1.125]
    np.testing.assert_allclose(got, want, atol=1e-12)
    np.testing.assert_allclose(R.resampler(d, warp).reshape(-1), want, atol=1e-12)
    rng = np.random.default_rng(3)
    data = rng.standard_normal((2, 5, 6, 3))
    w2 = rng.uniform(-2, 8, size=(2, 4, 7, 2))
    np.testing.assert_allclose(O.resampler(t64(data), t64(w2)).numpy(), R.resampler(data, w2), atol=1e-12)


@pytest.mark.parametrize('shape,new', [((4, 4), (8, 8)), ((8, 8), (4, 4)), ((5, 7), (9, 3)), ((6, 6), (6, 6))])
def test_resize_matches_loops_and_torch(shape, new):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, shape[0], shape[1], 3))
    got = O.resize_bilinear(t64(x), *new).numpy()
    np.testing.assert_allclose(got, R.resize_bilinear(x, *new), atol=1e-12)
    ref = torch.nn.functional.interpolate(t64(x).permute(0, 3, 1, 2), size=new, mode='bilinear',
                                          align_corners=False).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(got, ref, atol=1e-12)


def test_network_plan_matches_survey_table():
    plan, contr = O.network_plan(16, 256)
    assert [p[0] for p in plan] == ['conv1x1'] + ['down'] * 6 + ['up'] * 6 + ['conv1x1']
    assert contr == [True] * 7 + [False] * 7
    cfg = dict(depth0=16, depth=256, kernel=2, stride=2, use_obs=True)
    ch = O.model_channels(cfg, 5, 3)
    cin_q = [c[0][1] for c in ch['query']]
    assert cin_q == [5, 32, 32, 64, 128, 256, 512, 1024, 640, 320, 160, 80, 40, 36]
    p = O.init_params(cfg)
    assert sum(v.numel() for v in p.values()) == 3368071
    cfg['depth'] = 1024
    assert sum(v.numel() for v in O.init_params(cfg).values()) == 53841543


def test_amsgrad_matches_torch_adam():
    torch.manual_seed(0)
    p0 = torch.randn(50, dtype=torch.float64)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-2, betas=(0.9, 0.999), eps=1e-7, amsgrad=True)
    p, m, v, vh = p0.clone(), torch.zeros(50, dtype=torch.float64), torch.zeros(50, dtype=torch.float64), \
        torch.zeros(50, dtype=torch.float64)
    for t in range(1, 6):
        g = torch.randn(50, dtype=torch.float64)
        p_ref.grad = g.clone()
        opt.step()
        p, m, v, vh = O.amsgrad_step(p, g, m, v, vh, t, 1e-2)
    # torch puts eps outside the bias correction (sqrt(vhat/bc2)+eps) whereas
    # TF/Keras folds the correction into lr_t: equal up to O(eps)
    np.testing.assert_allclose(p.numpy(), p_ref.detach().numpy(), rtol=0, atol=1e-6)


def test_golden_model_vectors():
    """The committed fp64 oracle outputs (tests/golden/make_golden.py)."""
    from tests.golden import make_golden as G
    gold = np.load(os.path.join(GOLD, 'model_h64.npz'))
    out = G.compute(dtype=torch.float64)
    for k in gold.files:
        np.testing.assert_allclose(out[k], gold[k], rtol=1e-9, atol=1e-11, err_msg=k)
