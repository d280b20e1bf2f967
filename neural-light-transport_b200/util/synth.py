"""Synthetic NLT batches: no NLT data exists offline.

Shapes/dtypes follow the dataset 11-tuple (nlt/datasets/nlt.py:107-184):
images are k/255-quantised (uint8 PNG / 255), the uv2cam warp is an identity
grid plus a smooth distortion rounded through float16 (data_gen/util.py:67-70)
with a 35 % background mask set to exactly 0 (data_gen/render.py:155).
"""
import math

import torch


def _q255(x):
    return torch.round(x * 255.0) / 255.0


def as_uint8(batch):
    """The same batch with every k/255-quantised image tensor as uint8 (what a dataset with `uint8_inputs` emits);
    the warp stays float32."""
    out = []
    for i, t in enumerate(batch):
        if torch.is_tensor(t) and i != 4:
            q = torch.round(t * 255.0)
            assert float((q / 255.0 - t).abs().max()) == 0.0
            u = q.to(torch.uint8)
            out.append(u.pin_memory() if t.is_pinned() else u)
        else:
            out.append(t)
    return tuple(out)


def make_batch(B, uv, im, seed=1234, device='cpu', c_extra=0, pin=False, k_obs=1):
    """Returns the 11-tuple (id, base, cvis, lvis, warp, rgb, rgb_camspc,
    nn_id, nn_base, nn_rgb, nn_rgb_camspc).  c_extra > 0 widens cvis with
    extra U[0,1) maps (cfg4's 64-channel query stack: tf.concat is
    channel-count agnostic, nlt/models/nlt.py:95)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g, dtype=torch.float32)
    base, rgb = _q255(r(B, uv, uv, 3)), _q255(r(B, uv, uv, 3))
    if k_obs > 1:   # K neighbours, k-major [K,B,H,W,3] (cfg3's 6-neighbour observed-light stack)
        nn_base, nn_rgb = _q255(r(k_obs, B, uv, uv, 3)), _q255(r(k_obs, B, uv, uv, 3))
    else:
        nn_base, nn_rgb = _q255(r(B, uv, uv, 3)), _q255(r(B, uv, uv, 3))
    cos = lambda: _q255((r(B, uv, uv, 1) * 1.3 - 0.3).clamp(0, 1))
    cvis, lvis = cos(), cos()
    if c_extra:
        cvis = torch.cat((cvis, _q255(r(B, uv, uv, c_extra))), dim=3)
    ys, xs = torch.meshgrid(torch.arange(im, dtype=torch.float32), torch.arange(im, dtype=torch.float32),
                            indexing='ij')
    u = (xs + 0.5) / im
    v = (ys + 0.5) / im
    ph = r(B, 4) * 2 * math.pi
    wx = u[None] + 0.03 * torch.sin(2 * math.pi * v[None] * 1.5 + ph[:, 0, None, None]) \
        * torch.cos(2 * math.pi * u[None] + ph[:, 1, None, None])
    wy = v[None] + 0.03 * torch.cos(2 * math.pi * u[None] * 1.5 + ph[:, 2, None, None]) \
        * torch.sin(2 * math.pi * v[None] + ph[:, 3, None, None])
    warp = torch.stack((wx, wy), dim=3).to(torch.float16).to(torch.float32)
    # blocky 35 % background (8x8 blocks), warp == 0 there
    nb = (im + 7) // 8
    bg = (r(B, nb, nb) < 0.35).repeat_interleave(8, 1).repeat_interleave(8, 2)[:, :im, :im]
    warp = warp * (~bg)[..., None]
    rgb_camspc = _q255(r(B, im, im, 3))
    nn_rgb_camspc = _q255(r(B, im, im, 3))
    ids = [b'synth_%06d_%03d' % (seed, i) for i in range(B)]
    nn_ids = [b'synth_nn_%06d_%03d' % (seed, i) for i in range(B)]
    tensors = [base, cvis, lvis, warp, rgb, rgb_camspc, nn_base, nn_rgb, nn_rgb_camspc]
    if pin and torch.cuda.is_available():
        tensors = [t.pin_memory() for t in tensors]
    if device != 'cpu':
        tensors = [t.to(device) for t in tensors]
    base, cvis, lvis, warp, rgb, rgb_camspc, nn_base, nn_rgb, nn_rgb_camspc = tensors
    return (ids, base, cvis, lvis, warp, rgb, rgb_camspc, nn_ids, nn_base, nn_rgb, nn_rgb_camspc)
