"""Independent direct-loop numpy restatement of the four external ops on the
NLT hot path.  TEST INFRASTRUCTURE ONLY (small sizes; pure-Python loops).

Purpose: pin oracle/nlt_oracle.py (which leans on torch conv primitives)
against a second restatement written from the op *definitions*:

* SAME conv       : y[o] = sum_d w[d] x[o*s + d - pad_before]        (TF 'SAME')
* SAME conv^T     : defined as the input-gradient of the SAME conv whose input
                    has size n*s, i.e. out[j] = sum_{i,d: i*s+d-pad=j} w[d] in[i]
* resampler       : tfa.image.resampler 0.10 (bilinear, zero outside)
* resize          : tf.image.resize bilinear, half-pixel centres

Reference call sites: nlt/networks/elements.py:26-39, nlt/models/nlt.py:112-114,
nlt/util/img.py:113-116.
"""
import numpy as np


def same_pad_before(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2


def conv2d_same(x, w, b, s):
    """x [B,H,W,Ci], w [kh,kw,Ci,Co], b [Co]."""
    B, H, W, Ci = x.shape
    kh, kw, _, Co = w.shape
    Ho, Wo = -(-H // s), -(-W // s)
    pt, pl = same_pad_before(H, kh, s), same_pad_before(W, kw, s)
    y = np.zeros((B, Ho, Wo, Co), dtype=np.float64)
    for oy in range(Ho):
        for ox in range(Wo):
            for dy in range(kh):
                for dx in range(kw):
                    iy, ix = oy * s + dy - pt, ox * s + dx - pl
                    if 0 <= iy < H and 0 <= ix < W:
                        y[:, oy, ox, :] += x[:, iy, ix, :] @ w[dy, dx]
    return y + (0 if b is None else b)


def conv2d_transpose_same(x, w, b, s):
    """x [B,h,w,Ci], w [kh,kw,Co,Ci] (Keras Conv2DTranspose layout)."""
    B, h, wd, Ci = x.shape
    kh, kw, Co, _ = w.shape
    H, W = h * s, wd * s
    pt, pl = same_pad_before(H, kh, s), same_pad_before(W, kw, s)
    y = np.zeros((B, H, W, Co), dtype=np.float64)
    for iy in range(h):
        for ix in range(wd):
            for dy in range(kh):
                for dx in range(kw):
                    oy, ox = iy * s + dy - pt, ix * s + dx - pl
                    if 0 <= oy < H and 0 <= ox < W:
                        y[:, oy, ox, :] += x[:, iy, ix, :] @ w[dy, dx].T
    return y + (0 if b is None else b)


def resampler(data, warp):
    B, H, W, C = data.shape
    _, h, w, _ = warp.shape
    out = np.zeros((B, h, w, C), dtype=np.float64)

    def get(b, ix, iy):
        if 0 <= ix <= W - 1 and 0 <= iy <= H - 1:
            return data[b, int(iy), int(ix)]
        return np.zeros(C)

    for b in range(B):
        for r in range(h):
            for c in range(w):
                x, y = float(warp[b, r, c, 0]), float(warp[b, r, c, 1])
                if x > -1 and y > -1 and x < W and y < H:
                    fx, fy = np.floor(x), np.floor(y)
                    cx, cy = fx + 1, fy + 1
                    dx, dy = cx - x, cy - y
                    out[b, r, c] = dx * dy * get(b, fx, fy) \
                        + (1 - dx) * (1 - dy) * get(b, cx, cy) \
                        + dx * (1 - dy) * get(b, fx, cy) \
                        + (1 - dx) * dy * get(b, cx, fy)
    return out


def resize_bilinear(x, new_h, new_w):
    B, H, W, C = x.shape
    out = np.zeros((B, new_h, new_w, C), dtype=np.float64)
    for oy in range(new_h):
        sy = (oy + 0.5) * (H / new_h) - 0.5
        y0 = max(int(np.floor(sy)), 0)
        y1 = min(int(np.ceil(sy)), H - 1)
        ly = sy - np.floor(sy)
        for ox in range(new_w):
            sx = (ox + 0.5) * (W / new_w) - 0.5
            x0 = max(int(np.floor(sx)), 0)
            x1 = min(int(np.ceil(sx)), W - 1)
            lx = sx - np.floor(sx)
            top = x[:, y0, x0] * (1 - lx) + x[:, y0, x1] * lx
            bot = x[:, y1, x0] * (1 - lx) + x[:, y1, x1] * lx
            out[:, oy, ox] = top * (1 - ly) + bot * ly
    return out
