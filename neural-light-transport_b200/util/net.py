"""Channel schedule -- same semantics as nlt/util/net.py:18-56."""
import math


def _ilog2(x):
    return int(math.floor(math.log2(x) + 1e-12))


def gen_feat_n(min_n, max_n, final_n=3):
    """Numbers of channels across the network, excluding the first layer that
    produces an original-resolution feature map,
    e.g. `[8, 16, 32, 64, 64, 32, 16, 8, 4, 3]`."""
    assert max_n >= min_n and max_n >= final_n, \
        ("Max number of channels must be greater than or equal to the final "
         "number of channel")
    up = [1 << e for e in range(_ilog2(min_n) + 1, _ilog2(max_n) + 1)]
    if not up or up[0] != min_n:
        up.insert(0, min_n)
    if up[-1] != max_n:
        up.append(max_n)
    seq = up + up[::-1]
    e = _ilog2(seq[-1]) - 1
    while e > _ilog2(final_n):
        seq.append(1 << e)
        e -= 1
    while seq and seq[-1] < final_n:
        seq.pop()
    seq.append(final_n)
    return seq
