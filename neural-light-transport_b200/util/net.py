"""Channel schedule of the encoder/decoder (semantics of nlt/util/net.py:18-56, checked against the reference
function's own outputs in tests/golden/gen_feat_n_reference.npz)."""


def _floor_log2(x):
    """floor(log2(x)) for x >= 1, the way int(np.log2(x)) truncates."""
    x = float(x)
    e = 0
    while 2.0 ** (e + 1) <= x:
        e += 1
    return e


def gen_feat_n(min_n, max_n, final_n=3):
    """Channel counts of every block after the full-resolution 1x1 conv: powers of two rising from `min_n` to
    `max_n` (both ends included even when not powers of two), the same list mirrored, then halving down to --
    and ending with -- `final_n`.  (16, 256) -> [16, 32, 64, 128, 256, 256, 128, 64, 32, 16, 8, 4, 3]."""
    assert max_n >= min_n and max_n >= final_n, \
        "Max number of channels must be greater than or equal to the final number of channel"
    rising = [min_n] + [2 ** e for e in range(_floor_log2(min_n) + 1, _floor_log2(max_n) + 1) if 2 ** e != min_n]
    if rising[-1] != max_n:
        rising.append(max_n)
    schedule = rising + rising[::-1]
    # halve below the last mirrored entry while strictly above final_n's power of two ...
    schedule += [2 ** e for e in range(_floor_log2(schedule[-1]) - 1, _floor_log2(final_n), -1)]
    # ... and never end on something smaller than final_n before appending it
    while schedule and schedule[-1] < final_n:
        schedule.pop()
    return schedule + [final_n]
