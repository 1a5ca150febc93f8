"""The multiply-high division of xg_common.cuh (XgFastDiv), restated in Python and checked against
`//` — the arithmetic claim only (exact for every dividend below 2^31); the CUDA code itself is
exercised by every GPU parity test, whose index decompositions all go through it."""

import numpy as np


def fastdiv_make(d):
    if d < 2 or d >= 1 << 31:
        return d, 0, 0
    l = 0
    while (1 << l) < d:
        l += 1
    mul = ((1 << (31 + l)) + d - 1) // d
    assert mul < 1 << 32
    return d, mul, l - 1


def fastdiv_q(x, f):
    d, mul, sh = f
    return ((x * mul) >> 32) >> sh if mul else x


def test_fastdiv_small_divisors_edges():
    top = (1 << 31) - 1
    for d in list(range(1, 3000)) + [3599, 3600, 180000, 2 ** 16 - 1, 2 ** 16, 2 ** 16 + 1, 2 ** 30, 2 ** 31 - 1]:
        f = fastdiv_make(d)
        xs = {0, 1, d - 1, d, d + 1, top, top - 1, top // 2}
        for k in (1, 2, 3, 7, top // d, max(top // d - 1, 0)):
            xs.update({k * d - 1, k * d, k * d + 1})
        for x in xs:
            if 0 <= x <= top:
                assert fastdiv_q(x, f) == x // d, (x, d)


def test_fastdiv_random():
    rng = np.random.default_rng(0)
    ds = np.concatenate([rng.integers(1, 1 << 31, 3000), rng.integers(1, 1 << 12, 3000)])
    for d in ds.tolist():
        f = fastdiv_make(d)
        for x in rng.integers(0, 1 << 31, 64).tolist():
            assert fastdiv_q(x, f) == x // d, (x, d)
