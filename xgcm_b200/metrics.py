"""Enumeration of axis-set partitions used to compose metrics (reference ``xgcm/metrics.py``).

``get_metric`` needs, for axes ``(X, Y, Z)``, candidates such as ``({X,Y,Z},)``,
``({X,Y},{Z})``, ``({X},{Y},{Z})`` ... in the reference's preference order:
the whole set first, then splits with the largest "left" subset first.
"""

from __future__ import annotations

import itertools


def iterate_axis_combinations(items):
    """Yield tuples of frozensets that together cover ``items``."""
    whole = frozenset(items)
    yield (whole,)
    n = len(items)
    for n_left in range(n - 1, 0, -1):
        n_right = n - n_left
        sizes = range(min(n_right, n_left), 0, -1)
        for size, left in itertools.product(sizes, itertools.combinations(whole, n_left)):
            left = frozenset(left)
            rest = whole - left
            yield (left,) + tuple(frozenset(c) for c in itertools.combinations(rest, size))
