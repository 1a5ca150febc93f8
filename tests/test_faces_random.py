"""Randomised face topologies (CPU, mock backend): the stride derivation of xgcm_b200/padding.py
against the oracle's literal restatement of the reference algorithm, on link sets, dim orders and
width combinations the hand-written cases do not reach.  Deterministic seeds."""

import itertools

import numpy as np
import pytest
import torch

import xgcm_b200 as xg
from _mock_backend import install
from oracle.faces import pad_face_connections
from xgcm_b200.padding import pad

N = 5
AXES = {"X": ("x", "xl"), "Y": ("y", "yl")}
COORDS = {"X": {"center": "x", "left": "xl"}, "Y": {"center": "y", "left": "yl"}}


@pytest.fixture(autouse=True)
def mock_backend(monkeypatch):
    if torch.cuda.is_available():
        pytest.skip("real kernels available: covered by the gpu-marked modules")
    install(monkeypatch)


def _random_links(rng, nface):
    """Pair up random free edges (face, axis, side); a link between equal sides is reversed
    (grid.py:366-371: a reversed link arrives at the same side of the neighbour)."""
    edges = [(f, ax, side) for f in range(nface) for ax in "XY" for side in (0, 1)]
    rng.shuffle(edges)
    links = {f: {} for f in range(nface)}
    n_pairs = int(rng.integers(1, len(edges) // 2 + 1))
    for k in range(n_pairs):
        (fa, axa, sa), (fb, axb, sb) = edges[2 * k], edges[2 * k + 1]
        if fa == fb and axa == axb:
            continue  # an axis of one face wrapped onto itself: both directions of the same slot pair
        rev = sa == sb
        for (f, ax, s), (g, bx) in (((fa, axa, sa), (fb, axb)), ((fb, axb, sb), (fa, axa))):
            pair = list(links[f].get(ax, (None, None)))
            pair[s] = (g, bx, rev)
            links[f][ax] = tuple(pair)
    return {f: v for f, v in links.items()}


def _ds(nface):
    return xg.Dataset(coords={"face": np.arange(nface), "t": np.arange(2.0), "x": np.arange(N) + 0.0, "xl": np.arange(N) - 0.5,
                              "y": np.arange(N) + 0.0, "yl": np.arange(N) - 0.5})


WIDTHS = [{"X": (1, 0)}, {"Y": (0, 1)}, {"X": (2, 1)}, {"X": (1, 1), "Y": (1, 1)}, {"X": (0, 2), "Y": (1, 0)},
          {"X": (2, 2), "Y": (1, 2)}]


@pytest.mark.parametrize("seed", range(24))
def test_random_topology_scalar_and_vector_pad(seed):
    rng = np.random.default_rng(seed)
    nface = int(rng.integers(2, 5))
    links = _random_links(rng, nface)
    fc = {"face": links}
    grid = xg.Grid(_ds(nface), coords=COORDS, face_connections=fc, autoparse_metadata=False)
    mode = ["fill", "extend", "periodic"][seed % 3]
    fill = float(rng.choice([0.0, 1.5, np.nan]))
    # scalar with a random dim order (face not first, an extra dim somewhere)
    base = ["face", "y", "x"]
    dims = list(rng.permutation(base)) if seed % 2 else base
    dims.insert(int(rng.integers(0, 4)), "t")
    shape = [nface if d == "face" else 2 if d == "t" else N for d in dims]
    da = xg.DataArray(rng.random(shape), dims=tuple(dims))
    for pw in WIDTHS:
        got = pad(da, grid, padding_width=dict(pw), padding=mode, fill_value=fill)
        want = pad_face_connections(da.values, da.dims, AXES, "face", links, dict(pw),
                                    {"X": mode, "Y": mode}, {"X": fill, "Y": fill})
        assert got.dims == da.dims
        np.testing.assert_array_equal(got.values, want, err_msg=f"scalar {pw} {mode} {links}")
    # vector components with DIFFERENT dim orders (u: face, xl, y / v: y-major)
    u = xg.DataArray(rng.random((nface, N, N)), dims=("face", "xl", "y"))
    v = xg.DataArray(rng.random((N, nface, N)), dims=("yl", "face", "x"))
    for pw in WIDTHS:
        for comp, vax, partner, pax in ((u, "X", v, "Y"), (v, "Y", u, "X")):
            got = pad({vax: comp}, grid, padding_width=dict(pw), padding=mode, fill_value=fill, other_component={pax: partner})
            want = pad_face_connections(comp.values, comp.dims, AXES, "face", links, dict(pw), {"X": mode, "Y": mode},
                                        {"X": fill, "Y": fill}, vector_axis=vax, partner=partner.values, partner_dims=partner.dims)
            np.testing.assert_array_equal(got.values, want, err_msg=f"vector {vax} {pw} {mode} {links}")


@pytest.mark.parametrize("seed", range(12))
def test_random_topology_operators(seed):
    from oracle import stencil as so

    rng = np.random.default_rng(100 + seed)
    nface = int(rng.integers(2, 5))
    links = _random_links(rng, nface)
    grid = xg.Grid(_ds(nface), coords=COORDS, face_connections={"face": links}, autoparse_metadata=False)
    mode = ["fill", "extend", "periodic"][seed % 3]
    da = xg.DataArray(rng.random((2, nface, N, N)), dims=("t", "face", "y", "x"))
    for op, ax in itertools.product(["diff", "interp", "max"], ["X", "Y"]):
        out = getattr(grid, op)(da, ax, padding=mode, fill_value=0.5)
        padded = pad_face_connections(da.values, da.dims, AXES, "face", links, {ax: (1, 0)},
                                      {"X": mode, "Y": mode}, {"X": 0.5, "Y": 0.5})
        k = da.dims.index("x" if ax == "X" else "y")
        want = np.moveaxis(so.KERNELS[op](np.moveaxis(padded, k, -1)), -1, k)
        np.testing.assert_array_equal(out.values, want, err_msg=f"{op} {ax} {mode} {links}")
