"""Face connections (cubed sphere / LLC topology): xgcm_b200 against oracle/faces.py and the seam
assertions of the reference's tests (xgcm/test/test_faceconnections.py, test_padding.py:341-1205).

The halo of a connected edge is written by ``xg_strided_copy``; these bodies also run on CPU
against the mock backend (tests/test_host_logic.py) for the host logic that derives the strides.
"""

import itertools
import warnings

import numpy as np
import pytest

import xgcm_b200 as xg
from oracle.faces import pad_face_connections
from xgcm_b200.padding import pad

pytestmark = pytest.mark.gpu

N = 9
COORDS = {"X": {"center": "x", "left": "xl"}, "Y": {"center": "y", "left": "yl"}}
AXES = {"X": ("x", "xl"), "Y": ("y", "yl")}

X_TO_X = {"face": {0: {"X": (None, (1, "X", False))}, 1: {"X": ((0, "X", False), None)}}}
X_TO_X_REV = {"face": {0: {"X": (None, (1, "X", True))}, 1: {"X": (None, (0, "X", True))}}}
X_TO_Y = {"face": {0: {"X": (None, (1, "Y", False))}, 1: {"Y": ((0, "X", False), None)}}}
X_TO_Y_REV = {"face": {0: {"X": (None, (1, "Y", True))}, 1: {"Y": (None, (0, "X", True))}}}

# xgcm/test/test_faceconnections.py:99-127
CUBED_SPHERE = {
    "face": {
        0: {"X": ((3, "X", False), (1, "X", False)), "Y": ((4, "Y", False), (5, "Y", False))},
        1: {"X": ((0, "X", False), (2, "X", False)), "Y": ((4, "X", False), (5, "X", True))},
        2: {"X": ((1, "X", False), (3, "X", False)), "Y": ((4, "Y", True), (5, "Y", True))},
        3: {"X": ((2, "X", False), (0, "X", False)), "Y": ((4, "X", True), (5, "X", False))},
        4: {"X": ((3, "Y", True), (1, "Y", False)), "Y": ((2, "Y", True), (0, "Y", False))},
        5: {"X": ((3, "Y", False), (1, "Y", True)), "Y": ((0, "Y", False), (2, "Y", True))},
    }
}

PADDING_WIDTHS = [
    {"X": (1, 1)},
    {"X": (1, 2)},
    {"X": (0, 1)},
    {"X": (1, 1), "Y": (1, 1)},
    {"X": (2, 2), "Y": (2, 2)},
    {"X": (0, 1), "Y": (1, 0)},
    {"X": (0, 2), "Y": (1, 0)},
]


def _ds(nface=2, dtype=np.float64, seed=0, extra_dim=False):
    """xgcm/test/test_faceconnections.py:10-36: note u is (face, xl, y) and v (face, x, yl)."""
    rng = np.random.default_rng(seed)
    lead = (3,) if extra_dim else ()
    ld = ("time",) if extra_dim else ()
    coords = {"x": np.arange(N) + 0.0, "xl": np.arange(N) - 0.5, "y": np.arange(N) + 0.0,
              "yl": np.arange(N) - 0.5, "face": np.arange(nface)}
    if extra_dim:
        coords["time"] = np.arange(3.0)
    return xg.Dataset(
        data_vars={
            "data_c": (ld + ("face", "y", "x"), rng.random(lead + (nface, N, N)).astype(dtype)),
            "u": (ld + ("face", "xl", "y"), rng.random(lead + (nface, N, N)).astype(dtype)),
            "v": (ld + ("face", "x", "yl"), rng.random(lead + (nface, N, N)).astype(dtype)),
        },
        coords=coords,
    )


def _oracle(da, fc, pw, padding, fill_value, vector_axis=None, partner=None):
    pads = {ax: padding for ax in AXES} if not isinstance(padding, dict) else padding
    fills = {ax: fill_value for ax in AXES}
    return pad_face_connections(
        da.values, da.dims, AXES, "face", fc["face"], dict(pw), pads, fills,
        vector_axis=vector_axis,
        partner=None if partner is None else partner.values,
        partner_dims=None if partner is None else partner.dims,
    )


# ------------------------------------------------------------------ construction / validation
def test_create_connected_grid():
    """test_faceconnections.py:134-152: the Axis objects carry the checked links."""
    ds = _ds()
    grid = xg.Grid(ds, coords=COORDS, face_connections=X_TO_X)
    xaxis = grid.axes["X"]
    assert xaxis._facedim == "face"
    assert xaxis._face_connections[0][1][0] == 1
    assert xaxis._face_connections[0][1][1] is xaxis
    assert xaxis._face_connections[1][0][0] == 0
    assert xaxis._face_connections[1][0][1] is xaxis
    xg.Grid(_ds(6), coords=COORDS, face_connections=CUBED_SPHERE)  # test_faceconnections.py:406-407


def test_connection_errors():
    """grid.py:334-409: wrong face dim, dangling link, link that is not mirrored."""
    ds = _ds()
    with pytest.raises(ValueError, match="Face dimension nope does not exist in the dataset."):
        xg.Grid(ds, coords=COORDS, face_connections={"nope": X_TO_X["face"]})
    with pytest.raises(ValueError, match="Only one face dimension"):
        xg.Grid(ds, coords=COORDS, face_connections={"face": {}, "tile": {}})
    with pytest.raises(KeyError, match="Couldn't find a face link"):
        xg.Grid(ds, coords=COORDS, face_connections={"face": {0: {"X": (None, (1, "X", False))}, 1: {}}})
    with pytest.raises(ValueError, match="Face link mismatch"):
        xg.Grid(ds, coords=COORDS, face_connections={
            "face": {0: {"X": (None, (1, "X", False))}, 1: {"X": ((0, "X", True), None)}}})


# ------------------------------------------------------------------ pad against the oracle
@pytest.mark.parametrize("fill_value", [np.nan, 0.0])
@pytest.mark.parametrize("padding_width", PADDING_WIDTHS)
@pytest.mark.parametrize("fc", [X_TO_X, X_TO_X_REV, X_TO_Y, X_TO_Y_REV], ids=["xx", "xx_rev", "xy", "xy_rev"])
def test_pad_scalar_matches_oracle(fc, padding_width, fill_value):
    ds = _ds()
    grid = xg.Grid(ds, coords=COORDS, face_connections=fc)
    out = pad(ds["data_c"], grid, padding_width=dict(padding_width), padding="fill", fill_value=fill_value)
    expect = _oracle(ds["data_c"], fc, padding_width, "fill", fill_value)
    assert out.dims == ("face", "y", "x")
    np.testing.assert_array_equal(out.values, expect)


@pytest.mark.parametrize("fill_value", [np.nan, 0.0])
@pytest.mark.parametrize("padding_width", PADDING_WIDTHS)
@pytest.mark.parametrize("fc", [X_TO_X, X_TO_X_REV, X_TO_Y, X_TO_Y_REV], ids=["xx", "xx_rev", "xy", "xy_rev"])
def test_pad_vector_matches_oracle(fc, padding_width, fill_value):
    """Both components, through the dict form and the bare form (test_padding.py:939-1003)."""
    ds = _ds()
    grid = xg.Grid(ds, coords=COORDS, face_connections=fc)
    u, v = ds["u"], ds["v"]
    kw = dict(grid=grid, padding_width=dict(padding_width), padding="fill", fill_value=fill_value)
    u_out = pad({"X": u}, other_component={"Y": v}, **kw)
    v_out = pad({"Y": v}, other_component={"X": u}, **kw)
    np.testing.assert_array_equal(u_out.values, _oracle(u, fc, padding_width, "fill", fill_value, "X", v))
    np.testing.assert_array_equal(v_out.values, _oracle(v, fc, padding_width, "fill", fill_value, "Y", u))
    np.testing.assert_array_equal(pad(u, other_component={"Y": v}, **kw).values, u_out.values)
    np.testing.assert_array_equal(pad(v, other_component={"X": u}, **kw).values, v_out.values)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("padding", ["fill", "extend", "periodic"])
def test_pad_cubed_sphere_extra_dim_matches_oracle(dtype, padding):
    """Six faces, every edge connected, a leading batch dim, each basic padding underneath."""
    ds = _ds(6, dtype=dtype, extra_dim=True)
    grid = xg.Grid(ds, coords=COORDS, face_connections=CUBED_SPHERE)
    for pw in ({"X": (1, 1), "Y": (1, 1)}, {"X": (2, 0)}, {"Y": (0, 1)}):
        out = pad(ds["data_c"], grid, padding_width=dict(pw), padding=padding, fill_value=1.5)
        np.testing.assert_array_equal(out.values, _oracle(ds["data_c"], CUBED_SPHERE, pw, padding, 1.5))
        assert out.values.dtype == dtype


def test_cubed_sphere_scalar_pad_connected_halos():
    """test_faceconnections.py:445-478: every connected halo cell reads the declared neighbour."""
    ds = _ds(6)
    grid = xg.Grid(ds, coords=COORDS, face_connections=CUBED_SPHERE)
    face_field = xg.DataArray(
        np.broadcast_to(np.arange(6)[:, None, None], (6, N, N)).astype(float), dims=("face", "y", "x"))
    padded = pad(face_field, grid, {"X": (1, 1), "Y": (1, 1)}, padding={"X": "fill", "Y": "fill"},
                 fill_value=np.nan).values
    for f in range(6):
        (left_x, right_x), (down_y, up_y) = CUBED_SPHERE["face"][f]["X"], CUBED_SPHERE["face"][f]["Y"]
        np.testing.assert_array_equal(padded[f, 1:-1, 0], left_x[0])
        np.testing.assert_array_equal(padded[f, 1:-1, -1], right_x[0])
        np.testing.assert_array_equal(padded[f, 0, 1:-1], down_y[0])
        np.testing.assert_array_equal(padded[f, -1, 1:-1], up_y[0])


def test_bare_vector_pad_ambiguous_axis_raises():
    """test_padding.py:1005-1037"""
    ds = _ds()
    grid = xg.Grid(ds, coords=COORDS, face_connections=X_TO_X)
    tracer = xg.DataArray(np.zeros((2, N, N)), dims=("face", "x", "y"))
    with pytest.raises(ValueError, match="infer the axis"):
        pad(tracer, grid=grid, padding_width={"X": (1, 1)}, padding="fill", fill_value=0.0,
            other_component={"X": ds["u"]})


# ------------------------------------------------------------------ operators across seams
def test_diff_interp_connected_grid_x_to_x():
    """test_faceconnections.py:164-180"""
    ds = _ds()
    d = ds["data_c"].values
    grid = xg.Grid(ds, coords=COORDS, face_connections=X_TO_X, padding="fill")
    diff_x = grid.diff(ds["data_c"], "X", padding="fill")
    interp_x = grid.interp(ds["data_c"], "X", padding="fill")
    assert diff_x.dims == ("face", "y", "xl")
    np.testing.assert_array_equal(diff_x.values[1, :, 0], d[1, :, 0] - d[0, :, -1])
    np.testing.assert_array_equal(interp_x.values[1, :, 0], 0.5 * (d[1, :, 0] + d[0, :, -1]))
    np.testing.assert_array_equal(diff_x.values[0, :, 0], d[0, :, 0] - 0.0)
    np.testing.assert_array_equal(interp_x.values[0, :, 0], 0.5 * (d[0, :, 0] + 0.0))
    np.testing.assert_array_equal(diff_x.values[:, :, 1:], d[:, :, 1:] - d[:, :, :-1])
    np.testing.assert_array_equal(diff_x.coords["xl"].values, ds["xl"].values)


def test_connected_grid_rejects_unknown_padding_modes():
    """A typo / unsupported mode must raise on a connected grid exactly as on a simple one, not turn into a
    periodic halo on the unconnected edges."""
    ds = _ds()
    grid = xg.Grid(ds, coords=COORDS, face_connections=X_TO_X, padding="fill")
    with pytest.raises(ValueError, match="padding must be one of"):
        grid.diff(ds["data_c"], "X", padding="bogus")
    with pytest.raises(NotImplementedError, match="extrapolate"):
        grid.interp(ds["data_c"], "X", padding="extrapolate")
    # the three supported modes still differ from each other on the unconnected edge (face 0, left)
    outs = [grid.diff(ds["data_c"], "X", padding=p).values[0, :, 0] for p in ("fill", "extend", "periodic")]
    assert not np.array_equal(outs[0], outs[1]) and not np.array_equal(outs[1], outs[2])


def test_diff_interp_connected_grid_x_to_y():
    """test_faceconnections.py:183-202: a rotated seam."""
    ds = _ds()
    d = ds["data_c"].values
    grid = xg.Grid(ds, coords=COORDS, face_connections=X_TO_Y)
    diff_y = grid.diff(ds["data_c"], "Y", padding="fill")
    interp_y = grid.interp(ds["data_c"], "Y", padding="fill")
    np.testing.assert_array_equal(diff_y.values[1, 0, :], d[1, 0, :] - d[0, ::-1, -1])
    np.testing.assert_array_equal(interp_y.values[1, 0, :], 0.5 * (d[1, 0, :] + d[0, ::-1, -1]))


@pytest.mark.parametrize("padding", ["periodic", "fill"])
def test_vector_connected_grid_x_to_y(padding):
    """test_faceconnections.py:205-230: with u = (-2, -1) and v = (1, 1) per face every value of
    v interpolated along X is 1 only if the seam sign flip is right."""
    ds = _ds()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        grid = xg.Grid(ds, coords=COORDS, face_connections=X_TO_Y, padding=padding, fill_value=1)
    u = xg.DataArray(ds["u"].values * 0 + np.array([-2.0, -1.0])[:, None, None], dims=ds["u"].dims)
    v = xg.DataArray(ds["v"].values * 0 + 1.0, dims=ds["v"].dims)
    v_out = grid.interp({"Y": v}, "X", other_component={"X": u})
    np.testing.assert_array_equal(v_out.values, 1.0)


def test_vector_diff_interp_connected_grid_x_to_y():
    """test_faceconnections.py:233-293"""
    ds = _ds()
    u, v = ds["u"].values, ds["v"].values
    grid = xg.Grid(ds, coords=COORDS, face_connections=X_TO_Y)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        vc = grid.interp_2d_vector({"X": ds["u"], "Y": ds["v"]}, to="center", padding="fill", fill_value=100)
        vd = grid.diff_2d_vector({"X": ds["u"], "Y": ds["v"]}, to="center", padding="fill", fill_value=100)
        u_c_interp, u_c_diff = vc["X"].values, vd["X"].values
        np.testing.assert_array_equal(u_c_interp[0, 0, :], 0.5 * (u[0, 0, :] + u[0, 1, :]))
        np.testing.assert_array_equal(u_c_diff[0, 0, :], u[0, 1, :] - u[0, 0, :])
        np.testing.assert_array_equal(u_c_interp[0, -1, :], 0.5 * (u[0, -1, :] + v[1, ::-1, 0]))
        np.testing.assert_array_equal(u_c_diff[0, -1, :], -u[0, -1, :] + v[1, ::-1, 0])
        with pytest.raises(NotImplementedError):
            grid.interp_2d_vector({"X": ds["v"], "Y": ds["u"]}, to="left", padding="fill")
        with pytest.raises(NotImplementedError):
            grid.interp_2d_vector({"X": ds["v"], "Y": ds["u"]}, padding="fill")


def test_diff_interp_cubed_sphere():
    """test_faceconnections.py:410-428: no boundary condition needed on a closed topology."""
    ds = _ds(6)
    grid = xg.Grid(ds, coords=COORDS, face_connections=CUBED_SPHERE)
    face = xg.DataArray(np.broadcast_to(np.arange(6.0)[:, None, None], (6, N, N)).copy(), dims=("face", "y", "x"))
    face_diff_x = grid.diff(face, "X").values
    np.testing.assert_array_equal(face_diff_x[:, 0, 0], [-3, 1, 1, 1, 1, 2])
    np.testing.assert_array_equal(face_diff_x[:, -1, 0], [-3, 1, 1, 1, 1, 2])
    face_diff_y = grid.diff(face, "Y").values
    np.testing.assert_array_equal(face_diff_y[:, 0, 0], [-4, -3, -2, -1, 2, 5])
    np.testing.assert_array_equal(face_diff_y[:, 0, -1], [-4, -3, -2, -1, 2, 5])
    face_interp_x = grid.interp(face, "X").values
    np.testing.assert_array_equal(face_interp_x[:, 0, 0], [1.5, 0.5, 1.5, 2.5, 3.5, 4.0])


def test_cubed_sphere_operators_match_oracle():
    """diff / interp / min / max along both axes == oracle pad + pairwise kernel."""
    from oracle import stencil as so

    ds = _ds(6, dtype=np.float32)
    grid = xg.Grid(ds, coords=COORDS, face_connections=CUBED_SPHERE)
    d = ds["data_c"]
    for op, ax in itertools.product(["diff", "interp", "min", "max"], ["X", "Y"]):
        out = getattr(grid, op)(d, ax)
        padded = _oracle(d, CUBED_SPHERE, {ax: (1, 0)}, None, 0.0)
        axis_num = d.dims.index("x" if ax == "X" else "y")
        expect = np.moveaxis(so.KERNELS[op](np.moveaxis(padded, axis_num, -1)), -1, axis_num)
        np.testing.assert_array_equal(out.values, expect.astype(np.float32))


def test_unconnected_edge_without_boundary_raises():
    """test_faceconnections.py:431-442"""
    ds = _ds()
    grid = xg.Grid(ds, coords=COORDS, face_connections=X_TO_X)
    with pytest.raises(ValueError, match="No boundary condition was specified"):
        grid.diff(ds["data_c"], "X")
    grid.diff(ds["data_c"], "X", padding="fill")


def test_vector_missing_other_component():
    """test_faceconnections.py:481-490"""
    ds = _ds()
    grid = xg.Grid(ds, coords=COORDS, face_connections=X_TO_Y)
    with pytest.raises(ValueError, match="Padding vector components requires `other_component` input"):
        grid.diff({"X": ds["u"]}, "X", other_component=None)


def test_multi_axis_on_connected_grid_goes_axis_by_axis():
    """interp over ['X', 'Y'] on a cubed sphere == the two single-axis calls chained."""
    ds = _ds(6)
    grid = xg.Grid(ds, coords=COORDS, face_connections=CUBED_SPHERE)
    both = grid.interp(ds["data_c"], ["X", "Y"])
    chained = grid.interp(grid.interp(ds["data_c"], "X"), "Y")
    np.testing.assert_array_equal(both.values, chained.values)


def test_metric_weighted_and_derivative_on_connected_grid():
    """metric_weighted multiplies BEFORE the (face-connected) halo is taken and divides after
    (grid.py:806-808,830-832); derivative divides the difference by the metric at the new
    position (grid.py:1576-1578)."""
    from oracle import stencil as so

    ds = _ds(6)
    rng = np.random.default_rng(7)
    dx_c = 1.0 + rng.random((6, N, N))
    dx_l = 1.0 + rng.random((6, N, N))
    ds2 = xg.Dataset(
        data_vars={"data_c": (("face", "y", "x"), ds["data_c"].values)},
        coords={"x": ds["x"].values, "xl": ds["xl"].values, "y": ds["y"].values, "yl": ds["yl"].values,
                "face": np.arange(6), "dx_c": (("face", "y", "x"), dx_c), "dx_l": (("face", "y", "xl"), dx_l)},
    )
    grid = xg.Grid(ds2, coords=COORDS, face_connections=CUBED_SPHERE, metrics={("X",): ["dx_c", "dx_l"]})
    d = ds2["data_c"]
    out = grid.interp(d, "X", metric_weighted="X")
    weighted = xg.DataArray(d.values * dx_c, dims=d.dims)
    padded = _oracle(weighted, CUBED_SPHERE, {"X": (1, 0)}, None, 0.0)
    np.testing.assert_array_equal(out.values, so.interp_forward(padded) / dx_l)
    der = grid.derivative(d, "X")
    padded = _oracle(d, CUBED_SPHERE, {"X": (1, 0)}, None, 0.0)
    np.testing.assert_array_equal(der.values, so.diff_forward(padded) / dx_l)


@pytest.mark.parametrize("padding", ["fill", "extend", "periodic"])
@pytest.mark.parametrize("fc", [X_TO_X, X_TO_X_REV, X_TO_Y, X_TO_Y_REV], ids=["xx", "xx_rev", "xy", "xy_rev"])
def test_operator_halo_planes_match_padded_oracle(fc, padding):
    """The operator fast path (halo planes gathered from the neighbours, fused kernel) against
    oracle pad + pairwise kernel: lower halo (center -> left) and upper halo (left -> center),
    scalars and vector components, every basic padding on the unconnected edges."""
    from oracle import stencil as so

    ds = _ds()
    grid = xg.Grid(ds, coords=COORDS, face_connections=fc)
    d, u, v = ds["data_c"], ds["u"], ds["v"]
    for ax, dim in (("X", "x"), ("Y", "y")):
        out = grid.diff(d, ax, padding=padding, fill_value=2.5)
        padded = _oracle(d, fc, {ax: (1, 0)}, padding, 2.5)
        k = d.dims.index(dim)
        np.testing.assert_array_equal(
            out.values, np.moveaxis(so.diff_forward(np.moveaxis(padded, k, -1)), -1, k))
    out = grid.interp({"X": u}, "X", other_component={"Y": v}, padding=padding, fill_value=2.5)
    padded = _oracle(u, fc, {"X": (0, 1)}, padding, 2.5, "X", v)
    k = u.dims.index("xl")
    np.testing.assert_array_equal(
        out.values, np.moveaxis(so.interp_forward(np.moveaxis(padded, k, -1)), -1, k))
    out = grid.interp({"Y": v}, "Y", other_component={"X": u}, padding=padding, fill_value=2.5)
    padded = _oracle(v, fc, {"Y": (0, 1)}, padding, 2.5, "Y", u)
    k = v.dims.index("yl")
    np.testing.assert_array_equal(
        out.values, np.moveaxis(so.interp_forward(np.moveaxis(padded, k, -1)), -1, k))


def test_cumsum_and_integrate_on_connected_grid():
    """cumsum pads the CUMSUM'D data with ``pad`` (grid.py:1385-1391): on a connected grid the
    boundary cell of a connected edge comes from the neighbour face's cumsum.  (Across an
    axis-swapping seam the trimmed array is no longer square, in the reference as here: same-axis
    seams only.)  integrate has no halo at all."""
    from oracle import stencil as so

    ds = _ds()
    d = ds["data_c"]
    for fc in (X_TO_X, X_TO_X_REV):
        grid = xg.Grid(ds, coords=COORDS, face_connections=fc)
        with pytest.raises(ValueError, match="No boundary condition was specified"):
            grid.cumsum(d, "X", to="left")  # the outer edges are not connected
        out = grid.cumsum(d, "X", to="left", padding="fill", fill_value=0.0)
        assert out.dims == ("face", "y", "xl")
        scanned = so.cumscan(d.values, 2, False, "drop_last", 0, 0, None)
        expect = _oracle(xg.DataArray(scanned, dims=d.dims), fc, {"X": (1, 0)}, "fill", 0.0)
        np.testing.assert_array_equal(out.values, expect)
        u = ds["u"]  # left -> center along X: the natural landing position, no padding involved
        np.testing.assert_array_equal(grid.cumsum(u, "X").values, so.cumscan(u.values, 1, False, "none", 0, 0, None))
    ds6 = _ds(6)
    grid6 = xg.Grid(ds6, coords=COORDS, face_connections=CUBED_SPHERE)
    with pytest.raises(ValueError, match="equal size"):
        grid6.cumsum(ds6["data_c"], "X", to="left")
