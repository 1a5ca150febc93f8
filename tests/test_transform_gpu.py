"""Grid.transform (linear / log) against the reference's golden `cases` and the oracle."""

import json
import os

import numpy as np
import torch
import pytest

import xgcm_b200 as xg
from oracle import stencil as oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _arr(v):
    return np.array([np.nan if x is None else x for x in v])


def _cases():
    cases = json.load(open(os.path.join(GOLDEN, "transform_cases.json")))
    return {k: v for k, v in cases.items()
            if "multidim_target" not in k and v["transform_kwargs"]["method"] != "conservative"}


def _conservative_cases():
    cases = json.load(open(os.path.join(GOLDEN, "transform_cases.json")))
    return {k: v for k, v in cases.items()
            if "multidim_target" not in k and v["transform_kwargs"]["method"] == "conservative"}


@pytest.mark.parametrize("name", sorted(_conservative_cases()))
def test_reference_conservative_cases(name):
    """xgcm/test/test_transform.py:428-640 + :1041-1069: Grid.transform(method="conservative")."""
    c = _conservative_cases()[name]
    sdim, scoord = c["source_coord"]
    data_vars = {c["source_data"][0]: ((sdim,), _arr(c["source_data"][1]))}
    if "source_additional_data" in c:
        data_vars[c["source_additional_data"][0]] = ((c["source_additional_data_coord"][0],), _arr(c["source_additional_data"][1]))
    coords = {sdim: _arr(scoord)}
    if "source_bounds_coord" in c:
        coords[c["source_bounds_coord"][0]] = _arr(c["source_bounds_coord"][1])
    ds = xg.Dataset(data_vars=data_vars, coords=coords)
    tdim, tvals = c["target_coord"]
    target = xg.DataArray(_arr(c["target_data"][1]), dims=(tdim,), coords={tdim: _arr(tvals)}, name=c["target_data"][0])
    kw = dict(c["transform_kwargs"])
    if kw.get("target_data"):
        kw["target_data"] = ds[kw["target_data"]]
    grid = xg.Grid(ds, **c["grid_kwargs"])
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = grid.transform(ds[c["source_data"][0]], "Z", target, **kw)
    want = _arr(c["expected_data"][1]).astype(float)
    assert got.dims == (c["expected_coord"][0],)
    np.testing.assert_allclose(got.values, want, rtol=1e-5, atol=1e-6, equal_nan=True)
    np.testing.assert_allclose(got.coords[c["expected_coord"][0]].values, _arr(c["expected_coord"][1]))
    # an extensive quantity is conserved (test_transform.py:1036-1038)
    np.testing.assert_allclose(np.nansum(got.values), np.nansum(_arr(c["source_data"][1])), rtol=1e-6)


def test_conservative_reference_numba_golden_vectors():
    """tests/golden/conservative_ref.npz: outputs of the reference's numba gufunc, bit-exact."""
    from xgcm_b200.transform import interp_1d_conservative

    g = np.load(os.path.join(GOLDEN, "conservative_ref.npz"))
    for tag in ("float32", "float64"):
        phi, theta, bins = g[f"phi|{tag}"], g[f"theta|{tag}"], g[f"bins|{tag}"]
        for d, b in (("up", bins), ("down", bins[::-1].copy())):
            got = interp_1d_conservative(phi, theta, b)
            assert got.dtype == g[f"out|{tag}|{d}"].dtype
            np.testing.assert_array_equal(got, g[f"out|{tag}|{d}"])
    with pytest.raises(ValueError, match="not monotonic"):
        interp_1d_conservative(g["phi|float64"], g["theta|float64"], np.array([0.0, 2.0, 1.0]))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_conservative_3d_field_vs_oracle(dtype):
    rng = np.random.default_rng(12)
    nz, ny, nx, m = 14, 6, 40, 11
    a = rng.random((nz, ny, nx)).astype(dtype)
    a[rng.random(a.shape) < 0.02] = np.nan
    bounds = np.cumsum(0.5 + rng.random((nz + 1, ny, nx)), axis=0).astype(dtype)
    bounds[:, 0, :5] = bounds[::-1, 0, :5]
    bounds[3, 1, 7] = np.nan
    ds = xg.Dataset(data_vars={"q": (("z", "y", "x"), a), "sig": (("zo", "y", "x"), bounds)},
                    coords={"z": np.arange(nz) + 0.5, "zo": np.arange(nz + 1.0)})
    grid = xg.Grid(ds, coords={"Z": {"center": "z", "outer": "zo"}})
    bins = np.linspace(0, float(np.nanmax(bounds)) + 1, m).astype(dtype)
    got = grid.transform(ds["q"], "Z", bins, target_data=ds["sig"], method="conservative")
    assert got.dims == ("y", "x", "sig") and got.shape == (ny, nx, m - 1)
    want = oracle.vinterp_conservative(a, bounds, bins, 0)
    np.testing.assert_array_equal(got.values, want)
    np.testing.assert_array_equal(got.coords["sig"].values, (bins[1:] + bins[:-1]) / 2)


@pytest.mark.parametrize("dtype,m", [(np.float32, 300), (np.float64, 300), (np.float64, 1700), (np.float32, 9000)])
def test_conservative_many_bins_multipass(dtype, m):
    """target grids with more bins than one shared-memory tile holds (round 1: NotImplementedError above
    ~193 fp64 / ~385 fp32 edges) run in passes over bin ranges; every bin still sums in source-cell order."""
    from xgcm_b200 import ops

    rng = np.random.default_rng(14)
    nz, ncol = 12, 96
    a = rng.random((nz, ncol)).astype(dtype)
    a[rng.random(a.shape) < 0.03] = np.nan
    bounds = np.cumsum(0.5 + rng.random((nz + 1, ncol)), axis=0).astype(dtype)
    bounds[:, :4] = bounds[::-1, :4]
    bins = np.linspace(0, float(np.nanmax(bounds)) + 1, m).astype(dtype)
    want = oracle.vinterp_conservative(a, bounds, bins, 0)
    got = ops.vinterp_conservative(torch.from_numpy(a).cuda(), torch.from_numpy(bounds).cuda(),
                                   torch.from_numpy(bins).cuda(), 0).cpu().numpy()
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("name", sorted(_cases()))
def test_reference_transform_cases(name):
    """xgcm/test/test_transform.py:41-683 + :951-1069 (high-level Grid.transform)."""
    c = _cases()[name]
    sdim, scoord = c["source_coord"]
    data_vars = {c["source_data"][0]: ((sdim,), _arr(c["source_data"][1]))}
    if "source_additional_data" in c:
        data_vars[c["source_additional_data"][0]] = ((c["source_additional_data_coord"][0],), _arr(c["source_additional_data"][1]))
    coords = {sdim: _arr(scoord)}
    if "source_bounds_coord" in c:
        coords[c["source_bounds_coord"][0]] = _arr(c["source_bounds_coord"][1])
    ds = xg.Dataset(data_vars=data_vars, coords=coords)
    tdim, tvals = c["target_coord"]
    target = xg.DataArray(_arr(c["target_data"][1]), dims=(tdim,), coords={tdim: _arr(tvals)}, name=c["target_data"][0])
    kw = dict(c["transform_kwargs"])
    if kw.get("target_data"):
        kw["target_data"] = ds[kw["target_data"]]
    grid = xg.Grid(ds, **c["grid_kwargs"])
    got = grid.transform(ds[c["source_data"][0]], "Z", target, **kw)
    want = _arr(c["expected_data"][1]).astype(float)
    for ii in c.get("expected_data_mask_index", []):
        want[ii] = np.nan
    assert got.dims == (c["expected_coord"][0],)
    assert got.name == c["source_data"][0]  # reference transform.py:455-466 never forwards `suffix`
    from xgcm_b200.transform import linear_interpolation

    theta = kw["target_data"] if kw.get("target_data") is not None else ds[sdim]
    mid = linear_interpolation(ds[c["source_data"][0]], theta, target, sdim, sdim, tdim,
                               mask_edges=kw.get("mask_edges", True), logarithmic=kw["method"] == "log",
                               suffix=kw.get("suffix", ""), grid=grid)
    assert mid.name == "data" + kw.get("suffix", "")  # test_transform.py:951-992 (mid level)
    np.testing.assert_allclose(mid.values, want, rtol=1e-5, atol=1e-6, equal_nan=True)
    np.testing.assert_allclose(got.values, want, rtol=1e-5, atol=1e-6, equal_nan=True)
    np.testing.assert_array_equal(got.coords[c["expected_coord"][0]].values, _arr(c["expected_coord"][1]))


@pytest.mark.parametrize("name", ["linear_depth_depth_nomask_multidim_target", "linear_depth_depth_multidim_target"])
def test_reference_multidim_target_cases(name):
    """xgcm/test/test_transform.py:122-215: a 2-D target (one level vector per horizontal point)."""
    c = json.load(open(os.path.join(GOLDEN, "transform_cases.json")))[name]
    sdim, scoord = c["source_coord"]
    ds = xg.Dataset(
        data_vars={c["source_data"][0]: ((sdim,), _arr(c["source_data"][1])),
                   c["source_additional_data"][0]: ((sdim,), _arr(c["source_additional_data"][1]))},
        coords={sdim: _arr(scoord)},
    )
    tvals = np.array(c["target_data"][1], dtype=float)
    target = xg.DataArray(tvals, dims=tuple(c["target_dims"]), name=c["target_data"][0])
    kw = dict(c["transform_kwargs"])
    kw["target_data"] = ds[kw["target_data"]]
    grid = xg.Grid(ds, **c["grid_kwargs"])
    got = grid.transform(ds[c["source_data"][0]], "Z", target, **kw)
    want = np.array(c["expected_data"][1], dtype=float)
    for ii in c.get("expected_data_mask_index", []):
        want[tuple(ii)] = np.nan
    assert got.dims == tuple(c["expected_dims"])
    np.testing.assert_allclose(got.values, want, rtol=1e-6, equal_nan=True)
    with pytest.raises(ValueError, match="target_dim"):
        kw2 = dict(kw)
        kw2.pop("target_dim")
        grid.transform(ds[c["source_data"][0]], "Z", target, **kw2)


def test_per_column_targets_3d():
    """Terrain-following style targets: (Y, X, m) level array against a (Z, Y, X) field."""
    rng = np.random.default_rng(8)
    nz, ny, nx, m = 12, 5, 37, 9
    a = rng.random((nz, ny, nx)).astype(np.float32)
    depth = np.cumsum(1 + rng.random(nz)).astype(np.float32)
    ds = xg.Dataset(data_vars={"t": (("z", "y", "x"), a)}, coords={"z": depth})
    grid = xg.Grid(ds, coords={"Z": {"center": "z"}})
    tg = (depth[0] + rng.random((ny, nx, m)) * (depth[-1] - depth[0]) * 1.2 - 0.5).astype(np.float32)
    tg.sort(axis=-1)
    got = grid.transform(ds["t"], "Z", xg.DataArray(tg, dims=("y", "x", "lev")), target_dim="lev")
    assert got.dims == ("y", "x", "lev")
    want = np.empty((ny, nx, m), np.float32)
    for j in range(ny):
        for i in range(nx):
            want[j, i] = oracle.vinterp_linear(a[:, j, i], depth, tg[j, i], 0, True)
    np.testing.assert_array_equal(got.values, want)


def test_reference_numba_golden_vectors():
    """tests/golden/interp1d_ref.npz: outputs of the reference's own numba gufunc, bit-exact."""
    from xgcm_b200.transform import interp_1d_linear

    g = np.load(os.path.join(GOLDEN, "interp1d_ref.npz"))
    for tag in ("float32", "float64"):
        phi, theta, target = g[f"phi|{tag}"], g[f"theta|{tag}"], g[f"target|{tag}"]
        for mask in (0, 1):
            for bypass in (0, 1):
                got = interp_1d_linear(phi, theta, target, mask_edges=bool(mask), bypass_checks=bool(bypass))
                want = g[f"out|{tag}|{mask}|{bypass}|0"]
                assert got.dtype == want.dtype
                np.testing.assert_array_equal(got, want)
        got = interp_1d_linear(phi, g[f"log_theta|{tag}"], g[f"log_target|{tag}"], mask_edges=True, logarithmic=True)
        if tag == "float32":
            # the golden file holds the reference's float32 log (numpy's SIMD algorithm); the device evaluates the
            # same algorithm operation for operation (csrc/xg_vinterp.cuh xg_log<float>): bit-identical
            np.testing.assert_array_equal(got, g[f"out|{tag}|1|0|1"])
        else:
            np.testing.assert_allclose(got, g[f"out|{tag}|1|0|1"], rtol=1e-12, atol=1e-12, equal_nan=True)


def test_analytic_interp(rtol=1e-4):
    """xgcm/test/test_transform.py:850-865: uniformly stratified scalar, rtol 1e-4."""
    from xgcm_b200.transform import interp_1d_linear

    nz, nx = 100, 1000
    z_vertex = np.linspace(0, 1, nz + 1)
    z = 0.5 * (z_vertex[:-1] + z_vertex[1:])
    x = 2 * np.pi * np.linspace(0, 1, nx)
    theta = z + 0.1 * np.cos(3 * x)[:, None]
    phi = np.sin(theta) + 0.1 * np.cos(5 * x)[:, None]
    levels = np.arange(0.2, 0.9, 0.025)
    expected = np.sin(levels) + 0.1 * np.cos(5 * x)[:, None]
    got = interp_1d_linear(phi, theta, levels, mask_edges=False)
    np.testing.assert_allclose(got, expected, rtol=rtol)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_config5_like_vertical_regrid(dtype):
    """BASELINE configs[4] in miniature: (Z, Y, X) field to new depth levels, output (Y, X, Znew)."""
    rng = np.random.default_rng(5)
    nz, ny, nx, m = 25, 12, 20, 33
    dz = 10 * 1.05 ** np.arange(nz)
    depth = (np.cumsum(dz) - dz / 2).astype(dtype)
    a = rng.random((nz, ny, nx)).astype(dtype)
    ds = xg.Dataset(data_vars={"temp": (("z", "y", "x"), a)}, coords={"z": depth, "y": np.arange(ny), "x": np.arange(nx)})
    grid = xg.Grid(ds, coords={"Z": {"center": "z"}})
    levels = np.linspace(depth[0] - 5, depth[-1] + 5, m).astype(dtype)
    got = grid.transform(ds["temp"], "Z", levels)
    assert got.dims == ("y", "x", "z") and got.shape == (ny, nx, m)
    assert got.name == "temp"
    want = oracle.vinterp_linear(a, depth.reshape(-1, 1, 1) * np.ones((1, ny, nx), dtype), levels, 0, True)
    assert got.dtype == want.dtype
    np.testing.assert_array_equal(got.values, want)
    assert np.isnan(got.values[..., 0]).all() and np.isnan(got.values[..., -1]).all()
    # target_data given as a 3-D tracer field (transform onto e.g. density)
    dens = np.cumsum(0.1 + rng.random((nz, ny, nx)), axis=0).astype(dtype)
    ds2 = xg.Dataset(data_vars={"temp": (("z", "y", "x"), a), "dens": (("z", "y", "x"), dens)}, coords={"z": depth})
    grid2 = xg.Grid(ds2, coords={"Z": {"center": "z"}})
    tg = xg.DataArray(np.linspace(0, dens.max(), 17).astype(dtype), dims=("dens_lev",))
    got = grid2.transform(ds2["temp"], "Z", tg, target_data=ds2["dens"])
    assert got.dims == ("y", "x", "dens_lev")
    np.testing.assert_array_equal(got.values, oracle.vinterp_linear(a, dens, tg.values, 0, True))


def test_transform_errors():
    ds = xg.Dataset(data_vars={"t": (("z",), np.arange(5.0))}, coords={"z": np.arange(5.0)})
    with pytest.raises(ValueError, match="non-periodic"):
        xg.Grid(ds, coords={"Z": {"center": "z"}}, padding="periodic").transform(ds["t"], "Z", np.arange(3.0))
    grid = xg.Grid(ds, coords={"Z": {"center": "z"}})
    with pytest.raises(ValueError):
        grid.transform(ds["t"], "Z", [1, 2, 3])  # list target: must be ndarray / DataArray
    with pytest.raises(RuntimeError, match="outer"):
        grid.transform(ds["t"], "Z", np.arange(3.0), method="conservative")


# ---------------------------------------------------------------- more of xgcm/test/test_transform.py:920-1430
def _construct(name):
    """test_transform.py:685-760 (construct_test_source_data) on the labelled arrays of this package:
    (source Dataset, grid kwargs, target DataArray, transform kwargs, expected values)."""
    c = json.load(open(os.path.join(GOLDEN, "transform_cases.json")))[name]
    sdim, scoord = c["source_coord"]
    data_vars = {c["source_data"][0]: ((sdim,), _arr(c["source_data"][1]).astype(float))}
    if "source_additional_data" in c:
        data_vars[c["source_additional_data"][0]] = (
            (c["source_additional_data_coord"][0],), _arr(c["source_additional_data"][1]).astype(float))
    coords = {sdim: _arr(scoord).astype(float)}
    if "source_bounds_coord" in c:
        coords[c["source_bounds_coord"][0]] = _arr(c["source_bounds_coord"][1]).astype(float)
    ds = xg.Dataset(data_vars=data_vars, coords=coords)
    tdim, tvals = c["target_coord"]
    target = xg.DataArray(_arr(c["target_data"][1]).astype(float), dims=(tdim,),
                          coords={tdim: _arr(tvals).astype(float)}, name=c["target_data"][0])
    kw = dict(c["transform_kwargs"])
    if kw.get("target_data"):
        kw["target_data"] = ds[kw["target_data"]]
    want = _arr(c["expected_data"][1]).astype(float)
    for ii in c.get("expected_data_mask_index", []):
        want[ii] = np.nan
    return ds, dict(c["grid_kwargs"]), target, kw, want


MULTIDIM = ["conservative_depth_dens_nonmono_edge", "linear_depth_dens", "linear_depth_depth", "conservative_depth_temp"]


def test_mid_level_rejects_numpy_targets():
    """test_transform.py:924-947: the labelled wrappers need a labelled target."""
    from xgcm_b200.transform import conservative_interpolation, linear_interpolation

    for name, fn in (("linear_depth_depth", linear_interpolation), ("conservative_depth_depth", conservative_interpolation)):
        ds, _, target, _, _ = _construct(name)
        (dim,) = ds["data"].dims
        (tdim,) = target.dims
        with pytest.raises((ValueError, AttributeError, TypeError)):
            fn(ds["data"], ds[dim], target.values, dim, dim, tdim)


def test_conservative_multidim_target_and_explicit_target_dim_and_bounds_warning():
    """test_transform.py:1056-1116"""
    import warnings

    c = json.load(open(os.path.join(GOLDEN, "transform_cases.json")))["conservative_depth_depth_multidim_target"]
    sdim, scoord = c["source_coord"]
    ds = xg.Dataset(data_vars={c["source_data"][0]: ((sdim,), _arr(c["source_data"][1]).astype(float))},
                    coords={sdim: _arr(scoord).astype(float),
                            c["source_bounds_coord"][0]: _arr(c["source_bounds_coord"][1]).astype(float)})
    target = xg.DataArray(np.array(c["target_data"][1], dtype=float), dims=tuple(c["target_dims"]), name=c["target_data"][0])
    kw = dict(c["transform_kwargs"])
    if kw.get("target_data"):
        kw["target_data"] = ds[kw["target_data"]]
    with pytest.raises(NotImplementedError):
        xg.Grid(ds, **c["grid_kwargs"]).transform(ds[c["source_data"][0]], "Z", target, **kw)

    ds, gk, target, kw, want = _construct("conservative_depth_depth_rename")
    (target_dim,) = target.dims
    assert len(target_dim) > 1
    got = xg.Grid(ds, **gk).transform(ds["data"], "Z", target, target_dim=target_dim, **kw)
    np.testing.assert_allclose(got.values, want, rtol=1e-5, atol=1e-6, equal_nan=True)

    ds, gk, target, kw, _ = _construct("conservative_depth_temp")
    with pytest.warns(UserWarning, match="The `target data` input is not located on the cell bounds"):
        xg.Grid(ds, **gk).transform(ds["data"], "Z", target, **kw)
    warnings.resetwarnings()


@pytest.mark.parametrize("name", MULTIDIM)
def test_grid_transform_names_and_auto_naming(name):
    """test_transform.py:1129-1205: an unnamed input gives an unnamed output; with a numpy target the
    new dimension is named after ``target_data`` (or the axis coordinate)."""
    import warnings

    ds, gk, target, kw, _ = _construct(name)
    grid = xg.Grid(ds, **gk)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        unnamed = ds["data"].copy()
        unnamed.name = None
        assert grid.transform(unnamed, "Z", target, **kw).name is None
        kw2 = dict(kw)
        target_data = kw2.setdefault("target_data", None)
        if target_data is None:
            expected_coord = grid.axes["Z"].coords["center" if kw2["method"] == "linear" else "outer"]
        else:
            expected_coord = target_data.name
        got = grid.transform(ds["data"], "Z", target.values, **kw2)
    assert expected_coord in got.coords


def test_grid_transform_noname_targetdata():
    """test_transform.py:1147-1171"""
    ds, gk, target, kw, _ = _construct("linear_depth_dens")
    target_data = kw.pop("target_data").copy()
    target_data.name = None
    with pytest.warns(UserWarning):
        got = xg.Grid(ds, **gk).transform(ds["data"], "Z", target.values, target_data=target_data, **kw)
    assert "TRANSFORMED_DIMENSION" in got.dims


@pytest.mark.parametrize("name", MULTIDIM)
def test_transform_error_periodic(name):
    """test_transform.py:1174-1185"""
    ds, gk, target, kw, _ = _construct(name)
    with pytest.raises(ValueError):
        xg.Grid(ds, padding="periodic", **gk).transform(ds["data"], "Z", target, **kw)


@pytest.mark.parametrize("bypass_checks", [True, False])
def test_grid_transform_bypass_checks(bypass_checks):
    """test_transform.py:1208-1235"""
    ds, gk, target, kw, want = _construct("linear_depth_dens")
    got = xg.Grid(ds, **gk).transform(ds["data"], "Z", target, bypass_checks=bypass_checks, **kw)
    np.testing.assert_allclose(got.values, want, rtol=1e-5, atol=1e-6, equal_nan=True)


@pytest.mark.parametrize("name", MULTIDIM)
def test_grid_transform_multidim_broadcast(name):
    """test_transform.py:1280-1313: the 1-D column broadcast against another dim gives the 1-D
    result in every column."""
    import warnings

    ds, gk, target, kw, want = _construct(name)
    na = 8
    data_vars = {k: (("a",) + ds[k].dims, np.broadcast_to(ds[k].values, (na,) + ds[k].shape).copy())
                 for k in ds.data_vars}
    coords = {k: ds[k].values for k in ds.dims}
    ds2 = xg.Dataset(data_vars=data_vars, coords=coords)
    target_data = kw.pop("target_data", None)
    if target_data is not None:
        target_data = ds2[target_data.name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = xg.Grid(ds2, **gk).transform(ds2["data"], "Z", target, target_data=target_data, **kw)
    assert got.shape == (na, want.size)
    np.testing.assert_allclose(got.values, np.broadcast_to(want, (na, want.size)), rtol=1e-5, atol=1e-6, equal_nan=True)


def test_grid_transform_other_dims_error_and_input_check():
    """test_transform.py:1339-1430: target_data on a differently named horizontal dim; Datasets
    where DataArrays are expected."""
    ds, gk, target, kw, _ = _construct("linear_depth_dens")
    na = 3
    src = xg.DataArray(ds["data"].values[:, None] * np.ones((1, na)), dims=("depth", "a"), name="data",
                       coords={"depth": ds["depth"].values})
    other = xg.DataArray(kw["target_data"].values[:, None] * np.ones((1, na)), dims=("depth", "a_other"), name="dens")
    grid = xg.Grid(ds, **gk)
    kw2 = dict(kw)
    kw2["target_data"] = other
    with pytest.raises(ValueError):
        grid.transform(src, "Z", target, **kw2)
    with pytest.raises(ValueError, match=r"`da` needs to be a"):
        grid.transform(ds, "Z", target, **kw)
    tds = xg.Dataset(data_vars={"dummy": (target.dims, target.values)})
    with pytest.raises(ValueError, match="needs to be a"):
        grid.transform(ds["data"], "Z", tds, **kw)
    kw3 = dict(kw)
    kw3["target_data"] = xg.Dataset(data_vars={"dummy": (kw["target_data"].dims, kw["target_data"].values)})
    with pytest.raises(ValueError, match="needs to be a"):
        grid.transform(ds["data"], "Z", target, **kw3)
