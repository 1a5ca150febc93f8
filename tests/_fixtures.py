"""Synthetic datasets for the Grid-level tests (same structure as the reference's
xgcm/test/datasets.py fixtures, rebuilt with xgcm_b200's labelled arrays)."""

import numpy as np

import xgcm_b200 as xg


def grid_metric_dataset(grid_type="C", seed=0, dtype=np.float64):
    """4 x 5 x 10 x 6 (x, y, time, z) B-/C-grid with dx, dy, dz, area, volume metrics whose areas
    are deliberately NOT dx*dy (cf. xgcm/test/datasets.py:554-724)."""
    rng = np.random.default_rng(seed)
    xt = np.arange(4.0)
    xu = xt + 0.5
    yt = np.arange(5.0)
    yu = yt + 0.5
    zt = np.arange(6.0)
    zw = zt + 0.5
    t = np.arange(10.0)

    def gen():
        return rng.random((4, 5, 10, 6)).astype(dtype)

    def hz(val):
        return np.full((4, 5), val, dtype=dtype)

    dims_t = ("xt", "yt", "time", "zt")
    data_vars = {
        "tracer": (dims_t, gen()),
        "wt": (("xt", "yt", "time", "zw"), gen()),
        "timeseries": (("time",), rng.random(10).astype(dtype)),
    }
    if grid_type == "B":
        data_vars["u"] = (("xu", "yu", "time", "zt"), gen())
        data_vars["v"] = (("xu", "yu", "time", "zt"), gen())
    else:
        data_vars["u"] = (("xu", "yt", "time", "zt"), gen())
        data_vars["v"] = (("xt", "yu", "time", "zt"), gen())
    dx, dy, dz = 0.3, 2.0, 20.0
    m = {
        "dx_ne": (("xu", "yu"), hz(dx - 0.1)), "dx_n": (("xt", "yu"), hz(dx - 0.2)),
        "dx_e": (("xu", "yt"), hz(dx - 0.3) + 1.0), "dx_t": (("xt", "yt"), hz(dx - 0.4) + 1.0),
        "dy_ne": (("xu", "yu"), hz(dy + 0.1)), "dy_n": (("xt", "yu"), hz(dy + 0.2)),
        "dy_e": (("xu", "yt"), hz(dy + 0.3)), "dy_t": (("xt", "yt"), hz(dy + 0.4)),
        "dz_t": (dims_t, gen() * dz + 1), "dz_w": (("xt", "yt", "time", "zw"), gen() * dz + 1),
        "dz_w_ne": (("xu", "yu", "time", "zw"), gen() * dz + 1),
        "dz_w_n": (("xt", "yu", "time", "zw"), gen() * dz + 1),
        "dz_w_e": (("xu", "yt", "time", "zw"), gen() * dz + 1),
    }
    # make the horizontal metrics non-uniform so that broadcasting mistakes show up
    for k in ("dx_ne", "dx_n", "dx_e", "dx_t", "dy_ne", "dy_n", "dy_e", "dy_t"):
        m[k] = (m[k][0], (m[k][1] + rng.random((4, 5)) * 0.05).astype(dtype))
    m["area_ne"] = (("xu", "yu"), m["dx_ne"][1] * m["dy_ne"][1] + 0.1)
    m["area_n"] = (("xt", "yu"), m["dx_n"][1] * m["dy_n"][1] + 0.2)
    m["area_e"] = (("xu", "yt"), m["dx_e"][1] * m["dy_e"][1] + 0.3)
    m["area_t"] = (("xt", "yt"), m["dx_t"][1] * m["dy_t"][1] + 0.4)
    m["volume_t"] = (dims_t, (m["dx_t"][1] * m["dy_t"][1])[:, :, None, None] * m["dz_t"][1] + 0.25)
    coords = {"xt": xt, "xu": xu, "yt": yt, "yu": yu, "zt": zt, "zw": zw, "time": t}
    coords.update(m)
    ds = xg.Dataset(data_vars=data_vars, coords=coords)
    grid_coords = {
        "X": {"center": "xt", "right": "xu"},
        "Y": {"center": "yt", "right": "yu"},
        "Z": {"center": "zt", "right": "zw"},
    }
    metrics = {
        ("X",): ["dx_t", "dx_n", "dx_e", "dx_ne"],
        ("Y",): ["dy_t", "dy_n", "dy_e", "dy_ne"],
        ("Z",): ["dz_t", "dz_w", "dz_w_ne", "dz_w_n", "dz_w_e"],
        ("X", "Y"): ["area_t", "area_n", "area_e", "area_ne"],
        ("X", "Y", "Z"): ["volume_t"],
    }
    return ds, grid_coords, metrics


def all_positions_1d(n=9, seed=0, dtype=np.float64):
    """One axis exposing all five positions (cf. xgcm/test/test_grid_ufunc.py:216-297)."""
    rng = np.random.default_rng(seed)
    coords = {
        "x_c": np.arange(n) + 0.5, "x_g": np.arange(n) + 0.0, "x_r": np.arange(n) + 1.0,
        "x_i": np.arange(1, n) + 0.0, "x_o": np.arange(n + 1) + 0.0,
    }
    data_vars = {f"a_{p}": ((f"x_{p}",), rng.random(len(coords[f"x_{p}"])).astype(dtype)) for p in "cgrio"}
    ds = xg.Dataset(data_vars=data_vars, coords=coords)
    grid_coords = {"X": {"center": "x_c", "left": "x_g", "right": "x_r", "inner": "x_i", "outer": "x_o"}}
    return ds, grid_coords


def all_positions_3d(shape=(6, 7, 8), seed=0, dtype=np.float32):
    """(Z, Y, X) grid with all five positions on every axis."""
    rng = np.random.default_rng(seed)
    coords, gc = {}, {}
    for ax, n in zip("ZYX", shape):
        a = ax.lower()
        coords.update({
            f"{a}_c": np.arange(n) + 0.5, f"{a}_g": np.arange(n) + 0.0, f"{a}_r": np.arange(n) + 1.0,
            f"{a}_i": np.arange(1, n) + 0.0, f"{a}_o": np.arange(n + 1) + 0.0,
        })
        gc[ax] = {"center": f"{a}_c", "left": f"{a}_g", "right": f"{a}_r", "inner": f"{a}_i", "outer": f"{a}_o"}
    ds = xg.Dataset(coords=coords)
    return ds, gc, rng


POS_SUFFIX = {"center": "c", "left": "g", "right": "r", "inner": "i", "outer": "o"}
