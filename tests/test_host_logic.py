"""Host-side logic on CPU: the Grid / Axis / grid-ufunc bookkeeping of the package with the CUDA
kernels replaced by the oracle (tests/_mock_backend.py).  The SAME test bodies run against the
real kernels on the GPU box (test_grid_gpu.py / test_transform_gpu.py / test_faces_gpu.py, marker `gpu`)."""

import numpy as np
import pytest
import torch

import test_apply_ufunc_gpu as A
import test_faces_gpu as F
import test_grid_gpu as G
import test_grid_more_gpu as M
import test_transform_gpu as T
import xgcm_b200 as xg
from _mock_backend import install

_NEEDS_REAL_GPU = {
    "test_device_resident_inputs_stay_on_device",  # calls .cuda()
    "test_config1_1e6_fp64_periodic",              # large; kernel-only value
    "test_gridops_raw_ufunc_attribute",
    "test_conservative_many_bins_multipass",       # exercises the kernel's pass structure: nothing to mock
}


@pytest.fixture(autouse=True)
def mock_backend(monkeypatch):
    if torch.cuda.is_available():
        pytest.skip("real kernels available: covered by the gpu-marked modules")
    install(monkeypatch)


for _mod in (G, T, F, A, M):
    for _name in dir(_mod):
        if _name.startswith("test_") and _name not in _NEEDS_REAL_GPU:
            globals()[f"{_name}__hostlogic"] = getattr(_mod, _name)


# ---- pure metadata: no arrays involved ----------------------------------------------------------
def _ds():
    return xg.Dataset(coords={"xc": np.arange(4) + 0.5, "xg": np.arange(4.0), "xo": np.arange(5.0), "yc": np.arange(3.0)})


def test_axis_default_shifts_and_validation():
    """xgcm/axis.py:11-17,126-171; xgcm/test/test_axis.py."""
    ds = _ds()
    ax = xg.Axis(ds, "X", {"center": "xc", "left": "xg", "outer": "xo"})
    assert ax.default_shifts == {"center": "left", "left": "center", "outer": "center"}
    assert ax.fill_value == 0.0 and ax.padding is None and not ax.periodic
    assert xg.Axis(ds, "X", {"center": "xc", "outer": "xo"}).default_shifts["center"] == "outer"
    assert xg.Axis(ds, "X", {"center": "xc", "left": "xg"}, default_shifts={"center": "left"}).default_shifts["center"] == "left"
    with pytest.raises(ValueError, match="Axis position must be one of"):
        xg.Axis(ds, "X", {"middle": "xc"})
    with pytest.raises(ValueError, match="Could not find dimension"):
        xg.Axis(ds, "X", {"center": "nope"})
    with pytest.raises(ValueError, match="multiple positions"):
        xg.Axis(ds, "X", {"center": "xc", "left": "xc"})
    with pytest.raises(ValueError, match="padding must be one of"):
        xg.Axis(ds, "X", {"center": "xc"}, padding="bogus")
    with pytest.raises(ValueError, match="Can't set the default shift"):
        xg.Axis(ds, "X", {"center": "xc"}, default_shifts={"center": "center"})
    with pytest.raises(TypeError):
        xg.Axis(ds, "X", {"center": "xc"}, fill_value="a")
    with pytest.raises(TypeError):
        xg.Axis(ds, 3, {"center": "xc"})
    with pytest.raises(TypeError):
        xg.Axis(np.zeros(3), "X", {"center": "xc"})
    with pytest.raises(ValueError, match="renamed to 'padding'"):
        xg.Axis(ds, "X", {"center": "xc"}, boundary="fill")
    with pytest.raises(AttributeError):
        ax.boundary
    pos, dim = ax._get_position_name(xg.DataArray(np.zeros((3, 5)), dims=("yc", "xo")))
    assert (pos, dim) == ("outer", "xo")
    with pytest.raises(KeyError):
        ax._get_position_name(xg.DataArray(np.zeros(3), dims=("yc",)))
    with pytest.raises(KeyError):
        ax._get_position_name(xg.DataArray(np.zeros((4, 4)), dims=("xc", "xg")))


def test_grid_kwargs_precedence_and_repr():
    """grid.py:229-332: scalar / dict kwargs mapped over axes; per-call beats Axis default."""
    ds = _ds()
    grid = xg.Grid(ds, coords={"X": {"center": "xc", "left": "xg"}, "Y": {"center": "yc"}},
                   padding={"X": "periodic"}, fill_value=None)
    assert grid.axes["X"].padding == "periodic" and grid.axes["Y"].padding is None
    assert grid._complete_user_kwargs_using_axis_defaults(None, "padding") == {"X": "periodic", "Y": None}
    assert grid._complete_user_kwargs_using_axis_defaults("fill", "padding") == {"X": "fill", "Y": "fill"}
    assert grid._complete_user_kwargs_using_axis_defaults({"Y": "extend"}, "padding") == {"X": "periodic", "Y": "extend"}
    assert grid._complete_user_kwargs_using_axis_defaults(None, "fill_value") == {"X": 0.0, "Y": 0.0}
    assert "X Axis (periodic, padding='periodic')" in repr(grid)
    with pytest.raises(ValueError, match="`periodic` argument has been removed"):
        xg.Grid(ds, coords={"X": {"center": "xc"}}, periodic=False)
    with pytest.raises(TypeError, match="unexpected keyword"):
        xg.Grid(ds, coords={"X": {"center": "xc"}}, bogus=1)
    with pytest.raises(ValueError, match="Could not determine Axis names"):
        xg.Grid(ds)
    with pytest.raises(TypeError):
        xg.Grid(np.zeros(3), coords={})
    with pytest.warns(DeprecationWarning):
        xg.Grid(ds, coords={"X": {"center": "xc"}}, fill_value=1.0)
    with pytest.raises(ValueError, match="Face dimension face does not exist"):
        xg.Grid(ds, coords={"X": {"center": "xc"}}, face_connections={"face": {}})
    with pytest.raises(KeyError):
        grid.set_metrics(("Q",), "xc")
    with pytest.raises(KeyError):
        grid.set_metrics(("X",), "missing")


def test_comodo_autoparse():
    """xgcm/comodo.py: axis / c_grid_axis_shift attributes."""
    ds = xg.Dataset(coords={
        "xc": (("xc",), np.arange(4) + 0.5, {"axis": "X"}),
        "xg": (("xg",), np.arange(4.0), {"axis": "X", "c_grid_axis_shift": -0.5}),
        "zc": (("zc",), np.arange(3.0), {"axis": "Z"}),
        "zo": (("zo",), np.arange(4.0), {"axis": "Z", "c_grid_axis_shift": -0.5}),
    })
    grid = xg.Grid(ds)
    assert grid.axes["X"].coords == {"center": "xc", "left": "xg"}
    assert grid.axes["Z"].coords == {"center": "zc", "outer": "zo"}


def test_signature_parsing_and_equivalence():
    """xgcm/test/test_grid_ufunc.py:20-213."""
    from xgcm_b200.grid_ufunc import _GridUFuncSignature as S

    s = S.from_string("(X:center,Y:left),(X:left)->(Y:center),()")
    assert s.in_ax_names == [("X", "Y"), ("X",)] and s.in_ax_positions == [("center", "left"), ("left",)]
    assert s.out_ax_names == [("Y",), ()] and s.out_ax_positions == [("center",), ()]
    assert str(s) == "(X:center,Y:left),(X:left)->(Y:center),()"
    assert S.from_string("( X:center )->( X:left )").equivalent(S.from_string("(Z:center)->(Z:left)"))
    assert not S.from_string("(X:center)->(X:left)").equivalent(S.from_string("(X:center)->(X:right)"))
    assert not S.from_string("(X:center,Y:center)->(X:left)").equivalent(S.from_string("(X:center,X:center)->(X:left)"))
    assert S.from_string("(X:center,Y:center)->(Y:left)").equivalent(S.from_string("(A:center,B:center)->(B:left)"))
    assert not S.from_string("(X:center,Y:center)->(Y:left)").equivalent(S.from_string("(A:center,B:center)->(A:left)"))
    for bad in ("(X:centre)->(X:left)", "X:center->X:left", "(X:center)(X:left)", "(X:center)->", "()->()->()"):
        with pytest.raises(ValueError):
            S.from_string(bad)


def test_as_grid_ufunc_decorator_and_type_hint_signatures():
    from typing import Annotated

    @xg.as_grid_ufunc(signature="(X:center)->(X:left)", padding_width={"X": (1, 0)}, padding="fill", fill_value=2.0)
    def f(a):
        return a[..., 1:] - a[..., :-1]

    assert isinstance(f, xg.GridUFunc) and f.padding == "fill" and f.fill_value == 2.0 and f.pad_before_func

    @xg.as_grid_ufunc()
    def g(a: Annotated[np.ndarray, "X:center"]) -> Annotated[np.ndarray, "X:left"]:
        return a

    assert str(g.signature) == "(X:center)->(X:left)"
    with pytest.raises(ValueError, match="only one of"):
        xg.as_grid_ufunc(signature="(X:center)->(X:left)")(g.ufunc)
    with pytest.raises(ValueError, match="either type hints or signature"):
        xg.as_grid_ufunc()(lambda a: a)
    with pytest.raises(TypeError, match="Unsupported keyword"):
        xg.as_grid_ufunc(signature="(X:center)->(X:left)", junk=1)
    with pytest.raises(ValueError, match="boundary_width"):
        xg.as_grid_ufunc(signature="(X:center)->(X:left)", boundary_width={"X": (1, 0)})
    ds = _ds()
    grid = xg.Grid(ds, coords={"X": {"center": "xc", "left": "xg"}}, padding="periodic")
    da = xg.DataArray(np.arange(4.0), dims=("xc",))
    with pytest.raises(ValueError, match="Must provide a grid"):
        f(None, da, axis=[("X",)])
    with pytest.raises(ValueError, match="Number of entries in `axis`"):
        f(grid, da, axis=[("X",), ("X",)])
    with pytest.raises(ValueError, match="Mismatch between signature"):
        f(grid, xg.DataArray(np.arange(4.0), dims=("xg",)), axis=[("X",)])
    with pytest.raises(NotImplementedError, match="map_overlap"):
        f(grid, da, axis=[("X",)], map_overlap=True)
    got = f(grid, da, axis=[("X",)])  # decorator-bound padding beats the Axis default
    np.testing.assert_array_equal(got.values, np.diff(np.concatenate([[2.0], np.arange(4.0)])))
    got = f(grid, da, axis=[("X",)], padding="periodic")  # per-call beats decorator
    np.testing.assert_array_equal(got.values, np.arange(4.0) - np.roll(np.arange(4.0), 1))


def test_select_grid_ufunc_dispatch():
    """grid.py:1779-1824 + xgcm/test/test_grid_ufunc.py:1341-1449."""
    from xgcm_b200 import gridops
    from xgcm_b200.grid import _select_grid_ufunc
    from xgcm_b200.grid_ufunc import _GridUFuncSignature as S

    gu, rest = _select_grid_ufunc("diff", S.from_string("(Q:center)->(Q:outer)"), gridops, padding="fill")
    assert gu is gridops.diff_center_to_outer and rest == {"padding": "fill"}
    with pytest.raises(NotImplementedError, match="Could not find any pre-defined"):
        _select_grid_ufunc("curl", S.from_string("(X:center)->(X:left)"), gridops)
    with pytest.raises(NotImplementedError, match="with signature"):
        _select_grid_ufunc("interp", S.from_string("(X:left)->(X:right)"), gridops)


def test_labeled_arithmetic_broadcasts_by_name():
    a = xg.DataArray(np.arange(6.0).reshape(2, 3), dims=("y", "x"), name="a")
    b = xg.DataArray(np.array([10.0, 20.0, 30.0]), dims=("x",))
    c = xg.DataArray(np.array([1.0, 2.0]), dims=("y",))
    np.testing.assert_array_equal((a * b).values, a.values * b.values)
    np.testing.assert_array_equal((a / c).values, a.values / c.values[:, None])
    np.testing.assert_array_equal((b * a).values, (a.values * b.values).T)
    assert (b * a).dims == ("x", "y")
    z = xg.DataArray(np.ones(4), dims=("z",))
    assert (a * z).dims == ("y", "x", "z")
    assert a.transpose("x", "y").dims == ("x", "y")
    assert a.isel(x=slice(0, 2)).shape == (2, 2)
    assert a.rename({"x": "xx"}).dims == ("y", "xx")
    with pytest.raises(ValueError):
        a * xg.DataArray(np.ones(5), dims=("x",))
    assert a.equals(a.copy()) and not a.equals(a + 1)


def test_deprecations_match_reference():
    """xgcm/test/test_deprecations.py: removed / renamed arguments and attributes raise the
    reference's messages."""
    ds = _ds()
    coords = {"X": {"center": "xc", "left": "xg"}}
    with pytest.raises(ValueError, match="The `periodic` argument has been removed"):
        xg.Grid(ds, coords=coords, autoparse_metadata=False, periodic=True)
    with pytest.raises(ValueError, match="Argument 'boundary' has been renamed to 'padding'"):
        xg.Grid(ds, coords=coords, autoparse_metadata=False, boundary="periodic")
    with pytest.raises(ValueError, match="Argument 'boundary_width' has been renamed to 'padding_width'"):
        xg.as_grid_ufunc("(X:center)->(X:left)", boundary_width={"X": (1, 0)})
    grid = xg.Grid(ds, coords=coords, autoparse_metadata=False, padding="periodic")
    with pytest.raises(AttributeError, match="Attribute 'boundary' has been renamed to 'padding'"):
        grid.axes["X"].boundary
    gf = xg.as_grid_ufunc("(X:center)->(X:left)")(lambda a: a)
    with pytest.raises(AttributeError, match="Attribute 'boundary' has been renamed to 'padding'"):
        gf.boundary
    with pytest.raises(AttributeError, match="Attribute 'boundary_width' has been renamed to 'padding_width'"):
        gf.boundary_width


def test_axis_reference_cases():
    """xgcm/test/test_axis.py:9-115"""
    ds = _ds()
    axis = xg.Axis(name="X", ds=ds, coords={"center": "xc", "left": "xg"})
    assert axis.name == "X" and axis.coords == {"center": "xc", "left": "xg"}
    assert axis.default_shifts == {"left": "center", "center": "left"}
    assert axis.padding is None
    assert repr(axis).startswith("<xgcm.Axis 'X'")
    axis = xg.Axis(name="foo", ds=ds, coords={"center": "xc", "left": "xg"},
                   default_shifts={"left": "inner", "center": "outer"}, padding="fill")
    assert axis.default_shifts == {"left": "inner", "center": "outer"} and axis.padding == "fill"
    with pytest.raises(ValueError, match="Could not find dimension"):
        xg.Axis(name="X", ds=ds, coords={"center": "lat", "left": "lon"})
    with pytest.raises(ValueError, match="same dimension cannot be assigned to multiple positions"):
        xg.Axis(name="X", ds=ds, coords={"center": "xc", "outer": "xc"})
    with pytest.raises(ValueError, match="Can't set the default"):
        xg.Axis(name="foo", ds=ds, coords={"center": "xc", "left": "xg"},
                default_shifts={"left": "left", "center": "center"})
    with pytest.raises(ValueError, match="padding must be one of"):
        xg.Axis(name="foo", ds=ds, coords={"center": "xc", "left": "xg"}, padding="blargh")
    da = xg.DataArray(np.zeros((3, 4)), dims=("yc", "xg"))
    axis = xg.Axis(name="X", ds=ds, coords={"center": "xc", "left": "xg"})
    assert axis._get_position_name(da) == ("left", "xg")
    assert axis._get_axis_dim_num(da) == da.get_axis_num("xg") == 1


# ---- xgcm/test/test_grid_ufunc.py:20-213, the parametrized lists verbatim -------------------------
@pytest.mark.parametrize(
    "sig_str, exp_in_ax_names, exp_in_ax_pos, exp_out_ax_names, exp_out_ax_pos",
    [
        ("()->()", [()], [()], [()], [()]),
        ("(X:center)->()", [("X",)], [()], [("center",)], [()]),
        ("()->(X:left)", [()], [("X",)], [()], [("left",)]),
        ("(X:center)->(X:left)", [("X",)], [("X",)], [("center",)], [("left",)]),
        ("(X:left)->(Y:center)", [("X",)], [("Y",)], [("left",)], [("center",)]),
        ("(X:left),(X:right)->(Y:center)", [("X",), ("X",)], [("Y",)], [("left",), ("right",)], [("center",)]),
        ("(X:center)->(Y:inner),(Y:outer)", [("X",)], [("Y",), ("Y",)], [("center",)], [("inner",), ("outer",)]),
        ("(X:center,Y:center)->(Z:center)", [("X", "Y")], [("Z",)], [("center", "center")], [("center",)]),
    ],
)
def test_parse_valid_signatures(sig_str, exp_in_ax_names, exp_out_ax_names, exp_in_ax_pos, exp_out_ax_pos):
    """test_grid_ufunc.py:21-67 (the reference's argument order: the 2nd / 4th lists are names / positions
    of the INPUTS, the 3rd / 5th of the outputs) and :84-103 (round trip through ``str``)."""
    from xgcm_b200.grid_ufunc import _GridUFuncSignature, _parse_signature_from_string

    in_ax_names, out_ax_names, in_ax_pos, out_ax_pos = _parse_signature_from_string(sig_str)
    assert in_ax_names == exp_in_ax_names
    assert in_ax_pos == exp_in_ax_pos
    assert out_ax_names == exp_out_ax_names
    assert out_ax_pos == exp_out_ax_pos
    assert str(_GridUFuncSignature.from_string(sig_str)) == sig_str


@pytest.mark.parametrize("signature", ["(x:left)(y:left)->()", "(x:left),(y:left)->", "((x:left))->(x:left)",
                                       "(x:left)->(x:left),(i)->(i)", "(X:centre)->()"])
def test_invalid_signatures(signature):
    """test_grid_ufunc.py:69-82"""
    from xgcm_b200.grid_ufunc import _parse_signature_from_string

    with pytest.raises(ValueError):
        _parse_signature_from_string(signature)


def test_signatures_from_type_hints():
    """test_grid_ufunc.py:106-213"""
    from typing import Annotated, Tuple

    with pytest.raises(ValueError, match="Must specify axis positions"):

        @xg.as_grid_ufunc()
        def nothing(): ...

    def sig_of(fn):
        return str(xg.as_grid_ufunc()(fn).signature)

    def f1(a: Annotated[np.ndarray, "X:center"]) -> Annotated[np.ndarray, "X:center"]: ...
    def f2(a: Annotated[np.ndarray, "X:center,Y:center"]) -> Annotated[np.ndarray, "X:center"]: ...
    def f3(a: Annotated[np.ndarray, "X:left"], b: Annotated[np.ndarray, "Y:right"]) -> Annotated[np.ndarray, "X:center"]: ...
    def f4(a: Annotated[np.ndarray, "X:center"]) -> Annotated[np.ndarray, "X:left,Y:right"]: ...
    def f5(a: Annotated[np.ndarray, "X:center"]) -> Tuple[Annotated[np.ndarray, "X:left"], Annotated[np.ndarray, "Y:right"]]: ...

    assert sig_of(f1) == "(X:center)->(X:center)"
    assert sig_of(f2) == "(X:center,Y:center)->(X:center)"
    assert sig_of(f3) == "(X:left),(Y:right)->(X:center)"
    assert sig_of(f4) == "(X:center)->(X:left,Y:right)"
    assert sig_of(f5) == "(X:center)->(X:left),(Y:right)"
    with pytest.raises(ValueError, match="only one of either type hints or signature kwarg"):
        xg.as_grid_ufunc(signature="(X:center)->(X:left)")(f1)
