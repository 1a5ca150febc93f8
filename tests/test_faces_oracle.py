"""Pin oracle/faces.py (face-connection padding) against the reference's own tests.

xgcm/test/test_padding.py:172-1205 builds the expected padded arrays BY HAND — pad each face on
its unconnected sides, cut the rim of the neighbour, flip / rename / negate it as the connection
demands, concatenate — and compares them with ``pad``.  Those constructions are transcribed here
call for call on the small named-array class of the oracle (``pad`` / ``isel`` / ``rename`` /
``flip`` / ``concat`` stand for the xarray methods of the same names) and compared with the
oracle's restatement of the reference ALGORITHM (padding.py:260-572).  CPU only.
"""

import numpy as np
import pytest

from oracle.faces import Named, concat, pad_face_connections, _swap_dimension_names

N = 7
AXES = {"X": ("x", "xl"), "Y": ("y", "yl")}

PADDING_WIDTHS = [
    {"X": (1, 1)},
    {"X": (1, 2)},
    {"X": (0, 1)},
    {"X": (1, 1), "Y": (1, 1)},
    {"X": (2, 2), "Y": (2, 2)},
    {"X": (0, 1), "Y": (1, 0)},
    {"X": (0, 2), "Y": (1, 0)},
]


def _ds(seed=0):
    """xgcm/test/test_faceconnections.py:10-36 (N reduced, as its TODO suggests)."""
    rng = np.random.default_rng(seed)
    return {
        "data_c": Named(rng.random((2, N, N)), ("face", "y", "x")),
        "u": Named(rng.random((2, N, N)), ("face", "xl", "y")),
        "v": Named(rng.random((2, N, N)), ("face", "x", "yl")),
    }


def _pad2(face: Named, x, y, wx, wy, fill_value):
    return face.pad(x, wx, "constant", constant_values=fill_value).pad(y, wy, "constant", constant_values=fill_value)


def _tail(w):
    # slice(-w, None if w > 0 else 0): nothing at all for w == 0 (test_padding.py:375-381)
    return slice(-w, None if w > 0 else 0)


def _run(da: Named, links, pw, fill_value, vector_axis=None, partner: Named = None):
    pw = dict(pw)
    out = pad_face_connections(
        da.data, da.dims, AXES, "face", links, pw,
        padding={"X": "fill", "Y": "fill"}, fill_value={"X": fill_value, "Y": fill_value},
        vector_axis=vector_axis,
        partner=None if partner is None else partner.data,
        partner_dims=None if partner is None else partner.dims,
    )
    return out


def _assert(result, expected: Named, dims):
    np.testing.assert_allclose(result, expected.transpose(*dims).data, equal_nan=True)


# ------------------------------------------------------------------ prepad helpers (test_padding.py:172-320)
def _prepad_right_left_same_axis(da, pw, fv, x="x", y="y"):
    f0, f1 = da.isel(face=0), da.isel(face=1)
    return (_pad2(f0, x, y, (pw["X"][0], 0), pw["Y"], fv), _pad2(f1, x, y, (0, pw["X"][1]), pw["Y"], fv))


def _prepad_right_right_same_axis(da, pw, fv, x="x", y="y"):
    f0, f1 = da.isel(face=0), da.isel(face=1)
    return (_pad2(f0, x, y, (pw["X"][0], 0), pw["Y"], fv), _pad2(f1, x, y, (pw["X"][0], 0), pw["Y"], fv))


def _prepad_right_left_swap_axis(da, pw, fv, x="x", y="y"):
    f0, f1 = da.isel(face=0), da.isel(face=1)
    return (
        _pad2(f0, x, y, (pw["X"][0], 0), pw["Y"], fv),
        _pad2(f1, x, y, pw["X"], (0, pw["Y"][1]), fv),
        _pad2(f0, x, y, (pw["Y"][0], 0), (pw["X"][1], pw["X"][0]), fv),
        _pad2(f1, x, y, (pw["Y"][1], pw["Y"][0]), (0, pw["X"][1]), fv),
    )


def _prepad_right_right_swap_axis(da, pw, fv, x="x", y="y"):
    f0, f1 = da.isel(face=0), da.isel(face=1)
    return (
        _pad2(f0, x, y, (pw["X"][0], 0), pw["Y"], fv),
        _pad2(f1, x, y, pw["X"], (pw["Y"][0], 0), fv),
        _pad2(f0, x, y, (pw["Y"][0], 0), pw["X"], fv),
        _pad2(f1, x, y, pw["Y"], (pw["X"][0], 0), fv),
    )


@pytest.mark.parametrize("fill_value", [np.nan, 0])
@pytest.mark.parametrize("padding_width", PADDING_WIDTHS)
class TestPaddingFaceConnection:
    def test_face_connections_right_left_same_axis(self, padding_width, fill_value):
        """test_padding.py:343-396"""
        links = {0: {"X": (None, (1, "X", False))}, 1: {"X": ((0, "X", False), None)}}
        data = _ds()["data_c"]
        pw = dict(padding_width)
        pw["Y"] = pw.get("Y", (0, 0))
        f0p, f1p = _prepad_right_left_same_axis(data, pw, fill_value)
        f0e = concat([f0p, f1p.isel(x=slice(0, pw["X"][1]))], "x")
        f1e = concat([f0p.isel(x=_tail(pw["X"][0])), f1p], "x")
        expected = concat([f0e, f1e], "face")
        _assert(_run(data, links, pw, fill_value), expected, data.dims)

    def test_face_connections_right_right_same_axis(self, padding_width, fill_value):
        """test_padding.py:398-460"""
        links = {0: {"X": (None, (1, "X", True))}, 1: {"X": (None, (0, "X", True))}}
        data = _ds()["data_c"]
        pw = dict(padding_width)
        pw["Y"] = pw.get("Y", (0, 0))
        f0p, f1p = _prepad_right_right_same_axis(data, pw, fill_value)
        f0a = f1p.isel(x=_tail(pw["X"][1])).flip("x")
        f1a = f0p.isel(x=_tail(pw["X"][1])).flip("x")
        expected = concat([concat([f0p, f0a], "x"), concat([f1p, f1a], "x")], "face")
        _assert(_run(data, links, pw, fill_value), expected, data.dims)

    def test_face_connections_right_left_swap_axis(self, padding_width, fill_value):
        """test_padding.py:462-535"""
        links = {0: {"X": (None, (1, "Y", False))}, 1: {"Y": ((0, "X", False), None)}}
        data = _ds()["data_c"].isel(y=slice(0, 2), x=slice(0, 2))
        pw = dict(padding_width)
        pw["Y"] = pw.get("Y", (0, 0))
        f0p, f1p, f0s, f1s = _prepad_right_left_swap_axis(data, pw, fill_value)
        f0a = _swap_dimension_names(f1s.isel(y=slice(0, pw["X"][1])).flip("x"), "y", "x")
        f1a = _swap_dimension_names(f0s.isel(x=_tail(pw["Y"][0])).flip("y"), "x", "y")
        expected = concat([concat([f0p, f0a], "x"), concat([f1a, f1p], "y")], "face")
        _assert(_run(data, links, pw, fill_value), expected, data.dims)

    def test_face_connections_right_right_swap_axis(self, padding_width, fill_value):
        """test_padding.py:537-616"""
        pw = {k: padding_width.get(k, (0, 0)) for k in ["X", "Y"]}
        links = {0: {"X": (None, (1, "Y", True))}, 1: {"Y": (None, (0, "X", True))}}
        data = _ds()["data_c"].isel(y=slice(0, 3), x=slice(0, 3))
        f0p, f1p, f0s, f1s = _prepad_right_right_swap_axis(data, pw, fill_value)
        f0a = _swap_dimension_names(f1s.isel(y=_tail(pw["X"][1])).flip("y"), "y", "x")
        f1a = _swap_dimension_names(f0s.isel(x=_tail(pw["Y"][1])).flip("x"), "x", "y")
        expected = concat([concat([f0p, f0a], "x"), concat([f1p, f1a], "y")], "face")
        _assert(_run(data, links, pw, fill_value), expected, data.dims)

    def test_vector_face_connections_right_left_same_axis(self, padding_width, fill_value):
        """test_padding.py:618-713: a same-axis, non-reversed seam moves each component as is."""
        links = {0: {"X": (None, (1, "X", False))}, 1: {"X": ((0, "X", False), None)}}
        ds = _ds()
        u, v = ds["u"], ds["v"]
        pw = dict(padding_width)
        pw["Y"] = pw.get("Y", (0, 0))
        u0, u1 = _prepad_right_left_same_axis(u, pw, fill_value, x="xl", y="y")
        v0, v1 = _prepad_right_left_same_axis(v, pw, fill_value, x="x", y="yl")
        u_exp = concat([concat([u0, u1.isel(xl=slice(0, pw["X"][1]))], "xl"),
                        concat([u0.isel(xl=_tail(pw["X"][0])), u1], "xl")], "face")
        v_exp = concat([concat([v0, v1.isel(x=slice(0, pw["X"][1]))], "x"),
                        concat([v0.isel(x=_tail(pw["X"][0])), v1], "x")], "face")
        _assert(_run(u, links, pw, fill_value, "X", v), u_exp, u.dims)
        _assert(_run(v, links, pw, fill_value, "Y", u), v_exp, v.dims)

    def test_vector_face_connections_right_left_swap_axis(self, padding_width, fill_value):
        """test_padding.py:824-937: across a rotated seam u is fed by v and vice versa, with the
        tangential flip and the sign change on the component it applies to."""
        links = {0: {"X": (None, (1, "Y", False))}, 1: {"Y": ((0, "X", False), None)}}
        ds = _ds()
        u, v = ds["u"], ds["v"]
        pw = dict(padding_width)
        pw["Y"] = pw.get("Y", (0, 0))
        u0, u1, u0s, u1s = _prepad_right_left_swap_axis(u, pw, fill_value, x="xl", y="y")
        v0, v1, v0s, v1s = _prepad_right_left_swap_axis(v, pw, fill_value, x="x", y="yl")

        u0a = v1s.isel(yl=slice(0, pw["X"][1])).rename({"x": "xl", "yl": "y"}).flip("xl")
        u0a = _swap_dimension_names(u0a, "y", "xl")
        u1a = v0s.isel(x=_tail(pw["Y"][0])).rename({"x": "xl", "yl": "y"}).flip("y").neg()
        u1a = _swap_dimension_names(u1a, "y", "xl")
        v0a = u1s.isel(y=slice(0, pw["X"][1])).rename({"xl": "x", "y": "yl"}).flip("x").neg()
        v0a = _swap_dimension_names(v0a, "yl", "x")
        v1a = u0s.isel(xl=_tail(pw["Y"][0])).rename({"xl": "x", "y": "yl"}).flip("yl")
        v1a = _swap_dimension_names(v1a, "yl", "x")

        u_exp = concat([concat([u0, u0a], "xl"), concat([u1a, u1], "y")], "face")
        v_exp = concat([concat([v0, v0a], "x"), concat([v1a, v1], "yl")], "face")
        _assert(_run(u, links, pw, fill_value, "X", v), u_exp, u.dims)
        _assert(_run(v, links, pw, fill_value, "Y", u), v_exp, v.dims)


def test_seam_values_x_to_y():
    """xgcm/test/test_faceconnections.py:183-202: the halo row of face 1 below y=0 is the last
    column of face 0 read backwards."""
    links = {0: {"X": (None, (1, "Y", False))}, 1: {"Y": ((0, "X", False), None)}}
    d = _ds()["data_c"]
    out = _run(d, links, {"Y": (1, 0)}, 0.0)
    np.testing.assert_array_equal(out[1, 0, :], d.data[0, ::-1, -1])
    np.testing.assert_array_equal(out[0, 0, :], np.zeros(N))  # unconnected: fill
    np.testing.assert_array_equal(out[:, 1:, :], d.data)
