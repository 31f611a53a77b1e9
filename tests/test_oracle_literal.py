"""Pins the oracle's closed-form spec path against a literal emulation of the reference's GLSL text
(oracle/glsl_literal.py).  The reference ships no tests or golden vectors, so this is
the strongest available pin: two independent restatements of SRC.h / LIB.h must agree."""
import numpy as np
import pytest

from oracle import glsl_literal as L


def _setup(O, S, n, w, h, theta, seed, sh=True, scale_mult=1.0, obj=None):
    cl = S.make_cloud(n, seed, sh=sh, scale_mult=scale_mult)
    fr = S.orbit_frame(w, h, theta)
    if obj is not None:
        fr = S.Frame(w, h, fr.view, fr.proj, S.colmajor(obj), S.colmajor(np.linalg.inv(obj)))
    cam = O.camera_from_view(fr.view)
    origin = cl.barycentre()
    return cl, fr, cam, origin


@pytest.mark.parametrize("sh_order", [0, 1, 2, 3])
def test_records_match_literal_glsl(oracle, scene, sh_order):
    O, S = oracle, scene
    cl, fr, cam, origin = _setup(O, S, 300, 640, 360, 33.0, 11, scale_mult=3.0)
    F = O.make_frame(fr, cam, origin, sh_order)
    pr = O.project(F, cl)
    checked = 0
    for i in range(cl.n):
        q = L.splat_quad(i, cl, fr, cam, origin, sh_order)
        if q is None:
            assert pr["vis"][i] == 0
            continue
        if not pr["vis"][i]:
            continue  # the spec additionally culls empty pixel rectangles / alpha < 1/255
        r = pr["recs"][i]
        assert abs(r["cx"] - q["centre"][0]) < 2e-3 and abs(r["cy"] - q["centre"][1]) < 2e-3
        # spec: q = M d ; literal: window = centre + qx*ax + qy*ay  =>  M = inv([ax ay])
        A = np.array([[q["ax"][0], q["ay"][0]], [q["ax"][1], q["ay"][1]]])
        Mi = np.linalg.inv(A)
        M = np.array([[r["m00"], r["m01"]], [r["m10"], r["m11"]]], np.float64)
        assert np.allclose(M, Mi, rtol=2e-3, atol=1e-6), (i, M, Mi)
        assert np.allclose([r["r"], r["g"], r["b"]], q["color"], rtol=1e-4, atol=2e-5), i
        assert r["alpha"] == np.float32(q["opacity"])
        checked += 1
    assert checked > 200


def test_records_match_literal_with_object_matrix(oracle, scene):
    """Formulas are followed literally even for OBJ-level transforms."""
    O, S = oracle, scene
    obj = np.eye(4)
    th = 0.4
    obj[:3, :3] = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]]) * 1.2
    obj[:3, 3] = [0.1, -0.2, 0.05]
    cl, fr, cam, origin = _setup(O, S, 120, 320, 180, -70.0, 5, scale_mult=3.0, obj=obj)
    F = O.make_frame(fr, cam, origin, 3)
    pr = O.project(F, cl)
    n = 0
    for i in range(cl.n):
        q = L.splat_quad(i, cl, fr, cam, origin, 3)
        if q is None or not pr["vis"][i]:
            continue
        r = pr["recs"][i]
        A = np.array([[q["ax"][0], q["ay"][0]], [q["ax"][1], q["ay"][1]]])
        M = np.array([[r["m00"], r["m01"]], [r["m10"], r["m11"]]], np.float64)
        assert np.allclose(M, np.linalg.inv(A), rtol=2e-3, atol=1e-6)
        assert np.allclose([r["r"], r["g"], r["b"]], q["color"], rtol=1e-4, atol=2e-5)
        n += 1
    assert n > 50


@pytest.mark.parametrize("sh_order,theta", [(0, 0.0), (3, 140.0)])
def test_frame_matches_literal_glsl(oracle, scene, sh_order, theta):
    """Whole-frame parity: literal VS+raster+FS+ROP vs the oracle's tiled spec path, no early-out."""
    O, S = oracle, scene
    cl, fr, cam, origin = _setup(O, S, 400, 96, 54, theta, 3, scale_mult=4.0)
    F = O.make_frame(fr, cam, origin, sh_order, eps_t=0.0)
    p = O.pipeline(F, cl)
    lit, unsafe = L.render(cl, fr, cam, origin, sh_order, p["order"])
    diff = np.abs(lit - p["rgba"].astype(np.float64))
    diff[unsafe] = 0.0
    assert unsafe.mean() < 0.05
    assert diff.max() < 1e-4, diff.max()
    assert p["rgba"][..., 3].max() > 0.5  # the scene actually covers pixels


def test_early_out_error_bound(oracle, scene):
    """Stopping at T < eps changes any channel by at most eps * max(rgb)."""
    O, S = oracle, scene
    cl, fr, cam, origin = _setup(O, S, 30000, 160, 90, 10.0, 9, scale_mult=2.0)
    full = O.pipeline(O.make_frame(fr, cam, origin, 3, eps_t=0.0), cl)
    fast = O.pipeline(O.make_frame(fr, cam, origin, 3, eps_t=1e-5), cl)
    rgbmax = max(1.0, float(np.max([full["recs"]["r"].max(), full["recs"]["g"].max(), full["recs"]["b"].max()])))
    assert np.abs(full["rgba"] - fast["rgba"]).max() <= 1e-5 * rgbmax * 1.01
    assert fast["n_consumed"] < full["n_consumed"]
