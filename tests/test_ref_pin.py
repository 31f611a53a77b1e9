"""Pins the oracle to the REFERENCE'S OWN SHADER TEXT: oracle/_ref/libgsplat_ref.so is the unmodified GLSL of
/root/reference/gsplat_plugin/shaders/{GSplatShaderCoreLib.h, GSplatShaderSource.h} compiled as C++ (oracle/build_ref.py,
oracle/glsl_cxx.h) behind an ideal rasteriser + the blend state of GSplatRenderer.C:613-621 (oracle/ref_harness.cpp).

The GLSL leaves evaluation order and rounding to the driver, so the comparison is within a float tolerance, not bitwise:
per-splat geometry to 1e-3 relative, colours to 2e-5, whole frames to 2e-5 on every pixel that is not within 2e-5 (in the
quad's own coordinates) of a support edge or of the 1/255 discard ring — there one ulp decides coverage and two valid
evaluations of the same shader differ by up to alpha; those pixels are counted and bounded, never silently dropped.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
import make_golden as G  # noqa: E402

from oracle import ref as RF  # noqa: E402

pytestmark = pytest.mark.skipif(not RF.available(), reason="oracle/_ref not built and /root/reference absent")

UNSAFE_TOL = 2e-5


def _bound(O, S, n, w, h, theta, seed, sh_order, sh=True, mult=1.0, obj=None, cam=None, eps=0.0):
    cl = S.make_cloud(n, seed, sh=sh, scale_mult=mult)
    fr = S.orbit_frame(w, h, theta)
    if obj is not None:
        fr = S.Frame(w, h, fr.view, fr.proj, S.colmajor(obj), S.colmajor(np.linalg.inv(obj)))
    cam = O.camera_from_view(fr.view) if cam is None else np.asarray(cam, np.float32)
    origin = cl.barycentre()
    F = O.make_frame(fr, cam, origin, sh_order if sh else 0, eps_t=eps)
    p = O.pipeline(F, cl)
    b = RF.Bound(cl, fr, cam, origin, sh_order)
    return cl, fr, F, p, b


def _check_vertices(p, b, w, h, min_checked):
    """Every visible splat: centre, the affine map M of its quad, colour and opacity from the reference's vertex shader."""
    v = b.vertices().astype(np.float64)                     # [n, 6, 12] in z-order-texture order
    order = b.order
    gw = v[:, :, 3:4]
    win = (v[:, :, 0:2] / gw * 0.5 + 0.5) * np.array([w, h])  # window coordinates of the 6 vertices
    checked = 0
    for k in range(order.shape[0]):
        i = int(order[k])
        if not p["vis"][i]:
            continue
        r = p["recs"][i]
        assert v[k, 0, 3] > 0
        # CalculateQuadPos: vertex 0 = (2,-2), 1 = (2,2), 2 = (-2,-2), 3 = (-2,2), 4 = vertex 2, 5 = vertex 1
        assert np.array_equal(v[k, :, 4:6], [[2, -2], [2, 2], [-2, -2], [-2, 2], [-2, -2], [2, 2]])
        c = 0.5 * (win[k, 1] + win[k, 2])
        ax = (win[k, 0] - win[k, 2]) / 4.0
        ay = (win[k, 1] - win[k, 0]) / 4.0
        scale = max(1.0, float(np.abs(c).max()))
        assert abs(r["cx"] - c[0]) < 2e-6 * scale * 16 + 1e-3 and abs(r["cy"] - c[1]) < 2e-6 * scale * 16 + 1e-3, (i, r, c)
        M = np.array([[r["m00"], r["m01"]], [r["m10"], r["m11"]]], np.float64)
        Mi = np.linalg.inv(np.array([[ax[0], ay[0]], [ax[1], ay[1]]]))
        assert np.allclose(M, Mi, rtol=2e-3, atol=2e-6), (i, M, Mi)
        assert np.allclose([r["r"], r["g"], r["b"]], v[k, 0, 8:11], rtol=1e-4, atol=2e-5), (i, r, v[k, 0, 8:11])
        assert np.float32(r["alpha"]) == np.float32(v[k, 0, 11])
        # all six vertices carry the centre's z and w (one depth per quad) and the same colour / opacity
        assert np.all(v[k, :, 2] == v[k, 0, 2]) and np.all(v[k, :, 3] == v[k, 0, 3])
        assert np.all(v[k, :, 8:12] == v[k, 0, 8:12])
        checked += 1
    assert checked >= min_checked
    return checked


@pytest.mark.parametrize("sh_order", [0, 1, 2, 3])
def test_vertex_shader_matches_oracle_records(oracle, scene, sh_order):
    cl, fr, F, p, b = _bound(oracle, scene, 2000, 640, 360, 33.0, 11, sh_order, mult=3.0)
    _check_vertices(p, b, 640, 360, 1500)


def test_vertex_shader_with_object_matrix_and_explicit_camera(oracle, scene):
    obj = np.eye(4); th = 0.4
    obj[:3, :3] = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]]) * 1.2
    obj[:3, 3] = [0.1, -0.2, 0.05]
    cl, fr, F, p, b = _bound(oracle, scene, 1500, 320, 180, -70.0, 5, 3, mult=3.0, obj=obj, cam=[0.5, 1.0, 2.0])
    _check_vertices(p, b, 320, 180, 500)


def test_behind_camera_is_degenerate_in_the_reference_too(oracle, scene):
    """clip.w <= 0 -> gl_Position = 0 and opacity 0 (SRC.h:209-214); the oracle culls exactly those."""
    O, S = oracle, scene
    cl = S.make_cloud(500, 8, sh=False)
    cl.pos[::2, 2] += np.float32(6.0)
    fr = S.orbit_frame(200, 120, 0.0)
    cam = O.camera_from_view(fr.view)
    p = O.pipeline(O.make_frame(fr, cam, cl.barycentre(), 0), cl)
    b = RF.Bound(cl, fr, cam, cl.barycentre(), 0)
    v = b.vertices()
    behind = np.all(v[:, :, 0:4] == 0, axis=(1, 2))
    assert behind.sum() >= 200
    assert not p["vis"][b.order[behind]].any()
    assert np.all(v[behind][:, :, 11] == 0)


def _compare_frames(ref_rgba, unsafe, other, tol, what):
    d = np.abs(ref_rgba.astype(np.float64) - other.astype(np.float64)).max(axis=2)
    assert unsafe.mean() < 0.05, (what, unsafe.mean())
    assert d[~unsafe].max() <= tol, (what, float(d[~unsafe].max()))
    over = d > 1e-3
    assert not (over & ~unsafe).any()
    assert over.sum() <= 0.002 * over.size, (what, int(over.sum()))        # isolated edge pixels, all flagged
    return float(d[~unsafe].max()), int(over.sum())


@pytest.mark.parametrize("name", list(G.CASES))
def test_oracle_frame_matches_reference_glsl(oracle, scene, name):
    """Whole frames of the golden scenes, reference semantics (no early termination): the reference's shaders through an
    ideal rasteriser and the ROP blend equation vs the oracle's tiled spec path."""
    n, seed, sh, mult, w, h, theta, order = G.CASES[name]
    cl, fr, F, p, b = _bound(oracle, scene, n, w, h, theta, seed, order, sh=sh, mult=mult, eps=0.0)
    # argsortByDistance sorts every splat; restricted to the visible ones it is the spec's depth order
    assert np.array_equal(b.order[p["vis"][b.order] > 0], p["order"][:p["n_visible"]])
    rgba, st, unsafe = b.draw(unsafe_tol=UNSAFE_TOL)
    assert st["fragments_kept"] > 10 * w * h
    _compare_frames(rgba, unsafe, p["rgba"], 2e-5, name)
    # with the default early-out (T < 1e-5) the frame moves by at most eps * max(rgb)
    F2 = oracle.make_frame(fr, F.cam[:], cl.barycentre(), order if sh else 0, eps_t=1e-5)
    p2 = oracle.pipeline(F2, cl)
    rgbmax = max(1.0, float(max(p2["recs"]["r"].max(), p2["recs"]["g"].max(), p2["recs"]["b"].max())))
    _compare_frames(rgba, unsafe, p2["rgba"], 2e-5 + 1e-5 * rgbmax, name + " eps 1e-5")


@pytest.mark.parametrize("func", [1, 2])
def test_depth_test_matches_reference_draw(oracle, scene, func):
    """Depth test on / depth writes off against an existing scene depth (R.C:608-610), one depth per quad."""
    O, S = oracle, scene
    cl = S.make_cloud(4000, 21, sh=True, scale_mult=2.0)
    fr = S.orbit_frame(160, 96, 15.0)
    cam = O.camera_from_view(fr.view); origin = cl.barycentre()
    F = O.make_frame(fr, cam, origin, 3, eps_t=0.0)
    zw = O.window_depth(F, cl)
    vals = np.sort(zw[zw > 0])
    sd = np.full((96, 160), 1.0, np.float32)
    sd[:, 80:] = vals[len(vals) // 2]                      # a wall through the middle of the cloud on the right half
    sd[40:60, 20:60] = vals[len(vals) // 5]
    p = O.pipeline(F, cl, scene_depth=sd, depth_func=func)
    b = RF.Bound(cl, fr, cam, origin, 3)
    rgba, st, unsafe = b.draw(scene_depth=sd, depth_func=func, unsafe_tol=UNSAFE_TOL)
    _compare_frames(rgba, unsafe, p["rgba"], 2e-5, f"depth func {func}")
    free, _ = b.draw()[0:2]
    assert np.abs(free - rgba).max() > 0.05                # the wall really occludes


def test_fragment_shader_known_answers():
    """SRC.h:304-312 through the compiled text: falloff, clamp, 1/255 discard, premultiplied output."""
    out = RF.fragment([0.0, 0.0, 0, 1], [0.2, 0.4, 0.6], 0.5)
    assert np.allclose(out, [0.1, 0.2, 0.3, 0.5], atol=1e-7)
    out = RF.fragment([1.0, 1.0, 0, 1], [1, 1, 1], 1.0)
    assert np.allclose(out, [np.exp(-2.0)] * 4, rtol=1e-6)
    assert np.allclose(RF.fragment([0, 0, 0, 1], [1, 2, 3], 7.0), [1, 2, 3, 1])             # clamp(alpha, 0, 1)
    assert RF.fragment([0, 0, 0, 1], [1, 1, 1], 1.0 / 256.0) is None                          # alpha < 1/255 -> discard
    assert RF.fragment([0, 0, 0, 1], [1, 1, 1], 1.0 / 254.0) is not None
    r = np.sqrt(np.log(255.0 * 0.8))                                                          # discard ring |q|^2 = ln(255 alpha)
    assert RF.fragment([r * 0.999, 0, 0, 1], [1, 1, 1], 0.8) is not None
    assert RF.fragment([r * 1.001, 0, 0, 1], [1, 1, 1], 0.8) is None


def test_texture_dimension_function():
    """closestSqrtPowerOf2 (R.C:155-163; SURVEY A.7 ix)."""
    assert [RF.closest_sqrt_pow2(n) for n in (0, 1, 2, 4, 5, 16, 17, 1 << 20, (1 << 20) + 1)] == [2, 2, 2, 2, 4, 4, 8, 1024, 2048]
    assert RF.closest_sqrt_pow2(4 * 1_000_000) == 2048 and RF.closest_sqrt_pow2(8 * 1_000_000) == 4096


# ---------------------------------------------------------------------------------------------- GPU vs the reference's text
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.CASES))
def test_cuda_frame_matches_reference_glsl(oracle, scene, name):
    """The CUDA path against the compiled reference shaders directly (no oracle in between), reference semantics
    (eps_t = 0) and the production default (eps_t = 1e-5, bounded K1, auto chunks)."""
    from houdini_gsplat_renderer_b200 import renderer as R
    n, seed, sh, mult, w, h, theta, order = G.CASES[name]
    cl, fr, F, p, b = _bound(oracle, scene, n, w, h, theta, seed, order, sh=sh, mult=mult, eps=0.0)
    ref_rgba, st, unsafe = b.draw(unsafe_tol=UNSAFE_TOL)
    r = R.GSplatRenderer(0)
    rid = r.registerUpdate(0x5EF, (1, 0, 0, 0), 0, cl)
    r.setSphericalHarmonicsOrder(order)
    host = np.zeros((h, w, 4), np.float32)
    r.set_option(R.OPT_EPS_T, 0.0)
    r.draw([rid], fr, host_rgba=host)
    _compare_frames(ref_rgba, unsafe, host, 2e-5, name + " cuda eps 0")
    r.set_option(R.OPT_EPS_T, 1e-5)
    r.draw([rid], fr, host_rgba=host)
    rgbmax = max(1.0, float(max(p["recs"]["r"].max(), p["recs"]["g"].max(), p["recs"]["b"].max())))
    _compare_frames(ref_rgba, unsafe, host, 2e-5 + 1e-5 * rgbmax, name + " cuda eps 1e-5")
    r.close()


@pytest.mark.gpu
def test_cuda_frame_matches_reference_glsl_100k_640x360(oracle, scene):
    from houdini_gsplat_renderer_b200 import renderer as R
    cl, fr, F, p, b = _bound(oracle, scene, 100_000, 640, 360, 140.0, 77, 3, mult=1.0, eps=0.0)
    ref_rgba, st, unsafe = b.draw(unsafe_tol=UNSAFE_TOL)
    r = R.GSplatRenderer(0)
    rid = r.registerUpdate(0x5F0, (1, 0, 0, 0), 0, cl)
    r.setSphericalHarmonicsOrder(3)
    host = np.zeros((360, 640, 4), np.float32)
    r.draw([rid], fr, host_rgba=host)
    rgbmax = max(1.0, float(max(p["recs"]["r"].max(), p["recs"]["g"].max(), p["recs"]["b"].max())))
    _compare_frames(ref_rgba, unsafe, host, 2e-5 + 1e-5 * rgbmax, "100k cuda")
    r.close()
