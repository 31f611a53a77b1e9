"""Scene-depth occlusion in the CPU oracle: analytic known-answer tests.

The reference draws its quads with the depth test on and depth writes off (src/GSplatRenderer.C:608-610) and gives
all vertices of a quad the centre's clip z and w (shaders/GSplatShaderSource.h:278-282): one window depth per splat,
tested per fragment against the scene depth already in the viewport."""
import numpy as np
import pytest

from test_oracle_kat import frame, one


def _scene(O, S, n=4000, w=160, h=90):
    cl = S.make_cloud(n, 4242, sh=True, scale_mult=3.0)
    fr = S.orbit_frame(w, h, 20.0)
    F = O.make_frame(fr, O.camera_from_view(fr.view), cl.barycentre(), 3, eps_t=1e-5)
    return cl, fr, F


def test_window_depth_formula_and_depth_range(oracle, scene):
    O, S = oracle, scene
    pos = np.array([[0, 0, 0], [0.3, -0.2, 0.5], [-0.7, 0.4, -0.9]], np.float32)
    cl = one(S, pos)
    fr, F = frame(O, S, cl)
    zw = O.window_depth(F, cl)
    ov = np.asarray(fr.obj_view, np.float64).reshape(4, 4).T          # column-major -> matrix
    pm = np.asarray(fr.proj, np.float64).reshape(4, 4).T
    for i, p in enumerate(pos.astype(np.float64)):
        vc = ov @ np.append(p, 1.0)
        clip = pm @ np.array([vc[0], -vc[1], vc[2], 1.0])
        assert abs(zw[i] - (clip[2] / clip[3] * 0.5 + 0.5)) < 2e-6
        assert 0.0 < zw[i] < 1.0
    # nearer splat -> smaller window depth (camera at radius 3 on +z looking at the origin)
    assert zw[1] < zw[0] < zw[2]
    zr = O.window_depth(F, cl, (0.25, 0.75))
    assert np.allclose(zr, zw * 0.5 + 0.25, atol=2e-7)


def test_far_wall_changes_nothing_near_wall_hides_everything(oracle, scene):
    O, S = oracle, scene
    cl, fr, F = _scene(O, S)
    plain = O.pipeline(F, cl)
    far = O.pipeline(F, cl, np.ones((fr.height, fr.width), np.float32), O.DEPTH_LEQUAL)
    assert np.array_equal(far["rgba"], plain["rgba"]) and far["n_consumed"] == plain["n_consumed"]
    near = O.pipeline(F, cl, np.zeros((fr.height, fr.width), np.float32), O.DEPTH_LESS)
    assert not near["rgba"].any()
    assert plain["rgba"][..., 3].max() > 0.9


def test_half_screen_wall(oracle, scene):
    """Scene depth 0 on the left half, 1 on the right: left half empty, right half untouched (tiles never interact)."""
    O, S = oracle, scene
    cl, fr, F = _scene(O, S)
    plain = O.pipeline(F, cl)["rgba"]
    sd = np.ones((fr.height, fr.width), np.float32); sd[:, : fr.width // 2] = 0.0
    got = O.pipeline(F, cl, sd, O.DEPTH_LEQUAL)["rgba"]
    assert not got[:, : fr.width // 2].any()
    assert np.array_equal(got[:, fr.width // 2:], plain[:, fr.width // 2:])


def test_less_versus_lequal_at_equality(oracle, scene):
    O, S = oracle, scene
    cl = one(S, [[0.07, 0.04, 0.0]], scale=(0.2, 0.15, 0.1))     # (an isotropic splat on the view axis has no eigenvector: culled)
    fr, F = frame(O, S, cl, w=128, h=72)
    zw = O.window_depth(F, cl)
    sd = np.full((72, 128), zw[0], np.float32)
    le = O.pipeline(F, cl, sd, O.DEPTH_LEQUAL)["rgba"]
    lt = O.pipeline(F, cl, sd, O.DEPTH_LESS)["rgba"]
    assert le[..., 3].max() > 0.5 and not lt.any()
    sd_eps = np.nextafter(sd, np.float32(2.0))
    assert np.array_equal(O.pipeline(F, cl, sd_eps, O.DEPTH_LESS)["rgba"], le)


def test_occluder_between_two_layers(oracle, scene):
    """Two opaque-ish splats on the view axis, a scene surface between them: only the near one is drawn, and the
    result equals rendering the near one alone ("under" blend of a hidden layer contributes nothing)."""
    O, S = oracle, scene
    both = one(S, [[0.03, 0.02, 0.5], [0.03, 0.02, -0.5]], scale=(0.15, 0.12, 0.1), alpha=0.6, cd=[(1, 0, 0), (0, 1, 0)])
    fr, F = frame(O, S, both, w=128, h=72)
    zw = O.window_depth(F, both)
    assert zw[0] < zw[1]
    sd = np.full((72, 128), 0.5 * (zw[0] + zw[1]), np.float32)
    got = O.pipeline(F, both, sd, O.DEPTH_LEQUAL)["rgba"]
    near_only = one(S, [[0.03, 0.02, 0.5]], scale=(0.15, 0.12, 0.1), alpha=0.6, cd=(1, 0, 0))
    _, Fn = frame(O, S, near_only, w=128, h=72)
    ref = O.pipeline(Fn, near_only)["rgba"]
    assert np.array_equal(got, ref)
    assert got[..., 1].max() == 0.0 and got[..., 0].max() > 0.3


@pytest.mark.parametrize("objmat", [False, True])
def test_window_depth_matches_the_literal_vertex_shader(oracle, scene, objmat):
    """The literal emulation of the reference's vertex shader (oracle/glsl_literal.py, SRC.h:190-288) gives all four
    corners of a quad the SAME gl_Position.z / gl_Position.w, and that ratio, mapped through glDepthRange, is the oracle's
    window depth: the "one depth per splat" the occlusion test relies on."""
    from oracle import glsl_literal as L
    O, S = oracle, scene
    cl = S.make_cloud(400, 23, sh=True, scale_mult=3.0)
    fr = S.orbit_frame(320, 180, 57.0)
    if objmat:
        a = 0.4; obj = np.eye(4)
        obj[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]) @ np.diag([1.2, 0.8, 1.1])
        obj[:3, 3] = [0.05, -0.02, 0.1]
        fr = S.Frame(320, 180, fr.view, fr.proj, S.colmajor(obj), S.colmajor(np.linalg.inv(obj)))
    cam = O.camera_from_view(fr.view); origin = cl.barycentre()
    F = O.make_frame(fr, cam, origin, 3)
    zw = O.window_depth(F, cl, (0.0, 1.0))
    zr = O.window_depth(F, cl, (0.1, 0.9))
    pr = O.project(F, cl)
    checked = 0
    for i in range(cl.n):
        vs = [L.vertex_shader(i, v, cl, fr, cam, origin, 3) for v in range(6)]
        if vs[0] is None or pr["vis"][i] == 0:
            continue
        ndc = [v["gl_Position"][2] / v["gl_Position"][3] for v in vs]
        assert max(ndc) - min(ndc) == 0.0                        # every vertex of the quad: the centre's z and w
        assert abs(zw[i] - (0.5 * ndc[0] + 0.5)) < 2e-6
        assert abs(zr[i] - (0.4 * ndc[0] + 0.5)) < 2e-6
        checked += 1
    assert checked > 100


@pytest.mark.parametrize("func", [1, 2])
def test_occluded_frame_matches_the_literal_render(oracle, scene, func):
    """Whole frames with a scene depth buffer: the oracle's depth-tested blend against the literal emulation of the
    reference's shaders + fixed-function depth test (R.C:608-621), no early-out, <= 1e-4 away from coverage edges."""
    from oracle import glsl_literal as L
    O, S = oracle, scene
    cl = S.make_cloud(400, 3, sh=True, scale_mult=4.0)
    fr = S.orbit_frame(96, 54, 20.0)
    cam = O.camera_from_view(fr.view); origin = cl.barycentre()
    F = O.make_frame(fr, cam, origin, 3, eps_t=0.0)
    zw = O.window_depth(F, cl)
    lo, hi = np.quantile(zw[zw > 0], [0.2, 0.8])
    x = np.linspace(0.0, 1.0, 96)[None, :]; y = np.linspace(0.0, 1.0, 54)[:, None]
    sd = (lo + (hi - lo) * (0.5 * x + 0.5 * y)).astype(np.float32)           # a tilted surface through the cloud
    p = O.pipeline(F, cl, sd, func)
    plain = O.pipeline(F, cl)
    assert np.abs(p["rgba"] - plain["rgba"]).max() > 0.05                    # the surface hides something
    lit, unsafe = L.render(cl, fr, cam, origin, 3, p["order"], scene_depth=sd.astype(np.float64), depth_func=func)
    diff = np.abs(lit - p["rgba"].astype(np.float64))
    diff[unsafe] = 0.0                                                        # (incl. pixels within 3e-7 of a depth tie)
    assert unsafe.mean() < 0.12
    assert diff.max() < 1e-4, diff.max()
