"""-m gpu: scene-depth occlusion through the C ABI vs the CPU oracle.
gsb_frame.scene_depth + depth_func reproduce the reference's "depth test on, depth writes off" draw state
(src/GSplatRenderer.C:608-610) with one window depth per quad (shaders/GSplatShaderSource.h:278-282)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(O, S, n=60_000, w=400, h=226, theta=15.0, mult=1.5, seed=909):
    cl = S.make_cloud(n, seed, sh=True, scale_mult=mult)
    fr = S.orbit_frame(w, h, theta)
    F = O.make_frame(fr, O.camera_from_view(fr.view), cl.barycentre(), 3, eps_t=1e-5)
    return cl, fr, F


def _scene_depth(O, F, cl, w, h):
    """A slanted plane through the cloud plus a nearer disc: window depths taken from the splats' own range."""
    zw = O.window_depth(F, cl)
    vis = zw > 0
    lo, hi = float(np.quantile(zw[vis], 0.1)), float(np.quantile(zw[vis], 0.9))
    x = np.linspace(0.0, 1.0, w, dtype=np.float32)[None, :]
    y = np.linspace(0.0, 1.0, h, dtype=np.float32)[:, None]
    sd = (lo + (hi - lo) * (0.25 + 0.5 * x + 0.0 * y)).astype(np.float32)
    disc = (x - 0.5) ** 2 + ((y - 0.5) * h / w) ** 2 < 0.02
    sd = np.where(disc, np.float32(lo * 0.999), sd).astype(np.float32)
    return np.ascontiguousarray(sd)


def _gpu(R, cl, fr, sd, func, chunks=1, eps_t=1e-5, depth_range=(0.0, 1.0)):
    import torch
    r = R.GSplatRenderer(0)
    r.set_option(R.OPT_DEPTH_CHUNKS, chunks); r.set_option(R.OPT_EPS_T, eps_t)
    rid = r.registerUpdate(0xD0, (1, 0, 0, 0), 0, cl); r.setSphericalHarmonicsOrder(3)
    d = torch.from_numpy(sd).cuda()
    host = np.zeros((fr.height, fr.width, 4), np.float32)
    r.draw([rid], fr, host_rgba=host, scene_depth=d.data_ptr(), depth_func=func, depth_range=depth_range)
    st = r.stats()
    r.close()
    return host, st


@pytest.mark.parametrize("func", [1, 2])
@pytest.mark.parametrize("chunks", [1, 3])
def test_depth_occlusion_matches_oracle(oracle, scene, func, chunks):
    from houdini_gsplat_renderer_b200 import renderer as R
    O, S = oracle, scene
    cl, fr, F = _setup(O, S)
    sd = _scene_depth(O, F, cl, fr.width, fr.height)
    o = O.pipeline(F, cl, sd, func)
    plain = O.pipeline(F, cl)
    assert np.abs(o["rgba"] - plain["rgba"]).max() > 0.05          # the occluder really hides something
    got, st = _gpu(R, cl, fr, sd, func, chunks)
    assert np.abs(got.astype(np.float64) - o["rgba"]).max() <= 2e-5
    assert st["n_visible"] >= o["n_visible"]                       # production path: the bounded K1 reports an upper bound
    if chunks == 1:
        assert st["n_instances"] == int(o["tile_start"][-1])
        assert abs(st["n_consumed"] - o["n_consumed"]) <= 0.01 * o["n_consumed"]


def test_depth_chunks_and_no_early_out_give_the_same_occluded_frame(oracle, scene):
    from houdini_gsplat_renderer_b200 import renderer as R
    O, S = oracle, scene
    cl, fr, F = _setup(O, S, n=30_000, w=320, h=180, mult=2.0)
    sd = _scene_depth(O, F, cl, fr.width, fr.height)
    a, _ = _gpu(R, cl, fr, sd, 2, chunks=1)
    for c in (2, 5):
        b, _ = _gpu(R, cl, fr, sd, 2, chunks=c)
        assert np.array_equal(a, b)
    F0 = O.make_frame(fr, O.camera_from_view(fr.view), cl.barycentre(), 3, eps_t=0.0)
    o0 = O.pipeline(F0, cl, sd, 2)
    g0, _ = _gpu(R, cl, fr, sd, 2, chunks=3, eps_t=0.0)
    assert np.abs(g0.astype(np.float64) - o0["rgba"]).max() <= 2e-5


def test_window_depth_bits_decide_equality_like_the_oracle(oracle, scene):
    """Scene depth set to EXACTLY one splat's window depth: LEQUAL keeps it, LESS drops it — only if the GPU's window
    depth has the oracle's bits.  Also a remapped glDepthRange."""
    from houdini_gsplat_renderer_b200 import renderer as R
    O, S = oracle, scene
    cl, fr, F = _setup(O, S, n=3000, w=256, h=144, mult=3.0, seed=31)
    for dr in ((0.0, 1.0), (0.25, 0.75)):
        zw = O.window_depth(F, cl, dr)
        order = O.sort(O.keys(cl.pos, np.asarray(F.cam[:], np.float32)))
        for pick in (int(order[k]) for k in range(len(order) // 8, len(order), 37)):      # a splat that is actually seen
            sd = np.full((fr.height, fr.width), zw[pick], np.float32)
            if np.abs(O.pipeline(F, cl, sd, 2, dr)["rgba"] - O.pipeline(F, cl, sd, 1, dr)["rgba"]).max() > 1e-3:
                break
        for func in (1, 2):
            o = O.pipeline(F, cl, sd, func, dr)
            g, _ = _gpu(R, cl, fr, sd, func, chunks=1, depth_range=dr)
            assert np.abs(g.astype(np.float64) - o["rgba"]).max() <= 2e-5, (dr, func)
        le = O.pipeline(F, cl, sd, 2, dr)["rgba"]; lt = O.pipeline(F, cl, sd, 1, dr)["rgba"]
        assert np.abs(le - lt).max() > 1e-3            # the picked splat is visible under LEQUAL only


def test_depth_func_without_a_depth_buffer_is_an_error(scene):
    from houdini_gsplat_renderer_b200 import renderer as R
    S = scene
    cl = S.make_cloud(100, 1, sh=False)
    fr = S.orbit_frame(64, 36, 0.0)
    r = R.GSplatRenderer(0)
    rid = r.registerUpdate(1, (1, 0, 0, 0), 0, cl)
    r.includeInRenderPass(rid); r.generateRenderGeometry()
    with pytest.raises(R.GsbError):
        r.render(fr, depth_func=2)
    with pytest.raises(R.GsbError):
        r.render(fr, depth_func=7, scene_depth=1)
    host = np.zeros((36, 64, 4), np.float32)
    r.render(fr, host_rgba=host)                       # still usable after the errors
    assert r.stats()["rendered"] == 1
    r.close()
