"""-m gpu: the GSplatRenderer state machine behind the C ABI behaves like the reference's
(R.C:141-153 active-set check, 218-320 registry, 336-376 cap, 403-418 origin, 536-549 early returns,
660-678 postRender)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_multi_prim_merge_eviction_cap_and_flags(oracle, scene):
    from houdini_gsplat_renderer_b200 import renderer as R
    O, S = oracle, scene
    a = S.make_cloud(3000, 1, sh=True, scale_mult=2.0)
    b = S.make_cloud(2000, 2, sh=True, scale_mult=2.0)
    b.pos[:] += np.float32(0.25)
    fr = S.orbit_frame(256, 144, 20.0)
    r = R.GSplatRenderer(0)
    r.set_option(R.OPT_KEEP_INTERMEDIATES, 1)
    ida = r.registerUpdate(0x1000, (1, 0, 0, 0), 0, a)
    idb = r.registerUpdate(0x2000, (1, 0, 0, 0), 0, b)
    assert ida == "0x1000__0__1_0_0_0" and idb == "0x2000__0__1_0_0_0"     # id text of R.C:241-243
    assert r.registry_size() == 2
    r.setSphericalHarmonicsOrder(3)

    # nothing included -> render is a no-op (R.C:541-549)
    r.generateRenderGeometry(); r.render(fr); assert r.stats()["rendered"] == 0; r.postRender()

    # both prims: packed in ascending-id order, origin = mean of barycentres (fp32)
    host = np.zeros((144, 256, 4), np.float32)
    r.draw([ida, idb], fr, host_rgba=host)
    st = r.stats()
    assert st["rendered"] == 1 and st["n_submitted"] == 5000 and st["repacked"] == 1
    org = ((a.barycentre() + b.barycentre()) / np.float32(2)).astype(np.float32)
    assert np.array_equal(np.asarray(st["origin"], np.float32), org)
    merged = S.SplatCloud(*[np.concatenate([getattr(a, f), getattr(b, f)]) for f in
                            ("pos", "cd_h", "alpha", "scale_h", "orient_h", "shx_h", "shy_h", "shz_h")])
    F = O.make_frame(fr, O.camera_from_view(fr.view), org, 3)
    o = O.pipeline(F, merged)
    assert np.array_equal(r.fetch(R.DBG_ORDER).astype(np.int64), o["order"].astype(np.int64))
    assert np.abs(host - o["rgba"]).max() < 2e-5

    # same active set next frame -> no repack (R.C:324-327)
    r.draw([ida, idb], fr); assert r.stats()["repacked"] == 0
    # only one prim -> repack with its own origin
    r.draw([idb], fr, host_rgba=host); st = r.stats()
    assert st["repacked"] == 1 and st["n_submitted"] == 2000
    assert np.array_equal(np.asarray(st["origin"], np.float32), b.barycentre())
    # rendering disabled (GR_GSplat.C:472 for non-beauty modes) -> no draw
    r.setRenderingEnabled(False); r.draw([idb], fr); assert r.stats()["rendered"] == 0; r.setRenderingEnabled(True)

    # new version of gdp 0x1000 evicts the old entry (R.C:246-265)
    ida2 = r.registerUpdate(0x1000, (2, 0, 0, 0), 0, a)
    assert ida2 != ida and r.registry_size() == 2
    # flush by id removes every entry of that gdp (R.C:293-311)
    r.flushEntriesForMatchingDetail(ida2); assert r.registry_size() == 1
    r.includeInRenderPass(ida2)       # unknown id: silently ignored like the reference

    # cap: reference budget semantics with a tiny cap (R.C:336-376,437-445)
    r.set_option(R.OPT_SPLAT_CAP, 1500)
    r.draw([idb], fr); assert r.stats()["n_submitted"] == 1500
    r.close()


def test_sh_present_only_if_all_active_prims_have_sh(oracle, scene):
    from houdini_gsplat_renderer_b200 import renderer as R
    S = scene
    a = S.make_cloud(500, 1, sh=True); b = S.make_cloud(500, 2, sh=False)
    fr = S.orbit_frame(128, 72, 0.0)
    r = R.GSplatRenderer(0)
    ia = r.registerUpdate(1, (0, 0, 0, 0), 0, a); ib = r.registerUpdate(2, (0, 0, 0, 0), 0, b)
    r.setSphericalHarmonicsOrder(3)
    r.draw([ia], fr); assert r.stats()["sh_order_used"] == 3
    r.draw([ia, ib], fr); assert r.stats()["sh_order_used"] == 0      # SURVEY B3 decision
    r.setSphericalHarmonicsOrder(0)
    r.draw([ia], fr); assert r.stats()["sh_order_used"] == 0
    r.close()


def test_bad_arguments_return_errors_not_crashes(scene):
    from houdini_gsplat_renderer_b200 import renderer as R
    S = scene
    r = R.GSplatRenderer(0)
    cl = S.make_cloud(10, 1, sh=False)
    rid = r.registerUpdate(1, (0, 0, 0, 0), 0, cl)
    fr = S.orbit_frame(128, 72, 0.0)
    r.includeInRenderPass(rid); r.generateRenderGeometry()
    with pytest.raises(R.GsbError):
        r.render(fr, row_rank=3, row_world=2)
    bad = S.orbit_frame(128, 72, 0.0); bad.width = 70000
    with pytest.raises(R.GsbError):
        r.render(bad)
    with pytest.raises(R.GsbError):
        r.set_option(99, 1.0)
    # CUDA<->GL hand-back without a GL context (this box has no OpenGL): a reported error, not a crash; the context
    # stays usable afterwards
    import ctypes as C
    fc = R.frame_to_c(fr); t = R.TargetC(None, None, 1234, 0, None)
    r.includeInRenderPass(rid)
    rc = r._lib.gsb_render(r._h, C.byref(fc), C.byref(t))
    assert rc != 0 and b"cudaGraphicsGLRegisterImage" in r._lib.gsb_last_error()
    host = np.zeros((72, 128, 4), np.float32)
    r.draw([rid], fr, host_rgba=host)
    assert r.stats()["rendered"] == 1
    r.close()
