"""Helpers shared by the -m gpu tests: run a cloud through the C ABI and fetch every intermediate."""
import numpy as np

from houdini_gsplat_renderer_b200 import renderer as R


def gpu_pipeline(cloud, frame, sh_order, eps_t=1e-5, row_rank=0, row_world=1, cap=0, renderer=None,
                 explicit_cam=None, origin=None, depth_chunks=1, compact=0, row_group=1):
    r = renderer or R.GSplatRenderer(0)
    r.set_option(R.OPT_KEEP_INTERMEDIATES, 1)
    r.set_option(R.OPT_SPLAT_CAP, cap)
    r.set_option(R.OPT_EPS_T, eps_t)
    r.set_option(R.OPT_COMPACT, compact)
    r.set_option(R.OPT_DEPTH_CHUNKS, depth_chunks)     # 1 = single pass: the fetched tile lists are the full lists
    rid = r.registerUpdate(0x7f00dead0000, (1, 2, 3, 4), 0, cloud, origin=origin)
    r.includeInRenderPass(rid)
    r.setSphericalHarmonicsOrder(sh_order)
    if explicit_cam is not None:
        r.setExplicitCameraPos(explicit_cam)
    r.generateRenderGeometry()
    host = np.zeros((frame.height, frame.width, 4), np.float32)
    r.render(frame, host_rgba=host, row_rank=row_rank, row_world=row_world, row_group=row_group)
    st = r.stats()
    r.postRender()
    out = dict(stats=st, rgba=host,
               keys=r.fetch(R.DBG_KEYS_UNSORTED), order=r.fetch(R.DBG_ORDER), keys_sorted=r.fetch(R.DBG_KEYS_SORTED),
               recs=r.fetch(R.DBG_RECORDS), rects=r.fetch(R.DBG_RECTS), ranges=r.fetch(R.DBG_TILE_RANGES),
               inst=r.fetch(R.DBG_INSTANCES), consumed=r.fetch(R.DBG_TILE_CONSUMED),
               fb=r.fetch(R.DBG_FRAMEBUFFER).reshape(frame.height, frame.width, 4))
    # the production path (GSB_OPT_LAZY_PROJECT: bounded K1, exact projection only for selected splats) must give the
    # same frame, the same instances and the same consumed counts as the exact-K1 path whose intermediates were fetched
    r.set_option(R.OPT_KEEP_INTERMEDIATES, 0)
    r.includeInRenderPass(rid)
    if explicit_cam is not None:
        r.setExplicitCameraPos(explicit_cam)
    r.generateRenderGeometry()
    host2 = np.full_like(host, -1.0)
    r.render(frame, host_rgba=host2, row_rank=row_rank, row_world=row_world, row_group=row_group)
    st2 = r.stats()
    r.postRender()
    assert np.array_equal(host2, host), "bounded-K1 frame differs from the exact-K1 frame"
    assert st2["n_consumed"] == st["n_consumed"], (st2, st)
    if depth_chunks == 1:                                # (with several chunks the plan, and with it D, follows the histogram)
        assert st2["n_instances"] == st["n_instances"], (st2, st)
    assert st2["n_visible"] >= st["n_visible"] and st2["n_live"] >= st["n_live"]
    out["stats_lazy"] = st2
    r.set_option(R.OPT_KEEP_INTERMEDIATES, 1)
    if renderer is None:
        r.close()
    return out


def assert_stage_parity(O, g, o, cloud_n, rgba_tol=2e-5):
    """g: gpu_pipeline output, o: oracle.pipeline output.  Bit-exact for integer/index work and for the
    fp32 records (same operation order); RGBA within rgba_tol (exp differs: MUFU.EX2 vs libm)."""
    vis = o["vis"] > 0
    assert g["stats"]["n_visible"] == int(vis.sum())
    assert np.array_equal(g["keys"], o["keys"]), "depth keys differ"
    ns = g["order"].shape[0]                       # V: only splats that reach a live tile are depth-sorted
    assert ns == int(vis.sum()) == g["stats"]["n_live"]
    assert np.array_equal(g["order"].astype(np.int64), o["order"][:ns].astype(np.int64)), "depth order differs"
    assert np.array_equal(g["keys_sorted"], o["keys"][o["order"]][:ns])
    gr, orc = g["rects"], o["rects"]
    assert np.array_equal(gr["x0"] > gr["x1"], ~vis), "cull decisions differ"
    for f in ("x0", "x1", "y0", "y1"):
        assert np.array_equal(gr[f][vis], orc[f][vis]), f"rect {f} differs"
    gb = g["recs"][vis].view(np.uint32).reshape(-1, 12)
    ob = o["recs"][vis].view(np.uint32).reshape(-1, 12)
    bad = np.nonzero((gb != ob).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} records differ, first {g['recs'][vis][bad[:1]]} vs {o['recs'][vis][bad[:1]]}"
    ts = o["tile_start"]
    nonempty = ts[1:] > ts[:-1]
    assert g["stats"]["n_instances"] == int(ts[-1]) == g["inst"].shape[0]
    assert np.array_equal(g["ranges"][nonempty, 0], ts[:-1][nonempty])
    assert np.array_equal(g["ranges"][nonempty, 1], ts[1:][nonempty])
    assert np.all(g["ranges"][~nonempty, 0] == g["ranges"][~nonempty, 1])
    assert np.array_equal(g["inst"].astype(np.int64), o["inst"].astype(np.int64)), "tile lists differ"
    assert np.array_equal(g["rgba"], g["fb"]), "host copy != device framebuffer"
    d = np.abs(g["rgba"].astype(np.float64) - o["rgba"].astype(np.float64))
    assert d.max() <= rgba_tol, f"RGBA max abs diff {d.max()}"
    # D_c: identical unless a pixel sits within exp-rounding of the T<eps threshold
    dc = np.abs(g["consumed"].astype(np.int64) - o["consumed"])
    assert (dc > 0).mean() < 0.02 and abs(g["stats"]["n_consumed"] - o["n_consumed"]) <= 0.01 * max(1, o["n_consumed"])
    return float(d.max())
