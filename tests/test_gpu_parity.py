"""-m gpu: the CUDA hot path vs the CPU oracle, stage by stage, through the C ABI (include/gsplat_b200.h).
Bit-exact: depth keys, depth order, cull decisions, pixel rectangles, 2-D records, tile ranges, tile lists.
Float tolerance: RGBA <= 2e-5 abs vs the oracle (north-star bound is 1e-3); exp() is the only difference."""
import numpy as np
import pytest

from gpu_util import assert_stage_parity, gpu_pipeline

pytestmark = pytest.mark.gpu


def _frame(O, S, cl, w, h, theta, sh_order, eps=1e-5, **kw):
    fr = S.orbit_frame(w, h, theta)
    cam = O.camera_from_view(fr.view)
    return fr, O.make_frame(fr, cam, cl.barycentre(), sh_order, eps_t=eps, **kw)


@pytest.mark.parametrize("n,w,h,theta,sh,order,mult", [
    (1, 64, 36, 0.0, True, 3, 30.0),
    (257, 200, 120, 15.0, False, 0, 3.0),          # ragged tiles (200x120 not multiples of 16)
    (5000, 320, 180, 75.0, True, 1, 2.0),
    (5000, 320, 180, 200.0, True, 2, 2.0),
    (40_000, 640, 360, 33.0, True, 3, 1.0),
    (200_000, 1920, 1080, 0.0, True, 3, 1.0),
    (100_000, 480, 270, 311.0, False, 3, 1.5),     # order 3 requested but no SH data -> order 0 (R.C:623)
])
def test_stage_parity(oracle, scene, n, w, h, theta, sh, order, mult):
    O, S = oracle, scene
    cl = S.make_cloud(n, 1000 + n, sh=sh, scale_mult=mult)
    fr, F = _frame(O, S, cl, w, h, theta, order if sh else 0)
    o = O.pipeline(F, cl)
    g = gpu_pipeline(cl, fr, order)
    assert g["stats"]["sh_order_used"] == (order if sh else 0)
    assert np.array_equal(np.asarray(g["stats"]["camera"], np.float32), np.asarray(F.cam[:], np.float32))
    assert np.array_equal(np.asarray(g["stats"]["origin"], np.float32), cl.barycentre())
    assert_stage_parity(O, g, o, n)


def test_culled_splats_never_reach_the_depth_sort(oracle, scene):
    """Only splats that reach a live tile are compacted and depth-sorted: with a third of the cloud behind the camera
    the sorted set is exactly the visible set (GSB_OPT_COMPACT is accepted and has no effect)."""
    O, S = oracle, scene
    cl = S.make_cloud(60_000, 808, sh=True, scale_mult=1.5)
    cl.pos[::3, 2] += np.float32(4.5)              # a third of the cloud behind the camera -> culled
    fr, F = _frame(O, S, cl, 400, 226, 5.0, 3)
    o = O.pipeline(F, cl)
    g = gpu_pipeline(cl, fr, 3, compact=1)
    assert g["order"].shape[0] == o["n_visible"] < cl.n
    assert_stage_parity(O, g, o, cl.n)
    g2 = gpu_pipeline(cl, fr, 3, compact=2, depth_chunks=4)
    assert np.array_equal(g["rgba"], g2["rgba"])


def test_no_early_out_matches_reference_semantics(oracle, scene):
    """eps_t = 0: full traversal like the reference's ROP blend (no termination)."""
    O, S = oracle, scene
    cl = S.make_cloud(30_000, 77, sh=True, scale_mult=2.0)
    fr, F = _frame(O, S, cl, 320, 180, 10.0, 3, eps=0.0)
    o = O.pipeline(F, cl)
    g = gpu_pipeline(cl, fr, 3, eps_t=0.0)
    assert_stage_parity(O, g, o, cl.n)
    assert g["stats"]["n_consumed"] == g["stats"]["n_instances"]


def test_edge_cases_big_splats_behind_camera_and_offscreen(oracle, scene):
    """Screen-filling splats (axis cap 4096 px), splats behind the camera, alpha below 1/255, NaNs."""
    O, S = oracle, scene
    cl = S.make_cloud(3000, 5, sh=True, scale_mult=1.0)
    cl.scale_h[:20] = np.float16(40.0)                     # capped at 4096 px
    cl.pos[20:40, 2] = 5.0                                 # behind the camera at z=+3
    cl.alpha[40:60] = np.float32(1.0 / 300.0)
    cl.alpha[60:70] = np.float32(np.nan)
    cl.pos[70:80, 0] = 50.0                                # far off screen
    cl.orient_h[80:90] = np.float16(0.0)                   # zero quaternion: Sigma = diag(scale^2)
    cl.alpha[90:100] = np.float32(1.7)                     # alpha > 1 exercises the clamp
    fr, F = _frame(O, S, cl, 256, 144, 0.0, 3)
    o = O.pipeline(F, cl)
    g = gpu_pipeline(cl, fr, 3)
    assert_stage_parity(O, g, o, cl.n)
    # non-finite positions: the depth-key range bound is disabled (full 32-bit sort), NaN splats are culled
    cl.pos[100:105] = np.float32(np.nan); cl.pos[105:108, 1] = np.float32(np.inf)
    fr, F = _frame(O, S, cl, 256, 144, 0.0, 3)
    F.origin[:] = [0.0, 0.0, 0.0]
    o = O.pipeline(F, cl)
    g = gpu_pipeline(cl, fr, 3, origin=(0.0, 0.0, 0.0))
    assert_stage_parity(O, g, o, cl.n)


def test_explicit_camera_and_object_matrix(oracle, scene):
    """gsplat__explicit_camera_pos override (R.C:552-555) and a non-identity object matrix."""
    O, S = oracle, scene
    cl = S.make_cloud(8000, 12, sh=True, scale_mult=2.0)
    obj = np.eye(4); obj[:3, :3] *= 1.1; obj[:3, 3] = [0.05, 0.02, -0.03]
    base = S.orbit_frame(400, 224, 40.0)
    fr = S.Frame(400, 224, base.view, base.proj, S.colmajor(obj), S.colmajor(np.linalg.inv(obj)))
    cam = np.array([0.5, 1.0, 2.0], np.float32)
    F = O.make_frame(fr, cam, cl.barycentre(), 3)
    o = O.pipeline(F, cl)
    g = gpu_pipeline(cl, fr, 3, explicit_cam=cam)
    assert_stage_parity(O, g, o, cl.n)


def test_tile_row_partition_reassembles_bit_exactly(oracle, scene):
    """Multi-GPU sharding rule exercised on one device: ranks render disjoint tile rows,
    their sum equals the single-context frame bit for bit, and each matches the oracle's shard."""
    O, S = oracle, scene
    cl = S.make_cloud(20_000, 31, sh=False, scale_mult=2.0)
    fr, F = _frame(O, S, cl, 328, 200, 25.0, 0)
    full = gpu_pipeline(cl, fr, 0)
    acc = np.zeros_like(full["rgba"])
    world = 4
    for rk in range(world):
        Fr = O.make_frame(fr, F.cam[:], cl.barycentre(), 0, row_rank=rk, row_world=world)
        o = O.pipeline(Fr, cl)
        g = gpu_pipeline(cl, fr, 0, row_rank=rk, row_world=world)
        assert_stage_parity(O, g, o, cl.n)
        assert g["stats"]["n_visible"] < full["stats"]["n_visible"]
        acc += g["rgba"]
    assert np.array_equal(acc, full["rgba"])


@pytest.mark.parametrize("chunks", [2, 5, 16])
def test_depth_chunks_leave_the_frame_bit_identical(oracle, scene, chunks):
    """Depth-chunked binning drops instances behind saturated tiles; the per-pixel sequence of blended instances is
    unchanged, so frame, per-tile consumed counts and D_c equal the single-pass result bit for bit (and the oracle)."""
    O, S = oracle, scene
    cl = S.make_cloud(150_000, 404, sh=True, scale_mult=2.5)
    fr, F = _frame(O, S, cl, 488, 270, 52.0, 3)
    one = gpu_pipeline(cl, fr, 3, depth_chunks=1)
    many = gpu_pipeline(cl, fr, 3, depth_chunks=chunks)
    assert many["stats"]["depth_chunks"] == chunks
    assert np.array_equal(one["rgba"], many["rgba"])
    assert np.array_equal(one["consumed"], many["consumed"])
    assert one["stats"]["n_consumed"] == many["stats"]["n_consumed"]
    assert many["stats"]["n_consumed"] <= many["stats"]["n_instances"] < one["stats"]["n_instances"]
    o = O.pipeline(F, cl)
    assert np.abs(many["rgba"] - o["rgba"]).max() <= 2e-5
    # also with tile-row sharding and without early-out
    a = gpu_pipeline(cl, fr, 3, depth_chunks=chunks, row_rank=1, row_world=3)
    b = gpu_pipeline(cl, fr, 3, depth_chunks=1, row_rank=1, row_world=3)
    assert np.array_equal(a["rgba"], b["rgba"])
    c = gpu_pipeline(cl, fr, 3, depth_chunks=chunks, eps_t=0.0)
    d = gpu_pipeline(cl, fr, 3, depth_chunks=1, eps_t=0.0)
    assert np.array_equal(c["rgba"], d["rgba"]) and c["stats"]["n_instances"] == d["stats"]["n_instances"]


def test_final_frame_target_receives_finished_tiles(scene):
    """gsb_target.final_rgba (the multi-GPU hand-off: peer memory in production, a second local buffer here):
    every finished tile lands there, bit-identical to the ordinary frame, for 1 and 4 depth chunks and for shards."""
    from houdini_gsplat_renderer_b200 import renderer as R
    S = scene
    cl = S.make_cloud(120_000, 515, sh=True, scale_mult=2.0)
    fr = S.orbit_frame(360, 200, 75.0)
    r = R.GSplatRenderer(0)
    rid = r.registerUpdate(7, (1, 0, 0, 0), 0, cl); r.setSphericalHarmonicsOrder(3)
    ref = np.zeros((200, 360, 4), np.float32)
    r.draw([rid], fr, host_rgba=ref)
    handle, shared = r.ipc_export_frame(360, 200)
    assert len(handle) == 64 and shared != 0
    for chunks in (1, 4):
        r.set_option(R.OPT_DEPTH_CHUNKS, chunks)
        got = np.full_like(ref, -1.0)
        r.draw([rid], fr, final_rgba=shared)
        r.copy_to_host(shared, got)
        assert np.array_equal(got, ref)
    # two shards writing into the same final frame reassemble it without any combine pass
    r.copy_to_host(shared, got)  # (dirty from the previous frame on purpose: every pixel is overwritten)
    for rk in range(2):
        r.draw([rid], fr, final_rgba=shared, row_rank=rk, row_world=2, row_group=2)
    r.copy_to_host(shared, got)
    assert np.array_equal(got, ref)
    r.close()


def test_pinned_host_target_is_written_by_the_blend_kernel(scene):
    """GSB_OPT_HOST_DIRECT: a pinned (device-addressable) host frame receives the finished tiles straight from the blend
    kernel; the bytes equal the staged-copy path (pageable target / option off) for 1 and 3 depth chunks."""
    import torch
    from houdini_gsplat_renderer_b200 import renderer as R
    S = scene
    cl = S.make_cloud(150_000, 616, sh=True, scale_mult=2.0)
    fr = S.orbit_frame(500, 281, 40.0)               # ragged tiles on both axes
    r = R.GSplatRenderer(0)
    rid = r.registerUpdate(9, (1, 0, 0, 0), 0, cl); r.setSphericalHarmonicsOrder(3)
    ref = np.zeros((281, 500, 4), np.float32)        # pageable: staged cudaMemcpyAsync
    r.draw([rid], fr, host_rgba=ref)
    assert ref[..., 3].max() > 0.5
    pinned = torch.full((281, 500, 4), -1.0, dtype=torch.float32).pin_memory()
    for chunks in (1, 3):
        for direct in (1, 0):
            r.set_option(R.OPT_DEPTH_CHUNKS, chunks); r.set_option(R.OPT_HOST_DIRECT, direct)
            pinned.fill_(-1.0)
            r.draw([rid], fr, host_rgba=pinned.numpy())
            assert np.array_equal(pinned.numpy(), ref), (chunks, direct)
            assert np.array_equal(r.fetch(R.DBG_FRAMEBUFFER).reshape(281, 500, 4), ref)
    r.close()


@pytest.mark.parametrize("case", ["aniso", "objmat", "bigsplats", "closeup"])
def test_shards_match_oracle_on_nasty_geometry(oracle, scene, case):
    """Row-partitioned shards (cull by owned rows, survivor compaction) on needles with |q| != 1, a non-rigid object
    matrix, screen-filling splats and a close-up: every stage equals the oracle's shard bit for bit."""
    from houdini_gsplat_renderer_b200 import renderer as R
    O, S = oracle, scene
    rng = np.random.default_rng(5)
    cl = S.make_cloud(30_000, 909, sh=False, scale_mult=2.0)
    obj = np.eye(4); theta = 25.0; w, h, world, group = 300, 420, 4, 1
    if case == "aniso":
        cl.scale_h[:, 1] = (cl.scale_h[:, 1].astype(np.float32) * 12).astype(np.float16)       # needles
        cl.orient_h[:] = (cl.orient_h.astype(np.float32) * rng.uniform(0.8, 1.25, (cl.n, 1))).astype(np.float16)  # |q| != 1
    elif case == "objmat":
        a = 0.7
        obj[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]) @ np.diag([1.3, 0.7, 1.1])
        obj[:3, 3] = [0.1, 0.05, -0.1]
    elif case == "bigsplats":
        cl.scale_h[::7] = np.float16(0.6); world, group = 3, 2
    elif case == "closeup":
        theta = 0.0; cl.pos[:, 2] = cl.pos[:, 2] * np.float32(0.2) + np.float32(2.6); world = 5  # just in front of the camera
    base = S.orbit_frame(w, h, theta)
    fr = S.Frame(w, h, base.view, base.proj, S.colmajor(obj), S.colmajor(np.linalg.inv(obj)))
    cam = O.camera_from_view(fr.view)
    nvis_pre = nvis_nopre = 0
    for rk in range(world):
        F = O.make_frame(fr, cam, cl.barycentre(), 0, row_rank=rk, row_world=world, row_group=group)
        o = O.pipeline(F, cl)
        g = gpu_pipeline(cl, fr, 0, row_rank=rk, row_world=world, row_group=group)
        assert_stage_parity(O, g, o, cl.n)
        nvis_pre += g["stats"]["n_visible"]
    assert nvis_pre > 0


def test_full_size_properties_1M_1080p(scene):
    """BASELINE config 2 size (1M splats, SH 0, 1080p) through size-independent properties: sorted keys,
    order is a permutation, tile lists are depth ordered, ranges partition D, alpha in [0,1], idempotence."""
    from houdini_gsplat_renderer_b200 import renderer as R
    S = scene
    cl = S.make_cloud(1_000_000, S.SEEDS["1M"], sh=False)
    fr = S.orbit_frame(1920, 1080, 0.0)
    r = R.GSplatRenderer(0)
    g = gpu_pipeline(cl, fr, 0, renderer=r)
    n = cl.n
    assert np.all(np.diff(g["keys_sorted"].astype(np.int64)) >= 0)
    V, D = g["stats"]["n_visible"], g["stats"]["n_instances"]
    assert 0.9 * n < V <= n and D > V and g["stats"]["n_live"] == V
    vis_idx = np.nonzero(g["rects"]["x0"] <= g["rects"]["x1"])[0].astype(np.uint32)
    assert np.array_equal(np.sort(g["order"]), vis_idx)          # the sorted set is exactly the visible set
    assert np.array_equal(g["keys_sorted"], g["keys"][g["order"]])
    rank = np.full(n, -1, np.int64); rank[g["order"]] = np.arange(V)
    rg = g["ranges"].astype(np.int64)
    ne = rg[:, 1] > rg[:, 0]
    assert (rg[ne, 1] - rg[ne, 0]).sum() == D
    starts = np.sort(rg[ne, 0]); ends = np.sort(rg[ne, 1])
    assert starts[0] == 0 and ends[-1] == D and np.array_equal(starts[1:], ends[:-1])
    inst_rank = rank[g["inst"]]
    seg_break = np.zeros(D, bool); seg_break[rg[ne, 0]] = True
    assert np.all((np.diff(inst_rank) > 0) | seg_break[1:])
    a = g["rgba"][..., 3]
    assert a.min() >= 0 and a.max() <= 1.0 + 1e-6 and a.mean() > 0.2
    assert 0 < g["stats"]["n_consumed"] <= D
    # idempotence: the same frame again gives the same bits
    host2 = np.zeros_like(g["rgba"])
    r.draw([r.registerUpdate(0x7f00dead0000, (1, 2, 3, 4), 0, cl)], fr, host_rgba=host2)
    assert np.array_equal(host2, g["rgba"])
    r.close()


def test_only_live_splats_get_sorted_and_shaded(oracle, scene):
    """Depth chunks + saturation feedback: deeper chunks sort / shade / bin only the splats that still touch an
    un-saturated tile.  L (n_live) < V, yet the frame equals the single-chunk frame bit for bit and the oracle."""
    O, S = oracle, scene
    cl = S.make_cloud(300_000, 2024, sh=True, scale_mult=2.5)
    fr, F = _frame(O, S, cl, 480, 270, 140.0, 3)
    one = gpu_pipeline(cl, fr, 3, depth_chunks=1)
    assert one["stats"]["n_live"] == one["stats"]["n_visible"]
    for chunks in (3, 6):
        many = gpu_pipeline(cl, fr, 3, depth_chunks=chunks)
        assert np.array_equal(one["rgba"], many["rgba"])
        assert np.array_equal(one["consumed"], many["consumed"])
        assert many["stats"]["n_live"] < many["stats"]["n_visible"]
        assert many["stats"]["n_consumed"] == one["stats"]["n_consumed"]
        # the live list of the last chunk is depth ordered (keys ascending, ties by ascending index)
        k, o_ = many["keys_sorted"].astype(np.int64), many["order"].astype(np.int64)
        assert np.all((np.diff(k) > 0) | ((np.diff(k) == 0) & (np.diff(o_) > 0)))
        assert np.array_equal(many["keys"][many["order"]], many["keys_sorted"])
    o = O.pipeline(F, cl)
    assert np.abs(one["rgba"] - o["rgba"]).max() <= 2e-5


def test_wide_splats_use_the_exact_rectangle(oracle, scene):
    """Splats spanning >= 127 tiles do not fit the packed tile rectangle and fall back to the exact pixel rectangle by
    splat index; that table is written sparsely unless intermediates are kept, so compare both modes and the oracle."""
    from houdini_gsplat_renderer_b200 import renderer as R
    O, S = oracle, scene
    cl = S.make_cloud(400, 9, sh=False, scale_mult=1.0)
    cl.scale_h[:6] = np.float16(30.0)                      # axis cap 4096 px: 256+ tiles wide at this resolution
    cl.alpha[:6] = np.float32(0.05)
    w, h = 4400, 2300
    fr, F = _frame(O, S, cl, w, h, 0.0, 0)
    o = O.pipeline(F, cl)
    g = gpu_pipeline(cl, fr, 0)
    assert_stage_parity(O, g, o, cl.n)
    tx = (g["rects"]["x1"].astype(int) // 16) - (g["rects"]["x0"].astype(int) // 16)
    assert (tx[g["rects"]["x0"] <= g["rects"]["x1"]] >= 127).sum() >= 3
    r = R.GSplatRenderer(0)                                # production mode: no intermediates, 2 depth chunks
    r.set_option(R.OPT_DEPTH_CHUNKS, 2)
    rid = r.registerUpdate(3, (1, 0, 0, 0), 0, cl)
    host = np.zeros((h, w, 4), np.float32)
    r.draw([rid], fr, host_rgba=host)
    assert np.array_equal(host, g["rgba"])
    r.close()


def _unpack_trects(t):
    culled = t == 0xFFFFFFFF
    tx0 = (t & 511).astype(np.int64); ty0 = ((t >> 9) & 511).astype(np.int64)
    w = ((t >> 18) & 127).astype(np.int64); h = ((t >> 25) & 127).astype(np.int64)
    return culled, tx0, ty0, w, h


@pytest.mark.parametrize("case", ["cloud", "aniso", "bigsplats", "closeup", "objmat", "8k"])
def test_bounded_k1_rectangles_contain_the_exact_ones(oracle, scene, case):
    """GSB_OPT_LAZY_PROJECT: for every splat the exact projection keeps, the bounded K1 keeps it too and its tile
    rectangle contains the exact one (otherwise a chunk could fail to select a splat that changes a pixel).  Checked
    on a million splats per case, including strongly anisotropic, huge, close-up and object-transformed ones, and
    the depth keys of the kept splats are the exact keys."""
    from houdini_gsplat_renderer_b200 import renderer as R
    O, S = oracle, scene
    rng = np.random.default_rng(77)
    n, w, h, theta, mult = 1_000_000, 1920, 1080, 40.0, 1.0
    if case == "8k":
        w, h = 7680, 4320
    if case == "bigsplats":
        n, mult = 20_000, 40.0                                               # screen-filling splats (D stays bounded)
    if case == "closeup":
        n, mult, theta = 50_000, 6.0, 0.0
    if case == "aniso":
        n = 100_000
    cl = S.make_cloud(n, 4321, sh=False, scale_mult=mult)
    if case == "aniso":                                                      # needles and flakes with |q| != 1
        sc = cl.scale_h.astype(np.float32); sc[:, 0] *= 30.0; sc[:, 2] *= 0.05; cl.scale_h = sc.astype(np.float16)
        cl.orient_h[:] = (cl.orient_h.astype(np.float32) * rng.uniform(0.8, 1.25, (cl.n, 1))).astype(np.float16)
    if case == "closeup":
        cl.pos[:, 2] = cl.pos[:, 2] * np.float32(0.2) + np.float32(2.6)      # a slab right in front of the camera
    fr = S.orbit_frame(w, h, theta)
    if case == "objmat":                                                     # non-rigid object matrix
        a = 0.7; obj = np.eye(4)
        obj[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]) @ np.diag([1.7, 0.6, 1.1])
        obj[:3, 3] = [0.1, 0.05, -0.1]
        fr = S.Frame(w, h, fr.view, fr.proj, S.colmajor(obj), S.colmajor(np.linalg.inv(obj)))
    r = R.GSplatRenderer(0)
    r.set_option(R.OPT_SPLAT_CAP, 0); r.set_option(R.OPT_DEPTH_CHUNKS, 1)
    rid = r.registerUpdate(0x51, (1, 0, 0, 0), 0, cl)
    r.set_option(R.OPT_KEEP_INTERMEDIATES, 1)                                # exact K1
    r.draw([rid], fr)
    exact = r.fetch(R.DBG_RECTS); keys_exact = r.fetch(R.DBG_KEYS_UNSORTED); fb_exact = r.fetch(R.DBG_FRAMEBUFFER)
    r.set_option(R.OPT_KEEP_INTERMEDIATES, 0)                                # bounded K1
    r.draw([rid], fr)
    bound = r.fetch(R.DBG_TRECTS); keys_bound = r.fetch(R.DBG_KEYS_UNSORTED); fb_bound = r.fetch(R.DBG_FRAMEBUFFER)
    assert np.array_equal(fb_exact, fb_bound)
    vis = exact["x0"] <= exact["x1"]
    assert vis.sum() > 0.2 * n
    culled, tx0, ty0, bw, bh = _unpack_trects(bound)
    assert not culled[vis].any(), f"{int(culled[vis].sum())} visible splats dropped by the bound"
    assert np.array_equal(keys_bound[vis], keys_exact[vis])
    assert np.all(keys_bound[culled] == 0xFFFFFFFF)
    ex0 = exact["x0"][vis].astype(np.int64) // 16; ex1 = exact["x1"][vis].astype(np.int64) // 16
    ey0 = exact["y0"][vis].astype(np.int64) // 16; ey1 = exact["y1"][vis].astype(np.int64) // 16
    bx0, by0, bw, bh = tx0[vis], ty0[vis], bw[vis], bh[vis]
    assert np.all(bx0 <= ex0) and np.all(by0 <= ey0)
    assert np.all((bx0 + bw >= ex1) | (bw == 127)) and np.all((by0 + bh >= ey1) | (bh == 127))
    # the bound is not vacuous: on the plain cloud it keeps well under twice the tiles the exact rectangles touch
    if case == "cloud":
        t_exact = ((ex1 - ex0 + 1) * (ey1 - ey0 + 1)).sum()
        t_bound = ((bw + 1) * (bh + 1)).sum()
        assert t_bound < 2.0 * t_exact, (t_bound, t_exact)
    r.close()
