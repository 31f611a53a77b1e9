"""-m gpu: SURVEY §8 f-1 — attribute ingestion on the GPU (gsb_update_from_attributes) equals the CPU restatement of
GR_PrimGsplat::update (oracle/ingest.py) bit for bit, for the three SH encodings, defaults and the Alpha rule, and
a prim ingested this way renders the same frame as one registered through registerUpdate."""
import numpy as np
import pytest

from oracle import ingest as I

pytestmark = pytest.mark.gpu


def raw_attrs(n, seed, enc):
    rng = np.random.default_rng(seed)
    a = {"P": (rng.random((n, 3), dtype=np.float32) * 2 - 1),
         "Cd": rng.random((n, 3), dtype=np.float32) * 1.2 - 0.1,
         "opacity": rng.random(n, dtype=np.float32),
         "scale": np.exp(rng.standard_normal((n, 3)).astype(np.float32) * 0.5) * np.float32(0.03),
         "orient": rng.standard_normal((n, 4)).astype(np.float32)}
    a["orient"] /= np.linalg.norm(a["orient"], axis=1, keepdims=True)
    sh = (rng.standard_normal((n, 15, 3)) * 0.15).astype(np.float32)
    sh[0, 0] = [70000.0, -70000.0, 1e-9]          # half overflow -> inf, underflow -> 0
    if enc == "array":
        a["sh_coefficients"] = sh
    elif enc == "array_short":
        a["sh_coefficients"] = sh[:, :8]          # fewer than 15 entries: the rest stay zero
    elif enc == "sh_attrs":
        for j in range(15):
            a[f"sh{j + 1}"] = sh[:, j].copy()
    elif enc == "f_rest":
        for j in range(15):
            for ch in range(3):
                a[f"f_rest_{j + 15 * ch}"] = sh[:, j, ch].copy()
    elif enc == "incomplete":
        for j in range(14):                       # sh15 missing: encoding does not count
            a[f"sh{j + 1}"] = sh[:, j].copy()
    return a


@pytest.mark.parametrize("enc", ["array", "array_short", "sh_attrs", "f_rest", "incomplete", "none"])
def test_ingest_matches_cpu_restatement(enc):
    from houdini_gsplat_renderer_b200 import renderer as R
    a = raw_attrs(5003, 3, enc)
    a["gsplat__sh_order"] = 2
    r = R.GSplatRenderer(0)
    res = r.update(0x55, (3, 0, 0, 1), 12, a)
    ref = I.update(a)
    assert res["id"] == "0x55__12__3_0_0_1" and res["sh_order"] == 2 and not res["sh_order_invalid"]
    assert res["sh_data_found"] == (ref["shx_h"] is not None) == (enc not in ("incomplete", "none"))
    assert np.array_equal(res["barycentre"], ref["barycentre"])
    n = a["P"].shape[0]
    assert np.array_equal(r.fetch_entry(res["id"], 0).reshape(n, 3), ref["pos"])
    assert np.array_equal(r.fetch_entry(res["id"], 1).reshape(n, 3), ref["cd_h"].view(np.uint16))
    assert np.array_equal(r.fetch_entry(res["id"], 2), ref["alpha"])
    assert np.array_equal(r.fetch_entry(res["id"], 3).reshape(n, 3), ref["scale_h"].view(np.uint16))
    assert np.array_equal(r.fetch_entry(res["id"], 4).reshape(n, 4), ref["orient_h"].view(np.uint16))
    for which, name in ((5, "shx_h"), (6, "shy_h"), (7, "shz_h")):
        got = r.fetch_entry(res["id"], which)
        if ref[name] is None:
            assert got.size == 0
        else:
            assert np.array_equal(got.reshape(n, 16), ref[name].view(np.uint16))
    r.close()


def test_defaults_alpha_rule_and_sh_order_validation():
    from houdini_gsplat_renderer_b200 import renderer as R
    rng = np.random.default_rng(1)
    n = 257
    a = {"P": rng.random((n, 3), dtype=np.float32), "opacity": np.full(n, 0.25, np.float32),
         "Alpha": np.full(n, 0.75, np.float32), "gsplat__sh_order": 7,
         "gsplat__explicit_camera_pos": (1.0, 2.0, 3.0)}
    r = R.GSplatRenderer(0)
    res = r.update(1, (0, 0, 0, 0), 0, a)
    assert res["sh_order"] == 0 and res["sh_order_invalid"]                    # GR.C:447-452
    assert res["set_explicit_camera"] and np.array_equal(res["explicit_camera"], [1, 2, 3])
    assert np.all(r.fetch_entry(res["id"], 2) == np.float32(0.75))             # Alpha preferred over opacity
    assert np.all(r.fetch_entry(res["id"], 1) == 0)                            # Cd default 0
    assert np.all(r.fetch_entry(res["id"], 3) == np.float16(1).view(np.uint16))   # scale default 1
    assert np.array_equal(r.fetch_entry(res["id"], 4).reshape(n, 4)[5], np.array([0, 0, 0, 1], np.float16).view(np.uint16))
    del a["Alpha"], a["opacity"], a["gsplat__sh_order"]
    res = r.update(1, (1, 0, 0, 0), 0, a)
    assert res["sh_order"] == 3 and r.registry_size() == 1                     # default order 3; old version evicted
    assert np.all(r.fetch_entry(res["id"], 2) == np.float32(1.0))              # alpha default 1
    r.close()


def test_ingested_prim_renders_like_a_registered_one(scene):
    from houdini_gsplat_renderer_b200 import renderer as R
    S = scene
    a = raw_attrs(30_000, 9, "f_rest")
    a["opacity"] = a["opacity"] * 0.9 + 0.05
    ref = I.update(a)
    cl = S.SplatCloud(ref["pos"], ref["cd_h"], ref["alpha"], ref["scale_h"], ref["orient_h"], ref["shx_h"], ref["shy_h"], ref["shz_h"])
    fr = S.orbit_frame(320, 180, 12.0)
    r = R.GSplatRenderer(0)
    res = r.update(2, (1, 0, 0, 0), 0, a)
    r.setSphericalHarmonicsOrder(res["sh_order"])
    f1 = np.zeros((180, 320, 4), np.float32); r.draw([res["id"]], fr, host_rgba=f1)
    rid = r.registerUpdate(3, (1, 0, 0, 0), 0, cl)
    f2 = np.zeros_like(f1); r.draw([rid], fr, host_rgba=f2)
    assert f1[..., 3].max() > 0.3 and np.array_equal(f1, f2)
    r.close()


def test_inria_ply_activation_on_the_gpu(tmp_path, oracle, scene):
    """a raw INRIA .ply goes to the GPU as is; the wrangle conversion (SH_C0 f_dc + 0.5, sigmoid, exp,
    quaternion reorder + normalise) and the half cast run in the ingestion kernel and equal the numpy restatement bit for
    bit (deterministic exp on both sides).  The scene is surface-like (nested shells), unlike the benchmark's cube; the
    production frame equals the oracle's frame of the activated attributes and the exact-K1 frame."""
    from houdini_gsplat_renderer_b200 import ply, renderer as R
    O, S = oracle, scene
    cols = S.make_inria_shell_columns(200_000, 5)
    cols["opacity"][:5] = [-200.0, 200.0, 0.0, -9.0, 40.0]           # saturating logits
    cols["scale_0"][:3] = [-30.0, 12.0, 0.0]                         # half underflow / overflow of exp(scale)
    cols["rot_0"][7] = cols["rot_1"][7] = cols["rot_2"][7] = cols["rot_3"][7] = 0.0   # zero quaternion: left as is
    path = tmp_path / "shells.ply"
    ply.write_ply(path, cols)
    raw = ply.load_inria(path)
    ref = I.update(I.activate_inria(raw))
    r = R.GSplatRenderer(0)
    res = r.update(0x9A, (1, 0, 0, 0), 0, raw, activation=R.ACT_INRIA)
    assert res["sh_data_found"]
    for which, name in ((0, "pos"), (1, "cd_h"), (2, "alpha"), (3, "scale_h"), (4, "orient_h"), (5, "shx_h"), (6, "shy_h"), (7, "shz_h")):
        got = r.fetch_entry(res["id"], which)
        want = np.ascontiguousarray(ref[name]).view(np.uint16 if which not in (0, 2) else np.float32).reshape(-1)
        assert np.array_equal(got.view(want.dtype), want), name
    # render it: production path == exact-K1 path == oracle (surface-like scene, camera outside the shells)
    cl = S.SplatCloud(ref["pos"], ref["cd_h"], ref["alpha"], ref["scale_h"], ref["orient_h"], ref["shx_h"], ref["shy_h"], ref["shz_h"])
    fr = S.orbit_frame(640, 360, 30.0)
    r.setSphericalHarmonicsOrder(3)
    f1 = np.zeros((360, 640, 4), np.float32)
    r.draw([res["id"]], fr, host_rgba=f1)
    st = r.stats()
    r.set_option(R.OPT_LAZY_PROJECT, 0)
    f2 = np.zeros_like(f1); r.draw([res["id"]], fr, host_rgba=f2)
    assert np.array_equal(f1, f2)
    F = O.make_frame(fr, O.camera_from_view(fr.view), np.asarray(res["barycentre"], np.float32), 3)
    o = O.pipeline(F, cl)
    assert np.abs(f1 - o["rgba"]).max() <= 2e-5 and st["n_consumed"] == o["n_consumed"]
    assert f1[..., 3].max() > 0.9
    r.close()


def test_config1_simple_scene_cook(oracle, scene):
    """BASELINE config 1 (hip/GSplatPlugin_simpleScene_v001.hip: GSplatSOP builds prims from 10 k synthetic points, no
    render), SURVEY.md §8d: the 10 000-point seed-1001 SH-3 cloud -> HDK-free model of cookMySop/build (one prim, vertex i
    <-> point i, barycentre = sequential fp32 sum / N, GEO_GSplat.C:338-351,413-431) -> update()-equivalent extraction
    on the GPU (fp16 quantisation, GR_GSplat.C:302-372) -> registry.  Vertex wiring, barycentre bits and every
    quantised array equal the oracle's; the cook is timed (reported, no roofline claim)."""
    import time
    from houdini_gsplat_renderer_b200 import renderer as R
    O, S = oracle, scene
    w = S.WORKLOADS["10k_sh3_cook"]
    cl = S.make_cloud(w["n"], w["seed"], sh=True)
    n = cl.n
    assert n == 10_000
    v2p, bary, bbox = O.build_prim(cl.pos)                      # cookMySop / GEO_PrimGsplat::build model
    assert np.array_equal(v2p, np.arange(n, dtype=np.int32))
    assert np.array_equal(bbox, np.r_[cl.pos.min(axis=0), cl.pos.max(axis=0)].astype(np.float32))
    f32 = np.float32
    sh = np.stack([cl.shx_h[:, :15].astype(f32), cl.shy_h[:, :15].astype(f32), cl.shz_h[:, :15].astype(f32)], axis=2)
    a = {"P": cl.pos, "Cd": cl.cd_h.astype(f32), "opacity": cl.alpha, "scale": cl.scale_h.astype(f32),
         "orient": cl.orient_h.astype(f32), "sh_coefficients": np.ascontiguousarray(sh)}
    ref = I.update(a)
    r = R.GSplatRenderer(0)
    r.update(0xC1, (1, 0, 0, 0), 0, a)                            # warm-up (allocations)
    t0 = time.perf_counter()
    res = r.update(0xC1, (2, 0, 0, 0), 0, a)                      # a new version evicts the old entry (R.C:246-265)
    dt = time.perf_counter() - t0
    assert r.registry_size() == 1
    assert res["sh_data_found"] and res["sh_order"] == 3 and not res["sh_order_invalid"]
    assert np.array_equal(res["barycentre"], bary) and np.array_equal(ref["barycentre"], bary)
    rid = res["id"]
    assert np.array_equal(r.fetch_entry(rid, 0).reshape(n, 3), cl.pos)
    assert np.array_equal(r.fetch_entry(rid, 1).reshape(n, 3), cl.cd_h.view(np.uint16))
    assert np.array_equal(r.fetch_entry(rid, 2), cl.alpha)
    assert np.array_equal(r.fetch_entry(rid, 3).reshape(n, 3), cl.scale_h.view(np.uint16))
    assert np.array_equal(r.fetch_entry(rid, 4).reshape(n, 4), cl.orient_h.view(np.uint16))
    for which, arr in ((5, cl.shx_h), (6, cl.shy_h), (7, cl.shz_h)):
        got = r.fetch_entry(rid, which).reshape(n, 16)
        assert np.array_equal(got[:, :15], arr.view(np.uint16)[:, :15])
    print(f"config 1: cook of {n} points (H2D of raw fp32 attributes + GPU quantisation + registry) {dt * 1e6:.0f} us")
    r.close()
