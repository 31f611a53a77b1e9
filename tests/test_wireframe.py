"""SURVEY §8 f-4 — wireframe / selection overlay (GR_GSplat.C:376-421, 474-483; wire shader SRC.h:22-110).
CPU: the oracle's restatement against the reference's OWN wire vertex shader compiled for the host (oracle/_ref), all 8
vertices of every splat.  GPU (-m gpu): the CUDA kernel through the C ABI, bit-exact vs the oracle (vertices, colours and
the rasterised overlay)."""
import numpy as np
import pytest

from oracle import ref as RF


def _scene(O, S, n=3000, w=640, h=360, theta=33.0, seed=11, mult=3.0, obj=None):
    cl = S.make_cloud(n, seed, sh=True, scale_mult=mult)
    fr = S.orbit_frame(w, h, theta)
    if obj is not None:
        fr = S.Frame(w, h, fr.view, fr.proj, S.colmajor(obj), S.colmajor(np.linalg.inv(obj)))
    F = O.make_frame(fr, O.camera_from_view(fr.view), cl.barycentre(), 0)
    return cl, fr, F


@pytest.mark.skipif(not RF.available(), reason="oracle/_ref not built and /root/reference absent")
@pytest.mark.parametrize("with_obj", [False, True])
def test_oracle_wire_vertices_match_the_reference_wire_shader(oracle, scene, with_obj):
    O, S = oracle, scene
    obj = None
    if with_obj:                                   # the wire shader ignores the object matrix in the covariance (SRC.h:74)
        obj = np.eye(4); obj[:3, :3] *= 1.3; obj[:3, 3] = [0.1, -0.05, 0.02]
    cl, fr, F = _scene(O, S, n=1500, obj=obj)
    verts, cols = O.wire_vertices(F, cl)
    b = RF.Bound(cl, fr, F.cam[:], cl.barycentre(), 0)       # binds the glH_* uniforms the wire shader reads
    cd = cl.cd_h.astype(np.float32); sc = cl.scale_h.astype(np.float32); orr = cl.orient_h.astype(np.float32)
    checked = 0
    for i in range(cl.n):
        for v in range(8):
            g, c = RF.wire_vertex(8 * i + v, cl.pos[i], cd[i], sc[i], orr[i])
            want = verts[8 * i + v]
            if not np.all(np.isfinite(g)):
                continue
            tol = 2e-5 * max(1.0, float(np.abs(g).max())) + 1e-3 * abs(float(g[3])) * 2.0 / 360      # 1e-3 px of the quad offset
            assert np.allclose(want, g, rtol=0, atol=tol), (i, v, want, g)
            assert np.array_equal(cols[8 * i + v], c)
            checked += 1
    assert checked > 8 * 1400
    # line topology: vertices (1,2) (3,4) (5,6) (7,0) coincide — four edges of one closed quad
    q = verts.reshape(cl.n, 8, 4)
    for a, b2 in ((1, 2), (3, 4), (5, 6), (7, 0)):
        assert np.array_equal(q[:, a], q[:, b2], equal_nan=True)


def test_oracle_overlay_draws_closed_outlines(oracle, scene):
    O, S = oracle, scene
    cl, fr, F = _scene(O, S, n=50, w=320, h=180, mult=8.0)
    verts, _ = O.wire_vertices(F, cl)
    img = O.wire_overlay(verts, cl, 320, 180)
    on = img[..., 3] > 0
    assert 200 < on.sum() < 0.5 * on.size
    assert np.all(img[on][:, 3] == 1.0)
    # every lit pixel carries the colour of some splat
    cds = cl.cd_h.astype(np.float32)
    lit = img[on][:, :3]
    assert all(np.any(np.all(cds == p, axis=1)) for p in lit[:200])


@pytest.mark.gpu
@pytest.mark.parametrize("with_obj", [False, True])
def test_cuda_wireframe_matches_oracle_bit_for_bit(oracle, scene, with_obj):
    import torch
    from houdini_gsplat_renderer_b200 import renderer as R
    O, S = oracle, scene
    obj = None
    if with_obj:
        obj = np.eye(4); obj[:3, :3] *= 1.3; obj[:3, 3] = [0.1, -0.05, 0.02]
    cl, fr, F = _scene(O, S, n=100_000, w=1280, h=720, mult=2.0, obj=obj)
    cl.pos[:50, 2] += np.float32(6.0)                          # behind the camera: vertices are still produced, the overlay skips them
    want_v, want_c = O.wire_vertices(F, cl)
    r = R.GSplatRenderer(0)
    rid = r.registerUpdate(0x31, (1, 0, 0, 0), 0, cl)
    frame = torch.zeros((720, 1280, 4), dtype=torch.float32, device="cuda")
    frame[..., 2] = 0.25                                       # a background the overlay must leave alone
    host = np.zeros((720, 1280, 4), np.float32)
    got_v, got_c = r.renderWireframe(rid, fr, cl.n, overlay_rgba=frame.data_ptr(), overlay_host=host)
    assert np.array_equal(got_v, want_v, equal_nan=True)
    assert np.array_equal(got_c, want_c)
    bg = np.zeros((720, 1280, 4), np.float32); bg[..., 2] = 0.25
    want_img = O.wire_overlay(want_v, cl, 1280, 720, rgba=bg)
    assert np.array_equal(host, want_img)
    assert (host[..., 3] == 1.0).sum() > 10_000
    with pytest.raises(R.GsbError):
        r.renderWireframe("no such id", fr, 1)
    r.close()
