"""CPU pin of the inequality behind the bounded K1 (csrc/project.cu project_bound_kernel).

For every splat the exact projection keeps, the exact pixel rectangle (oracle, fp32 spec path) must lie inside
centre ± h with h = min(rr, 2 sqrt 2) * sqrt(2 (|J|_2^2 |W|_2^2 lambda_max(Sigma) + 0.3)) (1 + 1e-4) + 0.01, where
|J|_2^2 = j0^2 (1 + rx^2 + ry^2) (clamped ratios, LIB.h:44-60).  The formula is restated here in float64 numpy WITHOUT the
kernel's safety margins, so the test fails if the mathematics (not just the rounding allowance) were wrong.  The GPU
kernel itself is checked splat by splat in tests/test_gpu_parity.py::test_bounded_k1_rectangles_contain_the_exact_ones."""
import numpy as np
import pytest


def _sigma_f64(cl, obj):
    s = cl.scale_h.astype(np.float64)
    q = cl.orient_h.astype(np.float64)          # (x, y, z, w), NOT re-normalised (LIB.h:10-35)
    x, y, z, r = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    Rt = np.empty((cl.n, 3, 3))
    Rt[:, 0, 0] = 1 - 2 * (y * y + z * z); Rt[:, 0, 1] = 2 * (x * y + r * z); Rt[:, 0, 2] = 2 * (x * z - r * y)
    Rt[:, 1, 0] = 2 * (x * y - r * z); Rt[:, 1, 1] = 1 - 2 * (x * x + z * z); Rt[:, 1, 2] = 2 * (y * z + r * x)
    Rt[:, 2, 0] = 2 * (x * z + r * y); Rt[:, 2, 1] = 2 * (y * z - r * x); Rt[:, 2, 2] = 1 - 2 * (x * x + y * y)
    M = s[:, :, None] * Rt                      # S * R^T
    M = M @ obj[:3, :3].T[None]                 # * transpose(mat3(ObjectMatrix))
    return np.swapaxes(M, 1, 2) @ M             # Sigma = M^T M


@pytest.mark.parametrize("case", ["cloud", "aniso", "bigsplats", "closeup", "objmat"])
def test_exact_rectangles_lie_inside_the_bound(oracle, scene, case):
    O, S = oracle, scene
    rng = np.random.default_rng(3)
    n, w, h, theta, mult = 150_000, 1920, 1080, 40.0, 1.0
    if case == "bigsplats":
        n, mult = 20_000, 40.0
    if case == "closeup":
        n, mult, theta = 50_000, 6.0, 0.0
    cl = S.make_cloud(n, 4321, sh=False, scale_mult=mult)
    if case == "aniso":
        sc = cl.scale_h.astype(np.float32); sc[:, 0] *= 30.0; sc[:, 2] *= 0.05; cl.scale_h = sc.astype(np.float16)
        cl.orient_h[:] = (cl.orient_h.astype(np.float32) * rng.uniform(0.8, 1.25, (cl.n, 1))).astype(np.float16)
    if case == "closeup":
        cl.pos[:, 2] = cl.pos[:, 2] * np.float32(0.2) + np.float32(2.6)
    obj = np.eye(4)
    if case == "objmat":
        a = 0.7
        obj[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]) @ np.diag([1.7, 0.6, 1.1])
        obj[:3, 3] = [0.1, 0.05, -0.1]
    base = S.orbit_frame(w, h, theta)
    fr = S.Frame(w, h, base.view, base.proj, S.colmajor(obj), S.colmajor(np.linalg.inv(obj)))
    F = O.make_frame(fr, O.camera_from_view(fr.view), cl.barycentre(), 0)
    pr = O.project(F, cl)
    vis = pr["vis"] > 0
    assert vis.sum() > 0.2 * n

    view = np.asarray(fr.view, np.float64).reshape(4, 4).T
    proj = np.asarray(fr.proj, np.float64).reshape(4, 4).T
    W3 = view[:3, :3]
    wnorm2 = np.linalg.eigvalsh(W3.T @ W3).max()
    lam = np.linalg.eigvalsh(_sigma_f64(cl, obj)).max(axis=1)
    p = cl.pos.astype(np.float64)
    t = (view @ np.c_[p, np.ones(n)].T).T
    aspect = proj[0, 0] / proj[1, 1]                        # LIB.h:44-49 as written: tanFovY = 1 / (P11 * aspect) = tanFovX
    tan_x, tan_y = 1.0 / proj[0, 0], 1.0 / (proj[1, 1] * aspect)
    rx = np.clip(t[:, 0] / t[:, 2], -1.3 * tan_x, 1.3 * tan_x)
    ry = np.clip(t[:, 1] / t[:, 2], -1.3 * tan_y, 1.3 * tan_y)
    j0 = (w * proj[0, 0] / 2.0) / t[:, 2]
    l1 = j0 * j0 * (1 + rx * rx + ry * ry) * wnorm2 * lam + 0.3
    s1 = np.minimum(np.sqrt(2 * l1), 4096.0)
    pmax = pr["recs"]["pmax"].astype(np.float64)
    rr = np.sqrt(np.maximum(pmax, 0.0))
    hb = np.minimum(rr, 2 * np.sqrt(2.0)) * s1 * (1 + 1e-4) + 0.01
    hb = hb * (1 + 1e-5) + 1e-3                 # fp32 rounding of the exact chain (the kernel allows 2e-3 and 0.06 px)
    cx = pr["recs"]["cx"].astype(np.float64); cy = pr["recs"]["cy"].astype(np.float64)
    r = pr["rects"]
    x0 = np.maximum(np.ceil(cx - hb - 0.5), 0); x1 = np.minimum(np.floor(cx + hb - 0.5), w - 1)
    y0 = np.maximum(np.ceil(cy - hb - 0.5), 0); y1 = np.minimum(np.floor(cy + hb - 0.5), h - 1)
    for name, lo_ok in (("x0", x0[vis] <= r["x0"][vis]), ("x1", x1[vis] >= r["x1"][vis]),
                        ("y0", y0[vis] <= r["y0"][vis]), ("y1", y1[vis] >= r["y1"][vis])):
        assert lo_ok.all(), f"{case}: {int((~lo_ok).sum())} exact rectangles stick out of the bound on {name}"
    if case == "cloud":                          # and the bound is not vacuous: a square around the major axis, < 2.5x the exact pixel area
        area_b = ((x1 - x0 + 1) * (y1 - y0 + 1))[vis].sum()
        area_e = ((r["x1"].astype(np.int64) - r["x0"] + 1) * (r["y1"].astype(np.int64) - r["y0"] + 1))[vis].sum()
        assert area_b < 2.5 * area_e, (area_b, area_e)
