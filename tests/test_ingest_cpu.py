"""CPU: the restatement of GR_PrimGsplat::update (oracle/ingest.py) and the PLY activation."""
import numpy as np


def test_f_rest_channel_split_and_matrix_layout():
    from oracle import ingest as I
    n = 4
    a = {"P": np.zeros((n, 3), np.float32)}
    for k in range(45):
        a[f"f_rest_{k}"] = np.full(n, k, np.float32)
    o = I.update(a)
    # coefficient j: R = f_rest_j, G = f_rest_{j+15}, B = f_rest_{j+30}  (GR_GSplat.C:357-366); element j at (j/4, j%4)
    assert np.array_equal(o["shx_h"][0, :15], np.arange(15, dtype=np.float16))
    assert np.array_equal(o["shy_h"][0, :15], np.arange(15, 30, dtype=np.float16))
    assert np.array_equal(o["shz_h"][0, :15], np.arange(30, 45, dtype=np.float16))
    assert o["shx_h"][0, 15] == 0 and o["shx_h"].reshape(n, 4, 4)[0, 2, 1] == 9
    assert o["sh_order"] == 3 and np.all(o["alpha"] == 1) and np.all(o["scale_h"] == 1)
    assert np.array_equal(o["orient_h"][0], np.array([0, 0, 0, 1], np.float16))


def test_encoding_priority_and_alpha_rule():
    from oracle import ingest as I
    n = 3
    a = {"P": np.ones((n, 3), np.float32), "opacity": np.full(n, .2, np.float32), "Alpha": np.full(n, .9, np.float32),
         "sh_coefficients": np.full((n, 15, 3), 2.0, np.float32)}
    for j in range(15):
        a[f"sh{j + 1}"] = np.full((n, 3), 5.0, np.float32)
    o = I.update(a)
    assert np.all(o["alpha"] == np.float32(.9)) and np.all(o["shx_h"][:, :15] == 2)     # array encoding wins
    del a["sh_coefficients"]
    assert np.all(I.update(a)["shx_h"][:, :15] == 5)
    del a["sh7"]
    assert I.update(a)["shx_h"] is None                                               # incomplete encoding: no SH
    a["gsplat__sh_order"] = -1
    o = I.update(a)
    assert o["sh_order"] == 0 and o["sh_order_invalid"]


def test_ply_round_trip_and_activation(tmp_path):
    from houdini_gsplat_renderer_b200 import ply
    rng = np.random.default_rng(0)
    n = 1000
    cols = {k: rng.standard_normal(n).astype(np.float32) for k in
            ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{j}" for j in range(45)] +
            ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]}
    p = tmp_path / "t.ply"
    ply.write_ply(p, cols)
    back = ply.read_ply(p)
    assert list(back) == list(cols) and all(np.array_equal(back[k], cols[k]) for k in cols)
    raw = ply.load_inria(p)                                   # raw columns: the activation runs on the GPU (csrc/ingest.cu)
    assert np.array_equal(raw["Cd"][:, 1], cols["f_dc_1"]) and np.array_equal(raw["orient"][:, 0], cols["rot_0"])
    assert np.array_equal(raw["f_rest_17"], cols["f_rest_17"]) and raw["P"].shape == (n, 3)
    # the restatement of the activation (what the ingestion kernel computes) against libm in double
    from oracle import ingest as I
    a = I.activate_inria(raw)
    assert np.allclose(a["Cd"][:, 1], 0.28209479177387814 * cols["f_dc_1"] + 0.5, atol=1e-7)
    assert np.array_equal(a["opacity"], (1 / (1 + np.exp(-cols["opacity"].astype(np.float64)))).astype(np.float32)) or \
        np.abs(a["opacity"].view(np.int32) - (1 / (1 + np.exp(-cols["opacity"].astype(np.float64)))).astype(np.float32).view(np.int32)).max() <= 1
    assert np.abs(a["scale"][:, 2].view(np.int32) - np.exp(cols["scale_2"].astype(np.float64)).astype(np.float32).view(np.int32)).max() <= 1
    q = a["orient"]
    assert np.allclose(np.linalg.norm(q, axis=1), 1, atol=1e-6)
    w = np.sqrt(cols["rot_0"] ** 2 + cols["rot_1"] ** 2 + cols["rot_2"] ** 2 + cols["rot_3"] ** 2)
    assert np.allclose(q[:, 3], cols["rot_0"] / w, atol=1e-6) and np.allclose(q[:, 0], cols["rot_1"] / w, atol=1e-6)
    # det_exp against libm over the whole useful range, and its special cases
    x = np.concatenate([np.linspace(-700, 700, 200001), rng.standard_normal(100000) * 5])
    rel = np.abs(I.det_exp(x) / np.exp(x) - 1.0)
    assert rel.max() < 4e-16
    assert I.det_exp(np.array([800.0]))[0] == np.inf and I.det_exp(np.array([-800.0]))[0] == 0.0 and np.isnan(I.det_exp(np.array([np.nan]))[0])
    # ascii variant
    pa = tmp_path / "a.ply"
    with open(pa, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\nend_header\n1 2 3\n4 5 6\n")
    assert np.array_equal(ply.read_ply(pa)["y"], np.array([2, 5], np.float32))
