"""A ctypes-free host: tests/c_driver/gsb_driver.c is plain C, includes include/gsplat_b200.h, links -lgsplat_b200 and
replays the call sequence of the reference's hooks.  CPU: it compiles and links against the header
as a C translation unit (so the header is valid C and every symbol it uses resolves).  GPU: it renders a golden scene and
the frame equals the reference-derived golden frame and the ctypes path's frame."""
import ctypes as C
import struct
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import make_golden as G  # noqa: E402

SRC = ROOT / "tests" / "c_driver" / "gsb_driver.c"
LIBDIR = ROOT / "houdini-gsplat-renderer_b200"


def _build(tmp_path) -> Path:
    exe = tmp_path / "gsb_driver"
    subprocess.run(["gcc", "-std=c11", "-O1", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(SRC), "-o", str(exe),
                    f"-L{LIBDIR}", "-lgsplat_b200", f"-Wl,-rpath,{LIBDIR}"], check=True)
    return exe


def test_c_driver_compiles_and_links_as_plain_c(tmp_path):
    exe = _build(tmp_path)
    assert exe.exists()
    out = subprocess.run(["nm", "-u", str(exe)], capture_output=True, text=True, check=True).stdout
    for sym in ("gsb_create", "gsb_register_update", "gsb_generate_render_geometry", "gsb_render", "gsb_post_render",
                "gsb_flush_entries_for_matching_detail", "gsb_destroy"):
        assert sym in out, sym


@pytest.mark.gpu
def test_c_driver_renders_a_golden_scene(tmp_path, scene):
    from houdini_gsplat_renderer_b200 import renderer as R
    S = scene
    name = "g3k_sh3_160x90"
    n, seed, sh, mult, w, h, theta, order = G.CASES[name]
    cl = S.make_cloud(n, seed, sh=sh, scale_mult=mult)
    fr = S.orbit_frame(w, h, theta)
    fc = R.frame_to_c(fr)
    blob = struct.pack("<qii3fi", cl.n, 1 if sh else 0, order, *[float(x) for x in cl.barycentre()], 0) + bytes(fc)
    arrays = [np.ascontiguousarray(cl.pos, np.float32), cl.cd_h.view(np.uint16), np.ascontiguousarray(cl.alpha, np.float32),
              cl.scale_h.view(np.uint16), cl.orient_h.view(np.uint16)]
    if sh:
        arrays += [cl.shx_h.view(np.uint16), cl.shy_h.view(np.uint16), cl.shz_h.view(np.uint16)]
    scene_bin = tmp_path / "scene.bin"
    with open(scene_bin, "wb") as f:
        f.write(blob)
        for a in arrays:
            f.write(np.ascontiguousarray(a).tobytes())
    exe = _build(tmp_path)
    out_bin = tmp_path / "frame.bin"
    p = subprocess.run([str(exe), str(scene_bin), str(out_bin), "2"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr + p.stdout
    raw = out_bin.read_bytes()
    rgba = np.frombuffer(raw[:w * h * 16], np.float32).reshape(h, w, 4)
    st = R.StatsC.from_buffer_copy(raw[w * h * 16:w * h * 16 + C.sizeof(R.StatsC)])
    assert st.rendered == 1 and st.n_submitted == n and st.width == w and st.height == h
    # the same frame through the ctypes path
    r = R.GSplatRenderer(0)
    rid = r.registerUpdate(0x7f00c0de0000, (1, 0, 0, 0), 0, cl); r.setSphericalHarmonicsOrder(order)
    host = np.zeros((h, w, 4), np.float32)
    r.draw([rid], fr, host_rgba=host)
    r.close()
    assert np.array_equal(rgba, host)
    # and the reference-derived golden frame (the reference's own GLSL compiled for the host)
    g = np.load(ROOT / "tests" / "golden" / f"ref_{name}.npz")
    shape = tuple(int(x) for x in g["shape"])
    unsafe = np.unpackbits(g["unsafe"])[:shape[0] * shape[1]].reshape(shape).astype(bool)
    d = np.abs(rgba.astype(np.float64) - g["rgba"]).max(axis=2)
    assert d[~unsafe].max() <= 2e-5 + 1e-5 * 4.0
