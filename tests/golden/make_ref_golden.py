"""Generates REFERENCE-DERIVED golden frames: the golden scenes of make_golden.py rendered by the reference's own shader
text compiled for the host (oracle/_ref/libgsplat_ref.so: oracle/build_ref.py reads the GLSL strings from
/root/reference/gsplat_plugin/shaders at build time; oracle/ref_harness.cpp supplies the ideal rasteriser and the blend
state of GSplatRenderer.C:613-621).  Needs /root/reference (or a prebuilt oracle/_ref); the fixtures travel without it.

    python tests/golden/make_ref_golden.py

Stored per scene: the frame (reference semantics: every fragment blended, no termination), the mask of pixels within
2e-5 (quad coordinates) of a support edge / the discard ring, where evaluation order decides coverage, and the fragment
counts.  tests/test_ref_golden.py checks the oracle (CPU) and the CUDA path (-m gpu) against them."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_golden as G  # noqa: E402

UNSAFE_TOL = 2e-5


def render_ref(name):
    from houdini_gsplat_renderer_b200 import scene as S
    from oracle import oracle as O, ref as RF
    n, seed, sh, mult, w, h, theta, order = G.CASES[name]
    cl = S.make_cloud(n, seed, sh=sh, scale_mult=mult)
    fr = S.orbit_frame(w, h, theta)
    cam = O.camera_from_view(fr.view)            # R.C:558-562 (double inverse -> f32); pinned separately in test_oracle_kat.py
    b = RF.Bound(cl, fr, cam, cl.barycentre(), order)
    rgba, st, unsafe = b.draw(unsafe_tol=UNSAFE_TOL)
    return rgba, st, unsafe


def main():
    out = Path(__file__).resolve().parent
    for name in G.CASES:
        rgba, st, unsafe = render_ref(name)
        np.savez_compressed(out / f"ref_{name}.npz", rgba=rgba, unsafe=np.packbits(unsafe), shape=np.array(unsafe.shape),
                            fragments=np.array([st["fragments_shaded"], st["fragments_kept"]], np.int64))
        print(name, st, "unsafe", float(unsafe.mean()), "bytes", (out / f"ref_{name}.npz").stat().st_size)


if __name__ == "__main__":
    main()
