"""Generates the golden vectors in this directory from the CPU oracle (oracle/gsplat_oracle.cpp).

    python tests/golden/make_golden.py

The reference ships no tests, fixtures or golden vectors, and cannot be run here, so these
are a REGRESSION PIN of the oracle (itself pinned by tests/test_oracle_literal.py and test_oracle_kat.py),
not reference outputs.  Inputs are regenerated from seeds by houdini-gsplat-renderer_b200/scene.py; only
outputs are stored (compressed .npz, a few hundred KB)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

CASES = {
    # name: (n, seed, sh, scale_mult, width, height, theta, sh_order)
    "g3k_sh3_160x90":   (3000, 2001, True, 2.0, 160, 90, 25.0, 3),
    "g5k_sh0_200x120":  (5000, 2002, False, 1.5, 200, 120, 200.0, 0),
    "g10k_sh2_192x108": (10000, 2003, True, 1.0, 192, 108, 310.0, 2),
}


def run_case(name):
    from houdini_gsplat_renderer_b200 import scene as S
    from oracle import oracle as O
    n, seed, sh, mult, w, h, theta, order = CASES[name]
    cl = S.make_cloud(n, seed, sh=sh, scale_mult=mult)
    fr = S.orbit_frame(w, h, theta)
    cam = O.camera_from_view(fr.view)
    F = O.make_frame(fr, cam, cl.barycentre(), order)
    return cl, fr, O.pipeline(F, cl)


def main():
    out = Path(__file__).resolve().parent
    for name in CASES:
        cl, fr, p = run_case(name)
        np.savez_compressed(out / f"{name}.npz", keys=p["keys"], order=p["order"].astype(np.int32), vis=p["vis"],
                            rects=p["rects"].view(np.uint16).reshape(-1, 4), recs=p["recs"].view(np.uint32).reshape(-1, 12),
                            tile_start=p["tile_start"], inst=p["inst"], rgba=p["rgba"], consumed=p["consumed"])
        print(name, "V", p["n_visible"], "D", p["inst"].shape[0], "bytes", (out / f"{name}.npz").stat().st_size)


if __name__ == "__main__":
    main()
