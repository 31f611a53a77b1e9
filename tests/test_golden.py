"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle).
CPU: the oracle still reproduces them bit for bit.  GPU (-m gpu): the CUDA path reproduces them too."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
import make_golden as G  # noqa: E402

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("name", list(G.CASES))
def test_oracle_reproduces_golden(name):
    g = np.load(GOLD / f"{name}.npz")
    _, _, p = G.run_case(name)
    assert np.array_equal(p["keys"], g["keys"]) and np.array_equal(p["order"], g["order"])
    assert np.array_equal(p["vis"], g["vis"])
    vis = g["vis"] > 0
    assert np.array_equal(p["rects"].view(np.uint16).reshape(-1, 4)[vis], g["rects"][vis])
    assert np.array_equal(p["recs"].view(np.uint32).reshape(-1, 12)[vis], g["recs"][vis])
    assert np.array_equal(p["tile_start"], g["tile_start"]) and np.array_equal(p["inst"], g["inst"])
    assert np.abs(p["rgba"] - g["rgba"]).max() <= 1e-6       # libm exp may differ between hosts by an ulp


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.CASES))
def test_cuda_reproduces_golden(name):
    from gpu_util import gpu_pipeline
    g = np.load(GOLD / f"{name}.npz")
    n, seed, sh, mult, w, h, theta, order = G.CASES[name]
    from houdini_gsplat_renderer_b200 import scene as S
    cl = S.make_cloud(n, seed, sh=sh, scale_mult=mult)
    out = gpu_pipeline(cl, S.orbit_frame(w, h, theta), order)
    vis = g["vis"] > 0
    assert np.array_equal(out["keys"], g["keys"])
    assert np.array_equal(out["order"].astype(np.int32), g["order"][:int(vis.sum())])    # culled splats are never sorted
    assert np.array_equal(out["rects"].view(np.uint16).reshape(-1, 4)[vis], g["rects"][vis])
    assert np.array_equal(out["recs"].view(np.uint32).reshape(-1, 12)[vis], g["recs"][vis])
    assert np.array_equal(out["inst"].astype(np.int32), g["inst"])
    assert np.abs(out["rgba"] - g["rgba"]).max() <= 2e-5
