"""Profiling helper: runs the device radix sort alone on N depth-like keys (for ncu captures)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from houdini_gsplat_renderer_b200 import renderer as R

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rng = np.random.default_rng(0)
if bits == 32:
    k = (rng.random(n, dtype=np.float32) * 20 + 1.5).view(np.uint32)
else:
    k = rng.integers(0, 1 << bits, n, dtype=np.uint64).astype(np.uint32)
v = np.arange(n, dtype=np.uint32)
r = R.GSplatRenderer(0)
for _ in range(3):
    ko, vo = r.sort_pairs(k, v, 0, bits)
assert np.all(np.diff(ko.astype(np.int64) & ((1 << bits) - 1)) >= 0)
print("sorted", n, bits)
