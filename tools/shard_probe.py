"""Profiling helper: one GPU renders ONE shard (rank r of world G) of the row-partitioned frame,
so the per-rank stage times of a G-GPU run can be measured without G GPUs."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from houdini_gsplat_renderer_b200 import multigpu as M, renderer as R, scene as S

wl = sys.argv[1] if len(sys.argv) > 1 else "20M_sh3_1080p"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
w = S.WORKLOADS[wl]
cl = S.make_cloud(w["n"], w["seed"], sh=w["sh"])
fr = S.orbit_frame(w["width"], w["height"], 0.0)
group = M.default_row_group(w["height"], world)
r = R.GSplatRenderer(0)
r.set_option(R.OPT_SPLAT_CAP, 0); r.set_option(R.OPT_STAGE_TIMING, 1)
rid = r.registerUpdate(1, (1, 0, 0, 0), 0, cl); r.setSphericalHarmonicsOrder(3 if w["sh"] else 0)
for pre in (0, 0):
    acc = {}
    for it in range(8):
        r.draw([rid], fr, row_rank=world // 2, row_world=world, row_group=group)
        st = r.stats()
        if it >= 3:
            for k in ("ms_project", "ms_sort", "ms_bin", "ms_blend", "ms_total"):
                acc[k] = acc.get(k, 0) + st[k] / 5
    print(f"{wl} shard {world // 2}/{world} group {group} precull={pre} V={st['n_visible']} " +
          " ".join(f"{k}={v:.3f}" for k, v in acc.items()), flush=True)
