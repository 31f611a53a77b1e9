"""Profiling helper: one GPU renders ONE shard (the middle rank of world G) of the row-partitioned frame, so the
per-rank stage times of a G-GPU run can be measured without G GPUs (the p2p hand-off and the 4-byte fence are not part
of it).  Workloads that share a cloud (the 20 M ones) are generated once.

    python tools/shard_probe.py 8 20M_sh3_1080p 20M_sh3_8k        (CHUNKS=0,1,2 in the environment sweeps the depth-chunk count)
"""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from houdini_gsplat_renderer_b200 import multigpu as M, renderer as R, scene as S

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
wls = sys.argv[2:] or ["20M_sh3_1080p"]
clouds = {}
r = R.GSplatRenderer(0)
r.set_option(R.OPT_SPLAT_CAP, 0); r.set_option(R.OPT_STAGE_TIMING, 1)
rid, cur = None, None
for wl in wls:
    w = S.WORKLOADS[wl]
    key = (w["n"], w["seed"], w["sh"])
    if key != cur:
        cl = S.make_cloud(w["n"], w["seed"], sh=w["sh"])
        rid = r.registerUpdate(1, (len(clouds) + 1, 0, 0, 0), 0, cl); r.setSphericalHarmonicsOrder(3 if w["sh"] else 0)
        clouds[key] = True; cur = key
    fr = S.orbit_frame(w["width"], w["height"], 0.0)
    group = int(os.environ["ROW_GROUP"]) if os.environ.get("ROW_GROUP") else M.default_row_group(w["height"], world)
    ranks = range(world) if os.environ.get("RANKS") == "all" else sorted({0, world // 2})
    for chunks, rank in [(c, k) for c in [int(x) for x in os.environ.get("CHUNKS", "0").split(",")] for k in ranks]:
        r.set_option(R.OPT_DEPTH_CHUNKS, chunks)
        acc = {}
        for it in range(10):
            r.draw([rid], fr, row_rank=rank, row_world=world, row_group=group)
            st = r.stats()
            if it >= 4:
                for k in ("ms_project", "ms_sort", "ms_records", "ms_bin", "ms_blend", "ms_total"):
                    acc[k] = acc.get(k, 0.0) + st[k] / 6
        print(json.dumps({"workload": wl, "world": world, "rank": rank, "depth_chunks": chunks, "row_group": group, "n_visible_bound": st["n_visible"],
                          "n_live": st["n_live"], "n_instances": st["n_instances"], "n_consumed": st["n_consumed"],
                          **{k: round(v, 4) for k, v in acc.items()}}), flush=True)
r.close()
