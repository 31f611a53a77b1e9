#!/usr/bin/env python
"""tools/sweep.py — one scene, many settings: per-stage CUDA-event times of the hot path for a list of depth-chunk counts.
Generates the cloud once (the 20 M scene takes ~16 s on the host), so a whole sweep costs one bench run of GPU time.

    python tools/sweep.py --workload 20M_sh3_1080p --chunks 1,2,3,4,5,6 --steps 10
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="20M_sh3_1080p")
    ap.add_argument("--chunks", default="1,2,3,4,5,6")
    ap.add_argument("--shifts", default="0", help="GSB_OPT_CHUNK_SHIFT values (first chunk = V / 2^shift; 0 = auto)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--env", default="", help="semicolon-separated experiment settings, each a comma-separated list of "
                                              "NAME=VALUE library knobs (GSB_K1_OCC, GSB_BLEND_OBB), e.g. 'GSB_K1_OCC=4;GSB_K1_OCC=6'")
    ap.add_argument("--lib", default="", help="experiment library built by tools/build_variants.sh (name in _exp/); default = the product library")
    a = ap.parse_args()
    import torch
    from houdini_gsplat_renderer_b200 import renderer as R, scene as S
    if a.lib:
        R.LIB_PATH = ROOT / "houdini-gsplat-renderer_b200" / "_exp" / f"lib_{a.lib}.so"
    w = S.WORKLOADS[a.workload]
    cloud = S.make_cloud(w["n"], w["seed"], sh=w["sh"])
    r = R.GSplatRenderer(0)
    r.set_option(R.OPT_SPLAT_CAP, 0)
    r.set_option(R.OPT_STAGE_TIMING, 1)
    stream = torch.cuda.current_stream()
    r.set_stream(stream.cuda_stream)
    r.setSphericalHarmonicsOrder(3 if w["sh"] else 0)
    rid = r.registerUpdate(0xB200, (1, 0, 0, 0), 0, cloud)
    fb = torch.zeros((w["height"], w["width"], 4), dtype=torch.float32, device="cuda")
    ref = None
    import os
    settings = [(c, sh, e) for e in (a.env.split(";") if a.env else [""]) for c in [int(x) for x in a.chunks.split(",")]
                for sh in [int(x) for x in a.shifts.split(",")]]
    for c, sh, envs in settings:
        r.set_option(R.OPT_CHUNK_SHIFT, sh)
        for kv in filter(None, envs.split(",")):
            k, v = kv.split("="); os.environ[k] = v
        r.set_option(R.OPT_DEPTH_CHUNKS, c)

        def step(i):
            fr = S.orbit_frame(w["width"], w["height"], float(i % 360) if w["orbit"] else 0.0)
            r.includeInRenderPass(rid); r.generateRenderGeometry()
            r.render(fr, device_rgba=fb.data_ptr())
            r.postRender()

        for i in range(a.warmup):
            step(i)
        acc = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(a.steps):
            step(a.warmup + i)
            st = r.stats()
            for k in ("ms_project", "ms_sort", "ms_records", "ms_bin", "ms_blend", "n_visible", "n_live", "n_instances",
                      "n_consumed", "launches", "depth_chunks"):
                acc[k] = acc.get(k, 0.0) + st[k]
        e1.record(); torch.cuda.synchronize()
        out = {k: round(v / a.steps, 4) for k, v in acc.items()}
        out["ms_frame"] = round(e0.elapsed_time(e1) / a.steps, 4)
        out["chunks_requested"] = c
        out["chunk_shift"] = sh
        out["env"] = envs
        out["workload"] = a.workload
        out["lib"] = a.lib or "product"
        if not w["orbit"]:
            cur = fb.cpu().numpy()
            if ref is None:
                ref = cur
            out["frame_equals_first_setting"] = bool(np.array_equal(ref, cur))
        print(json.dumps(out), flush=True)
    r.close()


if __name__ == "__main__":
    main()
