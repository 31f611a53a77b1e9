"""Profiling helper for the cold path (geometry change): what the box can do (pinned H2D rate, threaded host memcpy rate)
beside what gsb_register_update + gsb_generate_render_geometry achieve, first call and steady state (second registration of
the same cloud under a new version).  GSB_TRACE_COLD=1 makes the library print its own phases.

    GSB_TRACE_COLD=1 python tools/cold_probe.py [20M_sh3_1080p]
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from houdini_gsplat_renderer_b200 import renderer as R, scene as S

wl = sys.argv[1] if len(sys.argv) > 1 else "20M_sh3_1080p"
w = S.WORKLOADS[wl]
out = {"workload": wl, "host_cores": os.cpu_count()}

# the box: pinned -> device
pin = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
dev = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for _ in range(2):
    dev.copy_(pin, non_blocking=True)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(3):
    dev.copy_(pin, non_blocking=True)
torch.cuda.synchronize()
out["pinned_h2d_GBps"] = 3 * (1 << 30) / (time.perf_counter() - t) / 1e9
# the box: one-thread pageable -> pinned memcpy
src = np.ones(1 << 30, np.uint8)
dstn = pin.numpy()
np.copyto(dstn, src)
t = time.perf_counter(); np.copyto(dstn, src); out["memcpy_1thread_GBps"] = (1 << 30) / (time.perf_counter() - t) / 1e9
# pageable -> device through the driver
srct = torch.from_numpy(src)
dev.copy_(srct); torch.cuda.synchronize()
t = time.perf_counter(); dev.copy_(srct); torch.cuda.synchronize(); out["pageable_h2d_GBps"] = (1 << 30) / (time.perf_counter() - t) / 1e9
del pin, dev, src, srct, dstn

cl = S.make_cloud(w["n"], w["seed"], sh=w["sh"])
nbytes = w["n"] * (132 if w["sh"] else 36)
r = R.GSplatRenderer(0)
r.set_option(R.OPT_SPLAT_CAP, 0)
for k in range(3):
    t0 = time.perf_counter()
    rid = r.registerUpdate(0xC01D, (k + 1, 0, 0, 0), 0, cl)
    t1 = time.perf_counter()
    r.includeInRenderPass(rid); r.generateRenderGeometry(); r.synchronize()
    t2 = time.perf_counter()
    out[f"call{k}"] = {"register_ms": (t1 - t0) * 1e3, "generate_ms": (t2 - t1) * 1e3, "GBps": nbytes / (t2 - t0) / 1e9,
                       "register_GBps": nbytes / (t1 - t0) / 1e9}
    r.postRender()
r.close()
print(json.dumps(out))
