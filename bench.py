#!/usr/bin/env python
"""bench.py — Msplats/s and fps of the splat -> framebuffer hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A step = one full frame of the hot path (cull + key + bounded tile rectangle for every splat -> per depth chunk:
live-splat selection, depth sort, exact projection + records + SH, tile binning, blend) over
the synthetic cloud of SURVEY.md §8d.  Default workload: the north-star target, 20 M splats, SH degree 3,
1920x1080, one B200 (fits one GPU: 2.6 GB of attributes).  Prints ONE JSON line (rank 0).

  value        whole-job Msplats/s (= submitted splats x fps / 1e6), attributes resident in HBM, frame left on device
  e2e          same metric through the C ABI with HOST buffers: gsb_frame (368 B) in, RGBA32F frame copied to pinned
               host memory inside the timed call.  Geometry is NOT re-uploaded per frame — neither does the reference
               (its textures persist until the active set changes, GSplatRenderer.C:324-327); the cold cost
               (gsb_register_update H2D + pack) is reported separately as e2e_cold_ms.
  roofline     the blend kernel: algorithmic bytes D_c*(4+48)+W*H*16 / its CUDA-event time on the
               library stream, vs the measured HBM copy peak of MEASURED_PEAKS.json
  cpu_baseline oracle/_ref ("reference": the reference's own GLSL text compiled for the host + its CPU argsort; the plugin
               itself needs Houdini + OpenGL and cannot run here) on the host cores, one whole frame; a reported baseline,
               not the target.  Without oracle/_ref: the oracle port.
  parity       the production frame of this very run against the oracle's and the reference GLSL's frame of the workload

N > 1 (torchrun, one process per GPU): the frame is partitioned by interleaved tile rows; every rank
holds all splats, renders its rows, and the rows are combined on rank 0 with one NCCL reduction per frame
(exact: the other ranks contribute zeros).  Total work is fixed => "scaling": "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

RECORD_BYTES = 48
REF_BUDGET_S = 270.0   # wall budget of the whole --impl reference run (a few minutes)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="20M_sh3_1080p")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--port-baseline", action="store_true", help="cpu_baseline from the oracle port only (skip oracle/_ref)")
    ap.add_argument("--depth-chunks", type=int, default=0, help="0 = library default (auto)")
    ap.add_argument("--combine", default="p2p", choices=["p2p", "nccl", "host"],
                    help="N>1: p2p = the blend kernel stores finished tiles straight into rank 0's frame over NVLink peer "
                         "memory (CUDA IPC) + a 4-byte NCCL all-reduce as the per-frame completion fence; "
                         "nccl = every rank renders into its own frame and one NCCL reduction combines them")
    ap.add_argument("--verify", action="store_true", help="(accepted for compatibility: N>1 runs are always verified)")
    ap.add_argument("--sustain-s", type=float, default=2.0,
                    help="seconds of back-to-back frames for the sustained pass reported beside the burst (0 = skip)")
    ap.add_argument("--host-direct", type=int, default=1, choices=[0, 1],
                    help="e2e leg: 1 = finished tiles are stored straight into the pinned host frame by the blend kernel "
                         "(GSB_OPT_HOST_DIRECT, library default), 0 = staged cudaMemcpyAsync after the frame")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def wait_first_sample(self, timeout=5.0):
        """nvidia-smi takes a few hundred ms to start; do not open the timed region before it is sampling."""
        t0 = time.time()
        while self.p is not None and time.time() - t0 < timeout:
            if os.path.getsize(self.f.name) > 0:
                return
            time.sleep(0.05)

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(",") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_workload(name):
    from houdini_gsplat_renderer_b200 import scene as S
    if name not in S.WORKLOADS or S.WORKLOADS[name]["width"] == 0:
        raise SystemExit(f"unknown / non-rendering workload {name}; choose from "
                         f"{[k for k, v in S.WORKLOADS.items() if v['width']]}")
    w = S.WORKLOADS[name]
    t0 = time.time()
    cloud = S.make_cloud(w["n"], w["seed"], sh=w["sh"])
    return S, w, cloud, time.time() - t0


def frame_for(S, w, step):
    theta = float(step % 360) if w["orbit"] else 0.0
    return S.orbit_frame(w["width"], w["height"], theta)


def probe_software_gl() -> dict:
    """is there a software OpenGL on this box that could run the reference's GLSL unmodified
    (Mesa llvmpipe through OSMesa / EGL)?  dlopen only; the result is reported in cpu_baseline (expected: nothing)."""
    import ctypes
    found = {}
    for lib in ("libOSMesa.so.8", "libOSMesa.so", "libEGL.so.1", "libGL.so.1"):
        try:
            ctypes.CDLL(lib)
            found[lib] = True
        except OSError:
            found[lib] = False
    return found


# ------------------------------------------------------------------------------------ shared by both arms
def config_of(args, w, n_splats):
    """The workload description: identical keys and values in both arms (the driver compares them)."""
    return {"workload": args.workload, "splats": n_splats, "sh_degree": 3 if w["sh"] else 0,
            "width": w["width"], "height": w["height"], "camera": "orbit 1 deg/frame" if w["orbit"] else "static",
            "tile": 16, "eps_t": 1e-5, "splat_cap": "lifted (reference caps at 8388607)",
            "l2_policy": "inputs larger than L2 (%.2f GB packed attributes vs 126 MB L2)" % (n_splats * 160 / 1e9),
            "full_pipeline_every_frame": True}


# ------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_frame(O, S, w, cloud, n_sample, step=0):
    """One frame of the CPU port: the reference's own CPU stage (fp32 distances of every splat + parallel comparison
    argsort, R.C:176-216) supplies the depth order, then software vertex + fragment + blend stages over screen tiles."""
    sub = cloud if n_sample >= cloud.n else cloud.subset(slice(0, n_sample))
    fr = frame_for(S, w, step)
    F = O.make_frame(fr, O.camera_from_view(fr.view), sub.barycentre(), 3 if w["sh"] else 0, eps_t=1e-5)
    t0 = time.time()
    rgba, st = O.render(F, sub, time_reference_sort=True)
    return time.time() - t0, st, sub.n, rgba


def cpu_sample_note(st, ns, n, workload, w, cores):
    return (f"first {ns} of {n} splats of workload {workload}, same camera and {w['width']}x{w['height']} frame, {cores} host "
            f"threads (OpenMP, set explicitly); per frame: reference-style CPU argsort of every splat {st['ms_sort_reference']:.0f} ms "
            f"(R.C:176-216; its order is the one used) + project {st['ms_project']:.0f} + bin {st['ms_bin']:.0f} + blend "
            f"{st['ms_blend']:.0f} ms; oracle restatement with tiles and early-out, not the GLSL under llvmpipe (no OpenGL in the image)")


def ref_frame(RF, O, S, w, cloud, n_sample, step=0, unsafe_tol=None):
    """One frame of the REFERENCE path on the host: argsortByDistance (R.C:176-216) + the reference's own GLSL text compiled
    for the host (oracle/_ref: vertex shader 6x per splat, ideal rasteriser, fragment shader, "under" blend; no tiles, no
    termination — what the GL pipeline does).  Returns seconds, frame, stats[, unsafe mask]."""
    sub = cloud if n_sample >= cloud.n else cloud.subset(slice(0, n_sample))
    fr = frame_for(S, w, step)
    cam = O.camera_from_view(fr.view)
    t0 = time.time()
    b = RF.Bound(sub, fr, cam, sub.barycentre(), 3 if w["sh"] else 0)        # binds + argsortByDistance
    t_sort = time.time() - t0
    out = b.draw(unsafe_tol=unsafe_tol)
    t = time.time() - t0
    st = dict(out[1], ms_sort_reference=t_sort * 1e3, ms_draw=(t - t_sort) * 1e3)
    return t, out[0], st, (out[2] if unsafe_tol is not None else None), sub.n


def ref_sample_note(st, ns, n, workload, w, cores):
    return (f"first {ns} of {n} splats of workload {workload}, same camera and {w['width']}x{w['height']} frame, {cores} host "
            f"threads (OpenMP, set explicitly): the reference's CPU argsort {st['ms_sort_reference']:.0f} ms (R.C:176-216) + its own "
            f"GLSL text compiled for the host (oracle/_ref: g++ instead of llvmpipe's JIT; vertex shader 6x per splat, ideal "
            f"rasteriser, fragment shader, ROP blend; no tiles, no early termination) {st['ms_draw']:.0f} ms, "
            f"{st['fragments_shaded']} fragments shaded; no OpenGL / llvmpipe exists in the image")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    from oracle import ref as RF
    S, w, cloud, gen_s = load_workload(args.workload)
    use_ref = RF.available()
    cores = RF.set_num_threads() if use_ref else O.set_num_threads()      # every core, whatever OMP_NUM_THREADS torchrun exported
    O.set_num_threads()

    def one(n_sample, step):
        if use_ref:
            t, _, st, _, ns = ref_frame(RF, O, S, w, cloud, n_sample, step)
            return t, st
        t, st, ns, _ = cpu_frame(O, S, w, cloud, n_sample, step)
        return t, st
    # the whole cloud every step; only if warmup + steps of it cannot fit the budget is the sample cut (and said so)
    t_full, _ = one(cloud.n, 0)
    per_step = REF_BUDGET_S / max(1, args.steps + args.warmup + 1)
    n_sample = cloud.n if t_full <= per_step else int(max(100_000, cloud.n * per_step / t_full * 0.8))
    for i in range(args.warmup):
        one(n_sample, i)
    times, st = [], None
    for i in range(args.steps):
        t, st = one(n_sample, args.warmup + i)
        times.append(t)
    ms = 1e3 * sum(times) / len(times)
    val = n_sample / (ms * 1e-3) / 1e6
    note = ref_sample_note(st, n_sample, cloud.n, args.workload, w, cores) if use_ref else \
        cpu_sample_note(st, n_sample, cloud.n, args.workload, w, cores)
    line = {"impl": "reference", "metric": "Msplats/sec at %dx%d" % (w["width"], w["height"]), "value": val,
            "unit": "Msplats/s", "fps": 1e3 / ms, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_of(args, w, cloud.n), "sample_splats": n_sample,
            "cpu_baseline": {"value": val, "unit": "Msplats/s", "cores": cores, "kind": "reference" if use_ref else "port",
                             "sample": note, "software_gl_probe": probe_software_gl()},
            "e2e": {"value": val, "unit": "Msplats/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ our arm (CUDA)
def run_ours(args):
    import torch
    import torch.distributed as dist
    from houdini_gsplat_renderer_b200 import renderer as R
    from houdini_gsplat_renderer_b200 import multigpu as M

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")               # both levels print "NCCL version ..." on stdout: keep it to the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    S, w, cloud, gen_s = load_workload(args.workload)
    W, H, N = w["width"], w["height"], cloud.n
    sh_order = 3 if w["sh"] else 0
    r = R.GSplatRenderer(local)
    r.set_option(R.OPT_SPLAT_CAP, 0)          # the reference's 2^23-1 cap lifted for the 20 M configs
    r.set_option(R.OPT_DEPTH_CHUNKS, args.depth_chunks)
    r.set_option(R.OPT_HOST_DIRECT, args.host_direct if world == 1 else 1)
    stream = torch.cuda.current_stream()
    r.set_stream(stream.cuda_stream)
    r.setSphericalHarmonicsOrder(sh_order)

    # cold path: H2D of the prim's arrays + pack (geometry-change cost, GR_GSplat.C:302-372 + R.C:448-530).  The prim's
    # barycentre is an INPUT of registerUpdate (the prim computes it when it cooks, GEO_GSplat.C:338-351), so it is taken
    # before the clock starts.  Timed twice: the first registration of the process (device allocations of 7 GB, kernel
    # module load) and a re-registration of the same prim under a new version (what a recook costs; the old entry is erased
    # by registerUpdate like R.C:246-265) — the steady-state number.
    origin = cloud.barycentre()
    torch.cuda.synchronize()
    t0 = time.time()
    rid = r.registerUpdate(0xB200, (1, 0, 0, 0), 0, cloud, origin)
    r.includeInRenderPass(rid); r.generateRenderGeometry(); r.synchronize()
    cold_first_ms = (time.time() - t0) * 1e3
    r.postRender()
    t0 = time.time()
    rid = r.registerUpdate(0xB200, (2, 0, 0, 0), 0, cloud, origin)
    r.includeInRenderPass(rid); r.generateRenderGeometry(); r.synchronize()
    cold_upload_ms = (time.time() - t0) * 1e3
    h2d_cold = N * (132 if w["sh"] else 36)

    # the row partition, the shared frames, the fence and the hand-off live in the package (multigpu.py)
    mg = M.RowPartitionedRenderer(r, W, H, rank, world, combine=args.combine, stream=stream)
    row_group = mg.row_group
    frame_bytes = W * H * 16

    def step(i, to_host):
        r.includeInRenderPass(rid)
        r.generateRenderGeometry()
        mg.render(frame_for(S, w, i), to_host=to_host)
        r.postRender()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    STAGES = ("ms_project", "ms_sort", "ms_records", "ms_bin", "ms_blend", "ms_copy")
    COUNTS = ("n_visible", "n_instances", "n_consumed", "n_live", "launches", "depth_chunks")

    def timed(K, to_host, collect):
        acc = {k: 0.0 for k in STAGES}
        cnt = {k: 0 for k in COUNTS}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(K):
            step(args.warmup + i, to_host)
            if collect:            # reads pinned counters + event times of the frame that just ran
                st = r.stats()
                for k in acc: acc[k] += st[k]
                for k in cnt: cnt[k] += st[k]
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        return ms, acc, cnt

    # clocks are sampled (20 ms period) from before the warm-up to the end of the timed regions: the GPU is
    # under the same load throughout, so short timed regions still get a meaningful median
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler: sampler.wait_first_sample()
    for i in range(max(3, args.warmup)):
        step(i, False)
    ms_dev, _, _ = timed(args.steps, False, False)          # headline: K frames back to back, nothing read back
    for i in range(2):
        step(i, True)
    ms_e2e, _, _ = timed(args.steps, True, False)
    # per-stage CUDA events and counters: a third pass over the same K frames (gsb_get_stats synchronises after every
    # frame, so this pass is not the one that is timed for `value`)
    r.set_option(R.OPT_STAGE_TIMING, 1)
    step(0, False)
    _, acc, cnt = timed(args.steps, False, True)
    r.set_option(R.OPT_STAGE_TIMING, 0)
    clocks = sampler.stop() if sampler else None
    # sustained pass: >= args.sustain_s seconds of back-to-back frames (the burst above lasts tens of milliseconds, and for an
    # issue-bound pipeline the SM clock is the result), clocks sampled over exactly this region
    sustained = None
    if args.sustain_s > 0:
        Ks = max(args.steps, int(args.sustain_s * 1e3 / max(ms_dev / args.steps, 1e-3)) + 1)
        s2 = ClockSampler(local) if rank == 0 else None
        if s2: s2.wait_first_sample()
        ms_sus, _, _ = timed(Ks, False, False)
        c2 = s2.stop() if s2 else None
        sustained = {"seconds": ms_sus * 1e-3, "frames": Ks, "ms_per_step": ms_sus / Ks,
                     "value": N / (ms_sus / Ks * 1e-3) / 1e6, "unit": "Msplats/s", "fps": 1e3 * Ks / ms_sus, "clocks": c2}

    K = args.steps
    ms_step = ms_dev / K
    value = N / (ms_step * 1e-3) / 1e6
    e2e_ms = ms_e2e / K
    e2e_val = N / (e2e_ms * 1e-3) / 1e6

    peak, peak_src = peaks()
    sh_bytes = {0: 0, 1: 18, 2: 48, 3: 90}[sh_order]
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    tile_passes = max(1, -(-max(1, (tiles - 1).bit_length()) // 8))
    key_passes = 3                                        # 25 significant key bits at these camera distances: 9+8+8

    def stage_table(acc, cnt):
        """Algorithmic (compulsory) bytes of ONE rank per frame, SURVEY.md §8d per-unit figures, and that rank's stage
        times: every fraction is one GPU's bytes over one GPU's time against one GPU's peak."""
        V, D, Dc, L = cnt["n_visible"] / K, cnt["n_instances"] / K, cnt["n_consumed"] / K, cnt["n_live"] / K
        chunks = max(1.0, cnt["depth_chunks"] / K)
        owned_px = W * H / world
        sb = {
            # K1 (bounded, the library default): position + discard radius (16 B) + covariance eigenvalue bound (4 B) read,
            # (key, packed tile rectangle) written, for every submitted splat.  (SURVEY's unit for an exact cull phase is 30 B.)
            "project": N * 20 + N * 8,
            # per depth chunk: live selection streams key + tile rectangle of every splat (8 B); the L selected (key, index)
            # pairs are staged and gathered (8 + 12 B), then LSD-sorted (histogram read + passes x 24 B)
            "sort": chunks * N * (8 + 0.25) + L * (8 + 12) + L * (4 + key_passes * 24),
            # K2: index + 30 B geometry + colour/SH read, record written, per live splat
            "records": L * (4 + 30 + 6 + sh_bytes) + L * RECORD_BYTES,
            # K4: counts, emit (8 B per instance written), stable tile partition (histogram read + passes x 16 B), tile ranges
            "bin": L * (8 + 12 + 12) + D * 8 + D * 4 + tile_passes * D * 16 + D * 4 + tiles * 8,
            "blend": Dc * (4 + RECORD_BYTES) + owned_px * 16,
        }
        sm = {"project": acc["ms_project"] / K, "sort": acc["ms_sort"] / K, "records": acc["ms_records"] / K,
              "bin": acc["ms_bin"] / K, "blend": acc["ms_blend"] / K}
        out = {k: {"ms": sm[k], "algorithmic_bytes": sb[k],
                   "achieved_GBps": sb[k] / (sm[k] * 1e-3) / 1e9 if sm[k] > 0 else None,
                   "frac_of_peak": sb[k] / (sm[k] * 1e-3) / 1e9 / peak if sm[k] > 0 else None} for k in sm}
        return out, {"N": N, "V": V, "L": L, "D": D, "D_c": Dc, "depth_chunks": chunks, "launches": cnt["launches"] / K}

    my_stages, my_counters = stage_table(acc, cnt)
    per_rank = None
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, {"rank": rank, "stages": my_stages, "counters": my_counters})
        per_rank = gathered

    # multi-GPU frames are checked on every run: the combined frame against a single-rank render of the same frame
    verify = None
    if world > 1:
        step(0, True)
        barrier()
        if rank == 0:
            combined = np.array(mg.host_frame(), copy=True)
            solo = np.zeros_like(combined)
            r.includeInRenderPass(rid); r.generateRenderGeometry()
            r.render(frame_for(S, w, 0), host_rgba=solo); r.postRender()
            verify = "bit-identical to the single-rank frame" if np.array_equal(solo, combined) else \
                     "MISMATCH max|d|=%g" % float(np.abs(solo - combined).max())
        barrier()
    # the production frame of the benchmarked workload (frame 0, host target), kept for the parity check below
    gpu_frame0, gpu_stats0 = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        step(0, True)
        gpu_frame0 = np.array(mg.host_frame(), copy=True)
        gpu_stats0 = r.stats()
    if rank != 0:
        mg.close(); r.close()
        if world > 1: dist.destroy_process_group()
        return

    stages, counters = my_stages, my_counters
    blend_ach = stages["blend"]["achieved_GBps"] or 0.0
    # measured DRAM traffic of the blend launches of one frame, from the committed ncu --set full capture of this workload
    traffic, traffic_src = None, None
    try:
        tj = json.loads((ROOT / "profiles" / "r02_ncu_traffic.json").read_text()).get(args.workload)
        if tj and world == 1:
            traffic = tj["blend_kernel"]["dram_bytes_per_frame"]
            traffic_src = tj.get("source", "profiles/r02_ncu_full_summary.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, blend launches of one frame)")
    except Exception:
        pass
    # second roof of the blend: it is issue bound, not HBM bound: warp instructions per frame from the same
    # ncu capture (smsp__inst_executed.sum) against 148 SMs x 4 schedulers x 1 warp instruction per clock at the sampled SM clock
    issue = None
    try:
        ij = json.loads((ROOT / "profiles" / "r02_ncu_traffic.json").read_text()).get(args.workload)
        if ij and world == 1 and clocks and clocks.get("sm_mhz") and stages["blend"]["ms"] > 0:
            winst = float(ij["blend_kernel"]["warp_instructions_per_frame"])
            peak_issue = 148 * 4 * clocks["sm_mhz"] * 1e6
            ach = winst / (stages["blend"]["ms"] * 1e-3)
            issue = {"bound": "issue", "achieved": ach / 1e9, "peak": peak_issue / 1e9, "unit": "G warp-instr/s",
                     "frac": ach / peak_issue, "warp_instructions_per_frame": winst,
                     "note": "warp instructions from the committed ncu capture; time = CUDA events of this run; peak = 148 SMs x 4 "
                             "issue slots x the median SM clock sampled during this run"}
    except Exception:
        pass

    line = {
        "metric": "Msplats/sec at %dx%d" % (W, H), "value": value, "unit": "Msplats/s", "fps": 1e3 / ms_step,
        "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(args, w, N),
        "parallelism": f"tile-row bands of {row_group} x16 px interleaved over {world} GPU(s)",
        "depth_chunks": counters["depth_chunks"],
        "e2e": {"value": e2e_val, "unit": "Msplats/s", "fps": 1e3 / e2e_ms, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": 368 * world, "d2h_bytes_per_step": frame_bytes,
                "host_direct": True if world > 1 else bool(args.host_direct),
                "host_frame_on_gpu_local_cpus": (len(mg.host_cpus) if getattr(mg, "host_cpus", None) else None),
                "note": "gsb_render with a pinned host target: gsb_frame in, RGBA32F frame stored into host memory by the blend "
                        "kernels (every rank over its own PCIe link; N>1: one shared frame + a 4-byte all-reduce fence), the call "
                        "returns when the frame is there; geometry resident (the reference also re-uploads only on active-set change)"},
        "e2e_cold_ms": cold_upload_ms, "e2e_cold_h2d_bytes": h2d_cold,
        "e2e_cold_GBps": h2d_cold / (cold_upload_ms * 1e-3) / 1e9,
        "e2e_cold_first_ms": cold_first_ms, "e2e_cold_first_GBps": h2d_cold / (cold_first_ms * 1e-3) / 1e9,
        "e2e_cold_note": "registerUpdate (pageable host arrays -> HBM) + generateRenderGeometry (pack, Morton cells): first call of "
                         "the process, and a re-registration under a new version (steady state)",
        "gpu_launches": cnt["launches"],
        "roofline": {"kernel": "blend_kernel", "bound": "hbm", "achieved": blend_ach, "peak": peak, "unit": "GB/s",
                     "frac": blend_ach / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": stages["blend"]["algorithmic_bytes"], "ms_per_launch": stages["blend"]["ms"],
                     "launch": "the blend launches of one frame (one per depth chunk), bytes and time summed"
                               + ("; rank 0's share" if world > 1 else ""),
                     "formula": "D_c*(4+48) + W*H*16"},
        "roofline_issue": issue,
        "stages": stages,
        "stages_per_rank": per_rank,
        "counters_per_frame": counters,
        "counters_note": "bounded K1 (library default): V = splats that pass the cheap culls with a non-empty rectangle bound and "
                         "L = splats the depth chunks selected are upper bounds of the exact counts; D and D_c are exact"
                         + ("; rank 0's share" if world > 1 else ""),
        "clocks": clocks,
        "sustained": sustained,
        "combine": (args.combine if world > 1 else None), "verify": verify,
        "scene_gen_s": gen_s,
    }

    if not args.no_cpu_baseline and world == 1:
        from oracle import oracle as O
        from oracle import ref as RF
        cores = O.set_num_threads()
        # (1) the oracle's frame of the benchmarked workload (same early-out as the GPU): the tight check
        t, st, ns, orc_rgba = cpu_frame(O, S, w, cloud, N)
        d = np.abs(gpu_frame0.astype(np.float64) - orc_rgba.astype(np.float64))
        parity = {"frame": "frame 0 of the benchmarked workload, production options (auto depth chunks, bounded K1, host-direct delivery)",
                  "vs_oracle": {"max_abs": float(d.max()), "pixels_over_1e-3": int((d.max(axis=2) > 1e-3).sum()),
                                "pixels_over_2e-5": int((d.max(axis=2) > 2e-5).sum()),
                                "D_c_gpu": int(gpu_stats0["n_consumed"]), "D_c_oracle": int(st["n_consumed"]),
                                "D_c_equal": int(gpu_stats0["n_consumed"]) == int(st["n_consumed"])},
                  "tolerance": 1e-3}
        ok = bool(d.max() <= 1e-3)
        baseline = {"value": ns / t / 1e6, "unit": "Msplats/s", "cores": cores, "kind": "port",
                    "sample": "one whole frame: " + cpu_sample_note(st, ns, N, args.workload, w, cores),
                    "software_gl_probe": probe_software_gl()}
        # (2) the reference's own GLSL text compiled for the host (oracle/_ref), whole frame, reference semantics (every fragment
        # blended): the CPU baseline proper, and the frame the north-star tolerance (1e-3 abs) is stated against
        if RF.available() and not args.port_baseline:
            rcores = RF.set_num_threads()
            t2, ref_rgba, st2, unsafe, ns2 = ref_frame(RF, O, S, w, cloud, N, 0, unsafe_tol=2e-5)
            d2 = np.abs(gpu_frame0.astype(np.float64) - ref_rgba.astype(np.float64)).max(axis=2)
            over = d2 > 1e-3
            parity["vs_reference_glsl"] = {
                "max_abs_safe_pixels": float(d2[~unsafe].max()), "max_abs_all_pixels": float(d2.max()),
                "pixels_over_1e-3": int(over.sum()), "pixels_over_1e-3_not_flagged": int((over & ~unsafe).sum()),
                "flagged_pixels": int(unsafe.sum()), "pixels": int(unsafe.size),
                "note": "reference semantics has no early termination (GPU stops a pixel at T < 1e-5: error <= 1e-5 * max rgb); "
                        "flagged = a support edge or the 1/255 discard ring of a still-visible splat passes within tol_q of the "
                        "pixel centre, tol_q = 2e-5 + 2 ulp_fp32(screen extent) x |grad q| in the splat's own q units (a 2 px "
                        "splat at x = 1900 has a q uncertainty of 1e-4 under ANY fp32 evaluation), where one ulp of evaluation "
                        "order decides coverage (oracle/ref_harness.cpp ref_draw_ex)"}
            ok = ok and bool(d2[~unsafe].max() <= 1e-3) and int((over & ~unsafe).sum()) == 0
            baseline = {"value": ns2 / t2 / 1e6, "unit": "Msplats/s", "cores": rcores, "kind": "reference",
                        "sample": "one whole frame (incl. the edge-pixel diagnostic): " + ref_sample_note(st2, ns2, N, args.workload, w, rcores),
                        "port_value": ns / t / 1e6, "software_gl_probe": probe_software_gl()}
        parity["ok"] = ok
        line["cpu_baseline"] = baseline
        line["parity"] = parity
    print(json.dumps(line), flush=True)
    mg.close()
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
