#!/usr/bin/env python
"""tools/ncu_summary.py — turn an `ncu --set full` report into the per-launch summary CSV committed under profiles/ and the
per-frame DRAM-traffic / warp-instruction JSON that bench.py reads for `roofline.traffic` and `roofline_issue`.

    ncu --set full --clock-control none --import-source on --launch-skip S -c C -o gpurun_out/x/full python bench.py ...   (GPU box)
    python tools/ncu_summary.py gpurun_out/x/full.ncu-rep --workload 20M_sh3_1080p --chunks 2 \
           --csv profiles/r02_ncu_full_summary.csv --json profiles/r02_ncu_traffic.json --frame-kernel cell_project_kernel

--frame-kernel: a kernel that runs once per frame; the launches between two of its occurrences are taken as ONE frame for the
per-frame sums (without it every captured launch is summed).  --periodic: for a static camera, find the period of the launch
sequence and sum its last full period (the capture need not start on a frame boundary).  Runs on the CPU container: needs only `ncu -i`."""
import argparse
import csv
import io
import json
import subprocess
from pathlib import Path

METRICS = {
    "grid": "launch__grid_size", "regs": "launch__registers_per_thread", "time_us": "gpu__time_duration.sum",
    "dram_rd_MB": "dram__bytes_read.sum", "dram_wr_MB": "dram__bytes_write.sum",
    "sm_thr_%": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "mem_thr_%": "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "issue_act_%": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_act_%": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l2_hit_%": "lts__t_sector_hit_rate.pct", "warp_inst": "smsp__inst_executed.sum",
}


def short(name: str) -> str:
    """'void gsb::<unnamed>::os_pass_kernel<9, ...>(...)' -> 'os_pass_kernel<9>'"""
    import re
    m = re.search(r"(\w+_kernel)(?:<\(?\w*\)?\s*(\d+))?", name)
    if not m:
        return name.split("(")[0].strip()
    return m.group(1) + (f"<{m.group(2)}>" if m.group(2) else "")


def find_period(names, shortest=4):
    """Smallest P >= shortest with names[i] == names[i + P] for every i (the launch sequence of a static-camera run repeats
    frame by frame), or None if the capture holds less than two periods' worth of evidence."""
    return next((P for P in range(shortest, len(names)) if all(names[i] == names[i + P] for i in range(len(names) - P))), None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--workload", default="20M_sh3_1080p")
    ap.add_argument("--chunks", type=int, default=2)
    ap.add_argument("--csv", required=True)
    ap.add_argument("--json", required=True)
    ap.add_argument("--frame-kernel", default="")
    ap.add_argument("--note", default="")
    ap.add_argument("--periodic", action="store_true",
                    help="static camera: the per-frame sums are taken over the last full period of the launch sequence")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    col = {k: (hdr.index(m) if m in hdr else None) for k, m in METRICS.items()}

    def val(r, k):
        i = col[k]
        if i is None or r[i] == "":
            return 0.0
        v = float(r[i].replace(",", ""))
        u = units[i]
        if k == "time_us":
            v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        if k in ("dram_rd_MB", "dram_wr_MB"):
            v = {"byte": v / 1e6, "Kbyte": v / 1e3, "Mbyte": v, "Gbyte": v * 1e3}.get(u, v / 1e6)
        return v

    launches = [(short(r[ki]), {k: val(r, k) for k in METRICS}) for r in rows[2:] if len(r) > ki]
    frame = launches
    if a.frame_kernel:
        idx = [i for i, (n, _) in enumerate(launches) if a.frame_kernel in n]
        if len(idx) >= 2:
            frame = launches[idx[0]:idx[1]]
        elif len(idx) == 1:                       # one frame start in the capture: the frame ends with its last blend launch
            end, blends = idx[0], 0
            for i in range(idx[0], len(launches)):
                if "blend_kernel" in launches[i][0]:
                    blends += 1; end = i + 1
                    if blends == a.chunks:
                        break
            frame = launches[idx[0]:end]
    if a.periodic:
        # static camera: every frame launches the same sequence, so ANY window of one period is one frame's set of launches
        period = find_period([n for n, _ in launches])
        if period:
            frame = launches[len(launches) - period:]
    with open(a.csv, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on: {len(launches)} captured launches of bench.py, {a.workload}, "
                f"{a.chunks} depth chunks; per launch.  {a.note}\n")
        f.write("kernel," + ",".join(METRICS) + "\n")
        for n, m in launches:
            f.write(n + "," + ",".join(f"{m[k]:.1f}" if k not in ("grid", "regs", "warp_inst") else str(int(m[k])) for k in METRICS) + "\n")
    per = {}
    for n, m in frame:
        base = n.split("<")[0]
        d = per.setdefault(base, {"dram_bytes_per_frame": 0.0, "read": 0.0, "write": 0.0, "warp_instructions_per_frame": 0.0,
                                  "launches_per_frame": 0, "time_us_per_frame": 0.0})
        d["read"] += m["dram_rd_MB"] * 1e6; d["write"] += m["dram_wr_MB"] * 1e6
        d["dram_bytes_per_frame"] = d["read"] + d["write"]
        d["warp_instructions_per_frame"] += m["warp_inst"]; d["launches_per_frame"] += 1; d["time_us_per_frame"] += m["time_us"]
    out = {}
    p = Path(a.json)
    if p.exists():
        try:
            out = json.loads(p.read_text())
        except Exception:
            out = {}
    out["_comment"] = ("per-frame sums over the launches of ONE frame of an ncu --set full capture (dram__bytes_read.sum + dram__bytes_write.sum, "
                       "smsp__inst_executed.sum); bench.py reads blend_kernel for roofline.traffic and roofline_issue")
    out[a.workload] = dict(per, depth_chunks=a.chunks, source=f"{a.csv} (from {Path(a.report).name})")
    p.write_text(json.dumps(out, indent=1))
    for n, d in per.items():
        print(f"{n:28s} {d['launches_per_frame']:2d} launches {d['time_us_per_frame']:8.1f} us  {d['dram_bytes_per_frame'] / 1e6:8.1f} MB  "
              f"{d['warp_instructions_per_frame'] / 1e6:8.1f} M warp-instr")


if __name__ == "__main__":
    main()
