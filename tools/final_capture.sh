#!/bin/bash
# Final capture of a round: GPU tests, the bench lines of the five BASELINE configs, the launch list and one ncu --set full
# capture of the default workload.  Run on the GPU box from the repo root:   bash tools/final_capture.sh [outdir] [--with-reference]
set -x
O=${1:-gpurun_out/final}; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
python bench.py > $O/bench_1gpu_20M_sh3_1080p.json 2> $O/bench_default.err
if [ "$2" == "--with-reference" ]; then
  python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_20M_sh3_1080p.json 2> $O/bench_reference.err
fi
for wl in 1M_sh0_1080p 5M_sh3_1080p_orbit; do python bench.py --workload $wl --steps 50 > $O/bench_1gpu_$wl.json 2> $O/bench_$wl.err; done
for wl in 20M_sh3_4k 20M_sh3_8k; do python bench.py --workload $wl --steps 30 --no-cpu-baseline > $O/bench_1gpu_$wl.json 2> $O/bench_$wl.err; done
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 600 -c 120 --csv --log-file $O/launches_20M_sh3_1080p.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --sustain-s 0 > $O/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 633 -c 40 -o $O/full python bench.py --steps 6 --warmup 3 --no-cpu-baseline --sustain-s 0 > $O/ncu_full.log 2>&1
# one rank of an 8-way partition on this GPU: the per-kernel floor a rank of the 8-GPU run sees
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 200 -c 80 --csv --log-file $O/launches_shard_8way.csv python tools/shard_probe.py 8 20M_sh3_1080p > $O/shard_8way.log 2>&1
ls -la $O
