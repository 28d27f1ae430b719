#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines, ncu launch list and one full capture of the DXT1 kernel.
# Usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/smoke.log
for wl in dxt1_rgba8 dxt1_rgb8 dxt5_rgba8 etc1_rgb8 pvrtc2_rgba8; do
  echo "== bench $wl"; timeout 600 python bench.py --workload $wl --steps 30 --warmup 5 > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "exit $?"; cat $OUT/bench_$wl.json; tail -3 $OUT/bench_$wl.err
done
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>&1; cat $OUT/bench_reference.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_dxt1_rgba8.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "exit $?"
echo "== ncu full capture (DXT1 TMA kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:encode4x4_tma -s 3 -c 2 -o $OUT/prof_dxt1_rgba8 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "exit $?"
ls -la $OUT
