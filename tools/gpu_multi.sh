#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the multi-GPU tests, then bench.py at every N <= the box's GPU count the way the
# driver launches it (torchrun, one rank per GPU), headline workload and DXT5 (BASELINE config 3).
# Usage: bash tools/gpu_multi.sh <tag> "<N list>" [skip-tests]
TAG=${1:-multi}; NS=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1; nvidia-smi topo -m > $OUT/topo.txt 2>&1
if [ "$3" != skip-tests ]; then
  timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu > $OUT/pytest_multigpu.log 2>&1; echo "pytest multigpu exit $?"; tail -4 $OUT/pytest_multigpu.log
fi
port=29500
for n in $NS; do
  for wl in dxt1_rgba8 dxt5_rgba8; do
    port=$((port + 1))
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n --steps 50 --warmup 5 --workload $wl > $OUT/bench_${wl}_n$n.json 2> $OUT/bench_${wl}_n$n.err; echo "bench $wl n=$n exit $?"
    python - $OUT/bench_${wl}_n$n.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("  N=%d %s value %.0f Mpix/s (%.1f us)  parity %s  e2e %.0f Mpix/s (%.2f ms, ok %s)" % (d["n_gpus"], d["config"]["workload"], d["value"], d["ms_per_step"] * 1e3, d["parity"]["equal"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["output_equals_reference"]))
    dl = d.get("delivery")
    if dl:
        print("     delivery %.0f GB/s into root (%.2f of 900), %s, splits %s" % (dl["achieved"], dl["frac"], dl["partition"], d["config"]["job"][-60:]))
    for k in ("encode_only", "even_split_delivered", "nccl_gather", "weak_scaling_kernel_only"):
        if k in d:
            print("     %-26s %.1f us  %.0f Mpix/s %s" % (k, d[k]["ms_per_step"] * 1e3, d[k]["mpix_s"], d[k].get("equals_reference", "")))
except Exception as e:
    print("  parse failed", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
  done
done
