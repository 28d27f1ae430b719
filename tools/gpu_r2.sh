#!/bin/bash
# Round-2 GPU-box visit: GPU test suite, the default bench line (headline + other_workloads + parity), the reference
# arm, and an ncu metrics pass that counts integer-pipe instructions of the ETC1 kernel (profiles/pipe_counts.json).
# Usage (under gpurun, from the repo root):  bash tools/gpu_r2.sh <tag> [quick]
TAG=${1:-r2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("value %.0f Mpix/s  step %.1f us  frac %.3f  parity %s  e2e %.0f (%s)" % (d["value"], d["ms_per_step"] * 1e3, d["roofline"]["frac"], d["parity"], d["e2e"]["value"], d["e2e"].get("output_equals_reference")))
    for k, v in d.get("other_workloads", {}).items():
        if k == "unaligned_device_source":
            print("  unaligned source (generic kernel): %.1f us, parity %s" % (v["kernel_ms"] * 1e3, v["parity"]["equal"]))
        elif "kernel_ms" in v:
            print("  %-14s %.1f us  frac %.3f  parity %s" % (k, v["kernel_ms"] * 1e3, v["roofline"]["frac"], v["parity"]["equal"]), v.get("roofline_int_alu"))
        else:
            for c, r in v.items():
                print("  %s/%s %.1f us frac %.3f parity %s" % (k, c, r["kernel_ms"] * 1e3, r["roofline_frac"], r["parity"]["equal"]))
except Exception as e:
    print("bench parse failed", e)
PY
[ "$2" = quick ] && exit 0
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed.sum --clock-control none \
    -k regex:"encode4x4|pvrtc" -c 12 --csv --log-file $OUT/pipe_counts_etc1.csv python bench.py --workload etc1_rgb8 --steps 2 --warmup 3 --no-cpu-baseline --no-others > $OUT/ncu_etc1.log 2>&1; echo "ncu etc1 exit $?"
