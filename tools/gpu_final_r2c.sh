#!/bin/bash
# End-of-round evidence after the ETC1 rewrite (one GPU): full GPU suite, smoke, the default bench line, the reference
# arm, ETC1 as its own headline, launch list of the ETC1 bench, sanitizers over the ETC1 tests.
# Usage (under gpurun):  bash tools/gpu_final_r2c.sh <tag>
TAG=${1:-final_r2c}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 400 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke.log
timeout 400 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench exit $?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_arm.json 2> $OUT/bench_reference_arm.err; echo "reference arm exit $?"
timeout 200 python bench.py --workload etc1_rgb8 --no-others > $OUT/bench_etc1_rgb8.json 2> $OUT/bench_etc1_rgb8.err; echo "bench etc1 exit $?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $OUT/launches_etc1_rgb8.csv \
    python bench.py --workload etc1_rgb8 --steps 5 --warmup 3 --no-others --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu launch list exit $?"
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -x -q -m gpu -k "etc1_vs_oracle or golden_device or ragged" > $OUT/memcheck_etc1.log 2>&1; echo "memcheck exit $?"; tail -4 $OUT/memcheck_etc1.log
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -x -q -m gpu -k "etc1_vs_oracle" > $OUT/racecheck_etc1.log 2>&1; echo "racecheck exit $?"; tail -4 $OUT/racecheck_etc1.log
python - $OUT <<'PY'
import json, sys, os
out = sys.argv[1]
for f in sorted(os.listdir(out)):
    if f.startswith("bench_") and f.endswith(".json"):
        try:
            d = json.load(open(os.path.join(out, f)))
            print("%-28s value %.0f %s  step %.2f us  frac %s  parity %s  e2e %.0f" % (f, d["value"], d["unit"], d["ms_per_step"] * 1e3,
                  ("%.3f" % d["roofline"]["frac"]) if "roofline" in d else "-", (d.get("parity") or {}).get("equal"), d["e2e"]["value"]))
            for k, v in (d.get("other_workloads") or {}).items():
                if isinstance(v, dict) and "kernel_ms" in v:
                    print("    %-24s %.2f us  frac %s parity %s" % (k, v["kernel_ms"] * 1e3, ("%.3f" % v["roofline"]["frac"]) if "roofline" in v else "-", (v.get("parity") or {}).get("equal")))
        except Exception as e:
            print(f, "parse failed", e)
PY
