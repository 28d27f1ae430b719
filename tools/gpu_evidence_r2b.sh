#!/bin/bash
# Round 2 (second half) evidence visit, one GPU: ncu --set full of every workload's kernels, ETC1 pipe counts, then the
# end-of-round run (tools/gpu_final.sh: GPU suite, bench lines, reference arm, launch list, sanitizers).
# Usage (under gpurun):  bash tools/gpu_evidence_r2b.sh <tag>
TAG=${1:-fin}; OUT=gpurun_out/$TAG; mkdir -p $OUT
bash tools/gpu_profile_all.sh $TAG
timeout 600 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed.sum --clock-control none \
    -k regex:"encode4x4|pvrtc" -c 12 --csv --log-file $OUT/pipe_counts_etc1.csv python bench.py --workload etc1_rgb8 --steps 2 --warmup 3 --no-cpu-baseline --no-others > $OUT/ncu_etc1.log 2>&1; echo "ncu etc1 exit $?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke.log
bash tools/gpu_final.sh $TAG
bash tools/gpu_sanitize_r2.sh $TAG
