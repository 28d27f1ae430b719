#!/bin/bash
# compute-sanitizer over the small-image parity tests (memcheck, then racecheck on the TMA kernels).
OUT=gpurun_out/${1:-sanitize}; mkdir -p $OUT
K="golden_device or ragged or tma_path or decoders or pvrtc_vs or pvrtc_stripes or golden_ops or downsample_vs or pad_vs or solid"
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -x -q -m gpu -k "$K" > $OUT/memcheck.log 2>&1; echo "memcheck exit $?"; tail -6 $OUT/memcheck.log
timeout 240 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -x -q -m gpu -k "tma_path or medium or pvrtc_stripes or downsample_vs" > $OUT/racecheck.log 2>&1; echo "racecheck exit $?"; tail -6 $OUT/racecheck.log
grep -c "ERROR SUMMARY" $OUT/*.log
