#!/bin/bash
# compute-sanitizer memcheck over the tests added in round 2 (host-pipe pool, shard context, unaligned sources, fused
# PVRTC kernel, registered buffers, C++ classes through both header builds).
OUT=gpurun_out/${1:-sanitize_r2}; mkdir -p $OUT
K="unaligned_device_sources or shard_context_one_device or pipes_are_pooled or failed_host_call or registered_caller or pvrtc_fused or all_ten_virtuals or compress_matches_oracle"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -x -q -m gpu -k "$K" > $OUT/memcheck_r2_tests.log 2>&1; echo "memcheck exit $?"; tail -6 $OUT/memcheck_r2_tests.log
