OUT=gpurun_out/r2d; mkdir -p $OUT
for wl in dxt5_rgba8 dxt1_rgba8; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"encode4x4_(tma|ring)" -s 6 -c 1 -o $OUT/prof_$wl \
    python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-others > $OUT/ncu_$wl.log 2>&1; echo "ncu $wl exit $?"
done
