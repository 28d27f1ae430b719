#!/bin/bash
# One ncu --set full capture per workload (one launch of the hot kernel each) + the launch list of the headline
# bench command.  Usage (on the GPU box): bash tools/gpu_profile_all.sh <tag> [workloads...]
TAG=$1; shift
WLS=${@:-"dxt1_rgba8 dxt5_rgba8 dxt1_rgb8 etc1_rgb8 pvrtc2_rgba8"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
for wl in $WLS; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"encode4x4_(tma|ring)|pvrtc" -s 6 -c $([ $wl = pvrtc2_rgba8 ] && echo 3 || echo 1) -f -o $OUT/prof_$wl \
      python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_$wl.log 2>&1; echo "ncu $wl exit $?"
  python tools/ncu_summary.py $OUT/prof_$wl.ncu-rep > $OUT/summary_$wl.txt 2>&1
  ncu -i $OUT/prof_$wl.ncu-rep --page source --csv > $OUT/source_$wl.csv 2>/dev/null
  python tools/ncu_src_hist.py $OUT/source_$wl.csv > $OUT/hist_$wl.txt 2>&1
  # gpurun brings back at most 64 MiB: keep the text summaries, drop the bulky report and source table
  gzip -9 $OUT/source_$wl.csv; [ $(stat -c %s $OUT/source_$wl.csv.gz) -gt 4000000 ] && rm -f $OUT/source_$wl.csv.gz
  [ $(stat -c %s $OUT/prof_$wl.ncu-rep) -gt 8000000 ] && rm -f $OUT/prof_$wl.ncu-rep
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_dxt1_rgba8.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "launch list exit $?"
