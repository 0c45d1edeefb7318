#!/bin/bash
# A/B of two builds of the library on ONE box: ncu launch list (gpu__time_duration) of one 1024-walker local-energy pass
# (scripts/prof_le.py) per build, summarised per kernel.  usage: scripts/ab_launches.sh ab/base.so ab/new.so
set -u
mkdir -p gpurun_out
for lib in "$@"; do
  tag=$(basename "$lib" .so)
  cp "$lib" deephall_b200/libdeephall_b200.so
  timeout 300 python scripts/prof_le.py 1024 > gpurun_out/ab_${tag}_plain.log 2>&1 || { echo "plain run failed for $lib"; tail -5 gpurun_out/ab_${tag}_plain.log; continue; }
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ab_${tag}.csv \
      python scripts/prof_le.py 1024 > gpurun_out/ab_${tag}_ncu.log 2>&1
  echo "== $tag"
  python scripts/summarize_launches.py gpurun_out/ab_${tag}.csv 0.6667 | head -16
done
