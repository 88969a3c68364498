#!/bin/bash
# launch lists (ncu gpu__time_duration) of the small workloads
out=gpurun_out; tag=${1:-r02t}
mkdir -p $out
for w in jester ml20m ml1m; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches_$w.csv \
      python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --others none --no-scoring > $out/${tag}_ncu_$w.log 2>&1
  python scripts/launch_summary.py $out/${tag}_launches_$w.csv | tail -22
done
