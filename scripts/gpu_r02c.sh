#!/bin/bash
# round 2, call c: the whole GPU suite on the graph-replay build, the five bench workloads (graphs on / off),
# ncu --set full of K2 / K3 / K4a on the default bench workload
out=gpurun_out; tag=${1:-r02c}
mkdir -p $out
set -x
free -g | head -2 > $out/${tag}_host.txt; nproc >> $out/${tag}_host.txt
timeout 2400 python -m pytest tests -q -m gpu --durations=15 > $out/${tag}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_gpu_tests.log
tail -30 $out/${tag}_gpu_tests.log
for w in ml10m ml1m jester ml20m netflix; do
  timeout 900 python bench.py --workload $w --steps 30 --no-cpu-baseline > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err
  OCF_NO_GRAPH=1 timeout 900 python bench.py --workload $w --steps 30 --no-cpu-baseline > $out/${tag}_bench_${w}_nograph.json 2> $out/${tag}_bench_${w}_nograph.err
done
ncu --set full --clock-control none --import-source on -k regex:'k_enc_fwd|k_dec_fwd|k_sort' -s 30 -c 6 -f -o $out/${tag}_prof_k2k3k4a \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_k2k3.log 2>&1
ncu -i $out/${tag}_prof_k2k3k4a.ncu-rep --page raw --csv > $out/${tag}_prof_k2k3k4a_raw.csv 2>/dev/null
ls -la $out | tail
