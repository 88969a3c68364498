#!/bin/bash
# round 2, call y: fills on the batch's own stream (K1 overlaps the previous step): whole GPU suite, then workloads with it on / off
out=gpurun_out; tag=${1:-r02y}
mkdir -p $out
timeout 2400 python -m pytest tests -q -m gpu -k "not netflix" > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -6 $out/${tag}_tests.log
for w in ${WORKLOADS:-jester ml1m ml20m ml10m}; do
  timeout 600 python bench.py --workload $w --steps 40 --no-cpu-baseline --no-scoring --others none > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err
  OCF_SYNC_GATHER=1 timeout 600 python bench.py --workload $w --steps 40 --no-cpu-baseline --no-scoring --others none > $out/${tag}_bench_${w}_sync.json 2> $out/${tag}_bench_${w}_sync.err
done
python scripts/show_quick.py $tag "" _sync 2>/dev/null | grep -v "ERR\|GB/s"
