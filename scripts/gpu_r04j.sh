#!/bin/bash
# round 2, session 3, call J: work-item granularity again, now that K4a is off the step and K2 / K3 keep their loads in flight
out=gpurun_out; tag=r04j; mkdir -p $out
run() { name=$1; shift; env "$@" timeout 300 $B > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python scripts/show_line.py $out/${tag}_$name.json; }
B="python bench.py --others none --no-cpu-baseline --no-scoring"
for t in 592 740 888 1036 1332; do run ml10m_t$t OCF_TARGET_ITEMS=$t; done
B="python bench.py --workload ml1m --others none --no-cpu-baseline --no-scoring"
for t in 592 888 1332; do run ml1m_t$t OCF_TARGET_ITEMS=$t; done
