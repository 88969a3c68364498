#!/bin/bash
# round 2, session 3, call N: cluster split of the hidden-layer contractions at the small shapes (OCF_TC_SPLIT)
out=gpurun_out; tag=r04n; mkdir -p $out
run() { name=$1; shift; env "$@" timeout 120 $B > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python scripts/show_line.py $out/${tag}_$name.json; }
B="python bench.py --workload jester --steps 30 --others none --no-cpu-baseline --no-scoring"
run jester_auto X=1
for sp in 1 2 4; do run jester_split$sp OCF_TC_SPLIT=$sp; done
B="python bench.py --workload ml20m --steps 30 --others none --no-cpu-baseline --no-scoring"
run ml20m_auto X=1
for sp in 2 4; do run ml20m_split$sp OCF_TC_SPLIT=$sp; done
