#!/bin/bash
# round 2, session 3, call H: which updated weight rows should stay normal L2 lines for the next step's K2 / K3
out=gpurun_out; tag=r04h; mkdir -p $out
run() { name=$1; shift; env "$@" timeout 600 $B > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python scripts/show_line.py $out/${tag}_$name.json; }
B="python bench.py --others none --no-cpu-baseline --no-scoring"
for rep in a b; do for sv in 1 3 5 7; do run ml10m_s${sv}_$rep OCF_K4B_STREAM=$sv; done; done
B="python bench.py --workload netflix --steps 20 --others none --no-cpu-baseline --no-scoring"
for sv in 1 3 5; do run netflix_s$sv OCF_K4B_STREAM=$sv; done
