#!/bin/bash
# round 2, session 3, call F: K3 with the batch row's activations in shared memory (6 CTAs/SM), decoder tasks first in K4b
out=gpurun_out; tag=r04f; mkdir -p $out
timeout 900 python -m pytest tests -q -m gpu -x > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -4 $out/${tag}_tests.log
run() { name=$1; shift; env "$@" timeout 600 $B > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python scripts/show_line.py $out/${tag}_$name.json; }
B="python bench.py --others none --no-cpu-baseline --no-scoring"
run ml10m X=1
run ml10m_nodf OCF_DEC_FIRST=0
run ml10m_b X=1
run ml10m_nodf_b OCF_DEC_FIRST=0
B="python bench.py --workload netflix --steps 20 --others none --no-cpu-baseline --no-scoring"
run netflix X=1
run netflix_nodf OCF_DEC_FIRST=0
B="python bench.py --workload ml1m --others none --no-cpu-baseline --no-scoring"
run ml1m X=1
B="python bench.py --workload ml20m --others none --no-cpu-baseline --no-scoring"
run ml20m X=1
run ml20m_nodf OCF_DEC_FIRST=0
B="python bench.py --workload jester --others none --no-cpu-baseline --no-scoring"
run jester X=1
