#!/bin/bash
# round 2, session 3, call I: K2 / K3 with their row loads forced into flight together (dependency chain), rows per round
out=gpurun_out; tag=r04i; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_worklist.py tests/test_gpu_xl_sizes.py tests/test_gpu_score.py -q -m gpu -x > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -3 $out/${tag}_tests.log
run() { name=$1; shift; env "$@" timeout 600 $B > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python scripts/show_line.py $out/${tag}_$name.json; }
B="python bench.py --others none --no-cpu-baseline --no-scoring"
for r in 2 3 4 2; do run ml10m_rif$r OCF_K2_RIF=$r; done
B="python bench.py --workload ml1m --others none --no-cpu-baseline --no-scoring"
for r in 2 3 4; do run ml1m_rif$r OCF_K2_RIF=$r; done
B="python bench.py --workload jester --others none --no-cpu-baseline --no-scoring"
for r in 2 4; do run jester_rif$r OCF_K2_RIF=$r; done
B="python bench.py --workload ml20m --others none --no-cpu-baseline --no-scoring"
run ml20m X=1
