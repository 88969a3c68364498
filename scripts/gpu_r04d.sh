#!/bin/bash
# round 2, session 3, call D: dz fused into K3's row tail + bias gradients / metrics on the side stream, shorter
# prologue chains, RNG run-ahead cap; epoch-start anatomy
out=gpurun_out; tag=r04d; mkdir -p $out
timeout 900 python -m pytest tests -q -m gpu -x > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -4 $out/${tag}_tests.log
timeout 300 python scripts/e2e_trace.py ml10m 60 > $out/${tag}_e2e_trace_ml10m.txt 2>&1; grep -v "steps [1-5]" $out/${tag}_e2e_trace_ml10m.txt | tail -40
run() { name=$1; shift; env "$@" timeout 600 $B > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python scripts/show_line.py $out/${tag}_$name.json; }
B="python bench.py --others none --no-cpu-baseline --no-scoring"
run ml10m X=1
run ml10m_ovl3 OCF_OVERLAP=3

B="python bench.py --steps 20 --others none --no-cpu-baseline --no-scoring"
run ml10m_steps20 X=1
B="python bench.py --workload ml1m --others none --no-cpu-baseline --no-scoring"
run ml1m X=1
run ml1m_ovl3 OCF_OVERLAP=3
B="python bench.py --workload ml20m --others none --no-cpu-baseline --no-scoring"
run ml20m X=1
run ml20m_ovl3 OCF_OVERLAP=3
B="python bench.py --workload jester --others none --no-cpu-baseline --no-scoring"
run jester X=1
run jester_ovl3 OCF_OVERLAP=3
run jester_ovl0 OCF_OVERLAP=0
B="python bench.py --workload netflix --steps 20 --others none --no-cpu-baseline --no-scoring"
run netflix X=1
