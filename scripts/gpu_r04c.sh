#!/bin/bash
# round 2, session 3, call C: lean K4b with the task cursor, hidden-layer overlap bits, e2e host timeline
out=gpurun_out; tag=r04c; mkdir -p $out
timeout 900 python -m pytest tests -q -m gpu -x > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -4 $out/${tag}_tests.log
timeout 300 python scripts/e2e_trace.py ml10m 60 > $out/${tag}_e2e_trace_ml10m.txt 2>&1; tail -16 $out/${tag}_e2e_trace_ml10m.txt
run() { name=$1; shift; env "$@" timeout 600 $B > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python scripts/show_line.py $out/${tag}_$name.json; }
B="python bench.py --others none --no-cpu-baseline --no-scoring"
run ml10m X=1
B="python bench.py --workload netflix --steps 20 --others none --no-cpu-baseline --no-scoring"
run netflix X=1
B="python bench.py --workload ml1m --others none --no-cpu-baseline --no-scoring"
run ml1m X=1
B="python bench.py --workload ml20m --others none --no-cpu-baseline --no-scoring"
run ml20m_ovl3 X=1
run ml20m_ovl1 OCF_OVERLAP=1
run ml20m_ovl2 OCF_OVERLAP=2
run ml20m_ovl0 OCF_OVERLAP=0
B="python bench.py --workload jester --others none --no-cpu-baseline --no-scoring"
run jester_ovl3 X=1
run jester_ovl1 OCF_OVERLAP=1
run jester_ovl0 OCF_OVERLAP=0
timeout 300 python scripts/e2e_trace.py ml1m 60 > $out/${tag}_e2e_trace_ml1m.txt 2>&1; tail -8 $out/${tag}_e2e_trace_ml1m.txt
