#!/bin/bash
# (historical: ran on commit a1179e3 + the staged kernels of csrc/ocf_staged.cuh, OCF_DEC_STAGED / OCF_UPD_STAGED; removed since)
# round 2, session 3, call B: K3 / K4b staged through shared memory (cp.async.bulk), A/B on one box
out=gpurun_out; tag=r04b; mkdir -p $out
timeout 900 python -m pytest tests -q -m gpu -x > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -4 $out/${tag}_tests.log
run() { name=$1; shift; env "$@" timeout 600 $B > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python scripts/show_line.py $out/${tag}_$name.json; }
B="python bench.py --others none --no-cpu-baseline --no-scoring"
run ml10m_both X=1
run ml10m_noupd OCF_UPD_STAGED=0
run ml10m_lag3 OCF_BENCH_LAG=3 OCF_RING_DEPTH=4
run ml10m_lag4 OCF_BENCH_LAG=4 OCF_RING_DEPTH=6
run ml10m_nodec OCF_DEC_STAGED=0
run ml10m_none OCF_DEC_STAGED=0 OCF_UPD_STAGED=0
B="python bench.py --workload netflix --steps 20 --others none --no-cpu-baseline --no-scoring"
run netflix_both X=1
run netflix_noupd OCF_UPD_STAGED=0
run netflix_decst OCF_DEC_STAGED=1
B="python bench.py --workload ml1m --others none --no-cpu-baseline --no-scoring"
run ml1m_both X=1
run ml1m_nodec OCF_DEC_STAGED=0
B="python bench.py --workload ml20m --others none --no-cpu-baseline --no-scoring"
run ml20m_both X=1
run ml20m_nodec OCF_DEC_STAGED=0
run ml20m_noovl OCF_NO_OVERLAP=1
B="python bench.py --workload jester --others none --no-cpu-baseline --no-scoring"
run jester_both X=1
run jester_nodec OCF_DEC_STAGED=0
run jester_noovl OCF_NO_OVERLAP=1
