#!/bin/bash
# (historical: ran on commit a1179e3, whose Makefile built the OCF_K4B_VARIANT libraries v1 / v2 that OCF_LIB_VARIANT selected)
# round 2, session 3, call A: K4a ahead of the step (batch-side work list) + K4b scheduling variants, A/B on one box
out=gpurun_out; tag=r04a; mkdir -p $out
python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py tests/test_gpu_checkpoint.py tests/test_gpu_batches.py -q -m gpu -x > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -4 $out/${tag}_tests.log
B="python bench.py --others none --no-cpu-baseline --no-scoring"
run() { name=$1; shift; env "$@" timeout 600 $B > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python scripts/show_line.py $out/${tag}_$name.json; }
run ml10m_base X=1
run ml10m_noahead OCF_NO_AHEAD=1
run ml10m_v1 OCF_LIB_VARIANT=v1
run ml10m_v2 OCF_LIB_VARIANT=v2
run ml10m_base2 X=1
B="python bench.py --workload ml1m --others none --no-cpu-baseline --no-scoring"
run ml1m_base X=1
run ml1m_noahead OCF_NO_AHEAD=1
run ml1m_v2 OCF_LIB_VARIANT=v2
B="python bench.py --workload netflix --steps 20 --others none --no-cpu-baseline --no-scoring"
run netflix_base X=1
run netflix_v2 OCF_LIB_VARIANT=v2
B="python bench.py --workload ml20m --others none --no-cpu-baseline --no-scoring"
run ml20m_base X=1
run ml20m_v2 OCF_LIB_VARIANT=v2
B="python bench.py --workload jester --others none --no-cpu-baseline --no-scoring"
run jester_base X=1
