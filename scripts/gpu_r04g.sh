#!/bin/bash
# round 2, session 3, call G: lean K4b with evict-first state loads / row stores (what stays in L2 are K3's rows)
out=gpurun_out; tag=r04g; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_worklist.py tests/test_gpu_xl_sizes.py -q -m gpu -x > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -3 $out/${tag}_tests.log
run() { name=$1; shift; env "$@" timeout 600 $B > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python scripts/show_line.py $out/${tag}_$name.json; }
B="python bench.py --others none --no-cpu-baseline --no-scoring"
run ml10m X=1
run ml10m_nostream OCF_K4B_STREAM=0
run ml10m_b X=1
run ml10m_nostream_b OCF_K4B_STREAM=0
B="python bench.py --workload netflix --steps 20 --others none --no-cpu-baseline --no-scoring"
run netflix X=1
run netflix_nostream OCF_K4B_STREAM=0
