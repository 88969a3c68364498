#!/bin/bash
# round 2, call w: CTA-cooperative heavy update tasks: parity suite, then workloads with the heavy path on / off
out=gpurun_out; tag=${1:-r02w}
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_model.py tests/test_gpu_train.py tests/test_gpu_shards.py tests/test_gpu_checkpoint.py tests/test_gpu_full_configs.py tests/test_gpu_xl_sizes.py -q -m gpu -k "not netflix" > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -6 $out/${tag}_tests.log
for w in ${WORKLOADS:-jester ml1m ml20m ml10m}; do
  timeout 600 python bench.py --workload $w --steps 30 --no-cpu-baseline --no-scoring > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err
  OCF_NO_HEAVY=1 timeout 600 python bench.py --workload $w --steps 30 --no-cpu-baseline --no-scoring > $out/${tag}_bench_${w}_noheavy.json 2> $out/${tag}_bench_${w}_noheavy.err
done
python scripts/show_quick.py $tag "" _noheavy 2>/dev/null | grep -v "ERR"
