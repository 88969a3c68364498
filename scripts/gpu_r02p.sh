#!/bin/bash
# round 2, call p: tcgen05 hidden-layer contraction kernel (descriptor / split-K correctness), RNG service with the 2-launch jump,
# model parity with the tensor-core hidden layers, ml20m / jester bench with and without them
out=gpurun_out; tag=${1:-r02p}
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu > $out/${tag}_gemm_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_gemm_tests.log
tail -15 $out/${tag}_gemm_tests.log
timeout 900 python -m pytest tests/test_gpu_rng.py tests/test_gpu_model.py tests/test_gpu_train.py tests/test_gpu_score.py tests/test_gpu_full_configs.py -q -m gpu -k "not netflix" > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -15 $out/${tag}_tests.log
python scripts/mt_bench.py > $out/${tag}_mt_bench.txt 2>&1; tail -4 $out/${tag}_mt_bench.txt
for w in ml20m jester; do
  timeout 600 python bench.py --workload $w --steps 30 --no-cpu-baseline > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err
  OCF_NO_TC_HIDDEN=1 timeout 600 python bench.py --workload $w --steps 30 --no-cpu-baseline > $out/${tag}_bench_${w}_simt.json 2> $out/${tag}_bench_${w}_simt.err
done
python scripts/show_quick.py $tag 2>/dev/null | tail -12
