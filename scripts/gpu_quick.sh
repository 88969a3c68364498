#!/bin/bash
# quick A/B on one box: a subset of the GPU tests + the five bench workloads (value / e2e / per-kernel times)
out=gpurun_out; tag=${1:-quick}; shift
tests=${TESTS:-"tests/test_gpu_model.py tests/test_gpu_checkpoint.py tests/test_gpu_train.py tests/test_gpu_shards.py"}
mkdir -p $out
python -m pytest $tests -q -m gpu -x > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -5 $out/${tag}_tests.log
for w in ${WORKLOADS:-ml10m ml1m jester ml20m netflix}; do
  timeout 900 python bench.py --workload $w --steps ${STEPS:-30} --no-cpu-baseline > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err
done
