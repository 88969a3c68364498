#!/bin/bash
# N-GPU check on one box: real-NCCL parity (scripts/dist_check.py) + the column-sharded bench line (weak + strong + parity
# + per-rank kernel times; OTHERS=netflix adds that config)
# usage: gpurun --gpus N -- 'bash scripts/gpu_multi.sh N TAG [workload]'
N=$1; tag=${2:-multi}; w=${3:-ml10m}
out=gpurun_out; mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 scripts/dist_check.py > $out/${tag}_dist_check_n$N.txt 2>&1; echo "rc=$?" >> $out/${tag}_dist_check_n$N.txt
grep -E "dist_check|rc=|Error|error" $out/${tag}_dist_check_n$N.txt | tail -8
( time run 29512 bench.py --gpus $N --steps ${STEPS:-30} --workload $w --no-cpu-baseline ${OTHERS:+--others $OTHERS} > $out/${tag}_bench_${w}_n$N.json ) 2> $out/${tag}_bench_${w}_n$N.err
tail -c 1200 $out/${tag}_bench_${w}_n$N.json; tail -5 $out/${tag}_bench_${w}_n$N.err
if [ -n "$ROWS" ]; then
  run 29513 bench.py --gpus $N --steps ${STEPS:-30} --workload $ROWS --parallel rows --no-cpu-baseline > $out/${tag}_bench_${ROWS}_rows_n$N.json 2> $out/${tag}_bench_${ROWS}_rows_n$N.err
  tail -c 600 $out/${tag}_bench_${ROWS}_rows_n$N.json
fi
