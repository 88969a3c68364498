#!/bin/bash
# N-GPU check on one box: real-NCCL parity (scripts/dist_check.py) + the column-sharded bench line
# usage: gpurun --gpus N -- 'bash scripts/gpu_multi.sh N TAG [workload]'
N=$1; tag=${2:-multi}; w=${3:-ml10m}
out=gpurun_out; mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 scripts/dist_check.py > $out/${tag}_dist_check_n$N.txt 2>&1; echo "rc=$?" >> $out/${tag}_dist_check_n$N.txt
grep -E "dist_check|rc=|Error|error" $out/${tag}_dist_check_n$N.txt | tail -8
run 29512 bench.py --gpus $N --steps ${STEPS:-30} --workload $w --no-cpu-baseline > $out/${tag}_bench_${w}_n$N.json 2> $out/${tag}_bench_${w}_n$N.err
tail -c 1500 $out/${tag}_bench_${w}_n$N.json
OCF_NO_GRAPH=1 run 29513 bench.py --gpus $N --steps ${STEPS:-30} --workload $w --no-cpu-baseline > $out/${tag}_bench_${w}_n${N}_nograph.json 2> $out/${tag}_bench_${w}_n${N}_nograph.err
