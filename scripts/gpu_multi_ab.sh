#!/bin/bash
# N-GPU A/B of the column-sharded bench line (weak + strong): default, fills on the caller's stream, no dependent launches
N=$1; tag=${2:-ab}
out=gpurun_out; mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for v in ${VARIANTS:-default syncgather nopdl}; do
  unset OCF_SYNC_GATHER OCF_NO_PDL OCF_GATHER_PRIO OCF_REGATHER_SYNC
  case $v in syncgather) export OCF_SYNC_GATHER=1;; nopdl) export OCF_NO_PDL=1;; normalprio) export OCF_GATHER_PRIO=0;; regathersync) export OCF_REGATHER_SYNC=1;; esac
  run 29520 bench.py --gpus $N --steps 40 --no-cpu-baseline --others none > $out/${tag}_bench_n${N}_$v.json 2> $out/${tag}_bench_n${N}_$v.err
  python - <<PY
import json
d=json.loads(open("$out/${tag}_bench_n${N}_$v.json").read().strip().splitlines()[-1])
s=d["strong_scaling"]
print("$v: weak %.1f M (%.4f ms) e2e %.1f M | strong %.1f M (%.4f ms) e2e %.1f M" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, s["value"]/1e6, s["ms_per_step"], s["e2e"]["value"]/1e6))
PY
done
