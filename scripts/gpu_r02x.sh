#!/bin/bash
# work-item granularity sweep (OCF_TARGET_ITEMS) on the headline workload and ML-1M
out=gpurun_out; tag=${1:-r02x}
mkdir -p $out
for w in ml10m ml1m; do
for t in 444 592 740 888 1184 1480 2220 2960; do
  OCF_TARGET_ITEMS=$t timeout 600 python bench.py --workload $w --steps 40 --no-cpu-baseline --no-scoring --others none > $out/${tag}_bench_${w}_t$t.json 2> $out/${tag}_bench_${w}_t$t.err
  python - <<PY
import json
d=json.load(open("$out/${tag}_bench_${w}_t$t.json"))
k=d["kernels"]
print("$w items $t: %.1f M ratings/s  %.4f ms/step  e2e %.1f M | K1 %.1f K2 %.1f K3 %.1f K4a %.1f K4b %.1f us" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, *[1e3*v["ms"] for v in list(k.values())[:5]]))
PY
done
done
