"""Pretty-print the JSON line(s) bench.py wrote into a file."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.1f M %s  %.3f ms/step | e2e %.1f M/s %.3f ms/step | launches %d" % (
    d["value"] / 1e6, d["unit"], d["ms_per_step"], d["e2e"]["value"] / 1e6,
    d["e2e"].get("ms_per_step", d["e2e"].get("ms_per_call", 0)), d["gpu_launches"]))
r = d.get("roofline")
if r:
    print("roofline: %s %.0f %s = %.3f of %s (share of step %.2f)" % (r["kernel"], r["achieved"], r["unit"], r["frac"], r["peak"], r["share_of_step"]))
tot = 0
for k, v in (d.get("kernels") or d.get("kernels_rank0") or {}).items():
    tot += v["ms"]
    print("  %-46s %7.1f us  %6.0f GB/s" % (k, v["ms"] * 1e3, v.get("GB/s") or 0))
print("  tagged total %.1f us of %.1f us" % (tot * 1e3, d["ms_per_step"] * 1e3))
if "scoring" in d:
    s = d["scoring"]
    if "error" in s:
        print("scoring: ERROR", s["error"])
    else:
        print("scoring: %.2f M rows/s (%d rows/call, %.3f ms) | gemm %.0f TFLOP/s = %.3f of %.0f | e2e %.0f rows/s" % (
            s["value"] / 1e6, s["rows_per_call"], s["ms_per_call"], s["roofline"]["achieved"], s["roofline"]["frac"],
            s["roofline"]["peak"], s["e2e"]["value"]))
        if "topk" in s:
            print("         top-%d e2e %.2f M rows/s (%.3f ms/call)" % (s["topk"]["k"], s["topk"]["e2e_value"] / 1e6, s["topk"]["ms_per_call"]))
if "cpu_baseline" in d:
    print("cpu_baseline", d["cpu_baseline"])
print("clocks", d.get("clocks"))
