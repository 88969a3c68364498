import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.1f M ratings/s  %.3f ms/step | e2e %.1f M/s %.3f ms/step | launches %d" % (
    d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"].get("ms_per_step", 0), d["gpu_launches"]))
r = d["roofline"]
print("roofline: %s %.0f GB/s = %.3f of %s (share of step %.2f)" % (r["kernel"], r["achieved"], r["frac"], r["peak"], r["share_of_step"]))
tot = 0
for k, v in d.get("kernels", {}).items():
    tot += v["ms"]
    print("  %-24s %7.1f us  %6.0f GB/s" % (k, v["ms"] * 1e3, v["GB/s"] or 0))
print("  tagged total %.1f us of %.1f us" % (tot * 1e3, d["ms_per_step"] * 1e3))
if "cpu_baseline" in d:
    print("cpu_baseline", d["cpu_baseline"])
print("clocks", d.get("clocks"))
