"""One-line summaries of bench JSON lines: python scripts/show_line.py FILE..."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "ERR", e); continue
    r = d.get("roofline") or {}
    ks = " ".join("%s=%.1f" % (k.split(" ")[-1].strip("()"), 1e3 * v["ms"]) for k, v in (d.get("kernels") or {}).items())
    print("%-44s value %7.1fM e2e %7.1fM us/step %6.1f e2e_us %6.1f L/step %.1f frac %.3f step_frac %.3f | %s" % (
        f.split("/")[-1], d["value"] / 1e6, d["e2e"]["value"] / 1e6, 1e3 * d["ms_per_step"], 1e3 * d["e2e"]["ms_per_step"],
        d["gpu_launches"] / d["steps"], r.get("frac") or 0, r.get("step_frac") or 0, ks))
