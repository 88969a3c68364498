#!/usr/bin/env python
"""profiles/traffic.json: DRAM bytes per launch of the step's kernels from a raw `ncu --set full` page, so that
`bench.py`'s roofline.traffic is read from the committed capture instead of being a literal.

    python scripts/ncu_traffic.py WORKLOAD RAW.csv SOURCE_NAME     (RAW.csv = ncu -i X.ncu-rep --page raw --csv)

Per kernel (first word of its name up to '<' or '('): mean over the captured launches of
dram__bytes_read.sum + dram__bytes_write.sum, converted to bytes."""
import csv
import json
import os
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
workload, raw, source = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
acc = {}
for r in data:
    name = r[ix["Kernel Name"]].replace("void ", "").replace("ocf::", "")
    key = name.split("<")[0].split("(")[0].strip()
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[ix[m]].replace(",", "")) * UNIT[units[ix[m]]]
    acc.setdefault(key, []).append((tot, float(r[ix["gpu__time_duration.sum"]].replace(",", ""))))
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
out = json.load(open(path)) if os.path.exists(path) else {}
for key, vals in acc.items():
    out.setdefault(workload, {})[key] = {"dram_bytes": sum(v[0] for v in vals) / len(vals), "launches": len(vals),
                                         "ncu_duration_us": sum(v[1] for v in vals) / len(vals),
                                         "source": "profiles/%s (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, cold serialised replay)" % source}
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(out.get(workload), indent=1))
