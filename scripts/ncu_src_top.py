"""Top stall / instruction lines of an `ncu --page source --csv` dump (SASS or CUDA view)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
body = rows[2:]
i_src, i_smp, i_ins = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot_s = sum(float(r[i_smp] or 0) for r in body)
tot_i = sum(float(r[i_ins] or 0) for r in body)
print("total samples %d, instructions %d" % (tot_s, tot_i))
key = i_ins if len(sys.argv) > 2 and sys.argv[2] == "inst" else i_smp
for r in sorted(body, key=lambda r: -float(r[key] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print("%5.1f%% smp %5.1f%% ins  %s" % (100 * float(r[i_smp] or 0) / max(tot_s, 1), 100 * float(r[i_ins] or 0) / max(tot_i, 1), r[i_src][:120]))
