#!/usr/bin/env python
"""Static evidence from the built library, no GPU needed: per-kernel registers / shared memory / spills from
ptxas (`csrc/build.log`, written by `make`), and the SASS mnemonics that show which hardware paths the kernels
take (B200_PROFILING.md: UTCHMMA = tcgen05.mma, UTMALDG / UBLKCP = TMA tensor / bulk copies, UTCBAR = tcgen05
commit, SYNCS = mbarrier traffic, LDTM = tcgen05.ld).

    python scripts/static_report.py > profiles/rNN_static_kernels.txt
"""
import collections
import os
import re
import subprocess

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(HERE, "omnidirectional_collaborative_filtering_b200", "csrc")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", n).replace("void ", "") for n in out]


def main():
    log = open(os.path.join(CSRC, "build.log")).read()
    rows = []
    for m in re.finditer(r"Compiling entry function '([^']+)' for 'sm_100a'\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                         r"ptxas info\s*: Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", log):
        rows.append((m.group(1), int(m.group(5)), int(m.group(7) or 0), int(m.group(2)), int(m.group(3)), int(m.group(4))))
    names = demangle([r[0] for r in rows])
    print("ptxas -v, sm_100a (%d kernels)" % len(rows))
    print("%-64s %5s %9s %6s %12s" % ("kernel", "regs", "smem B", "stack", "spill st/ld"))
    for n, r in sorted(zip(names, rows)):
        print("%-64s %5d %9d %6d %7d/%d" % (n[:64], r[1], r[2], r[3], r[4], r[5]))
    so = os.path.join(CSRC, "libocf_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    per = collections.defaultdict(collections.Counter)
    cur = None
    want = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "LDTM", "STTM", "HMMA", "ELECT",
            "MUFU", "ATOMG", "RED", "LDG.E.128", "STG.E.128")
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur:
            for w in want:
                if re.search(r"\b" + re.escape(w), line):
                    per[cur][w] += 1
    print("\nSASS mnemonic counts (cuobjdump -sass), kernels that use the async / tensor paths or atomics")
    fn = list(per)
    for n, f in sorted(zip(demangle(fn), fn)):
        c = per[f]
        if any(c[w] for w in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "ATOMG", "RED", "SYNCS")):
            print("%-64s %s" % (n[:64], "  ".join("%s=%d" % (w, c[w]) for w in want if c[w])))


if __name__ == "__main__":
    main()
