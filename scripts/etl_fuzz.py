#!/usr/bin/env python
"""Mutation fuzzing of the native file readers / the splitter (`csrc/ocf_etl.cpp`) under AddressSanitizer and
UndefinedBehaviorSanitizer, host only:

    python scripts/etl_fuzz.py [seed] [cases]

builds `scripts/etl_fuzz_driver.cpp` (which includes ocf_etl.cpp) with -fsanitize=address,undefined into /tmp,
then feeds it the golden JSON / CSV files with random deletions, insertions of structural tokens, byte flips
and truncations. Every case must end in "ok" or a clean error message. Round 1: 3 300 cases, 0 findings."""
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORK = "/tmp/ocf_etl_fuzz"
os.makedirs(WORK, exist_ok=True)
subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-w", "-pthread", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", WORK + "/drv",
                os.path.join(ROOT, "scripts", "etl_fuzz_driver.cpp")], check=True)
os.chdir(WORK)
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
G = os.path.join(ROOT, "tests", "golden", "split")
os.makedirs("work/out", exist_ok=True)
vocab = G + "/ml/unique_items_list.json"
json_files = [(G + "/ml/ratingsByUser_dicts_train.json", 0), (G + "/ml/ratingsByUser_dicts_valid.json", 1),
              (G + "/ml_ts/ratingsByUser_dicts_withtimestamps_valid.json", 1), (G + "/amazon/ratingsByUser_dicts_test.json", 1)]
csv_files = [(G + "/%s/ratings.csv" % n, 3 if n.startswith("netflix") else 4) for n in ("ml", "amazon", "netflix_int", "amazon_ts_rev")]
tokens = [b"[", b"]", b"{", b"}", b",", b":", b'"', b"\\", b"null", b"NaN", b"-", b"1e999", b"\\u12", b"\\ud800", b"\n", b"\r\n", b'""', b"0x10", b".", b"e", b"\x00", b"\xff"]
def mutate(data):
    data = bytearray(data)
    for _ in range(random.randint(1, 6)):
        op = random.random()
        pos = random.randrange(len(data) + 1)
        if op < 0.3 and data:
            del data[pos:pos + random.randint(1, 20)]
        elif op < 0.6:
            data[pos:pos] = random.choice(tokens)
        elif op < 0.8 and data:
            data[min(pos, len(data) - 1)] = random.randrange(256)
        else:
            data = data[:pos]
    return bytes(data)
bad = 0
N = int(sys.argv[2]) if len(sys.argv) > 2 else 400
for it in range(N):
    if random.random() < 0.5:
        path, paired = random.choice(json_files)
        data = mutate(open(path, "rb").read())
        open("work/f.json", "wb").write(data)
        v = vocab
        if random.random() < 0.2:
            open("work/v.json", "wb").write(mutate(open(vocab, "rb").read())); v = "work/v.json"
        cmd = ["./drv", "json", v, "work/f.json", str(paired if random.random() < 0.8 else 1 - paired)]
    else:
        path, nc = random.choice(csv_files)
        data = mutate(open(path, "rb").read())
        open("work/f.csv", "wb").write(data)
        cmd = ["./drv", "csv", "work/f.csv", str(nc), "work/out/"]
    r = subprocess.run(cmd, capture_output=True)
    if r.returncode != 0 or b"ERROR" in r.stderr or b"runtime error" in r.stderr:
        bad += 1
        print("FAIL", cmd, r.returncode, r.stderr[-1500:].decode(errors="replace"))
        os.system("cp work/f.json work/fail_%d.json 2>/dev/null; cp work/f.csv work/fail_%d.csv 2>/dev/null" % (it, it))
        if bad > 3: break
print("done", N, "cases,", bad, "failures")
