"""SASS evidence of the shipped library (no GPU needed): per kernel, how many tcgen05 / tensor-memory / bulk-copy instructions it holds,
and a few of the lines themselves.     python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "depecher_b200", "lib", "libdepecher_b200.so")
PAT = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "SYNCS", "ELECT", "UCGABAR", "USETMAXREG", "F2F.F64.F32", "DADD", "ATOMG", "REDG", "HMMA"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
fn, counts, samples = None, collections.OrderedDict(), {}
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[fn] = collections.Counter()
        continue
    if fn is None or "/*" not in ln:
        continue
    body = ln.split("*/", 1)[-1]
    for p in PAT:
        if re.search(r"\b" + re.escape(p), body):
            counts[fn][p] += 1
            key = (fn, p)
            if key not in samples:
                samples[key] = re.sub(r"\s*/\*.*", "", body).strip()
print("cuobjdump -sass depecher_b200/lib/libdepecher_b200.so  (sm_100a; counts of instruction mnemonics per kernel; HMMA = legacy mma.sync: none expected)\n")
print(f"{'kernel':78s} " + " ".join(f"{p[:9]:>9s}" for p in PAT))
for fn, c in counts.items():
    if sum(c.values()) == 0:
        continue
    print(f"{fn[:78]:78s} " + " ".join(f"{c[p]:9d}" for p in PAT))
print("\nexamples (first occurrence per kernel):")
for (fn, p), body in samples.items():
    if p in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "UCGABAR", "USETMAXREG"):
        print(f"  {fn[:60]:60s} {body}")
