#!/usr/bin/env python
"""SASS instruction histogram of libflb.so (cuobjdump -sass), overall and per kernel for the tensor-core / TMA mnemonics
that prove the contraction path is tcgen05 + TMEM + TMA (B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld),
UTMALDG (cp.async.bulk.tensor), UTCBAR (tcgen05.commit), SYNCS (mbarrier).

    python scripts/sass_histogram.py > profiles/r02_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "federated-learning-for-privacy-preserving-image-classification_b200", "libflb.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
tot = collections.Counter()
per = collections.defaultdict(collections.Counter)
fn = None
KEY = ("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "SYNCS", "UTCCP", "REDG", "RED", "ATOMG", "STG", "LDG", "FFMA", "HMMA", "IMMA")
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if m and fn:
        op = m.group(1)
        tot[op.split(".")[0]] += 1
        base = op.split(".")[0]
        if base in KEY:
            per[fn][op if base in ("UTCHMMA", "UTMALDG", "LDTM") else base] += 1
print(f"# {os.path.basename(lib)}: {sum(tot.values())} SASS instructions in {len(per)} kernels that use the listed mnemonics")
print("# overall, the 40 most frequent mnemonics")
for op, n in tot.most_common(40):
    print(f"{n:8d}  {op}")
print("\n# tensor-core / TMA mnemonics, whole library")
for k in ("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "SYNCS", "HMMA", "IMMA"):
    print(f"{tot.get(k, 0):8d}  {k}")
print("\n# per kernel (only kernels with tcgen05 / TMA instructions)")
for f, c in sorted(per.items()):
    if any(k.startswith(("UTCHMMA", "UTMALDG", "LDTM")) for k in c):
        print(f[:150])
        print("    " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items()) if not k.startswith(("STG", "LDG", "FFMA"))))
