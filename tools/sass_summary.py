"""Per-kernel SASS mnemonic counts of the built objects (mmduet_b200/build/*.o): the instructions that prove the Blackwell
paths — UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA load / store / reduce),
plus HMMA (mma.sync), LDGSTS (cp.async), MUFU.EX2.  Writes profiles/r02_sass_summary.json.  CPU only (cuobjdump)."""
import collections
import glob
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "HMMA", "LDGSTS", "LDSM", "MUFU.EX2", "SYNCS"]
out = {}
for obj in sorted(glob.glob(os.path.join(ROOT, "mmduet_b200", "build", "*.o"))):
    if obj.endswith("_jitter.o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, counts = None, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
            cur, counts = name[:150], collections.Counter()
            out.setdefault(os.path.basename(obj), {})[cur] = counts
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for p in PAT:
            if op == p or op.startswith(p + "."):
                counts[p] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            counts["UTCHMMA.2CTA"] += 1
res = {f: {k: dict(c) for k, c in ks.items() if c} for f, ks in out.items()}
tot = collections.Counter()
for ks in res.values():
    for c in ks.values():
        tot.update(c)
res["_total"] = dict(tot)
json.dump(res, open(os.path.join(ROOT, "profiles", "r02_sass_summary.json"), "w"), indent=1)
print(json.dumps(res["_total"]))
for f, ks in res.items():
    if f.startswith("_"):
        continue
    for k, c in ks.items():
        if any(x in c for x in ("UTCHMMA", "UTMALDG", "HMMA")):
            print(f, "|", k[:90], "|", {x: c[x] for x in PAT if x in c})
