"""Static SASS census of so_b200/libsogpu.so (cuobjdump -sass; no GPU needed):
   python tools/sass_census.py > /tmp/census.md      # the table of profiles/r2_sass_census.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "so_b200", "libsogpu.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["UBLKCP", "SYNCS", "ATOMS", "ATOMG", "RED", "REDUX", "LDS", "STS", "LDG", "STG", "FFMA", "DFMA", "DMUL", "DADD", "MUFU",
        "F2I", "BAR", "CCTL", "MATCH"]
tensor = re.compile(r"\b(UTC\w*MMA|HMMA|IMMA|DMMA|QGMMA|HGMMA)\b")
per = collections.OrderedDict()
cur = None
n_tensor = 0
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["total"] += 1
        if tensor.search(line):
            n_tensor += 1
        for c in cols:
            if op == c or op.startswith(c + "."):
                per[cur][c] += 1
print("| kernel | total | " + " | ".join(cols) + " |")
print("|---|---:|" + "---:|" * len(cols))
tot = collections.Counter()
for k, c in sorted(per.items(), key=lambda kv: -kv[1]["total"]):
    print("| `%s` | %d | %s |" % (k, c["total"], " | ".join(str(c[x]) for x in cols)))
    tot.update(c)
print()
print("Whole library: " + ", ".join("%s %d" % (x, tot[x]) for x in cols) + "; tensor-core mnemonics: %d." % n_tensor)
