#!/usr/bin/env python
"""tools/ncu_src.py REPORT.ncu-rep LAUNCH_INDEX [min_samples] - per-instruction stall samples of one launch (ncu --page source --csv), the
instructions with the most samples in SASS order, plus the totals per stall reason.  Needs ncu (no GPU)."""
import csv, subprocess, sys
from collections import Counter
rep, idx = sys.argv[1], int(sys.argv[2])
minS = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
secs = [i for i, r in enumerate(rows) if len(r) > isamp and r[isamp] == "# Samples"]
data = [r for r in rows[secs[0] + 1:(secs[1] - 1 if len(secs) > 1 else len(rows))] if len(r) > isamp and r[isamp].isdigit()]
st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
print(rows[0][:2], "total samples", sum(int(r[isamp]) for r in data), "instructions", len(data), "executed", sum(int(r[iex]) for r in data))
tot = Counter()
for i, r in enumerate(data):
    for k in st:
        if r[k] not in ("0", ""):
            tot[hdr[k][6:]] += int(r[k])
    if int(r[isamp]) >= minS:
        print(i, r[isamp], r[iex], {hdr[k][6:]: r[k] for k in st if r[k] not in ("0", "")}, r[ia].strip()[:90])
print(tot.most_common())
