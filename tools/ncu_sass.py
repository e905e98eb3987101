#!/usr/bin/env python
"""Per-SASS-instruction view of an ncu report: executions, stall samples, shared-memory wavefronts.
    python tools/ncu_sass.py report.ncu-rep [min_executions]"""
import csv, io, subprocess, sys
path = sys.argv[1]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ie, ss, sc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
wf = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
stalls = {n: hdr.index(n) for n in hdr if n.startswith("stall_") and "Not Issued" not in n}
for i, r in enumerate(rows[2:]):
    n = int(r[ie] or 0)
    if n < thr:
        continue
    st = sorted(((int(r[j] or 0), k[6:]) for k, j in stalls.items()), reverse=True)[:2]
    st = " ".join(f"{k}:{v}" for v, k in st if v)
    print(f"{i:5d} {n:9d} {int(r[ss] or 0):5d} {(int(r[wf] or 0) if wf is not None else 0):9d}  {r[sc].strip():70s} {st}")
