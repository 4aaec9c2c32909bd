#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel in an .ncu-rep (needs ncu on PATH)."""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
si, ns = hdr.index("Source"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[2:]):
    try:
        n = int(r[ns])
    except Exception:
        continue
    st = {hdr[i]: int(r[i]) for i in stall_cols if r[i] not in ("", "0")}
    data.append((k, n, r[si], st))
tot = sum(d[1] for d in data)
print(rows[0][1][:90], "| samples", tot, "| instructions", len(data))
agg = {}
for _, n, _, st in data:
    for k, v in st.items():
        agg[k] = agg.get(k, 0) + v
print({k: f"{100*v/max(tot,1):.0f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:6]})
for k, n, s, st in sorted(sorted(data, key=lambda d: -d[1])[:top]):
    print(f"#{k:4d} {n:5d} {100*n/tot:5.1f}%  {s[:64]:64s} {dict(sorted(st.items(), key=lambda x: -x[1])[:2])}")
