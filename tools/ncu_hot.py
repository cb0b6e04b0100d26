#!/usr/bin/env python
"""Top stalled SASS instructions of an .ncu-rep (source page): python tools/ncu_hot.py rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ci["# Samples"]]) for r in body)
print(f"total samples {tot}, instructions {len(body)}")
agg = {h: sum(int(r[ci[h]]) for r in body) for h in stall_cols}
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for idx, r in sorted(enumerate(body), key=lambda t: -int(t[1][ci["# Samples"]]))[:top]:
    st = {h[6:]: int(r[ci[h]]) for h in stall_cols if int(r[ci[h]])}
    main = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{idx:4d} {int(r[ci['# Samples']]):6d} {100*int(r[ci['# Samples']])/tot:5.1f}%  {r[ci['Source']].strip():60s} {main}")
