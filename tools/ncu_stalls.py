#!/usr/bin/env python3
"""Top stall sites of a kernel from an .ncu-rep source page:  python tools/ncu_stalls.py rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
he = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))   # first kernel only
hdr, data = rows[hi], [r for r in rows[hi + 1:he] if len(r) == len(rows[hi])]
iS, isrc, iex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS] or 0) for r in data)
agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stall}
print("total samples", tot, " warp-instructions", sum(int(r[iex] or 0) for r in data))
print(sorted(agg.items(), key=lambda x: -x[1])[:8])
for k, r in enumerate(data):
    r.append(k)
for r in sorted(data, key=lambda r: -int(r[iS] or 0))[:N]:
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall), reverse=True)[:2]
    print(f"{int(r[iS]):8d} {100.0*int(r[iS])/tot:5.1f}% line {r[-1]:5d} ex {r[iex]:>9s}  {r[isrc][:70]:70s} {st}")
