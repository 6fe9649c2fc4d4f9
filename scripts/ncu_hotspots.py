#!/usr/bin/env python
"""Top stall hot spots (SASS level, with the CUDA source line when available) of every kernel in an .ncu-rep."""
import csv, io, subprocess, sys
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 18
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1][:70]
        hdr = rows[i + 1]
        ix = {h: k for k, h in enumerate(hdr)}
        j = i + 2
        data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr):
                data.append(rows[j])
            j += 1
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
        agg = sorted(((sum(int(r[ix[c]]) for r in data), c) for c in stall_cols), reverse=True)[:5]
        print(f"== {name}  samples={tot}  " + "  ".join(f"{c[6:]}={100*v/tot:.0f}%" for v, c in agg))
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top_n]:
            st = sorted(((int(r[ix[c]]), c[6:]) for c in stall_cols if int(r[ix[c]]) > 0), reverse=True)[:2]
            print(f"   {100*int(r[ix['# Samples']])/tot:5.1f}%  {r[ix['Source']].strip()[:64]:64s} {st}")
        i = j
    else:
        i += 1
