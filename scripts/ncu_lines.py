#!/usr/bin/env python
"""Per-source-line stall samples of an .ncu-rep captured with --import-source on (kernels compiled with -lineinfo)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != "":   # a source line row (sass rows have an empty line number)
        ix = {h: i for i, h in enumerate(hdr)}      # duplicate "Source" header: the later one wins, line rows carry the text in col 1
        try: smp = int(r[ix["# Samples"]] or 0)
        except ValueError: continue
        if smp:
            st = sorted(((int(r[i]), h[6:]) for h, i in ix.items() if h.startswith("stall_") and "Not Issued" not in h and r[i] not in ("", "0")), reverse=True)[:3]
            out.append((smp, (fname or "").split("/")[-1], r[0], r[1].strip()[:120], st))
tot = sum(o[0] for o in out) or 1
print("total samples", tot)
for smp, f, ln, src, st in sorted(out, key=lambda o: -o[0])[:top]:
    print(f"{100*smp/tot:5.1f}%  {f}:{ln:>4s}  {src:120s} {st}")
