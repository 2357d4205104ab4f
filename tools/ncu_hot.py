#!/usr/bin/env python
"""Top stall-sample SASS instructions of an ncu report (needs --import-source on / -lineinfo).
usage: tools/ncu_hot.py report.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hi]
ci = {h: i for i, h in enumerate(H)}
body = [r for r in rows[hi + 1:] if len(r) == len(H)]
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ci[s]] or 0) for r in body) for s in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for idx, r in sorted(enumerate(body), key=lambda ir: -int(ir[1][ci["# Samples"]] or 0))[:n]:
    top = sorted(((s, int(r[ci[s]] or 0)) for s in stalls), key=lambda kv: -kv[1])[:2]
    print("%5d %6.2f%% %-70s %s" % (idx, 100.0 * int(r[ci["# Samples"]] or 0) / max(tot, 1), r[ci["Source"]][:70], top))
