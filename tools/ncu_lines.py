#!/usr/bin/env python
"""Stall samples per CUDA source line of an ncu report (needs -lineinfo and --import-source on).
usage: tools/ncu_lines.py report.ncu-rep [N]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.OrderedDict(); fname = None; H = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": H = r; ci = {h: i for i, h in enumerate(H)}; continue
    if H is None or len(r) != len(H) or not r[0].isdigit(): continue
    key = (fname, int(r[0]))
    a = agg.setdefault(key, {"src": r[1], "samples": 0, "inst": 0, "stalls": collections.Counter()})
    a["samples"] += int(r[ci["# Samples"]] or 0); a["inst"] += int(r[ci["Instructions Executed"]] or 0)
    for h in H:
        if h.startswith("stall_") and "Not Issued" not in h:
            a["stalls"][h] += int(r[ci[h]] or 0)
tot = sum(a["samples"] for a in agg.values()); ti = sum(a["inst"] for a in agg.values())
print("total samples", tot, "warp instructions", ti)
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:n]:
    top = ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / max(a["samples"], 1)) for k, v in a["stalls"].most_common(3))
    print("%5.2f%% inst %5.2f%% %s:%-4d %-72s [%s]" % (100.0 * a["samples"] / max(tot, 1), 100.0 * a["inst"] / max(ti, 1), f[:14], ln, a["src"].strip()[:72], top))
