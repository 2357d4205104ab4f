#!/usr/bin/env python
"""Executed warp-instructions per SASS opcode of an ncu report (needs --import-source on).
usage: tools/ncu_opmix.py report.ncu-rep [N]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hi]; ci = {h: i for i, h in enumerate(H)}
body = [r for r in rows[hi + 1:] if len(r) == len(H)]
col = "# Warp Instructions Executed" if "# Warp Instructions Executed" in ci else [h for h in H if "Instructions Executed" in h][0]
agg = collections.Counter(); samp = collections.Counter()
for r in body:
    toks = r[ci["Source"]].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LD", "ST", "DMMA")) and "." in op else "")
    agg[op] += int(r[ci[col]] or 0); samp[op] += int(r[ci["# Samples"]] or 0)
tot = sum(agg.values()); ts = sum(samp.values())
print("total warp instructions", tot, "column:", col)
for op, c in agg.most_common(n):
    print("%-14s %14d %6.2f%%   samples %6.2f%%" % (op, c, 100.0 * c / tot, 100.0 * samp[op] / max(ts, 1)))
