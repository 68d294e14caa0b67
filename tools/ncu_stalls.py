#!/usr/bin/env python
"""Per-SASS-instruction stall reasons of an ncu report (--set full --import-source on):
top instructions by one stall reason with the instructions in front of them."""
import csv, subprocess, sys
rep, reason = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "stall_long_sb")
n = int(sys.argv[3]) if len(sys.argv) > 3 else 15
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
ins = [r for r in rows if r and r[0].startswith("0x") and len(r) >= len(hdr)]
tot = {k: sum(float(r[ix[k]] or 0) for r in ins) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
alls = sum(float(r[ix["# Samples"]] or 0) for r in ins)
print("samples", alls, {k: round(100 * v / alls, 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v / alls > 0.01})
order = sorted(range(len(ins)), key=lambda i: -float(ins[i][ix[reason]] or 0))[:n]
for i in order:
    r = ins[i]
    print("---- %s %.1f%% of all samples (%s of %s at this instruction)" % (reason, 100 * float(r[ix[reason]]) / alls, r[ix[reason]], r[ix["# Samples"]]))
    for j in range(max(0, i - 4), i + 1):
        print("   ", ins[j][ix["Source"]].strip()[:110], "| samp", ins[j][ix["# Samples"]])
