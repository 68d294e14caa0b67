#!/usr/bin/env python
"""Per source line: samples and the stall reasons behind them (ncu --set full --import-source on report)."""
import csv, subprocess, sys, collections, os
rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = cur_line = hdr = None
agg = collections.defaultdict(lambda: collections.Counter())
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] == 'Function Name': continue
    if r[0].isdigit(): cur_line = int(r[0]); continue
    if r[0] == '' and hdr and len(r) > 8 and r[2].startswith('0x'):
        off = len(r) - len(hdr)
        for i, h in enumerate(hdr):
            if h.startswith('stall_') and 'Not Issued' not in h:
                try: agg[(cur_file, cur_line)][h[6:]] += float(r[i + off] or 0)
                except ValueError: pass
tot = sum(sum(c.values()) for c in agg.values())
src = {}
for fn in ('ob_decode_pipe.cu', 'ob_decode_tile.cuh', 'ob_ptx.cuh', 'ob_cloud.cu'):
    p = '/root/repo/ouster-sdk_b200/csrc/' + fn
    if os.path.exists(p): src[fn] = open(p).read().split('\n')
allr = collections.Counter()
for c in agg.values(): allr.update(c)
print("total", tot, {k: round(100 * v / tot, 1) for k, v in allr.most_common(9)})
for (fn, ln), c in sorted(agg.items(), key=lambda kv: -sum(kv[1].values()))[:N]:
    s = sum(c.values())
    line = src[fn][ln - 1].strip()[:70] if fn in src and ln - 1 < len(src[fn]) else ''
    print("%-18s %4d %5.1f%%  %-44s %s" % (fn[:18], ln, 100 * s / tot, " ".join("%s:%d" % (k, 100 * v / s) for k, v in c.most_common(3)), line))
