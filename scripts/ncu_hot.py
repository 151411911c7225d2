#!/usr/bin/env python
"""Per-region instruction/stall-sample summary of an ncu source-page CSV (ncu -i rep --page source --csv --print-source sass).
Prints the instruction stream in runs of `step` instructions with executed-instruction and stall-sample sums."""
import csv, sys
f = sys.argv[1]; step = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(f)))
hdr = rows[1]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [(r[isrc].strip(), int(r[ismp] or 0), int(r[iex] or 0)) for r in rows[2:] if len(r) > iex]
tot_ex = sum(d[2] for d in data); tot_s = sum(d[1] for d in data)
print("total executed %d, samples %d, instrs %d" % (tot_ex, tot_s, len(data)))
for i in range(0, len(data), step):
    ch = data[i:i + step]
    ex = sum(d[2] for d in ch); sm = sum(d[1] for d in ch)
    if ex * 200 > tot_ex or sm * 200 > tot_s:
        print("%5d-%5d  exec %5.1f%%  samples %5.1f%%  | %s ... %s" % (i, i + len(ch), 100.0 * ex / tot_ex, 100.0 * sm / tot_s, ch[0][0][:40], ch[-1][0][:40]))
