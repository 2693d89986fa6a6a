"""Print the hot SASS instructions of one kernel from an `ncu --page source --csv` dump."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[hi + 1:]:
    if len(r) <= iex:
        continue
    try:
        data.append((r[ia], r[isrc], int(r[iex] or 0), int(r[ismp] or 0)))
    except ValueError:
        pass
tot = sum(d[2] for d in data)
print("total instr", tot, "samples", sum(d[3] for d in data))
mx = max(d[2] for d in data)
for i, d in enumerate(data):
    if d[2] > thr * mx:
        print(i, "%10d" % d[2], "%6d" % d[3], d[1][:110])
