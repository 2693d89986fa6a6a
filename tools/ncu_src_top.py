"""Top sampled SASS instructions of an `ncu --page source --csv` dump (stdin or file): samples, dominant stall, SASS."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ci, si, ei = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[2:]):
    try:
        data.append((int(r[ci]), k, r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for n, k, r in sorted(data, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print("%6d %5.1f%%  #%-5d exec=%-9s %-28s %s" % (n, 100.0 * n / tot, k, r[ei], ",".join("%s:%d" % (h[6:], c) for c, h in st if c), r[si].strip()[:90]))
