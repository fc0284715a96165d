"""Top stall sites of one launch from `ncu --page source --csv --print-source sass` output.
    python scripts/ncu_top_stalls.py file.csv [top] [context]"""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 15; ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(open(path)))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"] + [len(rows)]
hdr = rows[heads[0]]
si = hdr.index("# Samples")
secs = [[r for r in rows[a + 1:b] if len(r) == len(hdr) and r[0] != "Address"] for a, b in zip(heads[:-1], heads[1:])]
body = max(secs, key=lambda sec: sum(int(r[si] or 0) for r in sec))     # the function that actually ran
print("sections", len(secs))
si = hdr.index("# Samples"); src = hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:top]
for i in order:
    r = body[i]
    st = sorted(((int(r[c] or 0), hdr[c]) for c in stall_cols), reverse=True)[:2]
    print("%5d %5.1f%%  #%d %-70s %s" % (int(r[si]), 100.0 * int(r[si]) / tot, i, r[src].strip()[:70], st))
    for j in range(max(0, i - ctx), i):
        print("              .. #%d %s" % (j, body[j][src].strip()[:90]))
