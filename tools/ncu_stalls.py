"""Summarise the per-instruction stall samples of an ncu report (source page): python tools/ncu_stalls.py rep.ncu-rep [kernel-index] [top-n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
b = blocks[which]
hdr = b["rows"][0]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in b["rows"][1:] if len(r) > 10]
tot = sum(int(r[idx["# Samples"]]) for r in data)
print(b["name"][:100]); print("instructions", len(data), "samples", tot)
cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[idx[h]]) for r in data) for h in cols}
print("  ".join("%s %.1f%%" % (h[6:], 100.0 * v / tot) for h, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
order = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))[:topn]
for i in sorted(order):
    r = data[i]
    reasons = sorted([(int(r[idx[h]]), h[6:]) for h in cols], reverse=True)[:2]
    print("%5d %7s  %-64s %s" % (i, r[idx["# Samples"]], r[idx["Source"]].strip()[:64], reasons))
