"""SASS instruction counts per kernel of lib/libttx.so (runs anywhere: cuobjdump only).  python tools/sass_counts.py > profiles/r2_sass_counts.md"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "transformer-transducer_b200", "lib", "libttx.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "MUFU", "FFMA2", "HMMA", "REDG", "SYNCS"]
rows, cur = collections.OrderedDict(), None
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"^void ", "", name).split("(")[0].replace("ttx::", "").replace("(bool)0", "false").replace("(bool)1", "true")
        name = re.sub(r"\(int\)", "", name)
        cur = rows.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        base = op.split(".")[0]
        if base == "UTCHMMA":
            cur["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
        elif base in cols:
            cur[base] += 1
print("# SASS instruction counts per kernel of lib/libttx.so (cuobjdump -sass, sm_100a), final tree of round 2 (tools/sass_counts.py)\n")
print("UTC*MMA = tcgen05.mma (UTCHMMA: kind::f16 / tf32), UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk,\n"
      "LDTM / STTM = tcgen05.ld / st, SYNCS = mbarrier operations.  No HMMA (legacy mma.sync) anywhere.\n")
print("| kernel | " + " | ".join(cols) + " | total |")
print("|---|" + "---|" * (len(cols) + 1))
tot = collections.Counter()
for name, c in rows.items():
    print("| `%s` | " % name + " | ".join(str(c[k]) for k in cols) + " | %d |" % c["total"])
    tot.update(c)
print("| **all kernels** | " + " | ".join(str(tot[k]) for k in cols) + " | %d |" % tot["total"])
