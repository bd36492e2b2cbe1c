"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py [tag]     (needs gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep;
                                                  default tag r1_s2 = round 1, second session)
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1_s2"

shutil.copy(os.path.join(SRC, "launches_%s.csv" % TAG), os.path.join(OUT, "%s_launches.csv" % TAG))
rows = list(csv.reader(open(os.path.join(OUT, "%s_launches.csv" % TAG))))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
        continue
    agg.setdefault(r[kn].split("(")[0], []).append(float(r[mv].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
ours = sum(sum(v) for k, v in agg.items() if "ttx::" in k)
mma = sum(sum(v) for k, v in agg.items() if "joint_mma_kernel" in k or "joint_bwd_pair" in k or "joint_quad" in k)
lines = ["| kernel | launches | total ms | share |", "|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    if sum(v) / tot >= 0.002:
        lines.append("| `%s` | %d | %.3f | %.1f%% |" % (k[:80], len(v), sum(v) / 1e6, 100 * sum(v) / tot))

raw = subprocess.run(["ncu", "-i", os.path.join(SRC, "prof_%s.ncu-rep" % TAG), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("gpu__time_duration.sum", "duration"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active"),
        ("smsp__cycles_elapsed.avg.per_second", "SM clock"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__t_sector_hit_rate.pct", "L2 hit"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__grid_size", "grid"), ("launch__cluster_dim_x", "cluster"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts")]


def short(name):
    return name.split("(")[0].replace("void ttx::", "").replace("(int)", "").replace("(bool)", "")


tbl = ["| metric | " + " | ".join(short(d[hdr.index("Kernel Name")]) for d in data) + " |", "|---|" + "---|" * len(data)]
for key, label in want:
    if key in hdr:
        i = hdr.index(key)
        tbl.append("| %s (%s) | " % (label, units[i]) + " | ".join(d[i] for d in data) + " |")
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
names = {"joint_mma_kernel<0": "ttx_joint_lse_fwd", "joint_bwd_pair_kernel<3": "ttx_joint_fwd_grad",
         "joint_bwd_pair_kernel<1": "ttx_joint_grad[dA]", "joint_bwd_pair_kernel<2": "ttx_joint_grad[dW]",
         "joint_quad_kernel<1": "ttx_joint_grad[dA]", "joint_quad_kernel<2": "ttx_joint_grad[dW]"}
traffic = {}
for d in data:
    k = d[hdr.index("Kernel Name")].replace("(int)", "").replace("(bool)", "").replace(" ", "")
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    for pat, nm in names.items():
        if pat in k:
            traffic[nm] = float(d[ir]) * mult[units[ir]] + float(d[iw]) * mult[units[iw]]
json.dump(traffic, open(os.path.join(OUT, "traffic_bytes.json"), "w"), indent=1)

open(os.path.join(OUT, "%s_tables.md" % TAG), "w").write(
    "## Launch list (`%s_launches.csv`; cold-cache, serialised: compare shares)\n\n" % TAG +
    "Our kernels (`ttx::*`) are %.1f%% of the device time in the list; the tcgen05 kernels alone %.1f%%.\n\n%s\n\n"
    "## Full capture of the tensor-core kernels (one launch each, `--set full`)\n\n%s\n"
    % (100 * ours / tot, 100 * mma / tot, "\n".join(lines), "\n".join(tbl)))
print(open(os.path.join(OUT, "%s_tables.md" % TAG)).read())
print(traffic)
