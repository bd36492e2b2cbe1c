"""profiles/r2_tables.md from the ncu --set full reports brought back in gpurun_out/ (raw page, selected metrics).
    python tools/r2_tables.py gpurun_out/a.ncu-rep gpurun_out/b.ncu-rep ..."""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs/thread"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
        ("lts__t_sector_hit_rate.pct", "L2 hit"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("smsp__cycles_elapsed.avg.per_second", "SM clock")]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("ttx::", "")


out = ["# ncu `--set full --clock-control none` metrics per kernel launch (round 2; one launch per column)", "",
       "Captured with `python bench.py --no-cpu-baseline --steps 2 --warmup 3` (configs[1]) under ncu on one B200; times under a",
       "profiler are cold-cache and serialised -- the bench numbers come from CUDA events, these tables explain them.", ""]
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    out += ["## `%s`" % os.path.basename(rep), "", "| metric | " + " | ".join("`%s`" % short(d[kn]) for d in data) + " |",
            "|---|" + "---|" * len(data)]
    for key, label in WANT:
        if key in hdr:
            i = hdr.index(key)
            out.append("| %s (%s) | " % (label, units[i]) + " | ".join(d[i] for d in data) + " |")
    out.append("")
open(os.path.join(ROOT, "profiles", "r2_tables.md"), "w").write("\n".join(out))
print("\n".join(out)[:3000])
