"""Summarise gpurun_out/launches_*.csv and prof_*.ncu-rep into profiles/ (run in the build container)."""
import csv, subprocess, sys
from collections import defaultdict
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = []
rows = list(csv.reader(open("gpurun_out/launches_%s.csv" % tag)))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
t, n = defaultdict(float), defaultdict(int)
for r in rows[h + 1:]:
    if len(r) > vi:
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
        k = r[ki].split("(")[0]; t[k] += float(r[vi].replace(",", "")) * scale; n[k] += 1
tot = sum(t.values())
out.append("## launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)")
out.append("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k in sorted(t, key=t.get, reverse=True):
    out.append("| %s | %d | %.1f | %.1f | %.3f |" % (k, n[k], t[k], t[k] / n[k], t[k] / tot))
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_%s.ncu-rep" % tag, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg",
        "lts__t_sector_hit_rate.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]
tens = [x for x in hdr if "tensor" in x][:6]
out.append("\n## ncu --set full --clock-control none (one launch each)")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0]
    out.append("\n### %s" % name)
    for w in want + [x for x in tens if x not in want]:
        if w in hdr:
            out.append("- %s = %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
open("profiles/ncu_summary_%s.md" % tag, "w").write("\n".join(out) + "\n")
print("\n".join(out))
