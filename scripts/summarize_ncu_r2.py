"""Summarise gpurun_out/launches_r2.csv and prof_r2.ncu-rep (scripts/gpu_profile_r2.sh) into profiles/ncu_summary_r2.md and
profiles/ncu_traffic_r2.json (read by bench.py: roofline.traffic).  Run in the build container (ncu -i reads the report here)."""
import csv
import json
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
regime = ("bench window: bench.py --steps 1 --warmup 5 (BASELINE configs[2]: 16384 games, 500 sims/move, max_nodes 16384, net_tc, head-room collection), "
          "kernels of move 5 of the `value` pass")
out = ["# ncu evidence, round 2, build tag %s (commands: scripts/gpu_profile_r2.sh / gpu_profile_r2b.sh; r2b = the round-end build with the path cache on)" % tag, "", "Regime: " + regime + ".", ""]
rows = list(csv.reader(open("gpurun_out/launches_%s.csv" % tag)))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
t, n = defaultdict(float), defaultdict(int)
for r in rows[h + 1:]:
    if len(r) > vi:
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
        k = r[ki].split("(")[0]
        t[k] += float(r[vi].replace(",", "")) * scale
        n[k] += 1
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
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic, seen = {}, defaultdict(int)
out.append("\n## ncu --set full --clock-control none (two consecutive simulation steps)")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0]
    seen[name] += 1
    out.append("\n### %s (launch %d)" % (name, seen[name]))
    for w in want:
        if w in hdr:
            out.append("- %s = %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))

    def val(metric):
        i = hdr.index(metric)
        return float(r[i].replace(",", "")), units[i]
    rd, ru = val("dram__bytes_read.sum")
    wr, wu = val("dram__bytes_write.sum")
    dur, du = val("gpu__time_duration.sum")
    tp, _ = val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    rec = traffic.setdefault(name, {"dram_bytes": 0.0, "tensor_pipe_pct": 0.0, "duration_ms_under_ncu": 0.0, "launches": 0})
    rec["dram_bytes"] += rd * scale.get(ru, 1.0) + wr * scale.get(wu, 1.0)
    rec["tensor_pipe_pct"] += tp
    rec["duration_ms_under_ncu"] += dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(du, 1e-3)
    rec["launches"] += 1
for rec in traffic.values():
    for k in ("dram_bytes", "tensor_pipe_pct", "duration_ms_under_ncu"):
        rec[k] /= rec["launches"]
traffic["_regime"] = regime
traffic["_source"] = "profiles/ncu_summary_%s.md (scripts/gpu_profile_r2.sh)" % tag
open("profiles/ncu_summary_%s.md" % tag, "w").write("\n".join(out) + "\n")
json.dump(traffic, open("profiles/ncu_traffic_%s.json" % tag, "w"), indent=1)
print("\n".join(out))
