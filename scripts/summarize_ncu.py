"""Summarise gpurun_out/launches_*.csv and prof_*.ncu-rep into profiles/ (run in the build container)."""
import csv, subprocess, sys
from collections import defaultdict
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = ["# ncu evidence, round 1 (build: final state of the round; commands: scripts/gpu_final.sh)",
       "",
       "Both captures run `bench.py --steps 1 --warmup 1 --sims 150|100` (16384 games): the FIRST two moves of every game, where the trees are",
       "shallow (mean trace length 18-27 levels instead of the 65 of the default bench window, moves 4-9).  The tree kernels are therefore",
       "cheaper here than in the bench line: events inside the same early window give select+expand 16 %, k_tc_conv 60 %, k_tc_fc 11 %,",
       "backup 3 % (`phases_ms_per_step` of the bench line printed by the ncu --set full run, gpurun_out/ncu_full.log), against 15 / 70 / 11 / 3 %",
       "in the serialised, cold-cache launch list below; in the default bench window the shares are 36 / 43 / 8 / 9 % (profiles/bench_r1_final.json).",
       ""]
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

# per-launch DRAM traffic of the captured kernels -> profiles/ncu_traffic_<tag>.json (read by bench.py: roofline.traffic)
import json
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0]
    def val(metric):
        i = hdr.index(metric)
        return float(r[i].replace(",", "")), units[i]
    rd, ru = val("dram__bytes_read.sum"); wr, wu = val("dram__bytes_write.sum")
    dur, du = val("gpu__time_duration.sum")
    tp, _ = val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    traffic[name] = {"dram_bytes": rd * scale.get(ru, 1.0) + wr * scale.get(wu, 1.0), "tensor_pipe_pct": tp,
                     "duration_ms_under_ncu": dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(du, 1e-3)}
traffic["_source"] = "profiles/ncu_summary_%s.md: ncu --set full --clock-control none, bench.py --steps 1 --warmup 1 --sims 100 (16384 games), one launch per kernel (scripts/gpu_final.sh)" % tag
json.dump(traffic, open("profiles/ncu_traffic_%s.json" % tag, "w"), indent=1)
