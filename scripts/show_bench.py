import json, sys
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
print("same_workload", (d.get("same_workload") or {}).get("value_pass_vs_instrumented_vs_e2e"), "exchange", d.get("trajectory_allgather"))
print("phases", {k: round(v, 1) for k, v in d["phases_ms_per_step"].items()}, "instrumented", round(d.get("instrumented_ms_per_step", 0), 1))
print("roofline", {k: d["roofline"].get(k) for k in ("kernel", "achieved", "frac", "ms_per_launch")}, "tree", {k: d["roofline_select_backup"].get(k) for k in ("achieved", "frac", "mean_trace_len")})
print("clocks", d["clocks"], "counters", {k: int(v) for k, v in d["counters_per_step"].items() if v})
if "also_configs1_vanilla" in d:
    print("vanilla", d["also_configs1_vanilla"])
