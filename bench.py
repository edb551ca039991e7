#!/usr/bin/env python
"""bench.py — MCTS simulations/second of the per-move simulation loop (select -> step/expand -> value net -> backup)
over N concurrent games, on 1..8 B200 (one process per GPU), next to the reference's CPU path.

  python bench.py --gpus 1 --steps K --warmup W            B200 arm (this repo's CUDA engine through its C-ABI)
  python bench.py --impl reference --steps K --warmup W    reference arm: the reference's own compiled
                                                            agents/cppmodule/agent.cpp (oracle/_ref) on the host cores
A "step" is one move of every game: `sims` simulations per game, then get_action, play, update_root (play.py:118-177).
Workload (BASELINE.json configs[2], the one north_star's target sentence names: >= 16384 concurrent games, ValueSimLP +
value net, 500 sims/move).  `--workload vanilla` selects configs[1] (Vanilla MCTS, 4096 games, 300 sims/move).
Prints ONE JSON line on rank 0."""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONV_FLOP = 82944 + 1769472 + 1032192          # conv1 + conv2 + conv3 per board
FC_FLOP = 917504 + 1024                          # fc1 + fc_out per board
NET_FLOP = CONV_FLOP + FC_FLOP                   # 3 803 136
ENV_ARGS = ((20, 10), 1, 0, 0)                   # play.py:75 defaults
BASE_SEED = 123                                  # SURVEY §8d (echoes agent.cpp:23)


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic_r1.json")
    try:
        return json.load(open(p)).get(kernel, {}).get("dram_bytes")
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor=1400.0, src="fallback")   # B200_PROFILING.md fallback


# ----------------------------------------------------------------------------------------------- clocks sampler
class Clocks:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.p, self.index = [], None, index

    def __enter__(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None
        return self

    def _read(self):
        for line in self.p.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 7:
                self.rows.append(f)

    def __exit__(self, *a):
        if self.p:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i] == "Active" for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows)}


# ----------------------------------------------------------------------------------------------- reference arm
def _ref_worker(conn, wid, workload, sims):
    """One single-threaded worker = one play.py-equivalent process on the reference's compiled C++ agent."""
    try:
        os.environ["OMP_NUM_THREADS"] = "1"
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import torch
        torch.set_num_threads(1)
        import oracle_py as O
        from random import randint
        pt = O.load_ref_module("pyTetris")
        agent_mod = O.load_ref_module("agent")          # the reference's own agents/cppmodule/agent.cpp, compiled unchanged
        game = pt.Tetris(*ENV_ARGS)
        game.seed(BASE_SEED + wid)
        if workload == "vanilla":                        # agents/VanillaC.py:5-13
            def random_playout(g):
                while not g.end:
                    g.play(randint(0, 7))
                return g.score, 1e5
            agent = agent_mod.MCTSAgent(sims, 500000, True, 0.99, True, random_playout, 1, False)
        else:                                            # agents/ValueSimC.py:17-42 (LP=True, evaluator = Model_VV.inference)
            from ref_net import RefModel
            model = RefModel(O.weights_to_state_dict(O.seeded_weights(0)))
            agent = agent_mod.OnlineMCTSAgent(sims=sims, max_nodes=100000, online=False, accumulation_policy=3, memory_size=1,
                                              episodes_per_train=25, memory_growth_rate=5000, min_visit=25, projection=True, gamma=0.999,
                                              benchmark=True, evaluator=model.inference, evaluation_type=0, train=(lambda *a: None), LP=True)
        agent.update_root(game)
        conn.send("ready")
        while True:
            msg = conn.recv()
            if msg[0] == "stop":
                break
            done, t0 = 0, time.perf_counter()
            if msg[0] == "moves":
                budget_moves, budget_s = msg[1], 1e18
            else:
                budget_moves, budget_s = 1 << 30, msg[1]
            while done < budget_moves and time.perf_counter() - t0 < budget_s:
                a = agent.play()
                game.play(a)
                agent.update_root(game)
                if game.end:
                    game.reset()
                    agent.update_root(game)
                done += 1
            conn.send((done * sims, time.perf_counter() - t0))
    except Exception as ex:   # noqa
        conn.send(("error", repr(ex)))


class RefPool:
    def __init__(self, workload, sims, procs=None):
        self.P = procs or os.cpu_count() or 1
        ctx = mp.get_context("spawn")
        self.conns, self.procs = [], []
        for w in range(self.P):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_ref_worker, args=(b, w, workload, sims), daemon=True)
            p.start()
            self.conns.append(a)
            self.procs.append(p)
        for c in self.conns:
            r = c.recv()
            if r != "ready":
                raise RuntimeError("reference worker failed: %r" % (r,))

    def run(self, kind, amount):
        for c in self.conns:
            c.send((kind, amount))
        res = [c.recv() for c in self.conns]
        for r in res:
            if r[0] == "error":
                raise RuntimeError("reference worker failed: %s" % r[1])
        return sum(r[0] for r in res), max(r[1] for r in res)

    def close(self):
        for c in self.conns:
            try:
                c.send(("stop",))
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=10)


def cpu_baseline(workload, sims, seconds):
    pool = RefPool(workload, sims)
    try:
        pool.run("seconds", min(2.0, seconds / 4))          # warm-up (JIT of the scripted net, caches)
        n, t = pool.run("seconds", seconds)
    finally:
        pool.close()
    return {"value": n / t, "unit": "sims/s", "cores": pool.P, "kind": "reference",
            "sample": "%d single-threaded workers x ~%.0f s of %s at %d sims/move on the reference's compiled agent.cpp + torch fp32 CPU net "
                      "(env = oracle restatement: pyTetris is absent upstream); %d sims total" % (pool.P, seconds, workload, sims, n)}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    sims = cfg["sims"]
    pool = RefPool(cfg["workload_key"], sims)
    try:
        moves_per_step = args.ref_moves_per_step
        for _ in range(args.warmup):
            pool.run("moves", moves_per_step)
        tot_n, tot_t = 0, 0.0
        for _ in range(args.steps):
            n, t = pool.run("moves", moves_per_step)
            tot_n += n
            tot_t += t
    finally:
        pool.close()
    value = tot_n / tot_t
    sample = "%d workers x %d moves x %d sims per step" % (pool.P, moves_per_step, sims)
    line = {"impl": "reference", "metric": "mcts_sims_per_sec", "value": value, "unit": "sims/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg["config"],
            "cpu_baseline": {"value": value, "unit": "sims/s", "cores": pool.P, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "reference_arm": "the reference's own compiled agents/cppmodule/agent.cpp (one game per single-threaded worker process, max_nodes and "
                             "collection as the reference does them) + torch fp32 CPU value net; `config` is the B200 arm's, for the ratio"}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- B200 arm
def run_b200(args, cfg):
    import torch
    from tetris_mcts_b200 import distributed as D
    from tetris_mcts_b200 import pyTetris as PT
    from tetris_mcts_b200.engine import BatchedEngine
    from tetris_mcts_b200.model.model_vv import init_weights

    rank, local_rank, world = D.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    G, sims, M = cfg["games_per_gpu"], cfg["sims"], cfg["max_nodes"]
    seeds = D.shard_seeds(BASE_SEED, G * world, rank, world)
    recs = PT.new_games(G, ENV_ARGS, seeds)
    dist_w = None
    if cfg["mode"] == "dist":
        from tetris_mcts_b200.agents.DistValueSimOnline import init_dist_weights
        dist_w = init_dist_weights(0, 50)
    eng = BatchedEngine(G, max_nodes=M, mode=cfg["mode"], eval_kind=cfg["eval"], weights=init_weights(0) if cfg["mode"] in ("lp", "single") else None,
                        dist_weights=dist_w, env_args=ENV_ARGS, seed=BASE_SEED + 7919 * rank, device=local_rank, rollout_variance=1e3, overflow_reset=True)
    eng.set_games(recs)
    eng.set_gc_headroom(cfg["gc_headroom"])
    for _ in range(args.warmup):
        eng.play_move(sims, auto_reset=True, want_stats=False)
    eng.sync()
    # ---- device-timed region: inputs resident in HBM, no host buffers; the production path (one captured CUDA graph per
    # simulation step, no per-kernel events), timed with CUDA events on the engine's own stream
    c0 = eng.counters()
    D.barrier()
    torch.cuda.synchronize()
    with Clocks(local_rank) as clk:
        eng.timer_start()
        for _ in range(args.steps):
            eng.play_move(sims, auto_reset=True, want_stats=False)
        ms = eng.timer_stop()
    torch.cuda.synchronize()
    D.barrier()
    c1 = eng.counters()
    ms_max = D.max_over_ranks(ms, dev)
    delta = {k: c1[k] - c0[k] for k in c1}
    tot = D.sum_over_ranks(delta, dev)
    value = tot["sims"] / (ms_max / 1e3)
    # ---- instrumented pass: the same K steps again with a CUDA event pair around every kernel (direct launches), for the
    # per-kernel launch durations the roofline figures are computed from.  Not part of `value`.
    eng.set_timing(True)
    p0 = eng.counters()
    eng.timer_start()
    for _ in range(args.steps):
        eng.play_move(sims, auto_reset=True, want_stats=False)
    ms_instr = eng.timer_stop()
    phases = eng.phase_ms()
    eng.set_timing(False)
    p1 = eng.counters()
    pdelta = {k: p1[k] - p0[k] for k in p1}
    # ---- end-to-end region: the public API with HOST buffers, H2D of the games and D2H of the results every step
    pin_recs = torch.empty((G, 20), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    pin_recs[:] = eng.get_games()
    e0 = eng.counters()
    D.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.set_games(pin_recs)                                   # H2D G*80 B + update_root
        actions, stats = eng.play_move(sims, auto_reset=True)     # D2H G*(4+84) B
        pin_recs[:] = eng.get_games()                             # D2H G*80 B
    eng.sync()
    torch.cuda.synchronize()
    e2e_s = D.max_over_ranks(time.perf_counter() - t0, dev)
    D.barrier()
    e1 = eng.counters()
    e2e_sims = D.sum_over_ranks({"sims": e1["sims"] - e0["sims"]}, dev)["sims"]
    # ---- the one exchange step: all-gather of fixed-size replay-sample blocks (212-byte rows), outside the timed regions
    cap = 65536
    block = torch.zeros((cap, D.SAMPLE_BYTES), dtype=torch.uint8, device=dev)
    n_local = eng.collect_samples_into(block.data_ptr(), cap, 25)          # ValueSimLP.py:11 min_visits_to_store=25
    torch.cuda.synchronize()
    D.barrier()
    tg = time.perf_counter()
    rows, counts = D.allgather_samples(block, n_local)
    torch.cuda.synchronize()
    tg = D.max_over_ranks(time.perf_counter() - tg, dev)
    traj = {"samples_total": int(rows.shape[0]), "samples_per_rank": counts, "block_bytes_per_rank": cap * D.SAMPLE_BYTES,
            "ms": tg * 1e3, "backend": "nccl" if world > 1 else "none"}
    launches = sum(n for _, n in phases.values())
    peaks = measured_peaks()
    out = None
    if rank == 0:
        steps = max(args.steps, 1)
        roof, roof_tree = None, None
        D_mean = pdelta["trace_levels"] / max(pdelta["sims"], 1)
        sel_ms, sel_n = phases["select_expand"]
        bk_ms, bk_n = phases["backup"]
        tree_bytes = pdelta["sims"] * (200.0 * D_mean - 140.0) + pdelta["expansions"] * 1100.0
        tree_s = (sel_ms + bk_ms) / 1e3
        if tree_s > 0:
            ach = tree_bytes / tree_s / 1e9
            tr_s, tr_b = ncu_traffic("k_select_expand"), ncu_traffic("k_backup")
            roof_tree = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                         "traffic": (tr_s + tr_b) if tr_s and tr_b else None, "algorithmic_bytes_per_launch_pair": tree_bytes / max(sel_n, 1),
                         "kernels": "k_select_expand + k_backup", "mean_trace_len": D_mean, "ms_per_launch_pair": (sel_ms + bk_ms) / max(sel_n, 1),
                         "peak_src": peaks["src"]}
        if cfg["mode"] == "dist":
            conv_ms, conv_n = phases["conv"]
            flop = 121856 + 16 * 4 * 32 * 512 * 2          # conv1 + conv2 on the 22x10 input of model_distributional.py:27
            ach = pdelta["eval_requests"] * flop / max(conv_ms / 1e3, 1e-9) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": ach / peaks["tensor"], "traffic": None,
                    "kernel": "k_dn_conv", "ms_per_launch": conv_ms / max(conv_n, 1), "share_of_step": conv_ms / ms_instr,
                    "note": "fp32 CUDA-core kernel (the distributional head is not on tensor cores yet); bf16 peak shown as the driver-measured denominator"}
        elif cfg["mode"] != "vanilla":
            conv_ms, conv_n = phases["conv"]
            fc_ms, fc_n = phases["fc"]
            boards = pdelta["eval_requests"]
            if conv_ms > 0 and conv_n > 0:
                ach = boards * CONV_FLOP / (conv_ms / 1e3) / 1e12
                kname = "k_vn_conv" if cfg["eval"] == "net" else "k_tc_conv"
                roof = {"bound": "tensor", "achieved": ach, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": ach / peaks["tensor"],
                        "traffic": ncu_traffic(kname), "kernel": kname, "mma_flops_issued_per_launch": boards * 2 * 128 * 16 * (96 * 36 + 64 * 2) / conv_n, "ms_per_launch": conv_ms / conv_n,
                        "flops_per_launch": boards * CONV_FLOP / conv_n, "boards_per_launch": boards / conv_n,
                        "fc_kernel_tflops": boards * FC_FLOP / (fc_ms / 1e3) / 1e12 if fc_ms > 0 else None,
                        "share_of_step": conv_ms / ms_instr, "peak_src": peaks["src"] + " bf16 dense, sustained",
                        "note": "achieved counts ALGORITHMIC conv FLOPs (SURVEY 8d: 2 884 608 per board).  fp32-faithful arithmetic (north_star 1e-5): "
                                "eval=net is CUDA-core fp32 FMA; eval=net_tc is tcgen05 kind::f16 with every fp32 operand split into two scaled fp16 terms "
                                "(3 products per algorithmic product; M=128 pixel tiles on an 8-wide grid carry 25-56% halo rows; per board 2x18 MMAs of 128x96x16 "
                                "plus 2 of 128x64x16 for conv1), so the tensor pipe executes mma_flops_issued_per_launch.  scripts/probe/mma_probe.cu "
                                "measures SS-mode tcgen05.mma at (A+B operand bytes)/128 B/clk with a 44.7 clk floor: these small-N MMAs are shared-memory "
                                "operand-fetch bound, not tensor-rate bound; the bf16 peak is the driver-measured denominator, not this kernel's ceiling"}
        else:
            ro_ms, ro_n = phases["rollout"]
            roof = {"bound": "hbm", "achieved": 0.0, "peak": peaks["hbm"], "unit": "GB/s", "frac": 0.0, "traffic": None, "kernel": "k_rollout",
                    "ms_per_launch": ro_ms / max(ro_n, 1), "note": "rollouts are integer-issue bound in registers: 0 algorithmic HBM bytes "
                    "(SURVEY §8d); board steps/s = %.3g" % (pdelta["rollout_steps"] / max(ro_ms / 1e3, 1e-9))}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu = cpu_baseline(cfg["workload_key"], sims, args.cpu_seconds)
            except Exception as ex:   # the oracle/_ref modules are prebuilt; report rather than die
                cpu = {"value": None, "unit": "sims/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (ex,)}
        config = dict(cfg["config"])
        config.update({"parallelism": "games sharded x%d, no data-path collective" % world,
                       "l2": "inputs larger than L2: %.1f GB of arenas per GPU; %.2f GB of activations stream through L2 every sim-step"
                             % (G * M * 324 / 1e9, G * 7 * 1792 * 4 / 1e9)})
        out = {"metric": "mcts_sims_per_sec", "value": value, "unit": "sims/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": config, "roofline": roof, "roofline_select_backup": roof_tree, "cpu_baseline": cpu,
               "e2e": {"value": e2e_sims / e2e_s, "unit": "sims/s", "h2d_bytes_per_step": G * 80 * world, "d2h_bytes_per_step": G * (4 + 84 + 80) * world},
               "gpu_launches": int(launches), "clocks": clk.summary(), "trajectory_allgather": traj,
               "phases_ms_per_step": {k: v[0] / steps for k, v in phases.items()}, "instrumented_ms_per_step": ms_instr / steps,
               "phases_note": "value / ms_per_step: K steps on the production path (each simulation step replayed as one CUDA graph).  phases_ms_per_step, "
                              "roofline.*: a second pass of K steps with an event pair around every kernel (direct launches), instrumented_ms_per_step long",
               "counters_per_step": {k: v / steps for k, v in delta.items()}}
    eng.close()
    # ---- BASELINE configs[1] alongside (Vanilla MCTS, random rollouts, no value net): short, so it rides in the same line
    if cfg["mode"] == "lp" and not args.no_secondary:
        G2, sims2 = 4096, 300
        e2 = BatchedEngine(G2, max_nodes=8192, mode="vanilla", eval_kind="synthetic", env_args=ENV_ARGS, seed=BASE_SEED + 7919 * rank,
                           device=local_rank, rollout_variance=1e3, overflow_reset=True)
        e2.set_games(PT.new_games(G2, ENV_ARGS, D.shard_seeds(BASE_SEED, G2 * world, rank, world)))
        e2.set_gc_headroom(8192 * 5 // 32)
        for _ in range(max(args.warmup, 3)):
            e2.play_move(sims2, auto_reset=True, want_stats=False)
        k0 = e2.counters()
        D.barrier()
        torch.cuda.synchronize()
        e2.timer_start()
        for _ in range(max(args.steps, 1)):
            e2.play_move(sims2, auto_reset=True, want_stats=False)
        ms2 = D.max_over_ranks(e2.timer_stop(), dev)
        k1 = e2.counters()
        d2 = D.sum_over_ranks({k: k1[k] - k0[k] for k in k1}, dev)
        e2.close()
        if out is not None:
            out["also_configs1_vanilla"] = {"workload": "BASELINE configs[1]: Vanilla MCTS, %d games/GPU, %d sims/move" % (G2, sims2), "value": d2["sims"] / (ms2 / 1e3),
                                            "unit": "sims/s", "ms_per_step": ms2 / max(args.steps, 1), "rollout_steps_per_sim": d2["rollout_steps"] / max(d2["sims"], 1),
                                            "mean_trace_len": d2["trace_levels"] / max(d2["sims"], 1)}
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="valuesimlp", choices=["valuesimlp", "vanilla", "dist"])
    ap.add_argument("--games-per-gpu", type=int, default=None)
    ap.add_argument("--sims", type=int, default=None)
    ap.add_argument("--max-nodes", type=int, default=None)
    ap.add_argument("--gc-headroom", type=int, default=None, help="collect between moves every game with fewer free slots (default 5/32 of max_nodes; 0 = lazy collection only)")
    ap.add_argument("--eval", default=os.environ.get("B200_EVAL", "net_tc"), choices=["net", "net_tc", "synthetic"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--ref-moves-per-step", type=int, default=4)
    args = ap.parse_args()
    if args.workload == "vanilla":
        G, sims, M, mode = args.games_per_gpu or 4096, args.sims or 300, args.max_nodes or 8192, "vanilla"
        name = "BASELINE configs[1]: Vanilla MCTS (random rollouts, no value net), %d games/GPU, %d sims/move" % (G, sims)
    elif args.workload == "dist":
        G, sims, M, mode = args.games_per_gpu or 2048, args.sims or 1500, args.max_nodes or 32768, "dist"
        name = "BASELINE configs[4]: distributional head (agents/core_distributional.py), %d games/GPU (16384 over 8 GPUs), %d sims/move" % (G, sims)
        if args.eval == "net_tc":
            args.eval = "net"
    else:
        G, sims, M, mode = args.games_per_gpu or 16384, args.sims or 500, args.max_nodes or 16384, "lp"
        name = "BASELINE configs[2]: ValueSimLP + value net, %d games/GPU, %d sims/move" % (G, sims)
    headroom = args.gc_headroom if args.gc_headroom is not None else M * 5 // 32
    cfg = dict(games_per_gpu=G, sims=sims, max_nodes=M, mode=mode, eval=args.eval, workload_key=args.workload, gc_headroom=headroom,
               config={"workload": name, "games_per_gpu": G, "sims_per_move": sims, "max_nodes": M, "evaluator": args.eval if mode != "vanilla" else "rollout",
                       "env_args": "((20,10),1,0,0)", "weights": "default-init distribution, numpy PCG64 seed 0",
                       "garbage_collection": "TreeAgent.remove_nodes() between moves for every game with fewer than %d free slots, batched over the games "
                                             "(b200_set_gc_headroom; reference equivalent: the driver calling agent.remove_nodes(), agents/agent.py:246-257), "
                                             "plus the reference's own call inside new_node when a free list runs dry in the middle of a move" % headroom,
                       "arena_overflow": "reference semantics up to max_nodes per game; a game whose reachable set fills its arena (a collection recovers fewer "
                                         "than max_nodes/8 slots; the reference then collects at nearly every expansion and dies with IndexError) drops its tree "
                                         "and re-roots (counters_per_step.tree_resets)"})
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_b200(args, cfg)


if __name__ == "__main__":
    main()
