// kernels.cuh — the __global__ entry points of the engine (sm_100a).  Launch geometry: one 8-lane group per game,
// 128-thread blocks (16 games per block), so G games = G/16 blocks; at G >= 16384 that is >= 1024 blocks = ~7 CTAs
// per SM on 148 SMs.  The tree walk is pointer chasing (latency bound); occupancy, not ILP, hides HBM latency here.
#pragma once
#include "search_dev.cuh"

namespace b200 {

constexpr int GROUPS_PER_BLOCK = 16;
constexpr int TPB = GROUPS_PER_BLOCK * 8;

__device__ __forceinline__ void load_rec(const uint32_t *src, uint32_t (&w)[REC_WORDS]) {
#pragma unroll
    for (int q = 0; q < REC_WORDS / 4; ++q) {
        uint4 v = reinterpret_cast<const uint4 *>(src)[q];
        w[q * 4] = v.x; w[q * 4 + 1] = v.y; w[q * 4 + 2] = v.z; w[q * 4 + 3] = v.w;
    }
}
__device__ __forceinline__ void store_rec(uint32_t *dst, const uint32_t (&w)[REC_WORDS]) {
#pragma unroll
    for (int q = 0; q < REC_WORDS / 4; ++q)
        reinterpret_cast<uint4 *>(dst)[q] = make_uint4(w[q * 4], w[q * 4 + 1], w[q * 4 + 2], w[q * 4 + 3]);
}

// ---------------------------------------------------------------- arena initialisation (agent.py:58-88)
__global__ void k_init_arena(Arena A, const uint32_t *default_rec, uint32_t seed) {
    size_t n = (size_t)A.G * A.M;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int g = (int)(i / A.M), idx = (int)(i % A.M);
        uint32_t w[REC_WORDS];
#pragma unroll
        for (int q = 0; q < REC_WORDS; ++q) w[q] = default_rec[q];      // agent.py:70 every slot starts as env(*env_args)
        store_rec(A.rec + i * REC_WORDS, w);
        if (idx < A.M - 1) { A.nfree[i] = idx + 1; A.ofree[i] = idx + 1; }   // agent.py:72 deque(range(1, M))
        if (idx == 0) {
            A.n_nfree[g] = A.M - 1; A.n_ofree[g] = A.M - 1;
            A.root[g] = 0; A.episode[g] = 0; A.status[g] = ST_OK;
            uint32_t s = seed + 0x9E3779B9u * (uint32_t)(g + 1);
            A.srng[g] = s ? s : 0x2545F491u;
        }
    }
}

// ---------------------------------------------------------------- batched board step (SPEC §3, play.py:150)
__global__ void k_env_step(uint32_t *recs, const int32_t *actions, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[REC_WORDS];
    load_rec(recs + (size_t)i * REC_WORDS, w);
    Game g;
    unpack(g, w);
    play(g, actions[i]);
    pack(g, w);
    store_rec(recs + (size_t)i * REC_WORDS, w);
}

// observation int8[20,10] of packed games (SPEC §1; agent.py:116 getState())
__global__ void k_env_state(const uint32_t *recs, int8_t *out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[REC_WORDS];
    load_rec(recs + (size_t)i * REC_WORDS, w);
    Game g;
    unpack(g, w);
    uint32_t k[KEY_WORDS];
    obskey(g, k);
    int8_t *o = out + (size_t)i * 200;
    for (int r = 0; r < 20; ++r) {
        uint32_t row = (k[r >> 1] >> ((r & 1) * 16)) & 0x3ffu;
        for (int c = 0; c < 10; ++c) o[r * 10 + c] = (int8_t)((row >> c) & 1u);
    }
    for (int j = 0; j < 4; ++j) o[(k[10] >> (8 * j)) & 0xffu] = -1;
}

__device__ __forceinline__ void reset_game(Game &g) {   // SPEC §4 reset(): keeps the RNG stream
#pragma unroll
    for (int i = 0; i < 10; ++i) g.w[i] = 0;
    g.bag = 0x7fu; g.dropcnt = 0; g.end = 0; g.combo = 0; g.score = 0; g.lines = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) g.ls[i] = 0;
    spawn(g);
}

// ---------------------------------------------------------------- update_root (agent.py:296-301)
// auto_reset reproduces play.py:161-177: a finished game is counted, reset and re-rooted.
__global__ void __launch_bounds__(TPB) k_update_root(Arena A, int auto_reset, unsigned long long *game_stats) {
    Grp gp;
    int g = blockIdx.x * GROUPS_PER_BLOCK + (threadIdx.x >> 3);
    if (g >= A.G) return;
    int status = A.status[g];
    if (status == ST_ARENA_FULL && A.overflow_reset) reset_tree(A, gp, g, status);   // re-roots at the live game
    if (status != ST_OK) return;
    uint32_t w[REC_WORDS];
    load_rec(A.cur + (size_t)g * REC_WORDS, w);
    int o; float s;
    int idx = new_node(A, gp, g, w, o, s, status);
    bool ended = (w[10] >> 21) & 1u;
    if (gp.lane == 0) { A.root[g] = idx; if (ended) A.episode[g] += 1; }
    gp.sync();
    if (ended && auto_reset && status == ST_OK) {
        Game gm;
        unpack(gm, w);
        if (gp.lane == 0 && game_stats) {
            atomicAdd(&game_stats[0], 1ull);
            atomicAdd(&game_stats[1], (unsigned long long)gm.score);
            atomicAdd(&game_stats[2], (unsigned long long)gm.lines);
        }
        reset_game(gm);
        pack(gm, w);
        if (gp.lane == 0) store_rec(A.cur + (size_t)g * REC_WORDS, w);
        gp.sync();
        idx = new_node(A, gp, g, w, o, s, status);
        if (gp.lane == 0) A.root[g] = idx;
    }
    if (gp.lane == 0 && status != ST_OK) A.status[g] = status;
}

// ---------------------------------------------------------------- select + expand (ValueSimLP.py:45-57 etc.)
__global__ void __launch_bounds__(TPB) k_select_expand(Arena A) {
    Grp gp;
    int g = blockIdx.x * GROUPS_PER_BLOCK + (threadIdx.x >> 3);
    if (g >= A.G) return;
    int status = A.status[g];
    if (status == ST_ARENA_FULL && A.overflow_reset) reset_tree(A, gp, g, status);
    if (status != ST_OK) return;
    ArenaAcc acc{A, g};
    int D = 0;
    int leaf = select_trace(acc, gp, A.root[g], A.low, A.trace_max, D, status);
    if (status != ST_OK) { if (gp.lane == 0) A.status[g] = status; return; }
    uint32_t w[REC_WORDS];
    load_rec(A.rec + node_at(A, g, leaf) * REC_WORDS, w);
    bool ended = (w[10] >> 21) & 1u;
    int kind = ended ? LEAF_TERMINAL : LEAF_EXPANDED;
    unsigned need = 0;   // lanes that ask for an evaluation
    int my_o = 0;
    if (!ended) {
        if (A.mode == MODE_SINGLE) {          // ValueSim.py:83-88 evaluates the leaf itself, before expanding
            my_o = A.row[node_at(A, g, leaf) * ROW_WORDS + 15];
            need = 1u << 7;
        }
        int c, o; float s;
        expand_leaf(A, gp, g, leaf, w, c, o, s, status);
        if (status == ST_OK && A.mode == MODE_LP) {
            Uniq u = unique_children(gp, c, o, s);
            // ValueSimLP.py:55-60 evaluates every unique child; core.h:344 only uses results where visit == 0,
            // so only those boards are sent to the network
            bool ask = u.is_first && A.stat[node_at(A, g, o)].x == 0;
            need = gp.ballot(ask);
            my_o = o;
        }
    }
    if (gp.lane == 0) {
        A.trace_len[g] = D; A.leaf_kind[g] = kind;
        if (status != ST_OK) A.status[g] = status;
    }
    // compact the evaluation requests: one atomic per warp
    bool mine = status == ST_OK && ((need >> gp.lane) & 1u);
    unsigned wm = __ballot_sync(0xffffffffu, mine);
    int lane32 = threadIdx.x & 31;
    int base = 0;
    if (lane32 == 0 && wm) { base = atomicAdd(A.n_req, __popc(wm)); atomicAdd(&A.counters[2], (unsigned long long)__popc(wm)); }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (mine) A.req[base + __popc(wm & ((1u << lane32) - 1u))] = make_uint2((uint32_t)g, (uint32_t)my_o | ((uint32_t)gp.lane << 28));
    if (gp.lane == 0) { atomicAdd(&A.counters[0], 1ull); atomicAdd(&A.counters[4], (unsigned long long)D); }
}

// ---------------------------------------------------------------- test evaluator (shared definition with oracle/mcts_oracle.c)
__device__ __forceinline__ void synthetic_eval(const uint32_t *k, float &v, float &var) {
    uint32_t h = 2166136261u;
#pragma unroll
    for (int i = 0; i < 11; ++i) { h ^= k[i]; h *= 16777619u; }
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    v = (float)((h >> 7) & 0x3fffu) * 0.00390625f;
    var = 0.5f + (float)((h >> 21) & 0x3ffu) * 0.0625f;
}

__global__ void k_eval_synthetic(Arena A) {
    int n = *A.n_req;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint2 r = A.req[i];
        int g = (int)r.x, o = (int)(r.y & 0x0fffffffu), slot = (int)(r.y >> 28);
        float v, var;
        synthetic_eval(A.key + node_at(A, g, o) * KEY_WORDS, v, var);
        A.eval_out[(size_t)g * 8 + slot] = make_float2(v, var);
    }
}

// ---------------------------------------------------------------- random rollout (Vanilla.py:47-55), one thread per game
__global__ void k_rollout(Arena A) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= A.G || A.status[g] != ST_OK || A.leaf_kind[g] != LEAF_EXPANDED) return;
    int leaf = A.trace[(size_t)g * A.trace_max + A.trace_len[g] - 1];
    uint32_t w[REC_WORDS];
    load_rec(A.rec + node_at(A, g, leaf) * REC_WORDS, w);
    Game gm;
    unpack(gm, w);
    uint32_t sr = A.srng[g];
    unsigned steps = 0;
    while (!gm.end) { play(gm, (int)(rng_next(sr) % 7u)); ++steps; }
    A.srng[g] = sr;
    A.rollout_val[g] = (float)gm.score;
    atomicAdd(&A.counters[5], (unsigned long long)steps);
}

// ---------------------------------------------------------------- backup (core.h:226-381), one thread per game
__global__ void k_backup(Arena A) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= A.G || A.status[g] != ST_OK) return;
    ArenaAcc acc{A, g};
    int D = A.trace_len[g];
    int leaf = acc.get_trace(D - 1);
    int kind = A.leaf_kind[g];
    int lo; float leaf_score;
    acc.meta(leaf, lo, leaf_score);
    if (A.mode == MODE_LP) {
        int c_nodes[7], c_obs[7], slot[7]; float c_score[7], ev[7], evar[7]; bool cend[7];
        int k = 0;
        if (kind == LEAF_EXPANDED) {
            k = unique_scalar(acc, leaf, c_nodes, c_obs, c_score, slot);
            for (int i = 0; i < k; ++i) {
                float2 e = A.eval_out[(size_t)g * 8 + slot[i]];
                ev[i] = e.x; evar[i] = e.y;
                cend[i] = A.lp_end_from_obs ? (A.stat[node_at(A, g, c_obs[i])].w != 0) : false;
            }
        }
        lp_backup(acc, D, k, c_obs, c_score, ev, evar, cend, A.gamma, false, true, A.lp_var_gamma2 != 0, leaf_score);
    } else if (A.mode == MODE_SINGLE) {
        double v = (double)leaf_score, var = 0.0;
        if (kind == LEAF_EXPANDED) {
            float2 e = A.eval_out[(size_t)g * 8 + 7];
            v = (double)__fadd_rn(leaf_score, e.x);      // ValueSim.py:86 int + np.float32 -> float32 (numpy >= 2)
            var = (double)e.y;
        }
        backup_trace(acc, D, v, var, A.gamma);
    } else {
        double v = (double)leaf_score, var = 0.0;
        if (kind == LEAF_EXPANDED) { v = (double)A.rollout_val[g]; var = A.rollout_variance; }   // Vanilla.py:53-54
        backup_trace(acc, D, v, var, A.gamma);
    }
}

// ---------------------------------------------------------------- compute_stats / get_action (agent.py:153-185)
__global__ void k_root_stats(Arena A, float *stats, int32_t *action) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= A.G) return;
    ArenaAcc acc{A, g};
    int root = A.root[g];
    int ro; float rs;
    acc.meta(root, ro, rs);
    float best = 0.f; int arg = 0;
    for (int a = 0; a < 7; ++a) {
        int c, o; float s;
        acc.children(root, a, c, o, s);
        int4 st = acc.stat(o);                                  // child 0 -> obs 0 -> zeros, as in the reference
        float diff = __fsub_rn(s, rs);                          // agent.py:171
        float val = __fadd_rn(__int_as_float(st.y), diff);      // agent.py:175
        stats[(size_t)g * 21 + a] = (float)st.x;
        stats[(size_t)g * 21 + 7 + a] = val;
        stats[(size_t)g * 21 + 14 + a] = __int_as_float(st.z);
        if (a == 0 || val > best) { best = val; arg = a; }      // np.argmax: first maximum
    }
    action[g] = arg;
}

// ---------------------------------------------------------------- single-call twins of core.cpp:20-26 (one group / one thread)
struct TwinArgs {
    const int32_t *child; int32_t *visit; float *value; float *variance; const float *score; const int32_t *n2o;
    int32_t *trace; uint32_t *rng; int32_t *out;
};

__global__ void k_twin_select(Arena A, TwinArgs t, int index, int low, int max_trace) {
    Grp gp;
    if (threadIdx.x >= 8) return;
    RefAcc acc{t.child, t.visit, t.value, t.variance, t.score, t.n2o, t.trace, t.rng, &A};
    int D = 0, status = ST_OK;
    select_trace(acc, gp, index, low, max_trace, D, status);
    if (gp.lane == 0) { t.out[0] = D; t.out[1] = status; }
}

__global__ void k_twin_unique(Arena A, TwinArgs t, int index) {
    Grp gp;
    if (threadIdx.x >= 8) return;
    RefAcc acc{t.child, t.visit, t.value, t.variance, t.score, t.n2o, t.trace, t.rng, &A};
    int c, o; float s;
    acc.children(index, gp.lane, c, o, s);
    Uniq u = unique_children(gp, c, o, s);
    int pos = __popc(u.first_mask & ((1u << gp.lane) - 1u));
    if (u.is_first) { t.out[1 + pos] = u.rep_c; t.out[8 + pos] = o; }
    if (gp.lane == 0) t.out[0] = __popc(u.first_mask);
}

__global__ void k_twin_backup(Arena A, TwinArgs t, int D, double v, double var, double gamma, int mixture) {
    RefAcc acc{t.child, t.visit, t.value, t.variance, t.score, t.n2o, t.trace, t.rng, &A};
    if (mixture) backup_trace_mixture(acc, D, v, var, gamma);
    else backup_trace(acc, D, v, var, gamma);
}

__global__ void k_twin_backup_lp(Arena A, TwinArgs t, int D, const uint8_t *end, const int32_t *c_nodes, const int32_t *c_obs,
                                 int k, const float *v, const float *var, double gamma, int mixture, int averaged) {
    RefAcc acc{t.child, t.visit, t.value, t.variance, t.score, t.n2o, t.trace, t.rng, &A};
    int co[7]; float cs[7], ev[7], evar[7]; bool ce[7];
    for (int i = 0; i < k; ++i) { co[i] = c_obs[i]; cs[i] = t.score[c_nodes[i]]; ev[i] = v[i]; evar[i] = var[i]; ce[i] = end[c_nodes[i]] != 0; }
    float leaf_score = t.score[t.trace[D - 1]];
    lp_backup(acc, D, k, co, cs, ev, evar, ce, gamma, mixture != 0, averaged != 0, true, leaf_score);
}

// get_all_childs (core.h:32-50) on the reference layout: mark[] = 1 for every traversed index (0 included)
__global__ void k_twin_all_childs(const int32_t *child, int M, int index, uint8_t *mark, int32_t *queue) {
    if (threadIdx.x != 0) return;
    int head = 0, tail = 0;
    queue[tail++] = index; mark[index] = 1;
    while (head < tail) {
        int n = queue[head++];
        for (int a = 0; a < 7; ++a) {
            int c = child[(size_t)n * 7 + a];
            if (!mark[c]) { mark[c] = 1; queue[tail++] = c; }
        }
    }
}

}  // namespace b200
