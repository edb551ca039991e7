// replay_policy.cuh — the online agent's replay-memory policies, the step after the collection that fills the memory.
//
// Reference: OnlineMCTSAgent::remove_nodes agents/cppmodule/agent.cpp:619-708 (accumulation policies 0-3), weighted_trimming :710-749,
// random_trimming :751-775 (+ IntSampler :51-65, the global std::mt19937 mt(123) :29-30), store_nodes' policy-0 drop :800-801.
// The rows themselves are stored on the device by k_gc (kernels.cuh) in the reference's order; this file holds the host-side decision logic
// (the reference's is host C++ too) and the device compaction the two trimmings need.  Random draws use the same std::mt19937 /
// std::uniform_real_distribution / std::shuffle as the reference, seeded alike, so that a single-game engine reproduces the reference's
// memory bit for bit (tests/golden/agent_online_golden.npz, recorded from the compiled agent.cpp).
#pragma once
#include <algorithm>
#include <deque>
#include <numeric>
#include <random>
#include <vector>

namespace b200 {

struct ReplayPolicy {
    int policy = -1;                     // -1: no policy (plain ValueSim memory: b200_replay_drain_dev)
    int memory_size = 0, episodes_per_train = 25, memory_growth_rate = 5000;       // agent.cpp:839-842 defaults
    int memory_index = 0;
    std::deque<int> nodes_per_episode;
    int accumulated_nodes = 0, last_accumulation_episode = 0, last_training_episode = 0, n_trains = 0;
    double memory_drop_prob = 0.0;
    std::mt19937 mt{123};                                        // agent.cpp:22,29  #define SEED 123; std::mt19937 mt(SEED)
    std::uniform_real_distribution<double> unif{0., 1.};         // agent.cpp:30
    std::vector<int> sampler;                                    // IntSampler::indices (agent.cpp:51-65), a function-static in random_trimming
};

// out[dst(i)] = rows[i] for every kept row i in [lo, hi), dst = lo + number of kept rows before i; one CTA, chunk by chunk (rare operation)
__global__ void __launch_bounds__(1024) k_replay_compact(const uint8_t *rows, const uint8_t *keep, int lo, int hi, uint8_t *out, int *kept_out) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int c0 = lo; c0 < hi; c0 += 1024) {
        const int i = c0 + threadIdx.x;
        const int f = (i < hi && keep[i]) ? 1 : 0;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) s_warp[wid] = __popc(m);
        __syncthreads();
        int off = 0, tot = 0;
        for (int w = 0; w < 32; ++w) { const int v = s_warp[w]; if (w < wid) off += v; tot += v; }
        const int dst = lo + s_base + off + __popc(m & ((1u << lane) - 1u));
        if (f) for (int b = 0; b < 212; b += 4) *reinterpret_cast<uint32_t *>(out + (size_t)dst * 212 + b) = *reinterpret_cast<const uint32_t *>(rows + (size_t)i * 212 + b);
        __syncthreads();
        if (threadIdx.x == 0) s_base += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *kept_out = s_base;
}

__global__ void k_replay_visits(const uint8_t *rows, int n, float *out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float f;
        memcpy(&f, rows + (size_t)i * 212 + 208, 4);
        out[i] = f;
    }
}

}  // namespace b200
