// dist_dev.cuh — distributional search cores (BASELINE config 5), device side.
// Restated from the reference's numba functions agents/core_distributional.py:12-124 (the agent file that should call
// them, agents/DistValueSimOnline.py, is not runnable upstream).  Arrays in the reference layout:
//   node_stats f32[M][5] = {visit, mean, reward, variance, M2}, node_dist f32[M][bins], child int32[M][7].
// The reference compiles these with fastmath (not bit-stable); parity is to 1e-5 relative.  One documented choice:
// `list(set(children))` (core_distributional.py:90-91) is taken in ascending index order.
#pragma once
#include <stdint.h>
#include "search_dev.cuh"

namespace b200 {

// core_distributional.py:12-36
__device__ inline void dist_shift(const float *dist, int bins, double x, double vmin, double vmax, float *result) {
    const double delta = (vmax - vmin) / bins, bin_shift = x / delta, fraction = bin_shift - floor(bin_shift);
    for (int b = 0; b < bins; ++b) result[b] = 0.f;
    for (int b = 0; b < bins; ++b) {
        int b_lb = (int)(b + bin_shift);
        if (b_lb >= bins) b_lb = bins - 1;
        int b_ub = (b_lb + 1 >= bins) ? bins - 1 : b_lb + 1;
        result[b_lb] = (float)__dadd_rn((double)result[b_lb], __dmul_rn((double)dist[b], 1.0 - fraction));
        result[b_ub] = (float)__dadd_rn((double)result[b_ub], __dmul_rn((double)dist[b], fraction));
    }
}

// core_distributional.py:39-45
__device__ inline double dist_mean(const float *dist, int bins, double vmin, double vmax) {
    const double delta = (vmax - vmin) / bins;
    double s = 0;
    for (int b = 0; b < bins; ++b) s = __dadd_rn(s, __dmul_rn((double)dist[b], (b + 0.5) * delta));
    return s;
}

// core_distributional.py:48-63
__device__ inline void dist_mean_variance(const float *dist, int bins, double vmin, double vmax, double &mean, double &var) {
    const double delta = (vmax - vmin) / bins;
    double m = 0, m2 = 0;
    for (int b = 0; b < bins; ++b) {
        double center = (b + 0.5) * delta, tmp = __dmul_rn(center, (double)dist[b]);
        m = __dadd_rn(m, tmp);
        m2 = __dadd_rn(m2, __dmul_rn(center, tmp));
    }
    mean = m;
    var = __dsub_rn(m2, __dmul_rn(m, m));
}

__device__ inline double dist_norm_quantile(double t) {   // agents/special.py:55-66 == special.h:26-33
    double alpha = 1.0 - 1.0 / t;
    return 10.0 * log(1.0 - log(-log(alpha) / log(2.0)) / log(22.0)) / log(41.0);
}

// core_distributional.py:66-79
__device__ inline int dist_policy(const int32_t *cn, int k, const float *node_stats, double curr_reward) {
    double n = 0;
    float s0[7], s1[7];
    for (int i = 0; i < k; ++i) {
        const float *ns = node_stats + (size_t)cn[i] * 5;
        n += (double)ns[0];
        s0[i] = (float)((double)__fadd_rn(ns[1], ns[2]) - curr_reward);
        s1[i] = (float)((double)ns[3] / ((double)ns[0] + 1e-3));
    }
    const double z = dist_norm_quantile(n);
    int best = 0; double bq = 0;
    for (int i = 0; i < k; ++i) {
        double q = (double)s0[i] + z * (double)__fsqrt_rn(s1[i]);
        if (i == 0 || q > bq) { bq = q; best = i; }
    }
    return cn[best];
}

__device__ inline int dist_unique_sorted(const int32_t *row, int32_t *out) {
    int k = 0;
    for (int a = 0; a < 7; ++a) {
        int c = row[a];
        if (c == 0) continue;
        int j = 0;
        while (j < k && out[j] != c) ++j;
        if (j == k) out[k++] = c;
    }
    for (int i = 1; i < k; ++i) { int v = out[i], j = i - 1; while (j >= 0 && out[j] > v) { out[j + 1] = out[j]; --j; } out[j + 1] = v; }
    return k;
}

// core_distributional.py:82-106 (check_low: agents/core.py:462-468, draws from the injected xorshift stream)
__device__ inline int dist_select_trace(int index, const int32_t *child, const float *node_stats, int low, int32_t *trace,
                                        int max_trace, uint32_t *rng) {
    int D = 0;
    for (;;) {
        if (D >= max_trace) return -1;
        trace[D++] = index;
        int32_t cn[7], lowl[7];
        int k = dist_unique_sorted(child + (size_t)index * 7, cn), nl = 0;
        if (k == 0) break;
        const double r = node_stats[(size_t)index * 5 + 2];
        for (int i = 0; i < k; ++i) if (node_stats[(size_t)cn[i] * 5] < (float)low) lowl[nl++] = cn[i];
        if (nl) { uint32_t s = *rng; uint32_t x = rng_next(s); *rng = s; index = lowl[x % (uint32_t)nl]; }
        else index = dist_policy(cn, k, node_stats, r);
    }
    return D;
}

// core_distributional.py:109-124; scratch = bins floats
__device__ inline void dist_backup_trace(const int32_t *trace, int D, float *node_stats, float *node_dist, int bins, double r,
                                         const float *dist, double vmin, double vmax, float *scratch) {
    const double mean = dist_mean(dist, bins, vmin, vmax);
    for (int t = 0; t < D; ++t) {
        const int idx = trace[t];
        float *ns = node_stats + (size_t)idx * 5, *nd = node_dist + (size_t)idx * bins;
        const double _r = r - (double)ns[2];
        dist_shift(dist, bins, _r, vmin, vmax, scratch);
        for (int b = 0; b < bins; ++b) {
            float num = __fadd_rn(__fmul_rn(nd[b], ns[0]), scratch[b]);
            nd[b] = (float)((double)num / ((double)ns[0] + 1.0));
        }
        const double x = mean + _r;
        ns[0] = __fadd_rn(ns[0], 1.f);
        const double delta = x - (double)ns[1];
        ns[1] = (float)((double)ns[1] + delta / (double)ns[0]);
        const double delta2 = x - (double)ns[1];
        ns[4] = (float)((double)ns[4] + delta * delta2);
        if (ns[0] > 1.f) ns[3] = (float)((double)ns[4] / ((double)ns[0] - 1.0));
    }
}

// ---- single-call twins (one thread; these are operator-level parity entry points, not the throughput path)
__global__ void k_dist_shift(const float *dist, int bins, double x, double vmin, double vmax, float *out) { dist_shift(dist, bins, x, vmin, vmax, out); }
__global__ void k_dist_mean_variance(const float *dist, int bins, double vmin, double vmax, double *out) { dist_mean_variance(dist, bins, vmin, vmax, out[0], out[1]); }
__global__ void k_dist_select(int index, const int32_t *child, const float *node_stats, int low, int32_t *trace, int max_trace, uint32_t *rng, int32_t *len) {
    *len = dist_select_trace(index, child, node_stats, low, trace, max_trace, rng);
}
__global__ void k_dist_backup(const int32_t *trace, int D, float *node_stats, float *node_dist, int bins, double r, const float *dist, double vmin,
                              double vmax, float *scratch) {
    dist_backup_trace(trace, D, node_stats, node_dist, bins, r, dist, vmin, vmax, scratch);
}

}  // namespace b200
