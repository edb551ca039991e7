/*
 * dist_oracle.c — TEST INFRASTRUCTURE ONLY.  C restatement of the reference's distributional search cores
 * (agents/core_distributional.py:12-124, numba @jit(nopython, fastmath)); the agent that was meant to call them
 * (agents/DistValueSimOnline.py) is not runnable upstream, so these functions are the contract.
 * Pinned against the numba functions themselves run in the build container (tests/golden/dist_golden.npz; fastmath
 * makes the reference not bit-stable, so the pin is to 1e-5 relative).
 * Arrays: node_stats f32[M][5] = {visit, mean, reward, variance, M2} (core_distributional.py:109-124 field use),
 * node_dist f32[M][bins], child int32[M][7].
 * One documented choice: `list(set(children))` (core_distributional.py:90-91) iterates in the order of numba's typed-set
 * hash table; here the unique children are taken in ASCENDING INDEX order.  Only argmax ties and the random pick among
 * low-visit children depend on that order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint32_t (*mo_rand_fn)(void *ctx);
double mo_norm_quantile(double t);

/* core_distributional.py:12-36 */
void do_shift_distribution(const float *dist, int bins, double x, double vmin, double vmax, float *result) {
    double delta = (vmax - vmin) / bins;
    double bin_shift = x / delta;
    double fraction = bin_shift - floor(bin_shift);
    for (int b = 0; b < bins; ++b) result[b] = 0.f;
    for (int b = 0; b < bins; ++b) {
        int b_lb = (int)(b + bin_shift);                 /* int(): truncation */
        if (b_lb >= bins) b_lb = bins - 1;
        int b_ub = (b_lb + 1 >= bins) ? bins - 1 : b_lb + 1;
        result[b_lb] = (float)((double)result[b_lb] + (double)dist[b] * (1 - fraction));
        result[b_ub] = (float)((double)result[b_ub] + (double)dist[b] * fraction);
    }
}

/* core_distributional.py:39-45 (vmin is not added upstream either) */
double do_mean_dist(const float *dist, int bins, double vmin, double vmax) {
    double delta = (vmax - vmin) / bins, s = 0;
    for (int b = 0; b < bins; ++b) s += (double)dist[b] * ((b + 0.5) * delta);
    return s;
}

/* core_distributional.py:48-63 */
void do_mean_variance(const float *dist, int bins, double vmin, double vmax, double *mean_out, double *var_out) {
    double delta = (vmax - vmin) / bins, mean = 0, m2 = 0;
    for (int b = 0; b < bins; ++b) {
        double center = (b + 0.5) * delta;
        double tmp = center * dist[b];
        mean += tmp;
        m2 += center * tmp;
    }
    *mean_out = mean;
    *var_out = m2 - mean * mean;
}

/* core_distributional.py:66-79.  q = (mean + reward - curr_reward) + z(n) * sqrt(variance / (visit + 1e-3)) */
int do_policy_dist(const int32_t *child_nodes, int k, const float *node_stats, double curr_reward, double *q_out) {
    double n = 0;
    float s0[8], s1[8];
    for (int i = 0; i < k; ++i) {
        const float *ns = node_stats + (size_t)child_nodes[i] * 5;
        n += ns[0];
        s0[i] = (float)((double)(ns[1] + ns[2]) - curr_reward);
        s1[i] = (float)((double)ns[3] / ((double)ns[0] + 1e-3));
    }
    double z = mo_norm_quantile(n);
    int best = 0; double bq = 0;
    for (int i = 0; i < k; ++i) {
        double q = (double)s0[i] + z * (double)sqrtf(s1[i]);
        if (q_out) q_out[i] = q;
        if (i == 0 || q > bq) { bq = q; best = i; }      /* np.argmax: first maximum */
    }
    return child_nodes[best];
}

static int unique_sorted(const int32_t *row, int32_t *out) {
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

/* core_distributional.py:82-106 (check_low: agents/core.py:462-468) */
int do_select_trace_distributional(int index, const int32_t *child, const float *node_stats, int low, int32_t *trace,
                                   int max_trace, mo_rand_fn rnd, void *ctx) {
    int D = 0;
    for (;;) {
        if (D >= max_trace) return -1;
        trace[D++] = index;
        int32_t cn[7];
        int k = unique_sorted(child + (size_t)index * 7, cn);
        if (k == 0) break;
        double r = node_stats[(size_t)index * 5 + 2];
        int32_t lowl[7]; int nl = 0;
        for (int i = 0; i < k; ++i) if (node_stats[(size_t)cn[i] * 5] < (float)low) lowl[nl++] = cn[i];
        if (nl) index = lowl[rnd(ctx) % (uint32_t)nl];
        else index = do_policy_dist(cn, k, node_stats, r, NULL);
    }
    return D;
}

/* core_distributional.py:109-124 */
void do_backup_trace_distributional(const int32_t *trace, int D, float *node_stats, float *node_dist, int bins, double r,
                                    const float *dist, double vmin, double vmax) {
    double mean = do_mean_dist(dist, bins, vmin, vmax);
    float *nd_new = (float *)malloc(sizeof(float) * (size_t)bins);
    for (int t = 0; t < D; ++t) {                          /* forward order, as upstream */
        int idx = trace[t];
        float *ns = node_stats + (size_t)idx * 5;
        float *nd = node_dist + (size_t)idx * bins;
        double _r = r - (double)ns[2];
        do_shift_distribution(dist, bins, _r, vmin, vmax, nd_new);
        for (int b = 0; b < bins; ++b) {
            float num = nd[b] * ns[0] + nd_new[b];                           /* float32 array arithmetic */
            nd[b] = (float)((double)num / ((double)ns[0] + 1.0));
        }
        double x = mean + _r;
        ns[0] += 1.f;
        double delta = x - (double)ns[1];
        ns[1] = (float)((double)ns[1] + delta / (double)ns[0]);
        double delta2 = x - (double)ns[1];
        ns[4] = (float)((double)ns[4] + delta * delta2);
        if (ns[0] > 1.f) ns[3] = (float)((double)ns[4] / ((double)ns[0] - 1.0));
    }
    free(nd_new);
}
