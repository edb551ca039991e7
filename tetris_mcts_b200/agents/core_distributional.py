"""agents.core_distributional — the reference's distributional search cores (agents/core_distributional.py:12-124),
same names and argument order, executed on the GPU through the C-ABI (b200_dist_*).  node_stats is f32[M,5] =
{visit, mean, reward, variance, M2}; node_dist is f32[M,bins]; arrays are caller-owned and updated in place."""
import numpy as np

from .. import _lib as L

rng_state = np.array([0x2545F491], np.uint32)   # stands in for np.random inside check_low (agents/core.py:462-468)


def shift_distribution(dist, x, vmin, vmax):                      # core_distributional.py:12-36
    d = np.ascontiguousarray(dist, np.float32)
    out = np.zeros_like(d)
    L.check(L.lib().b200_dist_shift_distribution(L.ptr(d), len(d), float(x), float(vmin), float(vmax), L.ptr(out)))
    return out


def mean_variance(dist, vmin, vmax):                              # core_distributional.py:48-63
    d = np.ascontiguousarray(dist, np.float32)
    m, v = np.zeros(1, np.float64), np.zeros(1, np.float64)
    L.check(L.lib().b200_dist_mean_variance(L.ptr(d), len(d), float(vmin), float(vmax), L.ptr(m), L.ptr(v)))
    return float(m[0]), float(v[0])


def mean_dist(dist, vmin, vmax):                                  # core_distributional.py:39-45
    return mean_variance(dist, vmin, vmax)[0]


def select_trace_distributional(index, child, node_stats, node_dist, vmin, vmax, low=5):   # core_distributional.py:82-106
    child = np.ascontiguousarray(child, np.int32)
    ns = np.ascontiguousarray(node_stats, np.float32)
    tr, n = np.zeros(4096, np.int32), np.zeros(1, np.int32)
    L.check(L.lib().b200_dist_select_trace(int(index), L.ptr(child), L.ptr(ns), child.shape[0], int(low), L.ptr(rng_state), L.ptr(tr), len(tr), L.ptr(n)))
    return tr[:n[0]].copy()


def backup_trace_distributional(trace, node_stats, node_dist, r, dist, vmin, vmax):        # core_distributional.py:109-124
    for a, name in ((node_stats, "node_stats"), (node_dist, "node_dist")):
        if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags.c_contiguous):
            raise TypeError("%s must be a C-contiguous float32 array (updated in place)" % name)
    trace = np.ascontiguousarray(trace, np.int32)
    d = np.ascontiguousarray(dist, np.float32)
    L.check(L.lib().b200_dist_backup_trace(L.ptr(trace), len(trace), L.ptr(node_stats), L.ptr(node_dist), node_stats.shape[0], node_dist.shape[1],
                                           float(r), L.ptr(d), float(vmin), float(vmax)))
