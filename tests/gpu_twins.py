"""numpy-in / numpy-out wrappers over the C-ABI twins of agents/cppmodule/core.cpp:20-26 (used by the -m gpu tests)."""
import ctypes as C

import numpy as np

from tetris_mcts_b200 import _lib as L


def select_trace_obs(index, child, visit, value, variance, score, n2o, low, rng_state=None, max_trace=512):
    tr = np.zeros(max_trace, np.int32)
    n = np.zeros(1, np.int32)
    L.check(L.lib().b200_select_trace_obs(int(index), L.ptr(child), L.ptr(visit), L.ptr(value), L.ptr(variance), L.ptr(score),
                                          L.ptr(n2o), child.shape[0], int(low), L.ptr(rng_state), L.ptr(tr), max_trace, L.ptr(n)))
    return tr[:int(n[0])].copy()


def backup_trace_obs(trace, visit, value, variance, n2o, score, v, var, gamma):
    trace = np.ascontiguousarray(trace, np.int32)
    L.check(L.lib().b200_backup_trace_obs(L.ptr(trace), len(trace), L.ptr(visit), L.ptr(value), L.ptr(variance), L.ptr(n2o),
                                          L.ptr(score), len(visit), float(v), float(var), float(gamma)))


def backup_trace_obs_LP(trace, visit, value, variance, n2o, score, end, c_nodes, c_obs, v, var, gamma, mixture, averaged):
    trace = np.ascontiguousarray(trace, np.int32)
    cn = np.ascontiguousarray(c_nodes, np.int32)
    co = np.ascontiguousarray(c_obs, np.int32)
    v = np.ascontiguousarray(v, np.float32)
    var = np.ascontiguousarray(var, np.float32)
    end = np.ascontiguousarray(end, np.uint8)
    L.check(L.lib().b200_backup_trace_obs_LP(L.ptr(trace), len(trace), L.ptr(visit), L.ptr(value), L.ptr(variance), L.ptr(n2o),
                                             L.ptr(score), L.ptr(end), len(visit), L.ptr(cn), L.ptr(co), len(cn), L.ptr(v),
                                             L.ptr(var), float(gamma), int(mixture), int(averaged)))


def get_unique_child_obs(index, child, score, n2o):
    cn = np.zeros(7, np.int32)
    co = np.zeros(7, np.int32)
    k = np.zeros(1, np.int32)
    L.check(L.lib().b200_get_unique_child_obs(int(index), L.ptr(child), L.ptr(score), L.ptr(n2o), child.shape[0], L.ptr(cn),
                                              L.ptr(co), L.ptr(k)))
    return cn[:int(k[0])].copy(), co[:int(k[0])].copy()


def get_all_childs(index, child):
    mark = np.zeros(child.shape[0], np.uint8)
    L.check(L.lib().b200_get_all_childs(int(index), L.ptr(child), child.shape[0], L.ptr(mark)))
    return set(np.nonzero(mark)[0].tolist())
