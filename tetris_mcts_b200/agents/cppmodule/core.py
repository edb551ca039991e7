"""agents.cppmodule.core — the five functions the reference exports from agents/cppmodule/core.cpp:20-26, same names,
argument order and in-place semantics (C-contiguous numpy arrays owned by the caller), executed on the GPU through the
C-ABI twins (b200_select_trace_obs, ...).  check_low's random pick (core.h:76 libc rand()) draws from `rng_state`."""
import numpy as np

from ... import _lib as L

rng_state = np.array([0x2545F491], np.uint32)   # module-level stream, like libc's global rand() state


def _a(x, dtype):
    a = np.ascontiguousarray(x, dtype)
    return a


def get_all_childs(index, child):
    child = _a(child, np.int32)
    mark = np.zeros(child.shape[0], np.uint8)
    L.check(L.lib().b200_get_all_childs(int(index), L.ptr(child), child.shape[0], L.ptr(mark)))
    return set(np.nonzero(mark)[0].tolist())


def get_unique_child_obs(index, child, score, n_to_o):
    child, score, n_to_o = _a(child, np.int32), _a(score, np.float32), _a(n_to_o, np.int32)
    cn, co, k = np.zeros(7, np.int32), np.zeros(7, np.int32), np.zeros(1, np.int32)
    L.check(L.lib().b200_get_unique_child_obs(int(index), L.ptr(child), L.ptr(score), L.ptr(n_to_o), child.shape[0], L.ptr(cn), L.ptr(co), L.ptr(k)))
    return cn[:k[0]].tolist(), co[:k[0]].tolist()


def select_trace_obs(index, child, visit, value, variance, score, n_to_o, low):
    tr, n = np.zeros(4096, np.int32), np.zeros(1, np.int32)
    L.check(L.lib().b200_select_trace_obs(int(index), L.ptr(_a(child, np.int32)), L.ptr(_a(visit, np.int32)), L.ptr(_a(value, np.float32)),
                                          L.ptr(_a(variance, np.float32)), L.ptr(_a(score, np.float32)), L.ptr(_a(n_to_o, np.int32)),
                                          len(visit), int(low), L.ptr(rng_state), L.ptr(tr), len(tr), L.ptr(n)))
    return tr[:n[0]].copy()


def _inplace(arr, dtype, name):
    if not (isinstance(arr, np.ndarray) and arr.dtype == dtype and arr.flags.c_contiguous):
        raise TypeError("%s must be a C-contiguous %s array (it is updated in place, core.h:226-260)" % (name, np.dtype(dtype).name))
    return arr


def backup_trace_obs(trace, visit, value, variance, n_to_o, score, _value, _variance, gamma):
    trace = _a(trace, np.int32)
    L.check(L.lib().b200_backup_trace_obs(L.ptr(trace), len(trace), L.ptr(_inplace(visit, np.int32, "visit")), L.ptr(_inplace(value, np.float32, "value")),
                                          L.ptr(_inplace(variance, np.float32, "variance")), L.ptr(_a(n_to_o, np.int32)), L.ptr(_a(score, np.float32)),
                                          len(visit), float(_value), float(_variance), float(gamma)))


def backup_trace_obs_LP(trace, visit, value, variance, n_to_o, score, end, _child, _obs, _value, _variance, gamma, mixture, averaged):
    trace, cn, co = _a(trace, np.int32), _a(_child, np.int32), _a(_obs, np.int32)
    v, var, end = _a(_value, np.float32), _a(_variance, np.float32), _a(end, np.uint8)
    L.check(L.lib().b200_backup_trace_obs_LP(L.ptr(trace), len(trace), L.ptr(_inplace(visit, np.int32, "visit")), L.ptr(_inplace(value, np.float32, "value")),
                                             L.ptr(_inplace(variance, np.float32, "variance")), L.ptr(_a(n_to_o, np.int32)), L.ptr(_a(score, np.float32)),
                                             L.ptr(end), len(visit), L.ptr(cn), L.ptr(co), len(cn), L.ptr(v), L.ptr(var), float(gamma), int(bool(mixture)), int(bool(averaged))))
