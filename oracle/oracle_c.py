"""TEST INFRASTRUCTURE -- ctypes loader of oracle/liboracle_c.so (C restatement, see oracle_c.c)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(os.path.join(_HERE, "liboracle_c.so"))
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def gae(v_s, v_s_, rew, end_flag, gamma, lam):
    v_s, v_s_, rew = (np.ascontiguousarray(x, dtype=np.float64) for x in (v_s, v_s_, rew))
    end = np.ascontiguousarray(end_flag, dtype=np.uint8)
    out = np.empty_like(rew)
    lib().oracle_gae(_p(v_s), _p(v_s_), _p(rew), _p(end), C.c_int64(len(rew)), C.c_double(gamma), C.c_double(lam), _p(out))
    return out


def nstep_return(rew, end_flag, target_q, idx, gamma, n_step):
    rew = np.ascontiguousarray(rew, dtype=np.float64)
    end = np.ascontiguousarray(end_flag, dtype=np.uint8)
    tq = np.ascontiguousarray(target_q, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    I, A = tq.shape
    out = np.empty((I, A), dtype=np.float64)
    lib().oracle_nstep_return(_p(rew), _p(end), _p(tq), _p(idx), C.c_int64(I), C.c_int64(A), C.c_int32(n_step),
                              C.c_double(gamma), _p(out))
    return out


def _step(fn, index, offset, done, last_index, lengths):
    index = np.ascontiguousarray(index, dtype=np.int64)
    offset, last_index, lengths = (np.ascontiguousarray(x, dtype=np.int64) for x in (offset, last_index, lengths))
    done = np.ascontiguousarray(done, dtype=np.uint8)
    out = np.empty_like(index)
    fn(_p(index), C.c_int64(len(index)), _p(offset), C.c_int64(len(lengths)), _p(done), _p(last_index), _p(lengths), _p(out))
    return out


def next_index(index, offset, done, last_index, lengths):
    return _step(lib().oracle_next_index, index, offset, done, last_index, lengths)


def prev_index(index, offset, done, last_index, lengths):
    return _step(lib().oracle_prev_index, index, offset, done, last_index, lengths)
