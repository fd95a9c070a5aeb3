"""ctypes view of libqkdldpc_host.so -- the C++ host helpers (reference-compatible input generation, ...)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqkdldpc_host.so")
_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C qkd_ldpc_v_b200/host`")
        L = C.CDLL(LIB_PATH)
        L.qkdhost_gen_keys.restype = C.c_double
        L.qkdhost_gen_keys.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.c_int]
        L.qkdhost_trial_seeds.restype = None
        L.qkdhost_trial_seeds.argtypes = [C.c_uint64, C.c_int64, C.c_void_p]
        _LIB = L
    return _LIB


def trial_seeds(simulation_seed: int, count: int) -> np.ndarray:
    """seeds[] of QKD_LDPC_batch_simulation (simulation.cpp:713-719)."""
    s = np.zeros(count, np.uint64)
    lib().qkdhost_trial_seeds(int(simulation_seed), count, s.ctypes.data)
    return s


def gen_keys(seeds, n: int, qber: float, threads: int | None = None):
    """run_trial's Alice/Bob keys per seed (simulation.cpp:549-555), packed uint32 [count][words]."""
    seeds = np.ascontiguousarray(seeds, np.uint64)
    w = (n + 31) // 32
    a = np.zeros((seeds.size, w), np.uint32)
    b = np.zeros((seeds.size, w), np.uint32)
    acc = lib().qkdhost_gen_keys(seeds.ctypes.data, seeds.size, n, float(qber), a.ctypes.data, b.ctypes.data,
                                 threads or (os.cpu_count() or 1))
    return a, b, acc
