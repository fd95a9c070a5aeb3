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
        L.qkdhost_last_error.restype = C.c_char_p
        L.qkdhost_matrix_load.restype = C.c_void_p
        L.qkdhost_matrix_load.argtypes = [C.c_char_p, C.c_int]
        L.qkdhost_matrix_free.argtypes = [C.c_void_p]
        L.qkdhost_matrix_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 3 + [C.POINTER(C.c_int)]
        L.qkdhost_matrix_csr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.qkdhost_matrix_csc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.qkdhost_describe_config.restype = C.c_int64
        L.qkdhost_describe_config.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.POINTER(C.c_int)]
        L.qkdhost_describe_inputs.restype = C.c_int64
        L.qkdhost_describe_inputs.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64]
        L.qkdhost_csv_from_trials.restype = C.c_int64
        L.qkdhost_csv_from_trials.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int64,
                                              C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_char_p,
                                              C.c_int64]
        L.qkdhost_format_shortest.restype = C.c_int64
        L.qkdhost_format_shortest.argtypes = [C.c_double, C.c_char_p, C.c_int64]
        L.qkdhost_adapt_code_rate.restype = C.c_int
        L.qkdhost_adapt_code_rate.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_int64, C.c_double, C.c_double,
                                              C.c_double, C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p,
                                              C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]
        L.qkdhost_untainted.restype = C.c_int64
        L.qkdhost_untainted.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.qkdhost_gen_keys_rate_adapt.restype = C.c_double
        L.qkdhost_gen_keys_rate_adapt.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_void_p, C.c_int64,
                                                  C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def last_error() -> str:
    return lib().qkdhost_last_error().decode()


def _text(fn, *args, cap=1 << 22) -> str:
    buf = C.create_string_buffer(cap)
    k = fn(*args, buf, cap)
    if k < 0:
        raise RuntimeError(last_error())
    if k >= cap:
        return _text(fn, *args, cap=k + 1)
    return buf.value.decode()


class HostMatrix:
    """A parity-check matrix read by the C++ host loaders (host/matrix.cpp; formats 0-3 of the reference)."""

    def __init__(self, path: str, fmt: int):
        self.h = lib().qkdhost_matrix_load(os.fsencode(path), int(fmt))
        if not self.h:
            raise RuntimeError(last_error())
        n, m, nnz, reg = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        lib().qkdhost_matrix_info(self.h, C.byref(n), C.byref(m), C.byref(nnz), C.byref(reg))
        self.n, self.m, self.nnz, self.is_regular = n.value, m.value, nnz.value, bool(reg.value)

    def csr(self):
        rp, ci = np.zeros(self.m + 1, np.int32), np.zeros(self.nnz, np.int32)
        if lib().qkdhost_matrix_csr(self.h, rp.ctypes.data, ci.ctypes.data) != 0:
            raise RuntimeError(last_error())
        return rp, ci

    def csc(self):
        cp, ri = np.zeros(self.n + 1, np.int32), np.zeros(self.nnz, np.int32)
        lib().qkdhost_matrix_csc(self.h, cp.ctypes.data, ri.ctypes.data)
        return cp, ri

    def adapt_code_rate(self, seed, qber, delta, efficiency, untainted=False, untp=None, privacy_maintenance=False):
        """adapt_code_rate with a fresh generator: (punctured, shortened, bits_to_remove, [pi, sigma, R_adapted])."""
        untp = np.ascontiguousarray(untp if untp is not None else [], np.int32)
        p, s, r = (np.zeros(self.n, np.int32) for _ in range(3))
        n_p, n_s, n_r = C.c_int64(), C.c_int64(), C.c_int64()
        fr = np.zeros(3)
        if lib().qkdhost_adapt_code_rate(self.h, int(seed), int(untainted), untp.ctypes.data, untp.size, qber, delta, efficiency,
                                         int(privacy_maintenance), p.ctypes.data, C.byref(n_p), s.ctypes.data, C.byref(n_s),
                                         r.ctypes.data, C.byref(n_r), fr.ctypes.data) != 0:
            raise RuntimeError(last_error())
        return p[:n_p.value].copy(), s[:n_s.value].copy(), r[:n_r.value].copy(), fr

    def untainted(self, seed):
        out = np.zeros(self.n, np.int32)
        k = lib().qkdhost_untainted(self.h, int(seed), out.ctypes.data)
        if k < 0:
            raise RuntimeError(last_error())
        return out[:k].copy()

    def close(self):
        if self.h:
            lib().qkdhost_matrix_free(self.h)
            self.h = None

    def __del__(self):
        self.close()


def describe_config(path: str):
    """(canonical text of the parsed config, schema generation 1..4)."""
    ver = C.c_int()
    buf = C.create_string_buffer(1 << 20)
    k = lib().qkdhost_describe_config(os.fsencode(path), buf, 1 << 20, C.byref(ver))
    if k < 0:
        raise RuntimeError(last_error())
    return buf.value.decode(), ver.value


def describe_inputs(config_path: str, matrix_dir: str, untp_cache: str = "") -> str:
    return _text(lib().qkdhost_describe_inputs, os.fsencode(config_path), os.fsencode(matrix_dir), os.fsencode(untp_cache))


def csv_from_trials(config_path, iters, flags, name, n, m, is_regular, config_qber, accurate_qber, primary=0.0, secondary=0.0,
                    adapt5=None) -> str:
    iters = np.ascontiguousarray(iters, np.int32)
    flags = np.ascontiguousarray(flags, np.uint8)
    a5 = np.ascontiguousarray(adapt5, np.float64) if adapt5 is not None else None
    return _text(lib().qkdhost_csv_from_trials, os.fsencode(config_path), iters.ctypes.data, flags.ctypes.data, iters.size,
                 name.encode(), n, m, int(is_regular), config_qber, accurate_qber, primary, secondary,
                 a5.ctypes.data if a5 is not None else None)


def format_shortest(v: float) -> str:
    return _text(lib().qkdhost_format_shortest, float(v), cap=128)


def gen_keys_rate_adapt(seeds, n: int, qber: float, punct, short):
    """Extended (rate-adapted) frames per seed: run_trial's keys + QKD_LDPC_RATE_ADAPT's frame construction."""
    seeds = np.ascontiguousarray(seeds, np.uint64)
    punct = np.ascontiguousarray(punct, np.int32)
    short = np.ascontiguousarray(short, np.int32)
    w = (n + 31) // 32
    a = np.zeros((seeds.size, w), np.uint32)
    b = np.zeros((seeds.size, w), np.uint32)
    acc = lib().qkdhost_gen_keys_rate_adapt(seeds.ctypes.data, seeds.size, n, float(qber), punct.ctypes.data, punct.size,
                                            short.ctypes.data, short.size, a.ctypes.data, b.ctypes.data)
    return a, b, acc


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
