"""ctypes wrapper of oracle/_ref/libqkdref.so -- the UNMODIFIED reference, compiled by oracle/Makefile
(TEST INFRASTRUCTURE). Available wherever the prebuilt library travelled to (it is git-ignored but not
gpurun-ignored) or /root/reference exists to build it from."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libqkdref.so")
REFERENCE_ROOT = os.environ.get("QKD_REFERENCE_ROOT", "/root/reference")
_LIB = None

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build() -> None:
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref", f"REF={REFERENCE_ROOT}"])


def available() -> bool:
    return os.path.exists(_PATH)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not available():
            build()
        if not available():
            raise RuntimeError("oracle/_ref/libqkdref.so missing and /root/reference absent -- cannot build it")
        L = C.CDLL(_PATH)
        L.ref_last_error.restype = C.c_char_p
        L.ref_matrix_load.argtypes = [C.c_char_p, C.c_int]
        L.ref_matrix_load.restype = C.c_void_p
        L.ref_matrix_free.argtypes = [C.c_void_p]
        L.ref_matrix_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 4 + [C.POINTER(C.c_int)]
        L.ref_matrix_csr.argtypes = [C.c_void_p, _i32p, _i32p]
        L.ref_matrix_csc.argtypes = [C.c_void_p, _i32p, _i32p]
        L.ref_set_cfg.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_int]
        L.ref_gen_keys.argtypes = [C.c_uint64, C.c_int64, C.c_double, _i32p, _i32p]
        L.ref_gen_keys.restype = C.c_double
        L.ref_trial_seeds.argtypes = [C.c_uint64, C.c_int64, _u64p]
        L.ref_decode.argtypes = [C.c_void_p, C.c_int, _f64p, _i32p, C.c_int64, C.c_double, C.c_double, C.c_int,
                                 C.c_double, _i32p, C.POINTER(C.c_int)]
        L.ref_decode.restype = C.c_int64
        L.ref_syndrome.argtypes = [C.c_void_p, _i32p, _i32p]
        L.ref_run_trials.argtypes = [C.c_void_p, C.c_double, _u64p, C.c_int64, C.c_double, C.c_double, _i32p,
                                     C.c_int64, _i32p, C.c_int64, _i32p, C.c_int64, C.c_int, _i32p, _u8p, _f64p]
        L.ref_run_trials.restype = C.c_int
        L.ref_qkd_ldpc_batch.argtypes = [C.c_void_p, _i32p, _i32p, C.c_int64, C.c_double, C.c_double, C.c_double,
                                         C.c_int, _i32p, _u8p]
        L.ref_qkd_ldpc_batch.restype = C.c_int
        L.ref_adapt_code_rate.argtypes = [C.c_void_p, C.c_uint64, C.c_int, _i32p, C.c_int64, C.c_double, C.c_double,
                                          C.c_double, _i32p, C.POINTER(C.c_int64), _i32p, C.POINTER(C.c_int64),
                                          _i32p, C.POINTER(C.c_int64), _f64p]
        L.ref_adapt_code_rate.restype = C.c_int
        L.ref_describe_config.restype = C.c_int64
        L.ref_describe_config.argtypes = [C.c_char_p, C.c_char_p, C.c_int64]
        L.ref_describe_inputs.restype = C.c_int64
        L.ref_describe_inputs.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64]
        L.ref_csv_from_trials.restype = C.c_int64
        L.ref_csv_from_trials.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int64, C.c_int,
                                          C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_char_p, C.c_char_p,
                                          C.c_int64]
        L.ref_bits_to_remove_rate_adapt.restype = C.c_int64
        L.ref_bits_to_remove_rate_adapt.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.ref_untainted.restype = C.c_int64
        L.ref_untainted.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        _LIB = L
    return _LIB


def _arr(x):
    x = np.ascontiguousarray(x, np.int32)
    return x if x.size else np.zeros(1, np.int32)


class RefMatrix:
    """H_matrix loaded by the reference's own loaders (format: 0 dense, 1 alist, 2 sparse_1, 3 sparse_2)."""

    def __init__(self, path: str, fmt: int):
        L = lib()
        self._h = L.ref_matrix_load(os.fsencode(path), fmt)
        if not self._h:
            raise RuntimeError(L.ref_last_error().decode())
        n, m, a, b = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        reg = C.c_int()
        L.ref_matrix_info(self._h, C.byref(n), C.byref(m), C.byref(a), C.byref(b), C.byref(reg))
        self.n, self.m, self.nnz, self.nnz_cols, self.is_regular = n.value, m.value, a.value, b.value, bool(reg.value)
        self.row_ptr = np.zeros(self.m + 1, np.int32)
        self.col_idx = np.zeros(max(self.nnz, 1), np.int32)
        L.ref_matrix_csr(self._h, self.row_ptr, self.col_idx)
        self.col_idx = self.col_idx[: self.nnz]
        self.col_ptr = np.zeros(self.n + 1, np.int32)
        self.row_idx = np.zeros(max(self.nnz_cols, 1), np.int32)
        L.ref_matrix_csc(self._h, self.col_ptr, self.row_idx)
        self.row_idx = self.row_idx[: self.nnz_cols]

    def __del__(self):
        try:
            if self._h:
                lib().ref_matrix_free(self._h)
                self._h = None
        except Exception:
            pass

    def decode(self, alg, llr, syndrome, max_iter=100, primary=0.0, secondary=0.0, enable_thr=True, thr=100.0):
        out = np.zeros(self.n, np.int32)
        match = C.c_int()
        it = lib().ref_decode(self._h, alg, np.ascontiguousarray(llr, np.float64),
                              np.ascontiguousarray(syndrome, np.int32), max_iter, primary, secondary,
                              int(enable_thr), thr, out, C.byref(match))
        return int(it), bool(match.value), out

    def syndrome(self, bits):
        s = np.zeros(self.m, np.int32)
        lib().ref_syndrome(self._h, np.ascontiguousarray(bits, np.int32), s)
        return s

    def run_trials(self, qber, seeds, primary=0.0, secondary=0.0, punct=(), shortd=(), remove=(), threads=None):
        """The reference's run_trial per seed (set_cfg first). Returns iters, flags, accurate_qber."""
        seeds = np.ascontiguousarray(seeds, np.uint64)
        k = seeds.size
        iters = np.zeros(k, np.int32)
        flags = np.zeros(k, np.uint8)
        acc = np.zeros(k, np.float64)
        p, s, r = np.asarray(punct), np.asarray(shortd), np.asarray(remove)
        rc = lib().ref_run_trials(self._h, qber, seeds, k, primary, secondary, _arr(p), p.size, _arr(s), s.size,
                                  _arr(r), r.size, threads or (os.cpu_count() or 1), iters, flags, acc)
        if rc != 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return iters, flags, acc

    def qkd_ldpc_batch(self, alice, bob, qber, primary=0.0, secondary=0.0, threads=1):
        """QKD_LDPC on pre-generated int32 [F][n] keys, decode only (the CPU baseline's timed call)."""
        alice = np.ascontiguousarray(alice, np.int32)
        bob = np.ascontiguousarray(bob, np.int32)
        F = alice.shape[0]
        iters = np.zeros(F, np.int32)
        flags = np.zeros(F, np.uint8)
        lib().ref_qkd_ldpc_batch(self._h, alice, bob, F, qber, primary, secondary, threads, iters, flags)
        return iters, flags

    def adapt_code_rate(self, seed, qber, delta, efficiency, untainted=False, untp=()):
        n = self.n
        po, so, ro = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        fr = np.zeros(3, np.float64)
        u = np.asarray(untp)
        rc = lib().ref_adapt_code_rate(self._h, seed, int(untainted), _arr(u), u.size, qber, delta, efficiency, po,
                                       C.byref(a), so, C.byref(b), ro, C.byref(c), fr)
        if rc != 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return po[: a.value].copy(), so[: b.value].copy(), fr


def set_cfg(algorithm, max_iter=100, enable_thr=True, thr=100.0, privacy_maintenance=False, rate_adaptation=False):
    lib().ref_set_cfg(algorithm, max_iter, int(enable_thr), thr, int(privacy_maintenance), int(rate_adaptation))


def gen_keys(seed, n, qber):
    """fill_random_bits + inject_errors exactly as run_trial does (simulation.cpp:549-555)."""
    a = np.zeros(n, np.int32)
    b = np.zeros(n, np.int32)
    acc = lib().ref_gen_keys(int(seed), n, qber, a, b)
    return a, b, acc


def trial_seeds(simulation_seed, count):
    s = np.zeros(count, np.uint64)
    lib().ref_trial_seeds(int(simulation_seed), count, s)
    return s


def _text(fn, *args, cap=1 << 22) -> str:
    buf = C.create_string_buffer(cap)
    k = fn(*args, buf, cap)
    if k < 0:
        raise RuntimeError(lib().ref_last_error().decode())
    if k >= cap:
        return _text(fn, *args, cap=k + 1)
    return buf.value.decode()


def describe_config(path: str) -> str:
    """The reference's parse_config_data (config.cpp:89-403) as canonical text."""
    return _text(lib().ref_describe_config, os.fsencode(path))


def describe_inputs(config_path: str, matrix_dir: str) -> str:
    """The reference's prepare_sim_inputs (simulation.cpp:371-537): one line per (matrix, combination)."""
    return _text(lib().ref_describe_inputs, os.fsencode(config_path), os.fsencode(matrix_dir))


def csv_from_trials(config_path, iters, flags, name, n, m, is_regular, config_qber, accurate_qber, primary=0.0, secondary=0.0,
                    adapt5=None, tmp_dir="/tmp") -> str:
    """process_trials_results + write_file (simulation.cpp:580-690, 4-176) on given per-trial results -> CSV text."""
    iters = np.ascontiguousarray(iters, np.int32)
    flags = np.ascontiguousarray(flags, np.uint8)
    a5 = np.ascontiguousarray(adapt5, np.float64) if adapt5 is not None else None
    return _text(lib().ref_csv_from_trials, os.fsencode(config_path), iters.ctypes.data, flags.ctypes.data, iters.size, name.encode(),
                 n, m, int(is_regular), config_qber, accurate_qber, primary, secondary,
                 a5.ctypes.data if a5 is not None else None, os.fsencode(tmp_dir))


def untainted(matrix: "RefMatrix", seed: int) -> np.ndarray:
    out = np.zeros(matrix.n, np.int32)
    k = lib().ref_untainted(matrix._h, int(seed), out.ctypes.data)
    if k < 0:
        raise RuntimeError(lib().ref_last_error().decode())
    return out[:k].copy()


def bits_to_remove_rate_adapt(matrix: "RefMatrix", punct, short, pad: int = 1) -> np.ndarray:
    """get_bits_positions_to_remove_rate_adapt (array_and_matrix_operations.cpp:189-256). The reference reads
    shortened_bits[s] / punctured_bits[p] one past the end (:215,:220, undefined behaviour); pad=1 makes that read
    hit a -1 sentinel instead of heap garbage."""
    punct = np.ascontiguousarray(punct, np.int32)
    short = np.ascontiguousarray(short, np.int32)
    out = np.zeros(matrix.n, np.int32)
    k = lib().ref_bits_to_remove_rate_adapt(matrix._h, punct.ctypes.data, punct.size, short.ctypes.data, short.size, pad,
                                            out.ctypes.data)
    if k < 0:
        raise RuntimeError(lib().ref_last_error().decode())
    return out[:k].copy()
