"""ctypes wrapper of oracle/liboracle.so (TEST INFRASTRUCTURE; see oracle/ldpc_oracle.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build() -> str:
    """Compile the C restatement (gcc, a second or two)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return os.path.join(_HERE, "liboracle.so")


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        src = os.path.join(_HERE, "ldpc_oracle.c")
        inc = os.path.join(_HERE, "ldpc_oracle_body.inc")
        if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(src), os.path.getmtime(inc)):
            build()
        L = C.CDLL(path)
        L.oracle_csr_to_csc.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p]
        L.oracle_csr_to_csc.restype = None
        for sfx, fp in (("f64", _f64p), ("f32", _f32p)):
            f = getattr(L, f"oracle_decode_{sfx}")
            f.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, C.c_int, fp, _i32p, C.c_int, C.c_double,
                          C.c_double, C.c_int, C.c_double, _i32p, C.POINTER(C.c_int), C.c_void_p]
            f.restype = C.c_int
            g = getattr(L, f"oracle_frame_setup_{sfx}")
            g.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, C.c_double, _i32p, C.c_int, _i32p, C.c_int,
                          fp, _i32p]
            g.restype = None
        L.oracle_qkd_ldpc_batch.argtypes = [
            C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
            C.c_double, C.c_long, _u8p, _u8p, _f64p, _i32p, C.c_int, _i32p, C.c_int, C.c_void_p, _i32p, _u8p, C.c_int]
        L.oracle_qkd_ldpc_batch.restype = C.c_int
        _LIB = L
    return _LIB


class Code:
    """Tanner graph as CSR (check_nodes) + CSC (bit_nodes)."""

    def __init__(self, n, m, row_ptr, col_idx, col_ptr=None, row_idx=None):
        self.n, self.m = int(n), int(m)
        self.row_ptr = np.ascontiguousarray(row_ptr, np.int32)
        self.col_idx = np.ascontiguousarray(col_idx, np.int32)
        self.nnz = int(self.row_ptr[-1])
        if col_ptr is None:
            col_ptr = np.zeros(self.n + 1, np.int32)
            row_idx = np.zeros(max(self.nnz, 1), np.int32)
            lib().oracle_csr_to_csc(self.n, self.m, self.row_ptr, self.col_idx, col_ptr, row_idx)
            row_idx = row_idx[: self.nnz]
        self.col_ptr = np.ascontiguousarray(col_ptr, np.int32)
        self.row_idx = np.ascontiguousarray(row_idx, np.int32)


def decode(code: Code, alg: int, llr, syndrome, max_iter=100, primary=0.0, secondary=0.0, enable_thr=True,
           thr=100.0, precision=64, trace=False):
    """One frame through one of the six decoders. Returns (iters, match, bits[, trace array])."""
    L = lib()
    dt = np.float64 if precision == 64 else np.float32
    llr = np.ascontiguousarray(llr, dt)
    syn = np.ascontiguousarray(syndrome, np.int32)
    out = np.zeros(code.n, np.int32)
    match = C.c_int(0)
    tr = None
    trp = None
    if trace:
        tr = np.full((max_iter, code.n + code.nnz), np.nan, dt)
        trp = tr.ctypes.data_as(C.c_void_p)
    f = L.oracle_decode_f64 if precision == 64 else L.oracle_decode_f32
    it = f(code.n, code.m, code.row_ptr, code.col_idx, code.col_ptr, code.row_idx, alg, llr, syn, max_iter,
           primary, secondary, int(enable_thr), thr, out, C.byref(match), trp)
    if trace:
        return it, bool(match.value), out, tr
    return it, bool(match.value), out


def frame_setup(code: Code, alice, bob, qber, punct=(), shortd=(), precision=64):
    L = lib()
    dt = np.float64 if precision == 64 else np.float32
    llr = np.zeros(code.n, dt)
    syn = np.zeros(code.m, np.int32)
    p = np.ascontiguousarray(punct, np.int32)
    s = np.ascontiguousarray(shortd, np.int32)
    f = L.oracle_frame_setup_f64 if precision == 64 else L.oracle_frame_setup_f32
    f(code.n, code.m, code.row_ptr, code.col_idx, np.ascontiguousarray(alice, np.int32),
      np.ascontiguousarray(bob, np.int32), float(qber), p if p.size else np.zeros(1, np.int32), p.size,
      s if s.size else np.zeros(1, np.int32), s.size, llr, syn)
    return llr, syn


def qkd_ldpc_batch(code: Code, alg, alice, bob, qber, max_iter=100, primary=0.0, secondary=0.0, enable_thr=True,
                   thr=100.0, punct=(), shortd=(), precision=64, threads=None, want_bits=True):
    """Batch of EXTENDED frames (uint8 [frames][n]). Returns (iters int32[F], flags uint8[F], bits uint8[F][n])."""
    L = lib()
    alice = np.ascontiguousarray(alice, np.uint8)
    bob = np.ascontiguousarray(bob, np.uint8)
    F = alice.shape[0]
    q = np.ascontiguousarray(np.broadcast_to(np.asarray(qber, np.float64), (F,)))
    p = np.ascontiguousarray(punct, np.int32)
    s = np.ascontiguousarray(shortd, np.int32)
    iters = np.zeros(F, np.int32)
    flags = np.zeros(F, np.uint8)
    bits = np.zeros((F, code.n), np.uint8) if want_bits else None
    if threads is None:
        threads = os.cpu_count() or 1
    L.oracle_qkd_ldpc_batch(precision, code.n, code.m, code.row_ptr, code.col_idx, code.col_ptr, code.row_idx, alg,
                            max_iter, primary, secondary, int(enable_thr), thr, F, alice, bob, q,
                            p if p.size else np.zeros(1, np.int32), p.size, s if s.size else np.zeros(1, np.int32),
                            s.size, bits.ctypes.data_as(C.c_void_p) if want_bits else None, iters, flags,
                            int(threads))
    return iters, flags, bits
