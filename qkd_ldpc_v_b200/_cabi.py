"""ctypes binding of libqkdldpc_cuda.so (include/qkdldpc.h). No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqkdldpc_cuda.so")
_LIB = None


class Params(C.Structure):
    _fields_ = [("algorithm", C.c_int32), ("max_iterations", C.c_int32), ("primary", C.c_double),
                ("secondary", C.c_double), ("enable_threshold", C.c_int32), ("threshold", C.c_double),
                ("message_precision", C.c_int32)]


class Options(C.Structure):
    _fields_ = [("pool_bytes", C.c_int64), ("pool_slots", C.c_int32), ("steps_per_poll", C.c_int32),
                ("frames_per_lane_f32", C.c_int32), ("use_graph", C.c_int32), ("decoder_path", C.c_int32),
                ("onchip_threads", C.c_int32), ("tail_compaction", C.c_int32), ("compaction_max_ctas", C.c_int32),
                ("copy_chunks", C.c_int32), ("onchip_record_bytes", C.c_int32), ("vn_items_per_warp", C.c_int32),
                ("vn_ctas_per_sm", C.c_int32), ("compaction_fill_pct", C.c_int32)]


class Combination(C.Structure):
    _fields_ = [("qber", C.c_double), ("primary", C.c_double), ("secondary", C.c_double), ("punct_pos", C.c_void_p),
                ("n_punct", C.c_int32), ("short_pos", C.c_void_p), ("n_short", C.c_int32), ("seed_offset", C.c_uint64),
                ("remove_pos", C.c_void_p), ("n_remove", C.c_int32), ("reserved", C.c_int32)]


class Info(C.Structure):
    _fields_ = [("n", C.c_int32), ("m", C.c_int32), ("nnz", C.c_int64), ("device", C.c_int32),
                ("frames_per_tile", C.c_int32), ("pool_tiles", C.c_int32), ("pool_bytes", C.c_int64),
                ("kernel_launches", C.c_int64), ("decoder_steps", C.c_int64), ("last_batch_ms", C.c_double),
                ("last_cn_ms", C.c_double), ("last_vn_ms", C.c_double), ("last_sched_ms", C.c_double),
                ("last_path", C.c_int32), ("onchip_threads", C.c_int32), ("last_precision", C.c_int32),
                ("onchip_record_bytes", C.c_int32), ("tail_compactions", C.c_int64), ("last_steps_per_poll", C.c_int32),
                ("last_vn_items_per_warp", C.c_int32)]


# every symbol include/qkdldpc.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "qkdldpc_version": (C.c_int, []),
    "qkdldpc_last_error": (C.c_char_p, []),
    "qkdldpc_tally_len": (C.c_int64, [C.c_int32]),
    "qkdldpc_effective_precision": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32]),
    "qkdldpc_device_count": (C.c_int, []),
    "qkdldpc_code_create": (C.c_int, [C.POINTER(_VP), C.c_int32, C.c_int32, C.c_int64, _VP, _VP, C.c_int32,
                                      C.POINTER(Options)]),
    "qkdldpc_code_destroy": (None, [_VP]),
    "qkdldpc_code_set_stream": (C.c_int, [_VP, _VP]),
    "qkdldpc_decode_batch": (C.c_int, [_VP, C.POINTER(Params), C.c_int64, _VP, _VP, _VP, C.c_int32, _VP, C.c_int32,
                                       _VP, C.c_int32, _VP, _VP, _VP, _VP]),
    "qkdldpc_decode_batch_device": (C.c_int, [_VP, C.POINTER(Params), C.c_int64, _VP, _VP, _VP, C.c_int32, _VP,
                                              C.c_int32, _VP, C.c_int32, _VP, _VP, _VP, _VP]),
    "qkdldpc_generate_keys_device": (C.c_int, [_VP, C.c_int64, C.c_double, C.c_uint64, _VP, _VP,
                                               C.POINTER(C.c_double)]),
    "qkdldpc_bench_synthetic": (C.c_int, [_VP, C.POINTER(Params), C.c_int64, C.c_double, C.c_uint64, _VP,
                                          C.POINTER(C.c_double)]),
    "qkdldpc_generate_trial_inputs_device": (C.c_int, [_VP, C.c_int64, _VP, C.c_uint64, C.c_double, _VP, C.c_int32, _VP,
                                                       C.c_int32, _VP, _VP, C.POINTER(C.c_double)]),
    "qkdldpc_run_trials": (C.c_int, [_VP, C.POINTER(Params), C.c_int64, _VP, C.c_uint64, C.c_double, _VP, C.c_int32, _VP,
                                     C.c_int32, _VP, _VP, _VP, _VP, C.POINTER(C.c_double)]),
    "qkdldpc_run_trials_multi": (C.c_int, [_VP, C.POINTER(Params), C.c_int32, _VP, C.c_int64, _VP, _VP, _VP, _VP, _VP]),
    "qkdldpc_run_trials_multi_keys": (C.c_int, [_VP, C.POINTER(Params), C.c_int32, _VP, C.c_int64, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "qkdldpc_remove_bits": (C.c_int, [_VP, C.c_int64, _VP, _VP, C.c_int32, _VP]),
    "qkdldpc_comm_nccl_version": (C.c_int, []),
    "qkdldpc_comm_get_unique_id": (C.c_int, [_VP]),
    "qkdldpc_comm_init_rank": (C.c_int, [_VP, _VP, C.c_int32, C.c_int32]),
    "qkdldpc_comm_init_all": (C.c_int, [_VP, C.c_int32]),
    "qkdldpc_comm_size": (C.c_int, [_VP]),
    "qkdldpc_tally_allreduce": (C.c_int, [_VP, _VP, C.c_int64]),
    "qkdldpc_tally_allreduce_device": (C.c_int, [_VP, _VP, C.c_int64]),
    "qkdldpc_code_info": (C.c_int, [_VP, C.POINTER(Info)]),
    "qkdldpc_code_set_profiling": (C.c_int, [_VP, C.c_int32]),
    "qkdldpc_onchip_layout_model": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, _VP, _VP, _VP]),
}


def lib() -> C.CDLL:
    """Loads the CUDA library; raises (never falls back) if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C qkd_ldpc_v_b200/csrc` (or __graft_entry__.build()). "
                "qkd_ldpc_v_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


class QkdLdpcError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = lib().qkdldpc_last_error().decode(errors="replace")
        super().__init__(f"{where} failed with status {code}: {msg}")


def check(rc: int, where: str) -> None:
    if rc != 0:
        raise QkdLdpcError(rc, where)
