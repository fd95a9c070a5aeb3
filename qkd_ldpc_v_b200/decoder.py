"""Host-side mirror of the reference's decoding interface over the C ABI.

Reference surface mirrored here (names, argument meaning, result fields):
  * ``H_matrix`` (array_and_matrix_operations.hpp:60-77)                      -> :class:`LdpcCode`
  * ``CFG.DECODING_ALGORITHM / ..MAX_ITERATIONS / ..MSG_LLR_THRESHOLD``       -> :class:`DecoderConfig`
  * ``decoding_scaling_factors`` (config.hpp:50-54)                           -> ``scaling_factors=(primary, secondary)``
  * ``H_matrix_params.punctured_bits / shortened_bits`` (.hpp:44-48)          -> ``punctured_bits= / shortened_bits=``
  * ``QKD_LDPC`` / ``QKD_LDPC_RATE_ADAPT`` (qkd_ldpc_algorithm.cpp:1031,1121) over the trial loop
    (simulation.cpp:740-746)                                                  -> :meth:`LdpcCode.QKD_LDPC_batch`
  * ``LDPC_result{decoding_res{iterations_num, syndromes_match}, keys_match}``-> :class:`BatchResult`

Everything is computed by libqkdldpc_cuda on the GPU; there is no CPU path in this package.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _cabi
from ._cabi import Options, Params

DEC_SPA, DEC_SPA_APPROX, DEC_NMSA, DEC_OMSA, DEC_ANMSA, DEC_AOMSA = range(6)  # config.hpp:201
ALG_NAMES = ("SPA", "SPA_LIN_APPROX", "NMSA", "OMSA", "ANMSA", "AOMSA")

FLAG_SYNDROMES_MATCH = 1
FLAG_KEYS_MATCH = 2
TALLY_FRAMES, TALLY_SYNDROMES_MATCH, TALLY_KEYS_MATCH, TALLY_ITERATIONS, TALLY_HIST = 0, 1, 2, 3, 4


def pack_bits(bits) -> np.ndarray:
    """[F][n] 0/1 -> packed uint32 [F][ceil(n/32)], bit i of a frame at word i>>5, position i&31."""
    bits = np.ascontiguousarray(bits, np.uint8)
    if bits.ndim == 1:
        bits = bits[None, :]
    f, n = bits.shape
    w = (n + 31) // 32
    padded = np.zeros((f, w * 32), np.uint8)
    padded[:, :n] = bits
    return np.packbits(padded, axis=1, bitorder="little").view("<u4").reshape(f, w)


def unpack_bits(words, n: int) -> np.ndarray:
    words = np.ascontiguousarray(words, "<u4")
    f = words.shape[0]
    return np.unpackbits(words.view(np.uint8).reshape(f, -1), axis=1, bitorder="little")[:, :n]


@dataclass
class DecoderConfig:
    """The fields of the reference's global ``CFG`` that the hot path reads (config.hpp:103-198)."""
    decoding_algorithm: int = DEC_NMSA
    max_iterations: int = 100
    enable_msg_llr_threshold: bool = True
    msg_llr_threshold: float = 100.0
    # 0 = the library's precision policy (float64 state for OMSA / ANMSA / AOMSA and for SPA on n > 65536, float32
    # otherwise: include/qkdldpc.h); 32 / 64 force float32 / float64 messages
    message_precision: int = 0

    def to_params(self, scaling_factors=(0.0, 0.0)) -> Params:
        pri, sec = (tuple(scaling_factors) + (0.0, 0.0))[:2]
        return Params(int(self.decoding_algorithm), int(self.max_iterations), float(pri), float(sec),
                      int(bool(self.enable_msg_llr_threshold)), float(self.msg_llr_threshold),
                      int(self.message_precision))


@dataclass
class BatchResult:
    iterations_num: np.ndarray                 # int32 [F]   decoding_result.iterations_num
    flags: np.ndarray                          # uint8 [F]
    bob_solution: Optional[np.ndarray]         # packed uint32 [F][W] (last hard decision) or None
    tally: np.ndarray                          # uint64 [max_iter + 5]
    n: int = 0
    device_ms: float = 0.0
    info: dict = field(default_factory=dict)

    @property
    def syndromes_match(self) -> np.ndarray:
        return (self.flags & FLAG_SYNDROMES_MATCH) != 0

    @property
    def keys_match(self) -> np.ndarray:
        return (self.flags & FLAG_KEYS_MATCH) != 0

    def bits(self) -> np.ndarray:
        return unpack_bits(self.bob_solution, self.n)


def stats_from_tally(tally: np.ndarray, trials: int) -> dict:
    """process_trials_results (simulation.cpp:580-624,683-689) from the all-reducible tally vector."""
    hist = np.asarray(tally[TALLY_HIST:], np.float64)
    ok = int(tally[TALLY_SYNDROMES_MATCH])
    its = np.arange(hist.size, dtype=np.float64)
    out = dict(iter_success_max=0, iter_success_min=0, iter_success_mean=0.0, iter_success_std=0.0)
    if ok > 0:
        nz = np.nonzero(hist)[0]
        mean = float((hist * its).sum() / ok)
        out.update(iter_success_max=int(nz.max()), iter_success_min=int(nz.min()), iter_success_mean=mean,
                   iter_success_std=float(np.sqrt((hist * (its - mean) ** 2).sum() / ok)))
    out["ratio_trials_success_dec_alg"] = ok / trials
    out["ratio_trials_success_ldpc"] = int(tally[TALLY_KEYS_MATCH]) / trials
    out["FER"] = round((1.0 - out["ratio_trials_success_ldpc"]) * trials) / trials   # simulation.cpp:117-118
    return out


class LdpcCode:
    """One parity-check matrix resident on one GPU (handle of ``qkdldpc_code_create``)."""

    def __init__(self, n: int, m: int, row_ptr, col_idx, device: int = 0, pool_bytes: int = 0, pool_slots: int = 0,
                 steps_per_poll: int = 0, frames_per_lane_f32: int = 0, use_graph: int = 0, decoder_path: int = 0,
                 onchip_threads: int = 0, tail_compaction: int = 0, compaction_max_ctas: int = 0, copy_chunks: int = 0,
                 onchip_record_bytes: int = 0, vn_items_per_warp: int = 0, vn_ctas_per_sm: int = 0,
                 compaction_fill_pct: int = 0):
        L = _cabi.lib()
        self.n, self.m = int(n), int(m)
        self.row_ptr = np.ascontiguousarray(row_ptr, np.int32)
        self.col_idx = np.ascontiguousarray(col_idx, np.int32)
        self.nnz = int(self.row_ptr[-1]) if self.row_ptr.size else 0
        self.words = (self.n + 31) // 32
        self.device = device
        # decoder_path: 0 auto, 1 streaming kernels (messages in HBM), 2 on-chip min-sum (frame state in shared memory)
        opt = Options(int(pool_bytes), int(pool_slots), int(steps_per_poll), int(frames_per_lane_f32), int(use_graph),
                      int(decoder_path), int(onchip_threads), int(tail_compaction), int(compaction_max_ctas), int(copy_chunks),
                      int(onchip_record_bytes), int(vn_items_per_warp), int(vn_ctas_per_sm), int(compaction_fill_pct))
        h = C.c_void_p()
        _cabi.check(L.qkdldpc_code_create(C.byref(h), self.n, self.m, self.nnz, self.row_ptr.ctypes.data,
                                          self.col_idx.ctypes.data, device, C.byref(opt)), "qkdldpc_code_create")
        self._h = h

    # -- lifetime -------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            _cabi.lib().qkdldpc_code_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- plumbing -------------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr: int):
        _cabi.check(_cabi.lib().qkdldpc_code_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), "set_stream")

    def set_profiling(self, on: bool):
        _cabi.check(_cabi.lib().qkdldpc_code_set_profiling(self._h, int(on)), "set_profiling")

    def info(self) -> dict:
        i = _cabi.Info()
        _cabi.check(_cabi.lib().qkdldpc_code_info(self._h, C.byref(i)), "code_info")
        return {k: getattr(i, k) for k, _ in _cabi.Info._fields_}

    @staticmethod
    def _poslist(x):
        a = np.ascontiguousarray(x if x is not None else [], np.int32)
        return a, (a.ctypes.data if a.size else None), int(a.size)

    # -- multi-GPU: the tally all-reduce (the handle owns the NCCL communicator) ------------------------------
    @staticmethod
    def comm_unique_id() -> np.ndarray:
        """128 bytes from ``ncclGetUniqueId`` (rank 0 calls this, the caller broadcasts them)."""
        ident = np.zeros(128, np.uint8)
        _cabi.check(_cabi.lib().qkdldpc_comm_get_unique_id(ident.ctypes.data), "qkdldpc_comm_get_unique_id")
        return ident

    def comm_init_rank(self, unique_id, n_ranks: int, rank: int):
        ident = np.ascontiguousarray(unique_id, np.uint8)
        assert ident.size == 128
        _cabi.check(_cabi.lib().qkdldpc_comm_init_rank(self._h, ident.ctypes.data, int(n_ranks), int(rank)), "qkdldpc_comm_init_rank")

    @staticmethod
    def comm_init_all(codes):
        """One process, several devices: one communicator over ``codes`` (one handle per distinct device)."""
        arr = (C.c_void_p * len(codes))(*[c._h for c in codes])
        _cabi.check(_cabi.lib().qkdldpc_comm_init_all(C.cast(arr, C.c_void_p), len(codes)), "qkdldpc_comm_init_all")

    def comm_size(self) -> int:
        return int(_cabi.lib().qkdldpc_comm_size(self._h))

    def tally_allreduce(self, tally: np.ndarray) -> np.ndarray:
        """Sum of a HOST tally vector over all ranks, in place (``qkdldpc_tally_allreduce``)."""
        t = np.ascontiguousarray(tally, np.uint64)
        _cabi.check(_cabi.lib().qkdldpc_tally_allreduce(self._h, t.ctypes.data, t.size), "qkdldpc_tally_allreduce")
        return t

    def tally_allreduce_device(self, d_tally: int, count: int):
        """Same on a DEVICE vector, enqueued on the handle's stream behind the decode that filled it."""
        _cabi.check(_cabi.lib().qkdldpc_tally_allreduce_device(self._h, C.c_void_p(d_tally), int(count)), "qkdldpc_tally_allreduce_device")

    # -- the hot path ---------------------------------------------------------------------------------------
    def QKD_LDPC_batch(self, alice_bit_array, bob_bit_array, QBER, scaling_factors=(0.0, 0.0),
                       cfg: Optional[DecoderConfig] = None, punctured_bits: Sequence[int] = (),
                       shortened_bits: Sequence[int] = (), want_bits: bool = True, out=None) -> BatchResult:
        """Batched ``QKD_LDPC`` / ``QKD_LDPC_RATE_ADAPT`` with HOST buffers (copies inside).

        alice_bit_array / bob_bit_array: packed uint32 [F][W] (or 0/1 arrays [F][n], packed here). With
        punctured / shortened lists the frames must already be the EXTENDED ones. QBER: scalar or [F].
        out: optional (bits uint32 [F][W], iterations int32 [F], flags uint8 [F]) arrays to receive the results, e.g.
        views of pinned memory (the default allocates fresh pageable arrays on every call).
        """
        cfg = cfg or DecoderConfig()
        L = _cabi.lib()
        a = self._as_packed(alice_bit_array)
        b = self._as_packed(bob_bit_array)
        if a.shape != b.shape:
            raise ValueError("alice and bob batches differ in shape")
        F = a.shape[0]
        q = np.ascontiguousarray(np.atleast_1d(np.asarray(QBER, np.float64)))
        scalar = int(q.size == 1)
        if not scalar and q.size != F:
            raise ValueError("QBER must be a scalar or one value per frame")
        p = cfg.to_params(scaling_factors)
        pa, pp, np_ = self._poslist(punctured_bits)
        sa, sp, ns_ = self._poslist(shortened_bits)
        if out is not None:
            out_bits, iters, flags = out
            want_bits = out_bits is not None
            if (want_bits and (out_bits.dtype != np.uint32 or out_bits.shape != (F, self.words) or not out_bits.flags.c_contiguous)) or \
                    iters.dtype != np.int32 or iters.shape != (F,) or flags.dtype != np.uint8 or flags.shape != (F,):
                raise ValueError("out buffers must be (uint32 [F][W] or None, int32 [F], uint8 [F])")
        else:
            out_bits = np.zeros((F, self.words), np.uint32) if want_bits else None
            iters = np.zeros(F, np.int32)
            flags = np.zeros(F, np.uint8)
        tally = np.zeros(int(L.qkdldpc_tally_len(p.max_iterations)), np.uint64)
        _cabi.check(L.qkdldpc_decode_batch(self._h, C.byref(p), F, a.ctypes.data, b.ctypes.data, q.ctypes.data, scalar,
                                           pp, np_, sp, ns_, out_bits.ctypes.data if want_bits else None,
                                           iters.ctypes.data, flags.ctypes.data, tally.ctypes.data),
                    "qkdldpc_decode_batch")
        del pa, sa
        inf = self.info()
        return BatchResult(iters, flags, out_bits, tally, self.n, inf["last_batch_ms"], inf)

    def QKD_LDPC_RATE_ADAPT_batch(self, alice_bit_array_extended, bob_bit_array_extended, QBER, scaling_factors=(0.0, 0.0),
                                  cfg: Optional[DecoderConfig] = None, punctured_bits: Sequence[int] = (),
                                  shortened_bits: Sequence[int] = (), want_bits: bool = True) -> BatchResult:
        """Batched ``QKD_LDPC_RATE_ADAPT`` (qkd_ldpc_algorithm.cpp:1121-1258) on EXTENDED frames: the same C-ABI call
        as :meth:`QKD_LDPC_batch` with the position lists of ``H_matrix_params`` (sorted ascending)."""
        return self.QKD_LDPC_batch(alice_bit_array_extended, bob_bit_array_extended, QBER, scaling_factors, cfg,
                                   punctured_bits, shortened_bits, want_bits)

    def decode_batch_device(self, d_alice: int, d_bob: int, d_qber: int, n_frames: int, scaling_factors=(0.0, 0.0),
                            cfg: Optional[DecoderConfig] = None, qber_is_scalar: bool = True,
                            punctured_bits: Sequence[int] = (), shortened_bits: Sequence[int] = (),
                            d_out_bits: int = 0, d_out_iters: int = 0, d_out_flags: int = 0, d_tally: int = 0):
        """Same with raw DEVICE pointers (ints), e.g. ``tensor.data_ptr()`` of torch tensors on this GPU."""
        cfg = cfg or DecoderConfig()
        p = cfg.to_params(scaling_factors)
        pa, pp, np_ = self._poslist(punctured_bits)
        sa, sp, ns_ = self._poslist(shortened_bits)
        vp = lambda x: C.c_void_p(x) if x else None  # noqa: E731
        _cabi.check(_cabi.lib().qkdldpc_decode_batch_device(
            self._h, C.byref(p), int(n_frames), vp(d_alice), vp(d_bob), vp(d_qber), int(qber_is_scalar), pp, np_, sp,
            ns_, vp(d_out_bits), vp(d_out_iters), vp(d_out_flags), vp(d_tally)), "qkdldpc_decode_batch_device")
        del pa, sa

    def run_trials(self, trial_seeds, QBER: float, scaling_factors=(0.0, 0.0), cfg: Optional[DecoderConfig] = None,
                   seed_offset: int = 0, punctured_bits: Sequence[int] = (), shortened_bits: Sequence[int] = (),
                   want_bits: bool = True) -> BatchResult:
        """The batched ``run_trial`` (simulation.cpp:540-577): inputs of every trial are generated ON THE DEVICE from
        ``Xoshiro256++(trial_seeds[n] + seed_offset)`` exactly as the reference does on the CPU, then decoded.
        ``QBER`` is the configured QBER; the accurate one is in ``result.info["accurate_qber"]``."""
        cfg = cfg or DecoderConfig()
        L = _cabi.lib()
        seeds = np.ascontiguousarray(trial_seeds, np.uint64)
        F = int(seeds.size)
        p = cfg.to_params(scaling_factors)
        pa, pp, np_ = self._poslist(punctured_bits)
        sa, sp, ns_ = self._poslist(shortened_bits)
        out_bits = np.zeros((F, self.words), np.uint32) if want_bits else None
        iters = np.zeros(F, np.int32)
        flags = np.zeros(F, np.uint8)
        tally = np.zeros(int(L.qkdldpc_tally_len(p.max_iterations)), np.uint64)
        acc = C.c_double()
        _cabi.check(L.qkdldpc_run_trials(self._h, C.byref(p), F, seeds.ctypes.data, int(seed_offset), float(QBER), pp, np_, sp, ns_,
                                         out_bits.ctypes.data if want_bits else None, iters.ctypes.data, flags.ctypes.data,
                                         tally.ctypes.data, C.byref(acc)), "qkdldpc_run_trials")
        del pa, sa
        inf = self.info()
        inf["accurate_qber"] = acc.value
        return BatchResult(iters, flags, out_bits, tally, self.n, inf["last_batch_ms"], inf)

    def run_trials_multi(self, trial_seeds, combinations, cfg: Optional[DecoderConfig] = None, want_keys: bool = False):
        """Several combinations of a sweep in one call (``qkdldpc_run_trials_multi_keys``). ``combinations``: list of dicts with
        keys QBER, primary, secondary, punctured_bits, shortened_bits, seed_offset and, optionally, bits_to_remove
        (``H_matrix_params.bits_to_remove``: ``remove_bits`` then runs on the device as the last step of every trial).
        Returns (iterations [C][T], flags [C][T], tallies [C][tally_len], accurate_qber [C]) and, with ``want_keys``, two
        lists of packed final keys per combination (Alice's, Bob's; None where a combination has no removal list)."""
        cfg = cfg or DecoderConfig()
        L = _cabi.lib()
        seeds = np.ascontiguousarray(trial_seeds, np.uint64)
        T, Cn = int(seeds.size), len(combinations)
        p = cfg.to_params((0.0, 0.0))
        table = (_cabi.Combination * max(Cn, 1))()
        keep = []
        keys_a, keys_b = [None] * Cn, [None] * Cn
        ptr_a, ptr_b = (C.c_void_p * max(Cn, 1))(), (C.c_void_p * max(Cn, 1))()
        for k, cb in enumerate(combinations):
            pa, pp, np_ = self._poslist(cb.get("punctured_bits", ()))
            sa, sp, ns_ = self._poslist(cb.get("shortened_bits", ()))
            ra, rp, nr_ = self._poslist(cb.get("bits_to_remove", ()))
            keep += [pa, sa, ra]
            table[k] = _cabi.Combination(float(cb["QBER"]), float(cb.get("primary", 0.0)), float(cb.get("secondary", 0.0)), pp, np_, sp,
                                         ns_, int(cb.get("seed_offset", 0)), rp, nr_, 0)
            if want_keys and nr_ > 0:
                wo = (self.n - nr_ + 31) // 32
                keys_a[k], keys_b[k] = np.zeros((T, wo), np.uint32), np.zeros((T, wo), np.uint32)
                ptr_a[k], ptr_b[k] = keys_a[k].ctypes.data, keys_b[k].ctypes.data
        tl = int(L.qkdldpc_tally_len(p.max_iterations))
        iters = np.zeros((Cn, T), np.int32)
        flags = np.zeros((Cn, T), np.uint8)
        tallies = np.zeros((Cn, tl), np.uint64)
        acc = np.zeros(Cn, np.float64)
        _cabi.check(L.qkdldpc_run_trials_multi_keys(self._h, C.byref(p), Cn, C.byref(table), T, seeds.ctypes.data, iters.ctypes.data,
                                                    flags.ctypes.data, tallies.ctypes.data, acc.ctypes.data,
                                                    C.cast(ptr_a, C.c_void_p) if want_keys else None,
                                                    C.cast(ptr_b, C.c_void_p) if want_keys else None), "qkdldpc_run_trials_multi_keys")
        del keep
        if want_keys:
            return iters, flags, tallies, acc, keys_a, keys_b
        return iters, flags, tallies, acc

    def generate_trial_inputs_device(self, trial_seeds, qber: float, d_alice: int, d_bob: int, seed_offset: int = 0,
                                     punctured_bits: Sequence[int] = (), shortened_bits: Sequence[int] = ()) -> float:
        """Reference-compatible trial inputs written to DEVICE buffers (``qkdldpc_generate_trial_inputs_device``);
        returns the accurate QBER."""
        seeds = np.ascontiguousarray(trial_seeds, np.uint64)
        pa, pp, np_ = self._poslist(punctured_bits)
        sa, sp, ns_ = self._poslist(shortened_bits)
        acc = C.c_double()
        _cabi.check(_cabi.lib().qkdldpc_generate_trial_inputs_device(
            self._h, int(seeds.size), seeds.ctypes.data, int(seed_offset), float(qber), pp, np_, sp, ns_, C.c_void_p(d_alice),
            C.c_void_p(d_bob), C.byref(acc)), "qkdldpc_generate_trial_inputs_device")
        del pa, sa
        return acc.value

    def remove_bits(self, keys, bits_to_remove) -> np.ndarray:
        """``remove_bits`` (array_and_matrix_operations.cpp:259-287) for a batch of packed frames: returns packed frames
        of ``n - len(bits_to_remove)`` bits."""
        k = self._as_packed(keys)
        r = np.ascontiguousarray(bits_to_remove, np.int32)
        words_out = (self.n - r.size + 31) // 32
        out = np.zeros((k.shape[0], words_out), np.uint32)
        _cabi.check(_cabi.lib().qkdldpc_remove_bits(self._h, k.shape[0], k.ctypes.data, r.ctypes.data if r.size else None, r.size,
                                                    out.ctypes.data if out.size else None), "qkdldpc_remove_bits")
        return out

    def generate_keys_device(self, n_frames: int, qber: float, seed: int, d_alice: int, d_bob: int) -> float:
        acc = C.c_double()
        _cabi.check(_cabi.lib().qkdldpc_generate_keys_device(self._h, int(n_frames), float(qber), int(seed),
                                                             C.c_void_p(d_alice), C.c_void_p(d_bob), C.byref(acc)),
                    "qkdldpc_generate_keys_device")
        return acc.value

    def bench_synthetic(self, n_frames: int, qber: float, scaling_factors, config: "DecoderConfig", seed: int = 1):
        """qkdldpc_bench_synthetic: synthetic keys generated on the device, decoded, timed. Returns (tally, seconds)."""
        P = config.to_params(scaling_factors)
        tally = np.zeros(tally_len(config.max_iterations), np.uint64)
        sec = C.c_double()
        _cabi.check(_cabi.lib().qkdldpc_bench_synthetic(self._h, C.byref(P), int(n_frames), float(qber), int(seed),
                                                        tally.ctypes.data, C.byref(sec)), "qkdldpc_bench_synthetic")
        return tally, sec.value

    def _as_packed(self, x) -> np.ndarray:
        x = np.asarray(x)
        if x.dtype == np.uint32 and x.ndim == 2 and x.shape[1] == self.words:
            return np.ascontiguousarray(x)
        if x.ndim == 1:
            x = x[None, :]
        if x.shape[1] != self.n:
            raise ValueError(f"expected frames of {self.n} bits (or packed {self.words} words), got {x.shape}")
        return pack_bits(x)


def tally_len(max_iterations: int) -> int:
    return int(_cabi.lib().qkdldpc_tally_len(int(max_iterations)))
