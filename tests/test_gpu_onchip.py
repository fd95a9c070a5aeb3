"""GPU: the on-chip min-sum path (frame state in shared memory, onchip_minsum.cuh) against the f32 oracle and the
streaming float32 kernels: all three perform the same float operations in the same order, so per-frame iteration
counts, flags, decoded words and tallies must be IDENTICAL -- including the edge cases of the iteration accounting
(quirks Q9/Q10), tiny iteration limits, clamp off, rate adaptation, and ragged batches."""
import numpy as np
import pytest

import util
from oracle import cpu

pytestmark = pytest.mark.gpu

FACT = {2: (0.75, 0.0), 3: (0.3, 0.0), 4: (0.8, 0.6), 5: (0.3, 0.9)}


@pytest.fixture(scope="module")
def q(built):
    import qkd_ldpc_v_b200 as q
    return q


_handles = {}


def handle(q, name, **kw):
    key = (name, tuple(sorted(kw.items())))
    if key not in _handles:
        a = util.code_arrays(name)
        _handles[key] = q.LdpcCode(a["n"], a["m"], a["row_ptr"], a["col_idx"], device=0, **kw)
    return _handles[key]


def keys(name, seed, frames, qber):
    from qkd_ldpc_v_b200 import hostlib
    arr = util.code_arrays(name)
    return hostlib.gen_keys(hostlib.trial_seeds(seed, frames), arr["n"], qber)


def same(r1, r2):
    assert (r1.iterations_num == r2.iterations_num).all()
    assert (r1.flags == r2.flags).all()
    assert (r1.bob_solution == r2.bob_solution).all()
    assert (r1.tally == r2.tally).all()


@pytest.mark.parametrize("alg", [2, 3, 4, 5])
@pytest.mark.parametrize("name,qber,frames", [("K1_5", 0.02, 700), ("K1_4", 0.035, 500), ("K1_3", 0.06, 300), ("A79", 0.021, 400),
                                              ("I80", 0.017, 300), ("I65", 0.03, 200), ("N100", 0.03, 257), ("N6", 0.2, 33),
                                              ("K1_hi", 0.004, 600), ("I50", 0.07, 120)])
def test_onchip_equals_streaming_and_oracle(q, alg, name, qber, frames):
    arr = util.code_arrays(name)
    a, b, acc = keys(name, 99 + alg, frames, qber)
    cfg = q.DecoderConfig(decoding_algorithm=alg, message_precision=32)
    ro = handle(q, name, decoder_path=2).QKD_LDPC_batch(a, b, acc, FACT[alg], cfg)
    rs = handle(q, name, decoder_path=1).QKD_LDPC_batch(a, b, acc, FACT[alg], cfg)
    assert ro.info["last_path"] == 2 and rs.info["last_path"] == 1
    same(ro, rs)
    if arr["n"] <= 1024 or alg == 2:   # the oracle is a scalar CPU loop: keep the big codes to one algorithm
        ab, bb = q.unpack_bits(a, arr["n"]), q.unpack_bits(b, arr["n"])
        it, fl, bits = cpu.qkd_ldpc_batch(util.oracle_code(name), alg, ab, bb, acc, primary=FACT[alg][0], secondary=FACT[alg][1],
                                          precision=32)
        assert (ro.iterations_num == it).all() and (ro.flags == fl).all() and (ro.bits() == bits).all()
    if name not in ("N6", "I65", "I50"):
        assert ro.syndromes_match.any(), "operating point should converge at least sometimes"


@pytest.mark.parametrize("alg", [2, 3, 4, 5])
@pytest.mark.parametrize("name,qber,frames", [("K1_5", 0.02, 300), ("A79", 0.021, 300), ("A82", 0.0161, 300), ("I80", 0.017, 300),
                                              ("I65", 0.03, 150), ("I50", 0.07, 120), ("N6", 0.2, 33)])
def test_record_formats_agree(q, alg, name, qber, frames):
    """float32 on-chip min-sum: 8-byte records {c1, signs | argmin} + c2 array (the default when every row fits one record of 27
    edges; on request rows of 28..51 edges own two) against one 16-byte record per row -- the same arithmetic, so identical
    results."""
    a, b, acc = keys(name, 177 + alg, frames, qber)
    cfg = q.DecoderConfig(decoding_algorithm=alg, message_precision=32)
    r0 = handle(q, name, decoder_path=2).QKD_LDPC_batch(a, b, acc, FACT[alg], cfg)
    r8 = handle(q, name, decoder_path=2, onchip_record_bytes=8).QKD_LDPC_batch(a, b, acc, FACT[alg], cfg)
    r16 = handle(q, name, decoder_path=2, onchip_record_bytes=16).QKD_LDPC_batch(a, b, acc, FACT[alg], cfg)
    dc_max = int(np.diff(util.code_arrays(name)["row_ptr"]).max())
    assert r0.info["onchip_record_bytes"] == (8 if dc_max <= 27 else 16)
    assert r8.info["onchip_record_bytes"] == (8 if dc_max <= 51 else 16) and r16.info["onchip_record_bytes"] == 16
    same(r8, r16)
    same(r0, r16)


@pytest.mark.parametrize("alg", [2, 3, 4, 5])
@pytest.mark.parametrize("max_iter", [1, 2, 3, 7])
def test_iteration_limits(q, alg, max_iter):
    """Q9/Q10: a frame that needs exactly max_iter iterations; the adaptive variants never test the last decision."""
    a, b, acc = keys("K1_5", 5, 600, 0.022)
    cfg = q.DecoderConfig(decoding_algorithm=alg, message_precision=32, max_iterations=max_iter)
    ro = handle(q, "K1_5", decoder_path=2).QKD_LDPC_batch(a, b, acc, FACT[alg], cfg)
    rs = handle(q, "K1_5", decoder_path=1).QKD_LDPC_batch(a, b, acc, FACT[alg], cfg)
    same(ro, rs)
    arr = util.code_arrays("K1_5")
    it, fl, bits = cpu.qkd_ldpc_batch(util.oracle_code("K1_5"), alg, q.unpack_bits(a, arr["n"]), q.unpack_bits(b, arr["n"]), acc,
                                      max_iter=max_iter, primary=FACT[alg][0], secondary=FACT[alg][1], precision=32)
    assert (ro.iterations_num == it).all() and (ro.flags == fl).all() and (ro.bits() == bits).all()
    assert ro.iterations_num.max() <= max_iter


@pytest.mark.parametrize("alg,fac", [(2, (0.9, 0.0)), (4, (1.0, 0.5)), (3, (0.2, 0.0))])
def test_clamp_disabled(q, alg, fac):
    a, b, acc = keys("K1_4", 17, 300, 0.03)
    cfg = q.DecoderConfig(decoding_algorithm=alg, message_precision=32, enable_msg_llr_threshold=False)
    ro = handle(q, "K1_4", decoder_path=2).QKD_LDPC_batch(a, b, acc, fac, cfg)
    rs = handle(q, "K1_4", decoder_path=1).QKD_LDPC_batch(a, b, acc, fac, cfg)
    same(ro, rs)
    arr = util.code_arrays("K1_4")
    it, fl, bits = cpu.qkd_ldpc_batch(util.oracle_code("K1_4"), alg, q.unpack_bits(a, arr["n"]), q.unpack_bits(b, arr["n"]), acc,
                                      primary=fac[0], secondary=fac[1], precision=32, enable_thr=False)
    assert (ro.iterations_num == it).all() and (ro.flags == fl).all() and (ro.bits() == bits).all()


def test_small_threshold_and_per_frame_qber(q):
    """A clamp that actually bites (threshold 3) and one QBER per frame."""
    a, b, acc = keys("K1_5", 23, 400, 0.02)
    qb = np.full(400, acc)
    qb[::3] *= 1.5
    cfg = q.DecoderConfig(decoding_algorithm=2, message_precision=32, msg_llr_threshold=3.0)
    ro = handle(q, "K1_5", decoder_path=2).QKD_LDPC_batch(a, b, qb, (0.8, 0), cfg)
    rs = handle(q, "K1_5", decoder_path=1).QKD_LDPC_batch(a, b, qb, (0.8, 0), cfg)
    same(ro, rs)


@pytest.mark.parametrize("alg", [2, 5])
def test_rate_adaptation_paths_agree(q, tmp_path, alg):
    from qkd_ldpc_v_b200 import hostlib
    path = str(tmp_path / "I80.mtrx")
    util.write_sparse2(path, "I80")
    arr = util.code_arrays("I80")
    hm = hostlib.HostMatrix(path, 3)
    p, s, _, _ = hm.adapt_code_rate(5555, 0.0116, 0.09, 1.5, untainted=True, untp=arr["untp"])
    assert p.size and s.size
    seeds = hostlib.trial_seeds(31337, 200)
    a, b, acc = hostlib.gen_keys_rate_adapt(seeds, arr["n"], 0.0116, p, s)
    cfg = q.DecoderConfig(decoding_algorithm=alg, message_precision=32)
    ro = handle(q, "I80", decoder_path=2).QKD_LDPC_batch(a, b, acc, FACT[alg], cfg, punctured_bits=p, shortened_bits=s)
    rs = handle(q, "I80", decoder_path=1).QKD_LDPC_batch(a, b, acc, FACT[alg], cfg, punctured_bits=p, shortened_bits=s)
    same(ro, rs)
    assert ro.syndromes_match.mean() > 0.5
    # and a second batch WITHOUT positions on the same handle must not see stale masks
    a2, b2, acc2 = hostlib.gen_keys(seeds[:64], arr["n"], 0.015)
    same(handle(q, "I80", decoder_path=2).QKD_LDPC_batch(a2, b2, acc2, FACT[alg], cfg),
         handle(q, "I80", decoder_path=1).QKD_LDPC_batch(a2, b2, acc2, FACT[alg], cfg))


def test_ineligible_requests(q):
    """n = 100k codes and float64 sum-product cannot run on chip, nor can the float32 sum-product variants on a code whose
    per-edge messages exceed the shared memory (n = 10240, E = 60430): an explicit request fails, the automatic choice
    streams. (Rows of 33..64 edges are fine: they take two records, see K1_hi in the parity test above.)"""
    from qkd_ldpc_v_b200._cabi import QkdLdpcError
    a, b, acc = keys("L100k", 3, 8, 0.06)
    cfg = q.DecoderConfig(decoding_algorithm=2, message_precision=32, max_iterations=20)
    with pytest.raises(QkdLdpcError):
        handle(q, "L100k", decoder_path=2).QKD_LDPC_batch(a, b, acc, (0.72, 0), cfg)
    r = handle(q, "L100k").QKD_LDPC_batch(a, b, acc, (0.72, 0), cfg)
    assert r.info["last_path"] == 1
    a, b, acc = keys("K1_5", 3, 40, 0.02)
    c = q.DecoderConfig(decoding_algorithm=0, message_precision=64)
    with pytest.raises(QkdLdpcError):
        handle(q, "K1_5", decoder_path=2).QKD_LDPC_batch(a, b, acc, (0.8, 0), c)
    assert handle(q, "K1_5").QKD_LDPC_batch(a, b, acc, (0.8, 0), c).info["last_path"] == 1
    # negative scaling factors: the branch-free kernels (on-chip, FAST streaming) clamp magnitudes only, so they are
    # refused there and the EXACT streaming kernels reproduce threshold_matrix on both sides
    c = q.DecoderConfig(decoding_algorithm=2, message_precision=32)
    with pytest.raises(QkdLdpcError):
        handle(q, "K1_5", decoder_path=2).QKD_LDPC_batch(a, b, acc, (-0.8, 0), c)
    r = handle(q, "K1_5").QKD_LDPC_batch(a, b, acc, (-0.8, 0), c)
    arr = util.code_arrays("K1_5")
    it, fl, bits = cpu.qkd_ldpc_batch(util.oracle_code("K1_5"), 2, q.unpack_bits(a, arr["n"]), q.unpack_bits(b, arr["n"]), acc,
                                      primary=-0.8, precision=32)
    assert r.info["last_path"] == 1 and (r.iterations_num == it).all() and (r.flags == fl).all() and (r.bits() == bits).all()
    a, b, acc = keys("I80", 3, 40, 0.015)
    c = q.DecoderConfig(decoding_algorithm=0, message_precision=32, max_iterations=30)
    with pytest.raises(QkdLdpcError):
        handle(q, "I80", decoder_path=2).QKD_LDPC_batch(a, b, acc, (0, 0), c)
    assert handle(q, "I80").QKD_LDPC_batch(a, b, acc, (0, 0), c).info["last_path"] == 1


@pytest.mark.parametrize("threads", [32, 64, 256, 512])
def test_cta_sizes_agree(q, threads):
    a, b, acc = keys("K1_5", 8, 300, 0.02)
    cfg = q.DecoderConfig(decoding_algorithm=4, message_precision=32)
    same(handle(q, "K1_5", decoder_path=2, onchip_threads=threads).QKD_LDPC_batch(a, b, acc, FACT[4], cfg),
         handle(q, "K1_5", decoder_path=2).QKD_LDPC_batch(a, b, acc, FACT[4], cfg))


def test_out_bits_optional_and_single_frame(q):
    a, b, acc = keys("A79", 77, 1, 0.02)
    cfg = q.DecoderConfig(decoding_algorithm=2, message_precision=32)
    r1 = handle(q, "A79", decoder_path=2).QKD_LDPC_batch(a, b, acc, (0.71, 0), cfg, want_bits=False)
    r2 = handle(q, "A79", decoder_path=1).QKD_LDPC_batch(a, b, acc, (0.71, 0), cfg)
    assert r1.iterations_num[0] == r2.iterations_num[0] and r1.flags[0] == r2.flags[0] == 3


# ---- float64 state on chip (onchip_minsum64.cuh): bit-identical to the streaming float64 kernels and to the f64 oracle (= the
# ---- reference, tests/test_oracle_vs_ref.py) -----------------------------------------------------------------------------------

@pytest.mark.parametrize("alg", [2, 3, 4, 5])
@pytest.mark.parametrize("name,qber,frames", [("K1_5", 0.02, 700), ("K1_4", 0.035, 500), ("K1_3", 0.06, 300), ("A79", 0.021, 300),
                                              ("I80", 0.017, 300), ("I65", 0.03, 200), ("N100", 0.03, 257), ("N6", 0.2, 33),
                                              ("K1_hi", 0.004, 600), ("I50", 0.07, 120)])
def test_onchip64_equals_streaming64_and_oracle(q, alg, name, qber, frames):
    arr = util.code_arrays(name)
    a, b, acc = keys(name, 199 + alg, frames, qber)
    cfg = q.DecoderConfig(decoding_algorithm=alg, message_precision=64)
    ro = handle(q, name, decoder_path=2).QKD_LDPC_batch(a, b, acc, FACT[alg], cfg)
    rs = handle(q, name, decoder_path=1).QKD_LDPC_batch(a, b, acc, FACT[alg], cfg)
    assert ro.info["last_path"] == 2 and rs.info["last_path"] == 1 and ro.info["last_precision"] == 64
    same(ro, rs)
    if arr["n"] <= 1024 or alg == 5:   # the oracle is a scalar CPU loop: keep the big codes to one algorithm
        ab, bb = q.unpack_bits(a, arr["n"]), q.unpack_bits(b, arr["n"])
        it, fl, bits = cpu.qkd_ldpc_batch(util.oracle_code(name), alg, ab, bb, acc, primary=FACT[alg][0], secondary=FACT[alg][1],
                                          precision=64)
        assert (ro.iterations_num == it).all() and (ro.flags == fl).all() and (ro.bits() == bits).all()


@pytest.mark.parametrize("alg", [2, 3, 4, 5])
@pytest.mark.parametrize("max_iter", [1, 2, 3, 7])
def test_onchip64_iteration_limits(q, alg, max_iter):
    a, b, acc = keys("K1_5", 5, 600, 0.022)
    cfg = q.DecoderConfig(decoding_algorithm=alg, message_precision=64, max_iterations=max_iter)
    ro = handle(q, "K1_5", decoder_path=2).QKD_LDPC_batch(a, b, acc, FACT[alg], cfg)
    arr = util.code_arrays("K1_5")
    it, fl, bits = cpu.qkd_ldpc_batch(util.oracle_code("K1_5"), alg, q.unpack_bits(a, arr["n"]), q.unpack_bits(b, arr["n"]), acc,
                                      max_iter=max_iter, primary=FACT[alg][0], secondary=FACT[alg][1], precision=64)
    assert ro.info["last_path"] == 2
    assert (ro.iterations_num == it).all() and (ro.flags == fl).all() and (ro.bits() == bits).all()


@pytest.mark.parametrize("alg,fac", [(2, (0.9, 0.0)), (4, (1.0, 0.5)), (3, (0.2, 0.0)), (5, (0.2, 0.7))])
def test_onchip64_clamp_disabled_and_small_threshold(q, alg, fac):
    a, b, acc = keys("K1_4", 17, 300, 0.03)
    arr = util.code_arrays("K1_4")
    ab, bb = q.unpack_bits(a, arr["n"]), q.unpack_bits(b, arr["n"])
    for kw_cfg, kw_or in ((dict(enable_msg_llr_threshold=False), dict(enable_thr=False)), (dict(msg_llr_threshold=3.0), dict(thr=3.0))):
        cfg = q.DecoderConfig(decoding_algorithm=alg, message_precision=64, **kw_cfg)
        ro = handle(q, "K1_4", decoder_path=2).QKD_LDPC_batch(a, b, acc, fac, cfg)
        it, fl, bits = cpu.qkd_ldpc_batch(util.oracle_code("K1_4"), alg, ab, bb, acc, primary=fac[0], secondary=fac[1], precision=64, **kw_or)
        assert (ro.iterations_num == it).all() and (ro.flags == fl).all() and (ro.bits() == bits).all()


@pytest.mark.parametrize("alg", [3, 5])
def test_onchip64_rate_adaptation(q, tmp_path, alg):
    from qkd_ldpc_v_b200 import hostlib
    path = str(tmp_path / "I80.mtrx")
    util.write_sparse2(path, "I80")
    arr = util.code_arrays("I80")
    hm = hostlib.HostMatrix(path, 3)
    p, s, _, _ = hm.adapt_code_rate(5555, 0.0116, 0.09, 1.5, untainted=True, untp=arr["untp"])
    seeds = hostlib.trial_seeds(31337, 200)
    a, b, acc = hostlib.gen_keys_rate_adapt(seeds, arr["n"], 0.0116, p, s)
    cfg = q.DecoderConfig(decoding_algorithm=alg, message_precision=64)
    ro = handle(q, "I80", decoder_path=2).QKD_LDPC_batch(a, b, acc, FACT[alg], cfg, punctured_bits=p, shortened_bits=s)
    rs = handle(q, "I80", decoder_path=1).QKD_LDPC_batch(a, b, acc, FACT[alg], cfg, punctured_bits=p, shortened_bits=s)
    same(ro, rs)
    assert ro.syndromes_match.mean() > 0.5


def test_precision_policy(q):
    """message_precision = 0: float64 state for OMSA / ANMSA / AOMSA (on chip), float32 for NMSA and the sum-product variants of
    the n <= 65536 codes, float64 for sum-product on n = 102400 (include/qkdldpc.h)."""
    from qkd_ldpc_v_b200 import _cabi
    L = _cabi.lib()
    assert [L.qkdldpc_effective_precision(alg, 10240, 0) for alg in range(6)] == [32, 32, 32, 64, 64, 64]
    assert [L.qkdldpc_effective_precision(alg, 102400, 0) for alg in range(6)] == [64, 64, 32, 64, 64, 64]
    assert L.qkdldpc_effective_precision(5, 10240, 32) == 32 and L.qkdldpc_effective_precision(0, 10240, 64) == 64
    a, b, acc = keys("K1_5", 8, 200, 0.02)
    for alg, prec in ((2, 32), (3, 64), (4, 64), (5, 64), (0, 32)):
        r = handle(q, "K1_5").QKD_LDPC_batch(a, b, acc, FACT.get(alg, (0, 0)), q.DecoderConfig(decoding_algorithm=alg))
        assert r.info["last_precision"] == prec and r.info["last_path"] == 2
        rx = handle(q, "K1_5").QKD_LDPC_batch(a, b, acc, FACT.get(alg, (0, 0)), q.DecoderConfig(decoding_algorithm=alg, message_precision=prec))
        same(r, rx)


def test_tail_compaction_with_capped_grid(q):
    """More moves than the copy kernels have CTAs (the grid cap forced down to 3): every straggler must still be moved."""
    a, b, acc = keys("K1_5", 77, 3000, 0.024)
    cfg = q.DecoderConfig(decoding_algorithm=2, message_precision=32)
    r0 = handle(q, "K1_5", decoder_path=1, tail_compaction=-1).QKD_LDPC_batch(a, b, acc, (0.75, 0), cfg)
    r1 = handle(q, "K1_5", decoder_path=1, compaction_max_ctas=3, steps_per_poll=2).QKD_LDPC_batch(a, b, acc, (0.75, 0), cfg)
    same(r0, r1)


@pytest.mark.parametrize("alg,prec", [(2, 32), (5, 64), (0, 32)])
def test_pipelined_host_batches(q, alg, prec):
    """qkdldpc_decode_batch cuts a host batch into pieces (copy-in of piece k+1 / copy-out of piece k-1 overlap the decoding
    of piece k; each piece is its own launch with its own frame queue, tallies are shared): per-frame results and tallies
    must not depend on the number of pieces -- ragged piece sizes, one QBER per frame, no decoded words requested."""
    frames = 5003
    a, b, acc = keys("K1_5", 4242, frames, 0.021)
    qb = np.full(frames, acc)
    qb[::5] *= 1.3
    cfg = q.DecoderConfig(decoding_algorithm=alg, message_precision=prec, max_iterations=50)
    fac = FACT.get(alg, (0.0, 0.0))
    r1 = handle(q, "K1_5", copy_chunks=1).QKD_LDPC_batch(a, b, qb, fac, cfg)
    for chunks in (0, 3, 16):
        rk = handle(q, "K1_5", copy_chunks=chunks).QKD_LDPC_batch(a, b, qb, fac, cfg)
        assert rk.info["last_path"] == 2
        same(r1, rk)
        rn = handle(q, "K1_5", copy_chunks=chunks).QKD_LDPC_batch(a, b, qb, fac, cfg, want_bits=False)
        assert (rn.iterations_num == r1.iterations_num).all() and (rn.flags == r1.flags).all() and (rn.tally == r1.tally).all()
    rs = handle(q, "K1_5", decoder_path=1).QKD_LDPC_batch(a, b, qb, fac, cfg)      # streaming: copy in, decode, copy out
    same(r1, rs)
    p = np.arange(0, 40, dtype=np.int32)                                            # rate adaptation through the pipeline
    r2 = handle(q, "K1_5", copy_chunks=5).QKD_LDPC_batch(a, b, acc, fac, cfg, punctured_bits=p)
    r3 = handle(q, "K1_5", decoder_path=1).QKD_LDPC_batch(a, b, acc, fac, cfg, punctured_bits=p)
    same(r2, r3)
