"""GPU: trial inputs generated on the device (qkdldpc_generate_trial_inputs_device / qkdldpc_run_trials) must be
bit-identical to the host generator, which drives the reference's own libstdc++ distributions and std::shuffle with the
same xoshiro256++ stream (tests/test_host_rng.py pins the host generator against keys produced by the compiled
reference). Covers even / odd / tiny block lengths, zero and large error counts, the rate-adaptation frame extension and
the batched run_trial."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


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


def device_keys(q, name, seeds, qber, offset=0, punct=(), short=()):
    import torch
    arr = util.code_arrays(name)
    words = (arr["n"] + 31) // 32
    da = torch.zeros((len(seeds), words), dtype=torch.int32, device="cuda:0")
    db = torch.zeros_like(da)
    acc = handle(q, name).generate_trial_inputs_device(seeds, qber, da.data_ptr(), db.data_ptr(), seed_offset=offset,
                                                       punctured_bits=punct, shortened_bits=short)
    torch.cuda.synchronize()
    return da.cpu().numpy().view(np.uint32), db.cpu().numpy().view(np.uint32), acc


@pytest.mark.parametrize("name,qber,frames", [("N6", 0.2, 64), ("N6", 0.5, 40), ("N7", 0.3, 64), ("N7", 0.15, 33), ("N100", 0.03, 200),
                                              ("N100", 0.005, 50), ("K1_5", 0.02, 300), ("K1_5", 0.118, 100), ("K1_3", 0.0009, 64),
                                              ("A79", 0.03, 150), ("A79", 0.002, 64), ("L100k", 0.06, 12)])
def test_device_keys_equal_host_keys(q, name, qber, frames):
    from qkd_ldpc_v_b200 import hostlib
    arr = util.code_arrays(name)
    seeds = hostlib.trial_seeds(20251018, frames)
    ha, hb, hacc = hostlib.gen_keys(seeds, arr["n"], qber)
    da, db, dacc = device_keys(q, name, seeds, qber)
    assert dacc == hacc
    assert (da == ha).all(), "Alice keys differ"
    assert (db == hb).all(), "Bob keys differ"
    if hacc > 0:
        flips = np.unpackbits((da ^ db).view(np.uint8), axis=1).sum(axis=1)
        assert (flips == round(hacc * arr["n"])).all()       # exactly floor(N * QBER) errors per frame


def test_seed_offset_is_added_to_every_seed(q):
    from qkd_ldpc_v_b200 import hostlib
    arr = util.code_arrays("K1_4")
    seeds = hostlib.trial_seeds(5, 100)
    with np.errstate(over="ignore"):
        shifted = seeds + np.uint64(12345)
    ha, hb, _ = hostlib.gen_keys(shifted, arr["n"], 0.03)
    da, db, _ = device_keys(q, "K1_4", seeds, 0.03, offset=12345)
    assert (da == ha).all() and (db == hb).all()


@pytest.mark.parametrize("name,qber,n_p,n_s", [("K1_5", 0.02, 60, 25), ("I80", 0.0196, 278, 29), ("I80", 0.0276, 54, 1072),
                                               ("N100", 0.05, 7, 0), ("N100", 0.05, 0, 9)])
def test_rate_adapted_frames_equal_host(q, name, qber, n_p, n_s):
    from qkd_ldpc_v_b200 import hostlib
    arr = util.code_arrays(name)
    rng = np.random.default_rng(n_p * 131 + n_s)
    pos = rng.permutation(arr["n"])
    p, s = np.sort(pos[:n_p]).astype(np.int32), np.sort(pos[n_p:n_p + n_s]).astype(np.int32)
    seeds = hostlib.trial_seeds(99, 80)
    ha, hb, hacc = hostlib.gen_keys_rate_adapt(seeds, arr["n"], qber, p, s)
    da, db, dacc = device_keys(q, name, seeds, qber, punct=p, short=s)
    assert dacc == hacc
    assert (da == ha).all() and (db == hb).all()


@pytest.mark.parametrize("alg,path", [(2, 0), (0, 0), (5, 1)])
def test_run_trials_equals_decode_of_host_keys(q, alg, path):
    from qkd_ldpc_v_b200 import hostlib
    arr = util.code_arrays("K1_5")
    seeds = hostlib.trial_seeds(777, 500)
    fac = {2: (0.75, 0.0), 0: (0.0, 0.0), 5: (0.3, 0.9)}[alg]
    cfg = q.DecoderConfig(decoding_algorithm=alg, message_precision=32)
    h = handle(q, "K1_5", decoder_path=path)
    with np.errstate(over="ignore"):
        a, b, acc = hostlib.gen_keys(seeds + np.uint64(3), arr["n"], 0.02)
    r_host = h.QKD_LDPC_batch(a, b, acc, fac, cfg)
    r_dev = h.run_trials(seeds, 0.02, fac, cfg, seed_offset=3)
    assert r_dev.info["accurate_qber"] == acc
    assert (r_dev.iterations_num == r_host.iterations_num).all() and (r_dev.flags == r_host.flags).all()
    assert (r_dev.bob_solution == r_host.bob_solution).all() and (r_dev.tally == r_host.tally).all()


def test_run_trials_rejects_qber_too_small_for_the_key(q):
    """run_trial throws "Key size ... is too small for QBER." when floor(N * QBER) == 0 (simulation.cpp:556-557)."""
    from qkd_ldpc_v_b200 import hostlib
    from qkd_ldpc_v_b200._cabi import QkdLdpcError
    with pytest.raises(QkdLdpcError, match="too small for QBER"):
        handle(q, "N100").run_trials(hostlib.trial_seeds(1, 4), 0.001, (0.8, 0), q.DecoderConfig())
    r = handle(q, "N100").run_trials(np.zeros(0, np.uint64), 0.05, (0.8, 0), q.DecoderConfig())
    assert r.iterations_num.size == 0 and r.tally.sum() == 0


@pytest.mark.parametrize("name,n_remove", [("K1_5", 250), ("N100", 0), ("N100", 99), ("I80", 2077), ("N7", 3)])
def test_remove_bits_equals_reference_semantics(q, name, n_remove):
    """remove_bits (array_and_matrix_operations.cpp:259-287): delete the listed positions, keep the order of the rest."""
    arr = util.code_arrays(name)
    rng = np.random.default_rng(n_remove + 5)
    bits = rng.integers(0, 2, (37, arr["n"]), dtype=np.uint8)
    rm = np.sort(rng.permutation(arr["n"])[:n_remove]).astype(np.int32)
    out = handle(q, name).remove_bits(bits, rm)
    expect = np.delete(bits, rm, axis=1)
    assert (q.unpack_bits(out, arr["n"] - n_remove) == expect).all()
    from qkd_ldpc_v_b200._cabi import QkdLdpcError
    if n_remove >= 2:
        with pytest.raises(QkdLdpcError):
            handle(q, name).remove_bits(bits, rm[::-1].copy())


@pytest.mark.parametrize("alg,name", [(5, "I80"), (2, "K1_5"), (0, "K1_5"), (1, "A82"), (0, "I80")])
def test_run_trials_multi_equals_one_call_per_combination(q, alg, name):
    """qkdldpc_run_trials_multi: a whole sweep (different QBER, scaling factors, punctured / shortened lists, seed offsets)
    in ONE generate + ONE decode launch (on-chip min-sum and sum-product kernels); SPA on the irregular n = 10240 code
    does not fit on chip and exercises the fallback (one combination after the other). The streaming path must give the
    same per-frame results."""
    from qkd_ldpc_v_b200 import hostlib
    arr = util.code_arrays(name)
    rng = np.random.default_rng(alg)
    seeds = hostlib.trial_seeds(4321, 70)
    combos = []
    for k in range(9):
        pos = rng.permutation(arr["n"])
        n_p, n_s = (0, 0) if k % 3 == 0 else (int(arr["n"] * 0.02 * (k % 3)), int(arr["n"] * 0.01 * k))
        combos.append(dict(QBER=0.012 + 0.002 * k, primary=0.6 + 0.02 * k, secondary=0.9 - 0.03 * k, seed_offset=1000 + k,
                           punctured_bits=np.sort(pos[:n_p]).astype(np.int32), shortened_bits=np.sort(pos[n_p:n_p + n_s]).astype(np.int32)))
    cfg = q.DecoderConfig(decoding_algorithm=alg, message_precision=32, max_iterations=60)
    h = handle(q, name)
    it, fl, tl, acc = h.run_trials_multi(seeds, combos, cfg)
    for k, cb in enumerate(combos):
        r = h.run_trials(seeds, cb["QBER"], (cb["primary"], cb["secondary"]), cfg, seed_offset=cb["seed_offset"],
                         punctured_bits=cb["punctured_bits"], shortened_bits=cb["shortened_bits"], want_bits=False)
        assert (it[k] == r.iterations_num).all() and (fl[k] == r.flags).all(), k
        assert (tl[k] == r.tally).all() and acc[k] == r.info["accurate_qber"], k
        if k % 4 == 1:
            rs = handle(q, name, decoder_path=1).run_trials(seeds, cb["QBER"], (cb["primary"], cb["secondary"]), cfg,
                                                            seed_offset=cb["seed_offset"], punctured_bits=cb["punctured_bits"],
                                                            shortened_bits=cb["shortened_bits"], want_bits=False)
            assert rs.info["last_path"] == 1
            assert (it[k] == rs.iterations_num).all() and (fl[k] == rs.flags).all(), k
    assert 0 < (fl & 1).sum() < fl.size


@pytest.mark.parametrize("alg,name", [(5, "I80"), (0, "I80"), (2, "K1_5")])
def test_run_trials_multi_final_keys(q, alg, name):
    """remove_bits as the last step of the batched run_trial (qkd_ldpc_algorithm.cpp:1089-1092, 1218-1220): combinations
    that carry H_matrix_params.bits_to_remove get Alice's (extended) key and bob_solution without those positions, built on
    the device inside the call. Checked against the per-combination call + numpy deletion; on-chip launch (AOMSA, NMSA)
    and the one-combination-after-the-other fallback (SPA on the irregular n = 10240 code)."""
    from qkd_ldpc_v_b200 import hostlib
    arr = util.code_arrays(name)
    n = arr["n"]
    rng = np.random.default_rng(10 + alg)
    seeds = hostlib.trial_seeds(777, 37)
    combos = []
    for k in range(5):
        pos = rng.permutation(n)
        n_p, n_s = (0, 0) if k == 0 else (int(n * 0.015 * k), int(n * 0.01 * k))
        p_, s_ = np.sort(pos[:n_p]).astype(np.int32), np.sort(pos[n_p:n_p + n_s]).astype(np.int32)
        # k = 0: privacy maintenance only (some positions); k = 3: no removal list at all; else punctured + shortened + extra
        rm = np.sort(np.concatenate([p_, s_, pos[n_p + n_s:n_p + n_s + 100 * (k % 2 == 0)]])).astype(np.int32) if k != 3 else np.zeros(0, np.int32)
        combos.append(dict(QBER=0.012 + 0.002 * k, primary=0.7, secondary=0.9, seed_offset=50 + k, punctured_bits=p_, shortened_bits=s_,
                           bits_to_remove=rm))
    cfg = q.DecoderConfig(decoding_algorithm=alg, max_iterations=40)
    h = handle(q, name)
    it, fl, tl, acc, ka, kb = h.run_trials_multi(seeds, combos, cfg, want_keys=True)
    it0, fl0, tl0, acc0 = h.run_trials_multi(seeds, [{k_: v for k_, v in cb.items() if k_ != "bits_to_remove"} for cb in combos], cfg)
    assert (it == it0).all() and (fl == fl0).all() and (tl == tl0).all() and (acc == acc0).all()
    for k, cb in enumerate(combos):
        if cb["bits_to_remove"].size == 0:
            assert ka[k] is None and kb[k] is None
            continue
        r = h.run_trials(seeds, cb["QBER"], (cb["primary"], cb["secondary"]), cfg, seed_offset=cb["seed_offset"],
                         punctured_bits=cb["punctured_bits"], shortened_bits=cb["shortened_bits"])
        assert (r.iterations_num == it[k]).all()
        a, _, _ = (hostlib.gen_keys_rate_adapt(seeds + np.uint64(cb["seed_offset"]), n, cb["QBER"], cb["punctured_bits"], cb["shortened_bits"])
                   if cb["punctured_bits"].size or cb["shortened_bits"].size else hostlib.gen_keys(seeds + np.uint64(cb["seed_offset"]), n, cb["QBER"]))
        keep = np.setdiff1d(np.arange(n), cb["bits_to_remove"])
        want_a = q.pack_bits(q.unpack_bits(a, n)[:, keep])
        want_b = q.pack_bits(r.bits()[:, keep])
        assert (ka[k] == want_a).all() and (kb[k] == want_b).all(), k
        assert (h.remove_bits(r.bob_solution, cb["bits_to_remove"]) == kb[k]).all()


@pytest.mark.parametrize("alg,fac", [(2, (0.75, 0.0)), (0, (0.0, 0.0))])
def test_bench_synthetic_entry_point(q, alg, fac):
    """qkdldpc_bench_synthetic: device-generated synthetic keys, decode, tally + seconds. The tally must
    be the one decode_batch_device gives on the same synthetic keys (same seed), and the time must be positive."""
    import torch
    arr = util.code_arrays("K1_5")
    h = handle(q, "K1_5")
    cfg = q.DecoderConfig(decoding_algorithm=alg, message_precision=32)
    frames, qber, seed = 700, 0.02, 12345
    tally, sec = h.bench_synthetic(frames, qber, fac, cfg, seed=seed)
    assert sec > 0 and tally[0] == frames and 0 < tally[1] <= frames
    words = (arr["n"] + 31) // 32
    d_a = torch.empty((frames, words), dtype=torch.int32, device="cuda:0")
    d_b = torch.empty_like(d_a)
    acc = h.generate_keys_device(frames, qber, seed, d_a.data_ptr(), d_b.data_ptr())
    r = h.QKD_LDPC_batch(d_a.cpu().numpy().view(np.uint32), d_b.cpu().numpy().view(np.uint32), acc, fac, cfg)
    assert (r.tally == tally).all()
    t0, s0 = h.bench_synthetic(0, qber, fac, cfg)
    assert s0 == 0 and t0.sum() == 0
