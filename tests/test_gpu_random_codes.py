"""GPU: randomly generated Tanner graphs through every float32 path. The golden codes are regular or mildly irregular;
the table builders of the on-chip kernels (degree classes, 32-lane groups, padding lanes, 4-edge index blocks, the item list
and the always-zero padding word of the sum-product kernel, rows of 33..64 edges, ELL records of the streaming variable
node) have many more corner cases than those codes exercise: degree-1 bits, degree-1 and degree-2 checks, a single wide
row, n not a multiple of 32, more degree classes than warps. For each random graph:
  * streaming (decoder_path 1) and on-chip (decoder_path 2) results must be identical, frame by frame;
  * the min-sum family and SPA-lin-approx must be bit-identical to the float32 oracle;
  * float64 messages must reproduce the float64 oracle (min-sum family, SPA-lin)."""
import numpy as np
import pytest

from oracle import cpu

pytestmark = pytest.mark.gpu


def random_graph(seed, n, m, col_degrees, wide_row=0):
    """CSR (row_ptr, col_idx) of a random m x n parity-check matrix: every bit gets a degree drawn from col_degrees and
    that many distinct checks; every check is guaranteed at least one bit; optionally one row of `wide_row` edges."""
    rng = np.random.default_rng(seed)
    rows = [set() for _ in range(m)]
    for bit in range(n):
        d = int(rng.choice(col_degrees))
        for r in rng.choice(m, size=min(d, m), replace=False):
            rows[int(r)].add(bit)
    for r in range(m):
        if not rows[r]:
            rows[r].add(int(rng.integers(n)))
    if wide_row:
        rows[0] = set(int(x) for x in rng.choice(n, size=min(wide_row, n), replace=False))
        covered = set().union(*rows)           # a bit that lost its only check gets one back
        for bit in set(range(n)) - covered:
            rows[1 + bit % (m - 1)].add(bit)
    row_ptr = np.zeros(m + 1, np.int32)
    col_idx = []
    for r in range(m):
        cols = sorted(rows[r])
        col_idx += cols
        row_ptr[r + 1] = len(col_idx)
    return row_ptr, np.asarray(col_idx, np.int32)


CASES = [
    # seed, n, m, column degrees, widest row, qber
    (1, 37, 20, (1, 2, 3), 0, 0.03),
    (2, 64, 32, (2, 3), 0, 0.03),
    (3, 203, 111, (1, 2, 3, 5, 9), 40, 0.02),
    (4, 1000, 300, (2, 3, 4, 7, 12), 64, 0.01),
    (5, 1500, 1000, (3,), 0, 0.06),
    (6, 999, 333, (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 17), 33, 0.01),
    (7, 900, 200, (2, 3, 6, 11), 51, 0.01),     # widest row the 8-byte record format takes (24 + 27 edges)
    (8, 700, 130, (3, 4, 8), 28, 0.01),         # rows on both sides of the 27-edge record limit
]
FACT = {0: (0.0, 0.0), 1: (0.0, 0.0), 2: (0.8, 0.0), 3: (0.4, 0.0), 4: (0.8, 0.6), 5: (0.3, 0.7)}


@pytest.mark.parametrize("seed,n,m,degs,wide,qber", CASES)
def test_random_graph_all_paths(built, seed, n, m, degs, wide, qber):
    import qkd_ldpc_v_b200 as q
    row_ptr, col_idx = random_graph(seed, n, m, degs, wide)
    assert int(np.diff(row_ptr).max()) <= 64
    oc = cpu.Code(n, m, row_ptr, col_idx)
    rng = np.random.default_rng(100 + seed)
    frames = 200
    alice = rng.integers(0, 2, (frames, n), dtype=np.uint8)
    bob = alice.copy()
    n_err = max(1, int(n * qber))
    for f in range(frames):
        bob[f, rng.choice(n, n_err, replace=False)] ^= 1
    acc = n_err / n
    with q.LdpcCode(n, m, row_ptr, col_idx, device=0, decoder_path=1) as stream, \
            q.LdpcCode(n, m, row_ptr, col_idx, device=0, decoder_path=2) as chip, \
            q.LdpcCode(n, m, row_ptr, col_idx, device=0, decoder_path=2, onchip_threads=96) as chip96, \
            q.LdpcCode(n, m, row_ptr, col_idx, device=0, decoder_path=2, onchip_record_bytes=16) as chip16, \
            q.LdpcCode(n, m, row_ptr, col_idx, device=0, decoder_path=2, onchip_record_bytes=8) as chip8:
        for alg in range(6):
            cfg = q.DecoderConfig(decoding_algorithm=alg, message_precision=32, max_iterations=40)
            rs = stream.QKD_LDPC_batch(alice, bob, acc, FACT[alg], cfg)
            rc = chip.QKD_LDPC_batch(alice, bob, acc, FACT[alg], cfg)
            r9 = chip96.QKD_LDPC_batch(alice, bob, acc, FACT[alg], cfg)
            r16 = chip16.QKD_LDPC_batch(alice, bob, acc, FACT[alg], cfg)
            r8 = chip8.QKD_LDPC_batch(alice, bob, acc, FACT[alg], cfg)
            assert rs.info["last_path"] == 1 and rc.info["last_path"] == 2
            if alg >= 2:   # min-sum family: 8-byte records by default when every row fits one (27 edges), on request up to 51 edges
                dc_max = int(np.diff(row_ptr).max())
                assert rc.info["onchip_record_bytes"] == (8 if dc_max <= 27 else 16)
                assert r8.info["onchip_record_bytes"] == (8 if dc_max <= 51 else 16)
                assert r16.info["onchip_record_bytes"] == 16
            for r in (rc, r9, r16, r8):
                assert (rs.iterations_num == r.iterations_num).all(), alg
                assert (rs.flags == r.flags).all() and (rs.bob_solution == r.bob_solution).all(), alg
                assert (rs.tally == r.tally).all(), alg
            it32, fl32, bits32 = cpu.qkd_ldpc_batch(oc, alg, alice, bob, acc, max_iter=40, primary=FACT[alg][0],
                                                    secondary=FACT[alg][1], precision=32)
            if alg >= 1:
                assert (rs.iterations_num == it32).all() and (rs.flags == fl32).all() and (rs.bits() == bits32).all(), alg
            else:
                assert (rs.iterations_num == it32).mean() >= 0.9, alg
            if alg >= 1:
                cfg64 = q.DecoderConfig(decoding_algorithm=alg, message_precision=64, max_iterations=40)
                r64 = stream.QKD_LDPC_batch(alice, bob, acc, FACT[alg], cfg64)
                it64, fl64, bits64 = cpu.qkd_ldpc_batch(oc, alg, alice, bob, acc, max_iter=40, primary=FACT[alg][0],
                                                        secondary=FACT[alg][1], precision=64)
                assert (r64.iterations_num == it64).all() and (r64.flags == fl64).all() and (r64.bits() == bits64).all(), alg


def test_random_graph_rate_adaptation_on_chip(built):
    """Punctured / shortened positions through both on-chip kernels and the streaming path on an irregular random graph."""
    import qkd_ldpc_v_b200 as q
    n, m = 600, 240
    row_ptr, col_idx = random_graph(11, n, m, (2, 3, 6), 0)
    rng = np.random.default_rng(7)
    pos = rng.permutation(n)
    punct, short = np.sort(pos[:30]).astype(np.int32), np.sort(pos[30:50]).astype(np.int32)
    frames = 150
    alice = rng.integers(0, 2, (frames, n), dtype=np.uint8)
    alice[:, short] = 0
    bob = alice.copy()
    for f in range(frames):
        bob[f, rng.choice(n, 6, replace=False)] ^= 1
    bob[:, short] = 0
    bob[:, punct] = rng.integers(0, 2, (frames, punct.size), dtype=np.uint8)
    with q.LdpcCode(n, m, row_ptr, col_idx, device=0, decoder_path=1) as stream, \
            q.LdpcCode(n, m, row_ptr, col_idx, device=0, decoder_path=2) as chip:
        for alg in (0, 1, 2, 5):
            cfg = q.DecoderConfig(decoding_algorithm=alg, message_precision=32, max_iterations=40)
            rs = stream.QKD_LDPC_batch(alice, bob, 0.01, FACT[alg], cfg, punctured_bits=punct, shortened_bits=short)
            rc = chip.QKD_LDPC_batch(alice, bob, 0.01, FACT[alg], cfg, punctured_bits=punct, shortened_bits=short)
            assert rc.info["last_path"] == 2
            assert (rs.iterations_num == rc.iterations_num).all() and (rs.flags == rc.flags).all(), alg
            assert (rs.bob_solution == rc.bob_solution).all(), alg
            assert rs.syndromes_match.mean() > 0.3, alg


def test_random_graph_wide_nodes_streaming(built):
    """Bits of 40 checks and rows of more than 64 edges: no on-chip kernel takes this graph (explicit request fails, the
    automatic choice streams); the streaming kernels' re-read variants (DCMAX = 0, DVMAX = 0) must match the oracles."""
    import qkd_ldpc_v_b200 as q
    from qkd_ldpc_v_b200._cabi import QkdLdpcError
    n, m = 300, 60
    row_ptr, col_idx = random_graph(21, n, m, (2, 3, 40), 0)
    assert int(np.diff(row_ptr).max()) > 64 and int(np.bincount(col_idx).max()) > 32
    oc = cpu.Code(n, m, row_ptr, col_idx)
    rng = np.random.default_rng(3)
    frames = 120
    alice = rng.integers(0, 2, (frames, n), dtype=np.uint8)
    bob = alice.copy()
    for f in range(frames):
        bob[f, rng.choice(n, 9, replace=False)] ^= 1
    with q.LdpcCode(n, m, row_ptr, col_idx, device=0) as auto, q.LdpcCode(n, m, row_ptr, col_idx, device=0, decoder_path=2) as chip:
        for alg in (0, 1, 2, 5):
            cfg = q.DecoderConfig(decoding_algorithm=alg, message_precision=32, max_iterations=30)
            with pytest.raises(QkdLdpcError):
                chip.QKD_LDPC_batch(alice, bob, 0.03, FACT[alg], cfg)
            r = auto.QKD_LDPC_batch(alice, bob, 0.03, FACT[alg], cfg)
            assert r.info["last_path"] == 1
            for prec, res in ((32, r), (64, auto.QKD_LDPC_batch(alice, bob, 0.03, FACT[alg],
                                                                q.DecoderConfig(decoding_algorithm=alg, message_precision=64,
                                                                                max_iterations=30)))):
                it, fl, bits = cpu.qkd_ldpc_batch(oc, alg, alice, bob, 0.03, max_iter=30, primary=FACT[alg][0],
                                                  secondary=FACT[alg][1], precision=prec)
                if alg >= 1:
                    assert (res.iterations_num == it).all() and (res.flags == fl).all() and (res.bits() == bits).all(), (alg, prec)
                else:
                    assert (res.iterations_num == it).mean() >= 0.9, (alg, prec)
