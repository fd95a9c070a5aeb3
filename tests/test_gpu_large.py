"""GPU: the north-star parity bar on LARGE batches at the reference's own operating points:
   (i) decoded words bit-exact on every frame where both implementations converge,
  (ii) per-frame iteration counts equal on >= 99 % of frames,
 (iii) FER of the GPU decoder inside the Wilson 95 % interval of the reference FER.
The decoder runs under the library's DEFAULT precision policy (message_precision = 0, include/qkdldpc.h): float32 state for
NMSA / SPA / SPA-lin-approx, float64 state (the on-chip float64 kernel) for OMSA / ANMSA / AOMSA. The bar is 0.99 on every row.
The reference side is the C oracle in float64 (bit-identical to the compiled reference: tests/test_oracle_golden.py,
tests/test_oracle_vs_ref.py); inputs come from the reference-compatible RNG so both sides see identical frames."""
import math
import os

import numpy as np
import pytest

import util
from oracle import cpu

pytestmark = pytest.mark.gpu

# QKD_PARITY_FULL=1 runs every point on the reference configs' own trial counts (30 000 for the NOPT configs)
# instead of the few thousand frames that keep the default suite short; tools/gpurun/parity_full.sh saves the printed lines.
FULL = os.environ.get("QKD_PARITY_FULL", "") == "1"
FULL_FRAMES = {("A79", 2): 30000, ("A82", 2): 30000, ("I80", 2): 10000, ("A82", 0): 30000, ("A82", 1): 30000,
               ("A82", 3): 10000, ("A82", 4): 10000, ("A82", 5): 10000}

# (code, alg, qber, primary, secondary, frames, sim_seed, min share of equal iteration counts)
POINTS = [
    ("A79", 2, 0.020, 0.71, 0.0, 10000, 10012025, 0.99),     # config 10k NMSA.json operating point
    ("A82", 2, 0.0159, 0.69, 0.0, 6000, 777, 0.99),          # NOPT_R=0,82_NMSA.json (FER ~ 0.01)
    ("I80", 2, 0.015, 0.70, 0.0, 6000, 10012025, 0.99),      # irregular R=0.8, converging
    ("A82", 0, 0.0162, 0.0, 0.0, 3000, 777, 0.99),           # NOPT_R=0,82_SPA.json
    ("A82", 1, 0.0162, 0.0, 0.0, 3000, 777, 0.99),           # NOPT_R=0,82_SPA_LIN_APPROX.json
    ("A82", 3, 0.0154, 0.81, 0.0, 3000, 777, 0.99),          # NOPT_R=0,82_OMSA.json   (float64 state by default)
    ("A82", 4, 0.0161, 0.80, 0.71, 3000, 777, 0.99),         # NOPT_R=0,82_ANMSA.json  (float64 state by default)
    ("A82", 5, 0.0161, 0.68, 1.25, 3000, 777, 0.99),         # NOPT_R=0,82_AOMSA .json (float64 state by default)
]
# float32 state forced on the offset / adaptive variants: chaotic near threshold (hard part 1) -- reported, and
# held to the level measured in round 1 so that a regression of the float32 kernels is still caught
F32_FLOOR = {3: 0.96, 4: 0.98, 5: 0.94}


def wilson(k, n, z=1.96):
    p = k / n
    d = 1 + z * z / n
    c = (p + z * z / (2 * n)) / d
    h = z * math.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / d
    return c - h, c + h


@pytest.mark.parametrize("name,alg,qber,pri,sec,frames,seed,bar", POINTS)
def test_operating_point(built, name, alg, qber, pri, sec, frames, seed, bar):
    import qkd_ldpc_v_b200 as q
    from qkd_ldpc_v_b200 import hostlib
    arr = util.code_arrays(name)
    oc = util.oracle_code(name)
    if FULL:
        frames = FULL_FRAMES[(name, alg)]
    seeds = hostlib.trial_seeds(seed, frames)
    a, b, acc = hostlib.gen_keys(seeds, arr["n"], qber)
    ab, bb = q.unpack_bits(a, arr["n"]), q.unpack_bits(b, arr["n"])
    it_ref, fl_ref, bits_ref = cpu.qkd_ldpc_batch(oc, alg, ab, bb, acc, primary=pri, secondary=sec, precision=64)
    prec = 64 if alg >= 3 else 32
    with q.LdpcCode(arr["n"], arr["m"], arr["row_ptr"], arr["col_idx"], device=0, pool_slots=2048) as code:
        cfg = q.DecoderConfig(decoding_algorithm=alg)         # message_precision = 0: the library's policy
        r = code.QKD_LDPC_batch(a, b, acc, (pri, sec), cfg)   # auto path: on-chip (all six algorithms fit these codes)
        assert r.info["last_path"] == 2 and r.info["last_precision"] == prec
        # the streaming kernels of the same precision are the same arithmetic: identical per-frame results
        with q.LdpcCode(arr["n"], arr["m"], arr["row_ptr"], arr["col_idx"], device=0, pool_slots=2048, decoder_path=1) as sc:
            rs = sc.QKD_LDPC_batch(a, b, acc, (pri, sec), q.DecoderConfig(decoding_algorithm=alg, message_precision=prec))
        assert rs.info["last_path"] == 1
        assert (rs.iterations_num == r.iterations_num).all() and (rs.flags == r.flags).all()
        assert (rs.bob_solution == r.bob_solution).all() and (rs.tally == r.tally).all()
        cfg64 = q.DecoderConfig(decoding_algorithm=alg, message_precision=64)
        r64 = code.QKD_LDPC_batch(a, b, acc, (pri, sec), cfg64)
        r32 = code.QKD_LDPC_batch(a, b, acc, (pri, sec), q.DecoderConfig(decoding_algorithm=alg, message_precision=32)) if alg >= 3 else None
    # float64 messages: identical to the reference (SPA: libm vs CUDA tanh/atanh, last-ulp differences only)
    if alg != 0:
        assert (r64.iterations_num == it_ref).all() and (r64.flags == fl_ref).all()
        assert (r64.bits() == bits_ref).all()
    else:
        assert (r64.iterations_num == it_ref).mean() >= 0.995
    if r32 is not None:
        agree32 = float((r32.iterations_num == it_ref).mean())
        print(f"\n{name} alg={alg}: float32 state forced: iterations equal {agree32:.4f}, FER {((r32.flags & 3) != 3).mean():.5f}")
        assert agree32 >= F32_FLOOR[alg]
        b32 = r32.syndromes_match & ((fl_ref & 1) != 0)
        assert (r32.bits()[b32] == bits_ref[b32]).all()
    # default policy: the north-star triple
    both = r.syndromes_match & ((fl_ref & 1) != 0)
    assert (r.bits()[both] == bits_ref[both]).all(), "words differ on co-converged frames"
    agree = float((r.iterations_num == it_ref).mean())
    fail_ref = int(((fl_ref & 3) != 3).sum())
    fail_gpu = int(((r.flags & 3) != 3).sum())
    lo, hi = wilson(fail_ref, frames)
    diff_words = int((r.bits()[both] != bits_ref[both]).any(axis=1).sum())
    print(f"\n{name} alg={alg} q={qber} frames={frames}: co-converged frames with different words {diff_words}, "
          f"iterations equal {agree:.4f}, FER gpu {fail_gpu / frames:.5f} "
          f"ref {fail_ref / frames:.5f} (Wilson95 [{lo:.5f}, {hi:.5f}]), mean it {it_ref.mean():.2f}")
    assert agree >= bar, agree
    assert lo - 1e-12 <= fail_gpu / frames <= hi + 1e-12
