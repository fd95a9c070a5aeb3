"""GPU: the two operating points where float32 sum-product decoding left the reference's FER interval in round 1
(profiles/r01_g_config_parity.md), against trial-level goldens of the UNMODIFIED reference (tests/golden/trials_*.npz,
made by tests/golden/make_golden_points.py with the configs' own simulation seed):

  * `config 100k.json` as shipped -- SPA on the n = 102400 code at QBER 8.4 %, 100 trials, 35 iterations per frame, 8 % from
    capacity. float32 fails 6 of the 100 frames (24-bit trajectories end in trapping sets; the reference's code in `float`
    fails the same 6), the reference none. The library's precision policy (message_precision = 0) decodes SPA on n > 65536
    in float64: FER inside the reference's Wilson interval, iteration counts equal on >= 95 of the 100 frames.
  * the R = 0.66 matrix of `config 1k.json` -- SPA at QBER 3 %, 20 000 trials, FER ~ 2e-4. The handful of failing frames are
    error-floor events around the reference's NaN quirk (Q3: 0/0 in the tanh product); float32 (the policy's choice for
    short codes) has its own handful. Bars: float64 two-sided inside the interval with >= 99.5 % equal iteration counts;
    float32 not worse than the interval's upper end (one-sided: fewer failures than the reference is not a defect) with
    >= 99 % equal iteration counts."""
import math
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def wilson(k, n, z=1.96):
    p = k / n
    d = 1 + z * z / n
    c = (p + z * z / (2 * n)) / d
    h = z * math.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / d
    return c - h, c + h


def run_point(case, precision, path=0):
    import qkd_ldpc_v_b200 as q
    from qkd_ldpc_v_b200 import hostlib
    g = np.load(os.path.join(util.GOLDEN, f"trials_{case}.npz"))
    arr = util.code_arrays(str(g["code"]))
    seeds = g["seeds"]
    a, b, acc = hostlib.gen_keys(seeds, arr["n"], float(g["qber"]))
    assert abs(acc - float(g["acc_qber"][0])) < 1e-15
    with q.LdpcCode(arr["n"], arr["m"], arr["row_ptr"], arr["col_idx"], device=0, decoder_path=path) as code:
        cfg = q.DecoderConfig(decoding_algorithm=int(g["alg"]), max_iterations=int(g["max_iter"]), message_precision=precision)
        r = code.QKD_LDPC_batch(a, b, acc, (0.0, 0.0), cfg)
    it_ref, fl_ref = g["iters"].astype(np.int32), g["flags"]
    k = seeds.size
    fail_ref, fail_gpu = int(((fl_ref & 3) != 3).sum()), int(((r.flags & 3) != 3).sum())
    agree = float((r.iterations_num == it_ref).mean())
    lo, hi = wilson(fail_ref, k)
    print(f"\n{case} precision {precision} -> {r.info['last_precision']} path {r.info['last_path']}: iterations equal {agree:.4f}, "
          f"failures gpu {fail_gpu} ref {fail_ref} of {k} (Wilson95 [{lo * k:.1f}, {hi * k:.1f}]), mean it {it_ref.mean():.2f}")
    return r, agree, fail_gpu, (lo, hi), k


def test_config_100k_spa_default_policy(built):
    r, agree, fail_gpu, (lo, hi), k = run_point("L100k_spa_q084", 0)
    assert r.info["last_precision"] == 64, "SPA on n > 65536 must run in float64 under the default policy"
    assert lo - 1e-12 <= fail_gpu / k <= hi + 1e-12
    assert agree >= 0.95, agree


def test_config_1k_r066_spa_float64(built):
    r, agree, fail_gpu, (lo, hi), k = run_point("K1_4_spa_q030", 64)
    assert r.info["last_precision"] == 64
    assert lo - 1e-12 <= fail_gpu / k <= hi + 1e-12
    assert agree >= 0.995, agree


def test_config_1k_r066_spa_default_policy(built):
    r, agree, fail_gpu, (lo, hi), k = run_point("K1_4_spa_q030", 0)
    assert r.info["last_precision"] == 32 and r.info["last_path"] == 2   # on-chip sum-product kernel
    assert fail_gpu / k <= hi + 1e-12
    assert agree >= 0.99, agree
