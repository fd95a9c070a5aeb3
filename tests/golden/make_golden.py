#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libqkdref.so), in THIS container only
(needs /root/reference). The fixtures travel to the GPU box; nothing there reads /root/reference.

  codes.npz      parity-check matrices of the reference tree re-encoded as CSR (row_ptr/col_idx int32), exactly as
                 the reference's own loaders return them (check_nodes order preserved) + the .untp puncturing lists.
  rng.npz        known answers for the reference's input generation (xoshiro256++ + libstdc++ distributions):
                 per-trial seeds, Alice/Bob keys (packed), accurate QBER.
  decode_*.npz   per operating point: seeds, the reference's iterations / flags (run_trial, simulation.cpp:540) and
                 the decoded words (bit_array_out of the six *_decoding functions), packed.

Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu, ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SM = os.path.join(ref.REFERENCE_ROOT, "sparse_matrices")

CODES = {
    # name: (relative path, matrix_format)
    "N6": ("matrices_uncompressed/(N=6,K=2,M=4,R=0.34).mtrx", 0),
    "N7": ("matrices_uncompressed/(N=7,K=4,M=3,R=0.57).mtrx", 0),
    "N10s1": ("matrices_1/(N=10,M=5,R=0.5).mtrx", 2),
    "N100": ("matrices_uncompressed/(N=100,M=50,R=0.5).mtrx", 0),
    "A79": ("matrices_alist_10k_all/(N=10240,M=2201,R=0.79,CW=4,SEED=777).mtrx", 1),
    "A82": ("matrices_alist_10k_all/(N=10240,M=1801,R=0.82,CW=4,SEED=777).mtrx", 1),
    "I80": ("matrices_2/(N=10240,M=2048,R=0.8).mtrx", 3),
    "I65": ("matrices_2/(N=10240,M=3584,R=0.65).mtrx", 3),
    "I50": ("matrices_2/(N=10240,M=5120,R=0.5).mtrx", 3),
}


def find_1k_100k():
    """Three 1k codes (one per column weight) and one 100k code, picked deterministically by name."""
    d1 = sorted(os.listdir(os.path.join(SM, "matrices_alist_1k_all")))
    pick = {}
    for cw in ("CW=3", "CW=4", "CW=5"):
        c = [f for f in d1 if cw in f and f.endswith(".mtrx")]
        pick[f"K1_{cw[-1]}"] = (os.path.join("matrices_alist_1k_all", c[len(c) // 2]), 1)
    # the widest rows in the 1k family (dc up to 62) -- exercises the large-degree check-node path
    hi = [f for f in d1 if "R=0.9" in f and f.endswith(".mtrx")]
    if hi:
        pick["K1_hi"] = (os.path.join("matrices_alist_1k_all", hi[-1]), 1)
    d100 = sorted(f for f in os.listdir(os.path.join(SM, "matrices_alist_100k_all")) if f.endswith(".mtrx"))
    pick["L100k"] = (os.path.join("matrices_alist_100k_all", d100[len(d100) // 2]), 1)
    return pick


def pack(bits):
    return np.packbits(np.asarray(bits, np.uint8), axis=-1, bitorder="little")


def main():
    CODES.update(find_1k_100k())
    mats, store = {}, {}
    for name, (rel, fmt) in CODES.items():
        m = ref.RefMatrix(os.path.join(SM, rel), fmt)
        mats[name] = m
        store[f"{name}.row_ptr"] = m.row_ptr
        store[f"{name}.col_idx"] = m.col_idx
        store[f"{name}.col_ptr"] = m.col_ptr
        store[f"{name}.row_idx"] = m.row_idx
        store[f"{name}.meta"] = np.array([m.n, m.m, m.nnz, int(m.is_regular), fmt], np.int64)
        store[f"{name}.file"] = np.array(os.path.basename(rel))
        untp = os.path.join(SM, rel[:-5] + ".untp")
        if os.path.exists(untp):
            store[f"{name}.untp"] = np.array(open(untp).read().split(), np.int32)
        print(name, m.n, m.m, m.nnz, "regular" if m.is_regular else "irregular", rel)
    np.savez_compressed(os.path.join(OUT, "codes.npz"), **store)

    # ---- RNG known answers -------------------------------------------------------------------------------
    rng = {}
    for sim_seed, n, q, k in ((10012025, 10240, 0.02, 4), (777, 10240, 0.0162, 3), (9012025, 1024, 0.05, 4),
                              (5555, 6, 0.2, 4)):
        seeds = ref.trial_seeds(sim_seed, k)
        a, b, acc = [], [], []
        for s in seeds:
            x, y, z = ref.gen_keys(s, n, q)
            a.append(x), b.append(y), acc.append(z)
        key = f"s{sim_seed}_n{n}"
        rng[f"{key}.seeds"] = seeds
        rng[f"{key}.qber"] = np.array([q])
        rng[f"{key}.alice"] = pack(a)
        rng[f"{key}.bob"] = pack(b)
        rng[f"{key}.acc"] = np.array(acc)
    np.savez_compressed(os.path.join(OUT, "rng.npz"), **rng)

    # ---- decode goldens -----------------------------------------------------------------------------------
    # (case, code, alg, qber, primary, secondary, sim_seed, frames, max_iter). Operating points follow the
    # reference's own configs (SURVEY.md 6 / BASELINE.md): config 10k NMSA.json, NOPT_R=0,82_*.json, config 1k.json.
    cases = [
        ("A79_nmsa_q020", "A79", 2, 0.020, 0.71, 0.0, 10012025, 48, 100),
        ("A79_nmsa_q030", "A79", 2, 0.030, 0.71, 0.0, 10012025, 8, 100),
        ("I80_nmsa_q015", "I80", 2, 0.015, 0.70, 0.0, 10012025, 48, 100),
        ("I80_nmsa_q030", "I80", 2, 0.030, 0.70, 0.0, 10012025, 8, 100),
        ("A82_spa_q0162", "A82", 0, 0.0162, 0.0, 0.0, 777, 32, 100),
        ("A82_spalin_q0162", "A82", 1, 0.0162, 0.0, 0.0, 777, 32, 100),
        ("A82_nmsa_q0159", "A82", 2, 0.0159, 0.69, 0.0, 777, 32, 100),
        ("A82_omsa_q0154", "A82", 3, 0.0154, 0.81, 0.0, 777, 32, 100),
        ("A82_anmsa_q0161", "A82", 4, 0.0161, 0.80, 0.71, 777, 32, 100),
        ("A82_aomsa_q0161", "A82", 5, 0.0161, 0.68, 1.25, 777, 32, 100),
        ("K1_3_spa", "K1_3", 0, 0.05, 0.0, 0.0, 9012025, 64, 100),
        ("K1_4_spa", "K1_4", 0, 0.03, 0.0, 0.0, 9012025, 64, 100),
        ("K1_5_nmsa", "K1_5", 2, 0.02, 0.75, 0.0, 9012025, 64, 100),
        ("K1_hi_nmsa", "K1_hi", 2, 0.005, 0.8, 0.0, 9012025, 64, 100),
        ("K1_hi_spa", "K1_hi", 0, 0.005, 0.0, 0.0, 9012025, 64, 100),
        ("L100k_nmsa", "L100k", 2, 0.06, 0.72, 0.0, 9012025, 4, 100),
        ("N100_all", "N100", 2, 0.05, 0.8, 0.0, 5555, 32, 20),
    ]
    for case, cname, alg, q, pri, sec, sim_seed, k, max_iter in cases:
        m = mats[cname]
        code = cpu.Code(m.n, m.m, m.row_ptr, m.col_idx, m.col_ptr, m.row_idx)
        seeds = ref.trial_seeds(sim_seed, k)
        ref.set_cfg(alg, max_iter, True, 100.0)
        it_rt, fl_rt, acc = m.run_trials(q, seeds, pri, sec)
        words, iters, flags = [], [], []
        for s in seeds:
            a, b, aq = ref.gen_keys(s, m.n, q)
            llr, syn = cpu.frame_setup(code, a, b, aq)        # same expression as qkd_ldpc_algorithm.cpp:1043-1052
            assert (syn == m.syndrome(a)).all()
            it, ok, z = m.decode(alg, llr, syn, max_iter, pri, sec, True, 100.0)
            words.append(z), iters.append(it)
            flags.append((1 if ok else 0) | (2 if (z == a).all() else 0))
        iters, flags = np.array(iters, np.int32), np.array(flags, np.uint8)
        assert (iters == it_rt).all() and (flags == fl_rt).all(), case   # decoder-level == run_trial-level
        np.savez_compressed(os.path.join(OUT, f"decode_{case}.npz"), code=np.array(cname), alg=np.array(alg),
                            qber=np.array(q), primary=np.array(pri), secondary=np.array(sec),
                            sim_seed=np.array(sim_seed, np.uint64), max_iter=np.array(max_iter), seeds=seeds,
                            acc_qber=acc, iters=iters, flags=flags, words=pack(words))
        print(case, "iters mean %.2f" % iters.mean(), "ok", int((flags & 1).sum()), "/", k)


if __name__ == "__main__":
    main()
