#!/usr/bin/env python
"""Trial-level goldens at two operating points where float32 sum-product decoding left the reference's FER interval in
round 1 (profiles/r01_g_config_parity.md): `config 100k.json` as shipped (SPA on the n = 102400 code at QBER 8.4 %, 100
trials) and the R = 0.66 matrix of `config 1k.json` (SPA at QBER 3 %, 20 000 trials). Generated in THIS container from the
UNMODIFIED reference (oracle/_ref/libqkdref.so: run_trial per seed, simulation.cpp:540) with the configs' own simulation
seed; only the per-trial iteration counts and flags are kept (the inputs are regenerated from the seeds at test time).

Usage: python tests/golden/make_golden_points.py        (about 3 minutes on 8 cores)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SM = os.path.join(ref.REFERENCE_ROOT, "sparse_matrices")

# (case, code name in codes.npz, alg, qber, sim_seed, trials, max_iter)
POINTS = [
    ("L100k_spa_q084", "L100k", 0, 0.084, 9012025, 100, 100),
    ("K1_4_spa_q030", "K1_4", 0, 0.030, 9012025, 20000, 100),
]


def main():
    codes = np.load(os.path.join(OUT, "codes.npz"))
    for case, cname, alg, q, sim_seed, k, max_iter in POINTS:
        fname = str(codes[f"{cname}.file"])
        sub = "matrices_alist_100k_all" if cname == "L100k" else "matrices_alist_1k_all"
        m = ref.RefMatrix(os.path.join(SM, sub, fname), 1)
        seeds = ref.trial_seeds(sim_seed, k)
        ref.set_cfg(alg, max_iter, True, 100.0)
        t0 = time.time()
        iters, flags, acc = m.run_trials(q, seeds, 0.0, 0.0)
        np.savez_compressed(os.path.join(OUT, f"trials_{case}.npz"), code=np.array(cname), alg=np.array(alg), qber=np.array(q),
                            sim_seed=np.array(sim_seed, np.uint64), max_iter=np.array(max_iter), seeds=seeds,
                            acc_qber=acc[:1], iters=iters.astype(np.int16), flags=flags)
        print(case, "trials", k, "mean iterations %.2f" % iters.mean(), "failures", int(((flags & 3) != 3).sum()),
              "%.0f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
