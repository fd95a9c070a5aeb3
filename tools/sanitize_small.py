"""Small end-to-end run of every kernel family (for compute-sanitizer): on-chip and streaming min-sum, SPA, float64,
rate adaptation, synthetic key generator. Usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402

import qkd_ldpc_v_b200 as q  # noqa: E402
from qkd_ldpc_v_b200 import hostlib  # noqa: E402


def main():
    for name, qber in (("K1_5", 0.02), ("N100", 0.03), ("I80", 0.017)):
        arr = util.code_arrays(name)
        frames = 40 if arr["n"] > 5000 else 150
        a, b, acc = hostlib.gen_keys(hostlib.trial_seeds(11, frames), arr["n"], qber)
        for path in (2, 1):
            with q.LdpcCode(arr["n"], arr["m"], arr["row_ptr"], arr["col_idx"], device=0, decoder_path=path, pool_slots=64) as code:
                for alg, fac in ((2, (0.75, 0.0)), (5, (0.3, 0.9))):
                    r = code.QKD_LDPC_batch(a, b, acc, fac, q.DecoderConfig(decoding_algorithm=alg, max_iterations=30))
                    print(name, "path", path, "alg", alg, "ok", int(r.syndromes_match.sum()), "/", frames)
                if arr["untp"] is not None:
                    p, s = arr["untp"][:200].copy(), np.setdiff1d(np.arange(arr["n"], dtype=np.int32), arr["untp"][:200])[:50]
                    p.sort()
                    r = code.QKD_LDPC_batch(a, b, acc, (0.7, 0), q.DecoderConfig(decoding_algorithm=2, max_iterations=20),
                                            punctured_bits=p, shortened_bits=s)
                    print(name, "path", path, "rate-adapted ok", int(r.syndromes_match.sum()))
        with q.LdpcCode(arr["n"], arr["m"], arr["row_ptr"], arr["col_idx"], device=0, pool_slots=64) as code:
            for alg, prec in ((0, 32), (1, 32), (0, 64), (3, 64)):
                r = code.QKD_LDPC_batch(a[:32], b[:32], acc, (0.3, 0.0), q.DecoderConfig(decoding_algorithm=alg, max_iterations=15,
                                                                                       message_precision=prec))
                print(name, "alg", alg, "fp", prec, "ok", int(r.syndromes_match.sum()))
    print("sanitize run done")


if __name__ == "__main__":
    main()
