"""Decoded Gbit/s of one code / algorithm / QBER on the default decoder path for several on-chip CTA sizes
(qkdldpc_bench_synthetic: keys generated on the device, decode only, CUDA events).

    python tools/quick_bench.py K1_4 0 0 0 0.03 262144 "0 128 256 512 1024"
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402

import qkd_ldpc_v_b200 as q  # noqa: E402


def main():
    name, alg, pri, sec, qber, frames = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), float(sys.argv[4]), float(sys.argv[5]), int(sys.argv[6])
    threads = [int(x) for x in sys.argv[7].split()] if len(sys.argv) > 7 else [0]
    prec = int(sys.argv[8]) if len(sys.argv) > 8 else 0
    arr = util.code_arrays(name)
    cfg = q.DecoderConfig(decoding_algorithm=alg, max_iterations=100, message_precision=prec)
    for t in threads:
        try:
            with q.LdpcCode(arr["n"], arr["m"], arr["row_ptr"], arr["col_idx"], device=0, onchip_threads=t) as code:
                code.bench_synthetic(frames, qber, (pri, sec), cfg, seed=3)
                best = min(code.bench_synthetic(frames, qber, (pri, sec), cfg, seed=3)[1] for _ in range(3))
                tally, _ = code.bench_synthetic(frames, qber, (pri, sec), cfg, seed=3)
                inf = code.info()
                st = q.stats_from_tally(tally, frames)
                print("%s alg %d q %.4f frames %d threads %d -> %d path %d precision %d: %.3f Gbit/s (%.2f ms) mean it %.2f FER %.5f"
                      % (name, alg, qber, frames, t, inf["onchip_threads"], inf["last_path"], inf["last_precision"], arr["n"] * frames / best / 1e9,
                         best * 1e3, float(tally[q.decoder.TALLY_ITERATIONS]) / frames, st["FER"]), flush=True)
        except Exception as e:   # a CTA size the kernel does not take
            print("%s threads %d: %s" % (name, t, str(e)[:100]), flush=True)


if __name__ == "__main__":
    main()
