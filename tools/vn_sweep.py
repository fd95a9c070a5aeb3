"""Streaming path, narrow variable-node kernel: sweep of qkdldpc_options.vn_items_per_warp / vn_ctas_per_sm (and of the
batch size against the resident pool) through qkdldpc_bench_synthetic -- keys generated on the device, decode only, CUDA
events. One process, one handle per setting; prints decoded Gbit/s and, from a second pass with the per-launch events on,
the check-node / variable-node milliseconds and their share of the measured HBM peak.

    python tools/vn_sweep.py L100k 4096 2 0.72 0.06 "1:0 2:0 4:0 8:0 4:5 8:5 4:4 8:4"
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402

import qkd_ldpc_v_b200 as q  # noqa: E402


def main():
    name, frames, alg, pri, qber = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), float(sys.argv[5])
    # a setting: items:ctas[:steps_per_poll[:compaction_fill_pct[:1 = no tail compaction]]]
    settings = [(tuple(int(x) for x in s.split(":")) + (0, 0, 0))[:5] for s in sys.argv[6].split()]
    reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
    pool_slots = int(sys.argv[8]) if len(sys.argv) > 8 else 0
    prec = int(sys.argv[9]) if len(sys.argv) > 9 else 32
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6534.5
    arr = util.code_arrays(name)
    n, nnz = arr["n"], arr["nnz"]
    cfg = q.DecoderConfig(decoding_algorithm=alg, max_iterations=100, message_precision=prec)
    for items, ctas, spp, fill, nocomp in settings:
        with q.LdpcCode(n, arr["m"], arr["row_ptr"], arr["col_idx"], device=0, decoder_path=1, vn_items_per_warp=items,
                        vn_ctas_per_sm=ctas, pool_slots=pool_slots, steps_per_poll=spp,
                        compaction_fill_pct=fill, tail_compaction=-1 if nocomp else 0) as code:
            code.bench_synthetic(frames, qber, (pri, 0.0), cfg, seed=5)
            best = min(code.bench_synthetic(frames, qber, (pri, 0.0), cfg, seed=5)[1] for _ in range(reps))
            code.set_profiling(True)
            tally, _ = code.bench_synthetic(frames, qber, (pri, 0.0), cfg, seed=5)
            inf = code.info()
            executed = float(tally[q.decoder.TALLY_ITERATIONS])   # sum of executed iterations over the batch
            half = executed * 8.0 * nnz * (prec // 32)          # one kernel: 4 B read + 4 B written per edge and executed iteration
            print("%s f%d frames %d items %d ctas %d spp %d fill %d%s pool_tiles %d: %.3f Gbit/s  (%.1f ms)  mean it %.2f  cn %.2f ms (%.3f)  vn %.2f ms (%.3f)"
                  % (name, prec, frames, items, ctas, spp, fill, " no compaction" if nocomp else "", inf["pool_tiles"], n * frames / best / 1e9, best * 1e3, executed / frames,
                     inf["last_cn_ms"], half / (inf["last_cn_ms"] * 1e-3) / 1e9 / peak,
                     inf["last_vn_ms"], (half + executed * 4.0 * n * (prec // 32)) / (inf["last_vn_ms"] * 1e-3) / 1e9 / peak), flush=True)
        if os.environ.get("VN_SWEEP_HIST"):
            h = tally[q.decoder.TALLY_HIST:]
            print("   iterations of the converged frames:", {int(i): int(v) for i, v in enumerate(h) if v}, "steps", inf["decoder_steps"], flush=True)


if __name__ == "__main__":
    main()
