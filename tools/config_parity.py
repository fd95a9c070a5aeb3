#!/usr/bin/env python
"""Config-level parity and timing: the reference's whole executable (oracle/_ref/qkd_ldpc_ref, unmodified sources) and
qkdldpc_sim (C++ host + libqkdldpc_cuda) run the SAME legacy-schema config on the SAME matrix directory; their CSV files
are compared field by field and both runs are timed.

Two of BASELINE.json's configs, restricted to the matrices whose graphs ship in tests/golden/codes.npz:
  config1k    `config 1k.json`      (schema v1 => SPA), 1k alist codes R = 0.47 / 0.66 / 0.76 / 0.92 at that file's QBERs
  config10k   `config 10k NMSA.json` (schema v1 => NMSA), 10k alist codes R = 0.79 / 0.82 at that file's alpha / QBER
The reference arm is run with --ref-trials (CPU time!), qkdldpc_sim additionally with the config's full trial count.

  python tools/config_parity.py --config config1k --ref-trials 20000 --gpus 1 --out gpurun_out/config1k.json
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "qkd_ldpc_ref")
SIM_BIN = os.path.join(ROOT, "qkd_ldpc_v_b200", "qkdldpc_sim")

COMMON = dict(threads_number=16, use_config_simulation_seed=True, interactive_mode=False, enable_privacy_maintenance=False,
              enable_throughput_measurement=False, throughput_measurement_parameters=dict(consider_RTT=True, RTT=20),
              decoding_algorithm_max_iterations=100, matrix_format=1, trace_qkd_ldpc=False, trace_decoding_algorithm=False,
              trace_decoding_algorithm_llr=False, enable_decoding_algorithm_msg_llr_threshold=True,
              decoding_algorithm_msg_llr_threshold=100.0)
# the entries of the reference's files that the golden matrices select (first code_rate >= R)
CONFIGS = {
    "config1k": dict(codes=["K1_3", "K1_4", "K1_5", "K1_hi"], trials=100000, cfg=dict(
        COMMON, simulation_seed=9012025, use_min_sum_normalized_algorithm=False,
        min_sum_normalized_parameters=dict(use_alpha_range=False, alpha_range=dict(begin=0.01, end=1.0, step=0.01),
                                           code_rate_alpha_maps=[dict(code_rate=0.36, alpha=0.55), dict(code_rate=0.58, alpha=0.72),
                                                                 dict(code_rate=0.95, alpha=0.76)]),
        code_rate_QBER_maps=[dict(code_rate=0.475, QBER_begin=0.069, QBER_end=0.069, QBER_step=0.001),
                             dict(code_rate=0.665, QBER_begin=0.03, QBER_end=0.03, QBER_step=0.001),
                             dict(code_rate=0.765, QBER_begin=0.016, QBER_end=0.016, QBER_step=0.001),
                             dict(code_rate=0.925, QBER_begin=0.001, QBER_end=0.001, QBER_step=0.001)])),
    "config10k": dict(codes=["A79", "A82"], trials=1000000, cfg=dict(
        COMMON, simulation_seed=10012025, use_min_sum_normalized_algorithm=True,
        min_sum_normalized_parameters=dict(use_alpha_range=False, alpha_range=dict(begin=0.01, end=1.0, step=0.01),
                                           code_rate_alpha_maps=[dict(code_rate=0.795, alpha=0.71), dict(code_rate=0.825, alpha=0.7)]),
        code_rate_QBER_maps=[dict(code_rate=0.795, QBER_begin=0.02, QBER_end=0.02, QBER_step=0.001),
                             dict(code_rate=0.825, QBER_begin=0.015, QBER_end=0.015, QBER_step=0.001)])),
}


def wilson(k, n, z=1.96):
    p = k / n
    d = 1 + z * z / n
    c = (p + z * z / (2 * n)) / d
    h = z * math.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / d
    return c - h, c + h


def to_v4(cfg):
    """The same run in the schema the reference's parser accepts today (config.cpp:89-403); the archived configs are
    schema v1, which only our parser reads."""
    nm = cfg["min_sum_normalized_parameters"]
    unused = dict(use_range=False, rng=dict(begin=0.1, end=1.0, step=0.1), maps=[dict(code_rate=0.99, v=0.5)])
    def block(prim, sec=None):
        b = {f"use_{prim}_range": unused["use_range"], f"{prim}_range": unused["rng"],
             f"code_rate_{prim}_maps": [{"code_rate": 0.99, prim: 0.5}]}
        if sec:
            b.update({f"use_{sec}_range": False, f"{sec}_range": unused["rng"], f"code_rate_{sec}_maps": [{"code_rate": 0.99, sec: 0.5}]})
        return b
    out = {k: v for k, v in cfg.items() if k not in ("use_min_sum_normalized_algorithm", "code_rate_QBER_maps", "interactive_mode")}
    out["decoding_algorithm"] = 2 if cfg["use_min_sum_normalized_algorithm"] else 0
    out["min_sum_normalized_parameters"] = nm
    out["min_sum_offset_parameters"] = block("beta")
    out["adaptive_min_sum_normalized_parameters"] = block("alpha", "nu")
    out["adaptive_min_sum_offset_parameters"] = block("beta", "sigma")
    out["code_rate_QBER_ranges"] = [dict(code_rate=m["code_rate"], QBER=dict(begin=m["QBER_begin"], end=m["QBER_end"], step=m["QBER_step"]))
                                    for m in cfg["code_rate_QBER_maps"]]
    out["enable_code_rate_adaptation"] = False
    out["code_rate_adaptation_parameters"] = dict(
        enable_untainted_puncturing=False, use_adaptation_parameters_ranges=True,
        code_rate_adaptation_parameters_ranges=[dict(code_rate=0.99, delta=dict(begin=0.05, end=0.1, step=0.05),
                                                     efficiency=dict(begin=1.3, end=1.3, step=0.1))],
        code_rate_QBER_adaptation_parameters_maps=[])
    return out


def setup(run_dir, spec, trials, legacy):
    """legacy=True: the archived schema-v1 file (qkdldpc_sim); False: its v4 translation (the reference executable)."""
    os.makedirs(os.path.join(run_dir, "configs"))
    mdir = os.path.join(run_dir, "sparse_matrices", "matrices_alist")
    os.makedirs(mdir)
    cfg = dict(spec["cfg"], trials_number=trials)
    with open(os.path.join(run_dir, "configs", "run.json"), "w") as f:
        json.dump(cfg if legacy else to_v4(cfg), f)
    for name in spec["codes"]:
        util.write_alist(os.path.join(mdir, util.code_arrays(name)["file"]), name)


def read_csv(directory):
    files = [f for f in os.listdir(directory) if f.endswith(".csv")]
    assert len(files) == 1, files
    rows = [ln.split(";") for ln in open(os.path.join(directory, files[0])).read().splitlines()]
    return rows[0], {r[1]: r for r in rows[1:]}


def num(x):
    return float(x.replace(",", "."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="config1k", choices=sorted(CONFIGS))
    ap.add_argument("--ref-trials", type=int, default=20000)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--full", action="store_true", help="also run qkdldpc_sim with the config's own trial count")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    spec = CONFIGS[args.config]
    report = {"config": args.config, "codes": spec["codes"], "ref_trials": args.ref_trials, "runs": {}}

    # reference executable vs qkdldpc_sim (float32 and float64) at the same trial count
    tmp_ref = tempfile.mkdtemp(prefix="cfgref_")
    setup(tmp_ref, spec, args.ref_trials, legacy=False)
    t0 = time.perf_counter()
    with open(os.devnull) as nul:
        subprocess.run([REF_BIN], cwd=tmp_ref, stdin=nul, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    t_ref = time.perf_counter() - t0
    hdr, ref = read_csv(os.path.join(tmp_ref, "results"))
    tmp = tempfile.mkdtemp(prefix="cfgpar_")
    setup(tmp, spec, args.ref_trials, legacy=True)
    report["runs"]["reference_cpu"] = {"seconds": t_ref, "threads": spec["cfg"]["threads_number"], "trials": args.ref_trials}
    for prec in (32, 64):
        out = os.path.join(tmp, f"results_gpu{prec}")
        t0 = time.perf_counter()
        subprocess.run([SIM_BIN, "--root", tmp, "--results-dir", out, "--precision", str(prec), "--gpus", str(args.gpus), "--quiet"],
                       check=True, stdout=subprocess.DEVNULL)
        dt = time.perf_counter() - t0
        _, ours = read_csv(out)
        rows = []
        for name, r in ref.items():
            o = ours[name]
            fails_ref = round(num(r[14]) * args.ref_trials)
            lo, hi = wilson(fails_ref, args.ref_trials)
            rows.append({"matrix": name, "qber": num(r[6]), "fer_ref": num(r[14]), "fer_gpu": num(o[14]), "fer_ci95": [lo, hi],
                         "fer_inside_ci": lo - 1e-12 <= num(o[14]) <= hi + 1e-12, "iter_mean_ref": num(r[8]), "iter_mean_gpu": num(o[8]),
                         "row_identical": o == r})
        report["runs"][f"qkdldpc_sim_fp{prec}"] = {"seconds": dt, "trials": args.ref_trials, "gpus": args.gpus, "rows": rows,
                                                  "csv_identical": all(x["row_identical"] for x in rows)}
    if args.full:
        tmp2 = tempfile.mkdtemp(prefix="cfgfull_")
        setup(tmp2, spec, spec["trials"], legacy=True)
        out = os.path.join(tmp2, "results_gpu")
        t0 = time.perf_counter()
        subprocess.run([SIM_BIN, "--root", tmp2, "--results-dir", out, "--gpus", str(args.gpus), "--quiet"], check=True, stdout=subprocess.DEVNULL)
        dt = time.perf_counter() - t0
        _, ours = read_csv(out)
        side = [f for f in os.listdir(out) if f.endswith(".gpu.json")][0]
        report["runs"]["qkdldpc_sim_full"] = {
            "seconds": dt, "trials": spec["trials"], "gpus": args.gpus,
            "rows": [{"matrix": k, "qber": num(v[6]), "fer": num(v[14]), "iter_mean": num(v[8])} for k, v in ours.items()],
            "sidecar": json.load(open(os.path.join(out, side)))["combinations"],
            "reference_cpu_seconds_extrapolated": t_ref * spec["trials"] / args.ref_trials}
    text = json.dumps(report, indent=1)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
