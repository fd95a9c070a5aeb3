#!/usr/bin/env python
"""Config-level parity and timing: the reference's whole executable (oracle/_ref/qkd_ldpc_ref, unmodified sources) and
qkdldpc_sim (C++ host + libqkdldpc_cuda) run the SAME legacy-schema config on the SAME matrix directory; their CSV files
are compared field by field and both runs are timed.

Two of BASELINE.json's configs, restricted to the matrices whose graphs ship in tests/golden/codes.npz:
  nopt_spa / nopt_spalin  `NOPT_R=0,82_SPA.json`, `NOPT_R=0,82_SPA_LIN_APPROX.json` (schema v2), 30 000 trials on A82
  config100k / config100k_nmsa  `config 100k.json` (schema v1 => SPA as shipped; NMSA variant), the n = 102400 R = 0.49 code
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


def adaptive_r(coarse):
    """`ADAPTIVE R.json` (schema v3, AOMSA, untainted puncturing) for the three irregular 10k codes of matrices_2.
    coarse=True thins the QBER / delta / f_EC grids so that the CPU reference finishes in about a minute."""
    q, d, e = (0.01, 0.1, 0.25) if coarse else (0.001, 0.05, 0.05)
    rates = [(0.805, 0.0046, 0.0446, 0.7, 0.99), (0.655, 0.0338, 0.0738, 0.74, 0.94), (0.505, 0.0714, 0.1114, 0.72, 0.85)]
    return dict(
        {k: v for k, v in COMMON.items() if k != "matrix_format"}, matrix_format=3, simulation_seed=5555,
        throughput_measurement_parameters=dict(consider_RTT=True, RTT=1.0), decoding_algorithm=5,
        adaptive_min_sum_offset_parameters=dict(
            use_beta_range=False, beta_range=dict(begin=0.01, end=1.0, step=0.01),
            code_rate_beta_maps=[dict(code_rate=r, beta=b) for r, _, _, b, _ in rates], use_sigma_range=False,
            sigma_range=dict(begin=0.01, end=1.0, step=0.01), code_rate_sigma_maps=[dict(code_rate=r, sigma=sg) for r, _, _, _, sg in rates]),
        code_rate_QBER_maps=[dict(code_rate=r, QBER=dict(begin=b0, end=b1, step=q)) for r, b0, b1, _, _ in rates],
        enable_code_rate_adaptation=True, enable_untainted_puncturing=True,
        code_rate_adaptation_parameters_maps=[dict(code_rate=r, delta=dict(begin=0.05, end=0.3 if not coarse else 0.25, step=d),
                                                   efficiency=dict(begin=1.0, end=2.0, step=e)) for r, *_ in rates])


def nopt(alg):
    """`NOPT_R=0,82_SPA.json` / `NOPT_R=0,82_SPA_LIN_APPROX.json` (schema v2: decoding_algorithm + every parameter block;
    30 000 trials, seed 777, the FER ~ 0.01 operating point QBER 1.62 % of the alist n = 10240 R = 0.82 code). The files
    say threads_number = 1; the reference arm here uses all cores of the box (results do not depend on it, quirk Q16)."""
    rng = dict(begin=0.01, end=1.0, step=0.01)
    return dict(
        COMMON, simulation_seed=777, decoding_algorithm=alg,
        min_sum_normalized_parameters=dict(use_alpha_range=False, alpha_range=rng, code_rate_alpha_maps=[dict(code_rate=0.995, alpha=0.8)]),
        min_sum_offset_parameters=dict(use_beta_range=False, beta_range=rng, code_rate_beta_maps=[dict(code_rate=0.995, beta=1.26)]),
        adaptive_min_sum_normalized_parameters=dict(use_alpha_range=False, alpha_range=rng, code_rate_alpha_maps=[dict(code_rate=0.995, alpha=0.88)],
                                                    use_nu_range=False, nu_range=rng, code_rate_nu_maps=[dict(code_rate=0.995, nu=0.79)]),
        adaptive_min_sum_offset_parameters=dict(use_beta_range=False, beta_range=rng, code_rate_beta_maps=[dict(code_rate=0.995, beta=0.91)],
                                                use_sigma_range=False, sigma_range=rng, code_rate_sigma_maps=[dict(code_rate=0.995, sigma=1.11)]),
        code_rate_QBER_maps=[dict(code_rate=0.825, QBER_begin=0.0162, QBER_end=0.0162, QBER_step=0.0005)])


def config_100k(nmsa):
    """`config 100k.json` (schema v1; SPA as shipped, NMSA as BASELINE.json describes it) on the one n = 102400 code of the
    golden set (R = 0.49 -> QBER 8.4 %, alpha 0.72), 100 trials, seed 9012025."""
    return dict(
        COMMON, simulation_seed=9012025, use_min_sum_normalized_algorithm=nmsa,
        min_sum_normalized_parameters=dict(use_alpha_range=False, alpha_range=dict(begin=0.01, end=1.0, step=0.01),
                                           code_rate_alpha_maps=[dict(code_rate=0.36, alpha=0.55), dict(code_rate=0.58, alpha=0.72),
                                                                 dict(code_rate=0.95, alpha=0.76)]),
        code_rate_QBER_maps=[dict(code_rate=0.475, QBER_begin=0.089, QBER_end=0.089, QBER_step=0.001),
                             dict(code_rate=0.495, QBER_begin=0.084, QBER_end=0.084, QBER_step=0.001),
                             dict(code_rate=0.515, QBER_begin=0.079, QBER_end=0.079, QBER_step=0.001)])


CONFIGS["config100k"] = dict(codes=["L100k"], trials=100, cfg=config_100k(False))
CONFIGS["config100k_nmsa"] = dict(codes=["L100k"], trials=100, cfg=config_100k(True))
CONFIGS["nopt_spa"] = dict(codes=["A82"], trials=30000, cfg=nopt(0))
CONFIGS["nopt_spalin"] = dict(codes=["A82"], trials=30000, cfg=nopt(1))
CONFIGS["adaptiveR"] = dict(codes=["I80", "I65", "I50"], trials=100, cfg=adaptive_r(False), ref_cfg=adaptive_r(True))


def wilson(k, n, z=1.96):
    p = k / n
    d = 1 + z * z / n
    c = (p + z * z / (2 * n)) / d
    h = z * math.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / d
    return c - h, c + h


def to_v4(cfg):
    """The same run in the schema the reference's parser accepts today (config.cpp:89-403); the archived configs are
    schema v1, which only our parser reads."""
    nm = cfg.get("min_sum_normalized_parameters")
    unused = dict(use_range=False, rng=dict(begin=0.1, end=1.0, step=0.1), maps=[dict(code_rate=0.99, v=0.5)])
    def block(prim, sec=None):
        b = {f"use_{prim}_range": unused["use_range"], f"{prim}_range": unused["rng"],
             f"code_rate_{prim}_maps": [{"code_rate": 0.99, prim: 0.5}]}
        if sec:
            b.update({f"use_{sec}_range": False, f"{sec}_range": unused["rng"], f"code_rate_{sec}_maps": [{"code_rate": 0.99, sec: 0.5}]})
        return b
    legacy_keys = ("use_min_sum_normalized_algorithm", "code_rate_QBER_maps", "interactive_mode", "enable_untainted_puncturing",
                   "code_rate_adaptation_parameters_maps")
    out = {k: v for k, v in cfg.items() if k not in legacy_keys}
    if "decoding_algorithm" not in cfg:
        out["decoding_algorithm"] = 2 if cfg["use_min_sum_normalized_algorithm"] else 0
    out["min_sum_normalized_parameters"] = nm or block("alpha")
    out.setdefault("min_sum_offset_parameters", block("beta"))
    out.setdefault("adaptive_min_sum_normalized_parameters", block("alpha", "nu"))
    out.setdefault("adaptive_min_sum_offset_parameters", block("beta", "sigma"))
    out["code_rate_QBER_ranges"] = [dict(code_rate=m["code_rate"], QBER=m["QBER"] if "QBER" in m else
                                         dict(begin=m["QBER_begin"], end=m["QBER_end"], step=m["QBER_step"])) for m in cfg["code_rate_QBER_maps"]]
    out["enable_code_rate_adaptation"] = cfg.get("enable_code_rate_adaptation", False)
    out["code_rate_adaptation_parameters"] = dict(
        enable_untainted_puncturing=cfg.get("enable_untainted_puncturing", False), use_adaptation_parameters_ranges=True,
        code_rate_adaptation_parameters_ranges=cfg.get("code_rate_adaptation_parameters_maps") or
        [dict(code_rate=0.99, delta=dict(begin=0.05, end=0.1, step=0.05), efficiency=dict(begin=1.3, end=1.3, step=0.1))],
        code_rate_QBER_adaptation_parameters_maps=[])
    return out


def setup(run_dir, spec, trials, legacy, coarse=False):
    """legacy=True: the archived legacy-schema file (qkdldpc_sim); False: its v4 translation (the reference executable).
    coarse: the thinned parameter grid of configs whose full grid is too much CPU work for the reference arm."""
    os.makedirs(os.path.join(run_dir, "configs"))
    fmt = spec["cfg"]["matrix_format"]
    mdir = os.path.join(run_dir, "sparse_matrices", {1: "matrices_alist", 3: "matrices_2"}[fmt])
    os.makedirs(mdir)
    cfg = dict(spec["ref_cfg"] if coarse and "ref_cfg" in spec else spec["cfg"], trials_number=trials)
    with open(os.path.join(run_dir, "configs", "run.json"), "w") as f:
        json.dump(cfg if legacy else to_v4(cfg), f)
    for name in spec["codes"]:
        (util.write_alist if fmt == 1 else util.write_sparse2)(os.path.join(mdir, util.code_arrays(name)["file"]), name)


def read_csv(directory):
    """header, {(matrix, config QBER, delta, f_EC): row}"""
    files = [f for f in os.listdir(directory) if f.endswith(".csv")]
    assert len(files) == 1, files
    rows = [ln.split(";") for ln in open(os.path.join(directory, files[0])).read().splitlines()]
    adapt = "DELTA" in rows[0]
    i_d = rows[0].index("DELTA") if adapt else None
    return rows[0], {(r[1], r[6]) + ((r[i_d], r[i_d + 1]) if adapt else ()): r for r in rows[1:]}


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
    setup(tmp_ref, spec, args.ref_trials, legacy=False, coarse=True)
    t0 = time.perf_counter()
    with open(os.devnull) as nul:
        subprocess.run([REF_BIN], cwd=tmp_ref, stdin=nul, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    t_ref = time.perf_counter() - t0
    hdr, ref = read_csv(os.path.join(tmp_ref, "results"))
    tmp = tempfile.mkdtemp(prefix="cfgpar_")
    setup(tmp, spec, args.ref_trials, legacy=True, coarse=True)
    report["runs"]["reference_cpu"] = {"seconds": t_ref, "threads": spec["cfg"]["threads_number"], "trials": args.ref_trials}
    for prec in (0, 32, 64):   # 0 = the library's precision policy (what a user gets without --precision)
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
            rows.append({"matrix": name[0], "qber": num(r[6]), "fer_ref": num(r[14]), "fer_gpu": num(o[14]), "fer_ci95": [lo, hi],
                         "fer_inside_ci": lo - 1e-12 <= num(o[14]) <= hi + 1e-12, "iter_mean_ref": num(r[8]), "iter_mean_gpu": num(o[8]),
                         "row_identical": o == r})
        report["runs"][f"qkdldpc_sim_fp{prec}" if prec else "qkdldpc_sim_default"] = {"seconds": dt, "trials": args.ref_trials, "gpus": args.gpus, "rows": rows,
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
            "combinations": len(ours),
            "rows": [{"matrix": k[0], "qber": num(v[6]), "fer": num(v[14]), "iter_mean": num(v[8])} for k, v in list(ours.items())[:64]],
            "sidecar": json.load(open(os.path.join(out, side)))["combinations"][:64],
            # the reference's time scales with trials x combinations (the coarse grid of the reference arm has fewer)
            "reference_cpu_seconds_extrapolated": t_ref * (spec["trials"] / args.ref_trials) * (len(ours) / max(1, len(ref)))}
    text = json.dumps(report, indent=1)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
