"""CSV-level parity: the reference's whole QKD_LDPC executable (oracle/_ref/qkd_ldpc_ref = its unmodified main.cpp and
sources, built by oracle/Makefile with -DUSE_CURRENT_DIR) and our qkdldpc_sim (C++ host + libqkdldpc_cuda) are run on the
same config and matrix directory; the results files must agree. With float64 messages the min-sum family is
bit-identical, so every CSV field must match; float32 runs are compared on FER."""
import json
import os
import subprocess

import pytest

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "qkd_ldpc_ref")
SIM_BIN = os.path.join(ROOT, "qkd_ldpc_v_b200", "qkdldpc_sim")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/qkd_ldpc_ref not built")]

BASE = dict(threads_number=4, use_config_simulation_seed=True, simulation_seed=20251018, enable_privacy_maintenance=False,
            enable_throughput_measurement=False, throughput_measurement_parameters=dict(consider_RTT=False, RTT=0.0),
            min_sum_normalized_parameters=dict(use_alpha_range=False, alpha_range=dict(begin=0.6, end=0.8, step=0.1),
                                               code_rate_alpha_maps=[dict(code_rate=0.95, alpha=0.75)]),
            min_sum_offset_parameters=dict(use_beta_range=False, beta_range=dict(begin=0.1, end=0.3, step=0.1),
                                           code_rate_beta_maps=[dict(code_rate=0.95, beta=0.2)]),
            adaptive_min_sum_normalized_parameters=dict(use_alpha_range=False, alpha_range=dict(begin=0.6, end=0.8, step=0.1),
                                                        code_rate_alpha_maps=[dict(code_rate=0.95, alpha=0.8)], use_nu_range=False,
                                                        nu_range=dict(begin=0.5, end=0.7, step=0.1),
                                                        code_rate_nu_maps=[dict(code_rate=0.95, nu=0.6)]),
            adaptive_min_sum_offset_parameters=dict(use_beta_range=False, beta_range=dict(begin=0.6, end=0.8, step=0.1),
                                                    code_rate_beta_maps=[dict(code_rate=0.95, beta=0.7)], use_sigma_range=False,
                                                    sigma_range=dict(begin=0.5, end=0.7, step=0.1),
                                                    code_rate_sigma_maps=[dict(code_rate=0.95, sigma=0.99)]),
            decoding_algorithm_max_iterations=100, trace_qkd_ldpc=False, trace_decoding_algorithm=False,
            trace_decoding_algorithm_llr=False, enable_decoding_algorithm_msg_llr_threshold=True,
            decoding_algorithm_msg_llr_threshold=100.0, enable_code_rate_adaptation=False,
            code_rate_adaptation_parameters=dict(enable_untainted_puncturing=False, use_adaptation_parameters_ranges=True,
                                                 code_rate_adaptation_parameters_ranges=[
                                                     dict(code_rate=0.95, delta=dict(begin=0.05, end=0.1, step=0.05),
                                                          efficiency=dict(begin=1.3, end=1.3, step=0.1))],
                                                 code_rate_QBER_adaptation_parameters_maps=[]))


def _setup(tmp_path, cfg, matrices):
    """matrices: list of (golden code name, format). Builds <tmp>/configs, <tmp>/sparse_matrices/<dir>."""
    (tmp_path / "configs").mkdir()
    (tmp_path / "configs" / "run.json").write_text(json.dumps(cfg))
    sub = {1: "matrices_alist", 3: "matrices_2"}[cfg["matrix_format"]]
    mdir = tmp_path / "sparse_matrices" / sub
    mdir.mkdir(parents=True)
    for name in matrices:
        fname = util.code_arrays(name)["file"]
        (util.write_alist if cfg["matrix_format"] == 1 else util.write_sparse2)(str(mdir / fname), name)
    return mdir


def _run_reference(tmp_path):
    with open(os.devnull) as nul:
        subprocess.run([REF_BIN], cwd=tmp_path, stdin=nul, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, check=True, timeout=900)
    files = [f for f in os.listdir(tmp_path / "results") if f.endswith(".csv")]
    assert len(files) == 1
    return files[0], (tmp_path / "results" / files[0]).read_text().splitlines()


def _run_ours(tmp_path, precision, extra=()):
    out = tmp_path / "results_gpu"
    r = subprocess.run([SIM_BIN, "--root", str(tmp_path), "--results-dir", str(out), "--precision", str(precision), "--quiet", *extra],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    files = [f for f in os.listdir(out) if f.endswith(".csv")]
    assert len(files) == 1
    assert os.path.exists(out / files[0].replace(".csv", ".gpu.json"))
    return files[0], (out / files[0]).read_text().splitlines()


def _strip_duration(name):
    return name.split(",sim_duration=")[0]


@pytest.mark.parametrize("alg", [2, 3, 4, 5])
def test_csv_identical_minsum_fp64(built, tmp_path, alg):
    cfg = dict(BASE, trials_number=400, decoding_algorithm=alg, matrix_format=1,
               code_rate_QBER_ranges=[dict(code_rate=0.7, QBER=dict(begin=0.03, end=0.05, step=0.01)),
                                      dict(code_rate=0.95, QBER=dict(begin=0.015, end=0.025, step=0.01))])
    _setup(tmp_path, cfg, ["K1_4", "K1_5"])
    rname, ref_lines = _run_reference(tmp_path)
    # 400 trials per combination: qkdldpc_sim decodes several combinations concurrently (own handle + stream each);
    # alg 2 is also run strictly one combination at a time -- the CSV must not depend on it
    # trial inputs are generated on the device by default (bit-identical to run_trial's); alg 3 uses the host generator
    oname, our_lines = _run_ours(tmp_path, 64, {2: ["--concurrent", "1"], 3: ["--host-keygen"]}.get(alg, []))
    assert _strip_duration(rname) == _strip_duration(oname)
    assert len(ref_lines) == 1 + 3 + 2
    # the reference enumerates matrices in directory order; so do we (same directory) -> rows line up
    assert our_lines == ref_lines


def test_csv_rate_adaptation_untainted_fp64(built, tmp_path):
    cra = dict(enable_untainted_puncturing=True, use_adaptation_parameters_ranges=False, code_rate_adaptation_parameters_ranges=[],
               code_rate_QBER_adaptation_parameters_maps=[dict(code_rate=0.805, QBER=0.0116, delta=0.09, efficiency=1.5),
                                                          dict(code_rate=0.805, QBER=0.0196, delta=0.03, efficiency=1.28),
                                                          dict(code_rate=0.805, QBER=0.0276, delta=0.11, efficiency=1.2)])
    cfg = dict(BASE, trials_number=48, decoding_algorithm=5, matrix_format=3, enable_code_rate_adaptation=True,
               code_rate_adaptation_parameters=cra, code_rate_QBER_ranges=[dict(code_rate=0.95, QBER=dict(begin=0.01, end=0.01, step=0.01))])
    _setup(tmp_path, cfg, ["I80"])
    rname, ref_lines = _run_reference(tmp_path)
    oname, our_lines = _run_ours(tmp_path, 64, ["--gpus", "1", "--chunk-frames", "20"])   # several chunks per combination
    assert _strip_duration(rname) == _strip_duration(oname)
    assert len(ref_lines) == 4 and "DELTA;EFFICIENCY;PUNCT_FRACTION" in ref_lines[0]
    assert our_lines == ref_lines


def test_csv_random_puncturing_and_spa(built, tmp_path):
    # SPA in float64: CUDA's tanh/atanh vs glibc's differ in the last ulp, so allow a small iteration-mean difference
    cra = dict(BASE["code_rate_adaptation_parameters"],
               code_rate_adaptation_parameters_ranges=[dict(code_rate=0.95, delta=dict(begin=0.05, end=0.1, step=0.05),
                                                            efficiency=dict(begin=1.8, end=2.0, step=0.2))])
    cfg = dict(BASE, trials_number=300, decoding_algorithm=0, matrix_format=1, enable_code_rate_adaptation=True,
               code_rate_adaptation_parameters=cra,
               code_rate_QBER_ranges=[dict(code_rate=0.95, QBER=dict(begin=0.01, end=0.015, step=0.005))])
    _setup(tmp_path, cfg, ["K1_5"])
    _, ref_lines = _run_reference(tmp_path)
    _, our_lines = _run_ours(tmp_path, 64)
    assert our_lines[0] == ref_lines[0] and len(our_lines) == len(ref_lines) >= 3
    for a, b in zip(our_lines[1:], ref_lines[1:]):
        fa, fb = a.split(";"), b.split(";")
        assert fa[:8] == fb[:8] and fa[15:] == fb[15:]          # identity columns and rate-adaptation columns
        assert abs(float(fa[8].replace(",", ".")) - float(fb[8].replace(",", "."))) <= 0.05   # ITER_SUCCESS_MEAN
        assert abs(float(fa[14].replace(",", ".")) - float(fb[14].replace(",", "."))) <= 0.01  # FER


def test_csv_fp32_fer(built, tmp_path):
    cfg = dict(BASE, trials_number=2000, decoding_algorithm=2, matrix_format=1,
               code_rate_QBER_ranges=[dict(code_rate=0.95, QBER=dict(begin=0.02, end=0.03, step=0.01))])
    _setup(tmp_path, cfg, ["K1_5"])
    _, ref_lines = _run_reference(tmp_path)
    _, our_lines = _run_ours(tmp_path, 32)
    assert our_lines[0] == ref_lines[0]
    for a, b in zip(our_lines[1:], ref_lines[1:]):
        fa, fb = a.split(";"), b.split(";")
        assert fa[:8] == fb[:8]
        assert abs(float(fa[14].replace(",", ".")) - float(fb[14].replace(",", "."))) <= 0.005
        if float(fb[12].replace(",", ".")) >= 0.5:   # the mean over a handful of successes is noise
            assert abs(float(fa[8].replace(",", ".")) - float(fb[8].replace(",", "."))) <= 0.1


def test_csv_trials_sharded_over_two_handles(built, tmp_path):
    """>= 2048 trials per combination: qkdldpc_sim splits the trials of a combination into contiguous ranges, one per
    device (here two handles on device 0, several chunks each) and sums the tallies; per-trial seeds are
    seeds[n] + combination index whatever the partition (quirk Q16), so the CSV equals the reference's."""
    cfg = dict(BASE, trials_number=2500, decoding_algorithm=2, matrix_format=1,
               code_rate_QBER_ranges=[dict(code_rate=0.95, QBER=dict(begin=0.02, end=0.025, step=0.005))])
    _setup(tmp_path, cfg, ["K1_5"])
    _, ref_lines = _run_reference(tmp_path)
    _, our_lines = _run_ours(tmp_path, 64, ["--devices", "0,0", "--chunk-frames", "700"])
    assert our_lines == ref_lines
    _, host_lines = _run_ours(tmp_path / "h", 64, ["--devices", "0,0", "--chunk-frames", "700", "--host-keygen", "--root", str(tmp_path)])
    assert host_lines == ref_lines


def _adaptive_t_config():
    """configs/ADAPTIVE T.json, the reference's one live config, restated: AOMSA, rate adaptation with
    untainted puncturing from the per-(rate, QBER) map, privacy maintenance, throughput measurement with RTT 0.4 ms, 10 trials
    per combination, 1 thread, seed 5555, the three irregular n = 10240 codes of matrices_2."""
    amap = [(0.905, 0.48, 0.88), (0.855, 0.82, 1.22), (0.805, 0.7, 0.99), (0.755, 0.8, 1.05), (0.705, 0.91, 1.11), (0.655, 0.74, 0.94),
            (0.605, 0.59, 0.8), (0.555, 0.61, 0.8), (0.505, 0.72, 0.85)]
    qmap = {0.805: [(0.0076, 0.1, 1.85), (0.0116, 0.09, 1.5), (0.0156, 0.06, 1.39), (0.0196, 0.03, 1.28), (0.0236, 0.01, 1.21),
                    (0.0276, 0.11, 1.2), (0.0316, 0.22, 1.22)],
            0.655: [(0.0368, 0.12, 1.22), (0.0408, 0.1, 1.2), (0.0448, 0.06, 1.18), (0.0488, 0.06, 1.17), (0.0528, 0.01, 1.16),
                    (0.0568, 0.06, 1.16), (0.0608, 0.11, 1.17), (0.0648, 0.16, 1.17), (0.0688, 0.23, 1.19)],
            0.505: [(0.0734, 0.12, 1.18), (0.0774, 0.09, 1.17), (0.0814, 0.08, 1.16), (0.0854, 0.04, 1.15), (0.0894, 0.01, 1.14),
                    (0.0934, 0.03, 1.14), (0.0974, 0.06, 1.14), (0.1014, 0.1, 1.15), (0.1054, 0.13, 1.16), (0.1094, 0.13, 1.15)]}
    cra = dict(enable_untainted_puncturing=True, use_adaptation_parameters_ranges=False,
               code_rate_adaptation_parameters_ranges=[
                   dict(code_rate=0.805, delta=dict(begin=0.01, end=0.25, step=0.01), efficiency=dict(begin=1.2, end=1.6, step=0.01)),
                   dict(code_rate=0.655, delta=dict(begin=0.05, end=0.2, step=0.05), efficiency=dict(begin=1.13, end=1.28, step=0.01)),
                   dict(code_rate=0.505, delta=dict(begin=0.05, end=0.2, step=0.05), efficiency=dict(begin=1.12, end=1.28, step=0.01))],
               code_rate_QBER_adaptation_parameters_maps=[dict(code_rate=r, QBER=qb, delta=d, efficiency=e)
                                                          for r, rows in qmap.items() for qb, d, e in rows])
    return dict(BASE, threads_number=1, trials_number=10, simulation_seed=5555, enable_privacy_maintenance=True,
                enable_throughput_measurement=True, throughput_measurement_parameters=dict(consider_RTT=True, RTT=0.4),
                decoding_algorithm=5, matrix_format=3, enable_code_rate_adaptation=True, code_rate_adaptation_parameters=cra,
                adaptive_min_sum_offset_parameters=dict(
                    use_beta_range=False, beta_range=dict(begin=0.01, end=1.0, step=0.01),
                    code_rate_beta_maps=[dict(code_rate=r, beta=b) for r, b, _ in amap], use_sigma_range=False,
                    sigma_range=dict(begin=0.01, end=1.0, step=0.01), code_rate_sigma_maps=[dict(code_rate=r, sigma=sg) for r, _, sg in amap]),
                code_rate_QBER_ranges=[dict(code_rate=0.805, QBER=dict(begin=0.0096, end=0.0196, step=0.002)),
                                       dict(code_rate=0.655, QBER=dict(begin=0.0338, end=0.0738, step=0.001)),
                                       dict(code_rate=0.505, QBER=dict(begin=0.0714, end=0.1114, step=0.001))])


@pytest.mark.parametrize("extra", [[], ["--trials", "10", "--concurrent", "1"]], ids=["sweep", "one_combination_per_call"])
def test_csv_adaptive_t_privacy_maintenance_and_throughput(built, tmp_path, extra):
    """configs/ADAPTIVE T.json (privacy maintenance + throughput columns + RTT): every column but the four THROUGHPUT_* ones
    must equal the reference executable's under the DEFAULT precision policy (AOMSA -> float64 state on chip); the
    throughput columns are the RTT model on the batch time: out_key_length / (batch_time / trials + RTT)
    (simulation.cpp:626-681), with out_key_length = n - |bits_to_remove| and remove_bits executed on the device inside
    the timed call."""
    cfg = _adaptive_t_config()
    _setup(tmp_path, cfg, ["I80", "I65", "I50"])
    rname, ref_lines = _run_reference(tmp_path)
    oname, our_lines = _run_ours(tmp_path, 0, extra)
    assert _strip_duration(rname) == _strip_duration(oname) and ",RTT=0.400ms," in oname and "priv_maint=ON" in oname
    hdr = ref_lines[0].split(";")
    assert our_lines[0] == ref_lines[0] and len(our_lines) == len(ref_lines) == 1 + 7 + 9 + 10
    thr = [hdr.index(k) for k in ("THROUGHPUT_MEAN", "THROUGHPUT_STD", "THROUGHPUT_MIN", "THROUGHPUT_MAX")]
    side = json.loads((tmp_path / "results_gpu" / oname.replace(".csv", ".gpu.json")).read_text())["combinations"]
    n = util.code_arrays("I80")["n"]
    for a, b, sc in zip(our_lines[1:], ref_lines[1:], side):
        fa, fb = a.split(";"), b.split(";")
        assert [x for k, x in enumerate(fa) if k not in thr] == [x for k, x in enumerate(fb) if k not in thr]
        assert 0 < sc["out_key_length"] < n     # bits were removed
        want = sc["out_key_length"] * 1e6 / (sc["batch_ms"] * 1e3 / 10 + 0.4 * 1000.)
        mean, std, lo, hi = (int(fa[k]) for k in thr)
        assert abs(mean - want) <= 1.0 + 1e-6 * want and lo == hi == mean and std == 0
        assert mean <= sc["out_key_length"] / 0.4e-3     # the RTT bounds the rate


def test_example_program(built):
    """example/qkdldpc_example.cpp = the reference's example (N = 6, Johnson ex. 2.5) through the C ABI: iteration counts
    traced from the unmodified reference: SPA 1, SPA-lin 1, NMSA 1, OMSA 2, ANMSA 3, AOMSA 2."""
    exe = os.path.join(ROOT, "qkd_ldpc_v_b200", "qkdldpc_example")
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 6
    for line, it in zip(lines, (1, 1, 1, 2, 3, 2)):
        assert f"iterations {it}," in line and "syndromes match, keys match" in line and line.endswith("0 0 1 0 1 1"), line
