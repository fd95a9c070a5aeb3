"""Host surface of the drop-in (C++: qkd_ldpc_v_b200/host) against the reference's own host code (oracle/_ref):
matrix loaders, config parser (schema v1-v4), combination builder incl. rate adaptation, statistics and the CSV
writer. CPU only."""
import glob
import json
import os

import numpy as np
import pytest

from qkd_ldpc_v_b200 import hostlib

REF_ROOT = "/root/reference"
needs_ref_tree = pytest.mark.skipif(not os.path.isdir(os.path.join(REF_ROOT, "src")), reason="reference tree absent")
HERE = os.path.dirname(os.path.abspath(__file__))


def _ref():
    from oracle import ref
    if not ref.available():
        ref.build()
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    return ref


def _first(pattern):
    files = sorted(glob.glob(os.path.join(REF_ROOT, "sparse_matrices", pattern)))
    assert files, pattern
    return files[0]


@needs_ref_tree
@pytest.mark.parametrize("fmt,pattern", [(0, "matrices_uncompressed/*.mtrx"), (1, "matrices_alist/*.mtrx"),
                                         (1, "matrices_alist_1k_all/*R=0.5*.mtrx"), (2, "matrices_1/*.mtrx"),
                                         (3, "matrices_2/*R=0.8*.mtrx")])
def test_loaders_match_reference(built, fmt, pattern):
    ref = _ref()
    path = _first(pattern)
    ours, theirs = hostlib.HostMatrix(path, fmt), ref.RefMatrix(path, fmt)
    assert (ours.n, ours.m, ours.nnz, ours.is_regular) == (theirs.n, theirs.m, theirs.nnz, theirs.is_regular)
    rp, ci = ours.csr()
    cp, ri = ours.csc()
    assert (rp == theirs.row_ptr).all() and (ci == theirs.col_idx).all()
    assert (cp == theirs.col_ptr).all() and (ri == theirs.row_idx).all()


def test_loader_rejects_bad_files(built, tmp_path):
    bad = tmp_path / "bad.mtrx"
    bad.write_text("3 2\n1 1\n1 1 1\n2 1\n1\n1\n1\n1 2 9\n3\n")   # alist with an out-of-range index
    with pytest.raises(RuntimeError):
        hostlib.HostMatrix(str(bad), 1)
    with pytest.raises(RuntimeError):
        hostlib.HostMatrix(str(tmp_path / "missing.mtrx"), 1)
    with pytest.raises(RuntimeError):
        hostlib.HostMatrix(str(bad), 7)


def test_unsorted_adjacency_is_refused(built, tmp_path):
    # sparse_2 file whose first row lists its bits in descending order: the reference would silently mis-pair
    # message slots (quirk Q1); the host refuses to hand such a graph to the GPU library
    f = tmp_path / "unsorted.mtrx"
    f.write_text("4 2\n1 0\n2 3\n0\n0\n1\n1\n")
    m = hostlib.HostMatrix(str(f), 3)
    with pytest.raises(RuntimeError, match="ascending"):
        m.csr()


@needs_ref_tree
def test_config_v4_matches_reference(built):
    ref = _ref()
    for path in sorted(glob.glob(os.path.join(REF_ROOT, "configs", "*.json"))):
        ours, ver = hostlib.describe_config(path)
        assert ver == 4
        assert ours == ref.describe_config(path)


LEGACY = {
    1: dict(threads_number=16, trials_number=1000, use_config_simulation_seed=True, simulation_seed=9012025,
            enable_privacy_maintenance=False, enable_throughput_measurement=False, use_min_sum_normalized_algorithm=True,
            min_sum_normalized_parameters=dict(use_alpha_range=False, alpha_range=dict(begin=0.1, end=1.0, step=0.1),
                                               code_rate_alpha_maps=[dict(code_rate=0.9, alpha=0.8), dict(code_rate=0.5, alpha=0.7)]),
            decoding_algorithm_max_iterations=100, matrix_format=1, trace_qkd_ldpc=False, trace_decoding_algorithm=False,
            trace_decoding_algorithm_llr=False, enable_decoding_algorithm_msg_llr_threshold=True,
            decoding_algorithm_msg_llr_threshold=100.0,
            code_rate_QBER_maps=[dict(code_rate=0.9, QBER_begin=0.01, QBER_end=0.01, QBER_step=0.001),
                                 dict(code_rate=0.5, QBER_begin=0.05, QBER_end=0.07, QBER_step=0.01)]),
}
LEGACY[2] = {**{k: v for k, v in LEGACY[1].items() if k != "use_min_sum_normalized_algorithm"}, "decoding_algorithm": 3,
             "min_sum_offset_parameters": dict(use_beta_range=True, beta_range=dict(begin=0.2, end=0.4, step=0.1),
                                               code_rate_beta_maps=[])}
LEGACY[3] = {**{k: v for k, v in LEGACY[2].items() if k != "code_rate_QBER_maps"}, "decoding_algorithm": 2,
             "code_rate_QBER_maps": [dict(code_rate=0.9, QBER=dict(begin=0.01, end=0.02, step=0.005))],
             "enable_code_rate_adaptation": True, "enable_untainted_puncturing": False,
             "code_rate_adaptation_parameters_maps": [dict(code_rate=0.9, delta=dict(begin=0.1, end=0.2, step=0.1),
                                                          efficiency=dict(begin=1.1, end=1.1, step=0.1))]}


@pytest.mark.parametrize("ver", [1, 2, 3])
def test_config_legacy_schemas(built, tmp_path, ver):
    f = tmp_path / f"v{ver}.json"
    f.write_text(json.dumps(LEGACY[ver]))
    text, got = hostlib.describe_config(str(f))
    assert got == ver
    first = text.splitlines()[0]
    assert {1: " alg=2 ", 2: " alg=3 ", 3: " alg=2 "}[ver] in first
    assert "trials=1000" in first and "seed=9012025" in first
    # maps are sorted ascending by code rate after parsing (config.cpp:43-47, 289-293)
    if ver == 1:
        assert "maps=0.5>0.69999999999999996,0.90000000000000002>0.80000000000000004," in text
        assert "qber_ranges=0.5>" in text
    if ver == 3:
        assert " adapt=1 " in first and " adapt_ranges=1" in first and "adapt_param_ranges=0.9" in text


def test_config_errors(built, tmp_path):
    f = tmp_path / "bad.json"
    f.write_text(json.dumps({**LEGACY[2], "decoding_algorithm": 9}))
    with pytest.raises(RuntimeError):
        hostlib.describe_config(str(f))
    f.write_text("{ not json")
    with pytest.raises(RuntimeError):
        hostlib.describe_config(str(f))
    f.write_text(json.dumps({**LEGACY[2], "trials_number": 0}))
    with pytest.raises(RuntimeError):
        hostlib.describe_config(str(f))


def _write_cfg(tmp_path, name, **over):
    with open(os.path.join(REF_ROOT, "configs", "ADAPTIVE T.json")) as fh:
        cfg = json.load(fh)
    cfg.update(over)
    p = tmp_path / name
    p.write_text(json.dumps(cfg))
    return str(p)


@needs_ref_tree
@pytest.mark.parametrize("case", ["maps_untainted", "ranges_random", "plain_nmsa_range", "privacy_plain"])
def test_combinations_match_reference(built, tmp_path, case):
    """prepare_sim_inputs: same combinations, same punctured / shortened / removed positions (hash of the lists),
    same scaling factors, in the same order, as the reference builds from the same config + matrix directory."""
    ref = _ref()
    mdir = os.path.join(REF_ROOT, "sparse_matrices", "matrices_2")
    if case == "maps_untainted":
        cfg = os.path.join(REF_ROOT, "configs", "ADAPTIVE T.json")
    elif case == "ranges_random":
        cra = dict(enable_untainted_puncturing=False, use_adaptation_parameters_ranges=True,
                   code_rate_adaptation_parameters_ranges=[
                       dict(code_rate=0.805, delta=dict(begin=0.05, end=0.15, step=0.05), efficiency=dict(begin=1.1, end=1.5, step=0.2)),
                       dict(code_rate=0.655, delta=dict(begin=0.1, end=0.1, step=0.05), efficiency=dict(begin=1.2, end=1.3, step=0.1)),
                       dict(code_rate=0.505, delta=dict(begin=0.05, end=0.2, step=0.15), efficiency=dict(begin=1.15, end=1.15, step=0.1))],
                   code_rate_QBER_adaptation_parameters_maps=[])
        cfg = _write_cfg(tmp_path, "rr.json", enable_privacy_maintenance=False, code_rate_adaptation_parameters=cra)
    elif case == "plain_nmsa_range":
        cfg = _write_cfg(tmp_path, "pn.json", enable_code_rate_adaptation=False, enable_privacy_maintenance=False, decoding_algorithm=2,
                         min_sum_normalized_parameters=dict(use_alpha_range=True, alpha_range=dict(begin=0.6, end=0.9, step=0.1),
                                                            code_rate_alpha_maps=[]))
    else:
        cfg = _write_cfg(tmp_path, "pp.json", enable_code_rate_adaptation=False, enable_privacy_maintenance=True, decoding_algorithm=4,
                         adaptive_min_sum_normalized_parameters=dict(
                             use_alpha_range=False, alpha_range=dict(begin=0.1, end=1.0, step=0.1),
                             code_rate_alpha_maps=[dict(code_rate=0.95, alpha=0.88)], use_nu_range=True,
                             nu_range=dict(begin=0.5, end=0.7, step=0.1), code_rate_nu_maps=[]))
    ours = hostlib.describe_inputs(cfg, mdir, str(tmp_path))
    theirs = ref.describe_inputs(cfg, mdir)
    assert ours.count("\n") > 3
    if case == "maps_untainted":
        # privacy maintenance + rate adaptation: the reference reads shortened_bits[s] / punctured_bits[p] one past
        # the end (array_and_matrix_operations.cpp:215,220), so its bits_to_remove list depends on heap garbage; the
        # list is compared in test_privacy_positions_match_reference with that read made defined
        import re
        mask = lambda t: re.sub(r" rm=\d+:\d+", " rm=*", t)
        ours, theirs = mask(ours), mask(theirs)
    assert ours == theirs


@needs_ref_tree
def test_privacy_positions_match_reference(built):
    ref = _ref()
    path = _first("matrices_2/*R=0.8*.mtrx")
    untp = np.array(open(path.replace(".mtrx", ".untp")).read().split(), dtype=np.int32)
    ours, theirs = hostlib.HostMatrix(path, 3), ref.RefMatrix(path, 3)
    for q, d, e, unt in [(0.0076, 0.1, 1.85, True), (0.0196, 0.03, 1.28, True), (0.0276, 0.11, 1.2, False), (0.0316, 0.22, 1.22, False)]:
        p, s, r, _ = ours.adapt_code_rate(5555, q, d, e, untainted=unt, untp=untp, privacy_maintenance=True)
        assert p.size and s.size
        assert (r == ref.bits_to_remove_rate_adapt(theirs, p, s)).all()


@needs_ref_tree
def test_untainted_selection_matches_reference(built):
    ref = _ref()
    path = _first("matrices_alist_1k_all/*R=0.5*.mtrx")
    ours, theirs = hostlib.HostMatrix(path, 1), ref.RefMatrix(path, 1)
    for seed in (1, 5555):
        assert (ours.untainted(seed) == ref.untainted(theirs, seed)).all()


def test_shortest_float_format(built):
    # fmt's "{:L}" with a decimal comma (simulation.cpp:120-135): shortest round-trip digits
    cases = {1.0: "1", 0.0: "0", 0.5: "0,5", 0.99: "0,99", 0.98765: "0,98765", 1e-05: "1e-05", 0.0001: "0,0001", 0.00012: "0,00012",
             3e-06: "3e-06", 1 / 3: "0,3333333333333333", 0.1 + 0.2: "0,30000000000000004", 123456.0: "123456", 2.5e-07: "2,5e-07"}
    for v, s in cases.items():
        assert hostlib.format_shortest(v) == s, v


@needs_ref_tree
@pytest.mark.parametrize("alg", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("adapt", [False, True])
def test_csv_matches_reference_writer(built, tmp_path, alg, adapt):
    """Statistics (from the tally vector) and CSV text vs process_trials_results + write_file of the reference."""
    ref = _ref()
    rng = np.random.default_rng(100 + alg)
    cfg = _write_cfg(tmp_path, "c.json", decoding_algorithm=alg, enable_code_rate_adaptation=adapt, enable_throughput_measurement=False)
    for count, p_ok in ((1000, 0.97), (37, 0.5), (10, 0.0), (100000, 0.99997)):
        ok = rng.random(count) < p_ok
        iters = np.where(ok, rng.integers(1, 60, count), 100).astype(np.int32)
        flags = (ok.astype(np.uint8)) | ((rng.random(count) < 0.9).astype(np.uint8) << 1)
        a5 = [0.1, 1.25, 0.04321, 0.05679, 0.7654321] if adapt else None
        args = (cfg, iters, flags, "(N=10240,M=2048,R=0.8).mtrx", 10240, 2048, False, 0.0196, 200 / 10240, 0.7, 0.99, a5)
        ours = hostlib.csv_from_trials(*args)
        theirs = ref.csv_from_trials(*args, tmp_dir=str(tmp_path))
        assert ours == theirs, (count, ours, theirs)
