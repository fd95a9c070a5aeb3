"""CPU, where the compiled reference is available (oracle/_ref/libqkdref.so: the UNMODIFIED reference sources):
pins the C oracle to the real thing on fresh inputs -- all six decoders, rate adaptation (punctured / shortened
frames through QKD_LDPC_RATE_ADAPT), the loaders' adjacency, and run_trial's flags."""
import os

import numpy as np
import pytest

import util
from oracle import cpu, ref

pytestmark = pytest.mark.skipif(not (ref.available() or os.path.isdir(os.path.join(ref.REFERENCE_ROOT, "src"))),
                                reason="needs oracle/_ref (built from /root/reference)")

CASES = [("A82", 0, 0.0162, 0.0, 0.0), ("A82", 1, 0.0162, 0.0, 0.0), ("A79", 2, 0.020, 0.71, 0.0),
         ("A82", 3, 0.0154, 0.81, 0.0), ("A82", 4, 0.0161, 0.80, 0.71), ("A82", 5, 0.0161, 0.68, 1.25),
         ("I80", 2, 0.018, 0.70, 0.0), ("K1_hi", 0, 0.006, 0.0, 0.0), ("K1_3", 3, 0.06, 0.5, 0.0)]


def write_sparse2(arr, path):
    with open(path, "w") as f:
        f.write(f"{arr['n']} {arr['m']}\n")
        for j in range(arr["m"]):
            f.write(" ".join(map(str, arr["col_idx"][arr["row_ptr"][j]:arr["row_ptr"][j + 1]])) + "\n")
        for i in range(arr["n"]):
            f.write(" ".join(map(str, arr["row_idx"][arr["col_ptr"][i]:arr["col_ptr"][i + 1]])) + "\n")


@pytest.fixture(scope="module")
def refmat(tmp_path_factory):
    cache = {}

    def get(name):
        if name not in cache:
            arr = util.code_arrays(name)
            p = tmp_path_factory.mktemp("m") / f"{name}.mtrx"
            write_sparse2(arr, p)
            cache[name] = ref.RefMatrix(str(p), 3)
        return cache[name]
    return get


@pytest.mark.parametrize("name,alg,qber,pri,sec", CASES)
def test_decoders_bit_identical(refmat, name, alg, qber, pri, sec):
    m = refmat(name)
    oc = util.oracle_code(name)
    assert (oc.col_ptr == m.col_ptr).all() and (oc.row_idx == m.row_idx).all()
    frames = 24 if m.n > 5000 else 96
    seeds = ref.trial_seeds(424242 + alg, frames)
    ref.set_cfg(alg, 100, True, 100.0)
    it_rt, fl_rt, acc = m.run_trials(qber, seeds, pri, sec)
    for k, s in enumerate(seeds):
        a, b, aq = ref.gen_keys(s, m.n, qber)
        llr, syn = cpu.frame_setup(oc, a, b, aq)
        it_r, ok_r, z_r = m.decode(alg, llr, syn, 100, pri, sec)
        it_o, ok_o, z_o = cpu.decode(oc, alg, llr, syn, 100, pri, sec)
        assert (it_r, ok_r) == (it_o, ok_o) and (z_r == z_o).all()
        assert it_rt[k] == it_o and (fl_rt[k] & 1) == int(ok_o) and ((fl_rt[k] >> 1) & 1) == int((z_o == a).all())


def test_threshold_disabled_and_small_max_iter(refmat):
    m = refmat("K1_4")
    oc = util.oracle_code("K1_4")
    for alg, pri, sec in ((0, 0, 0), (2, 0.8, 0), (5, 0.7, 1.1)):
        for s in ref.trial_seeds(99, 12):
            a, b, aq = ref.gen_keys(s, m.n, 0.04)
            llr, syn = cpu.frame_setup(oc, a, b, aq)
            for thr_on, mi in ((False, 100), (True, 3), (True, 1)):
                r = m.decode(alg, llr, syn, mi, pri, sec, thr_on, 100.0)
                o = cpu.decode(oc, alg, llr, syn, mi, pri, sec, thr_on, 100.0)
                assert r[:2] == o[:2] and (r[2] == o[2]).all()
