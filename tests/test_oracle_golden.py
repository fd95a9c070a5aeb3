"""CPU: the C oracle (oracle/ldpc_oracle.c) against golden outputs of the compiled, unmodified reference."""
import numpy as np
import pytest

import util
from oracle import cpu
from qkd_ldpc_v_b200 import hostlib, unpack_bits


def test_n6_known_answer():
    """example/qkd_ldpc_example.cpp:29-33 (Johnson ex. 2.5): values traced from the reference."""
    code = util.oracle_code("N6")
    alice = np.array([0, 0, 1, 0, 1, 1])
    bob = np.array([1, 0, 1, 0, 1, 1])
    llr, syn = cpu.frame_setup(code, alice, bob, 0.2)
    assert np.allclose(np.abs(llr), np.log(4.0)) and (syn == 0).all()
    expect_iters = {0: 1, 1: 1, 2: 1, 3: 2, 4: 3, 5: 2}
    factors = {0: (0, 0), 1: (0, 0), 2: (0.8, 0), 3: (0.8, 0), 4: (0.8, 0.5), 5: (0.8, 0.5)}
    for alg, it_expected in expect_iters.items():
        it, ok, z, tr = cpu.decode(code, alg, llr, syn, 100, *factors[alg], trace=True)
        assert ok and it == it_expected and (z == alice).all(), alg
        if alg == 0:
            assert np.allclose(tr[0][:6], [0.1212, 1.386, -2.894, 1.386, -1.386, -1.386], atol=2e-3)
            assert np.allclose(np.abs(tr[0][6:8]), 0.7538, atol=1e-4)
        if alg == 2:
            assert np.allclose(tr[0][:6], [0.8318, 1.386, -3.604, 1.386, -1.386, -1.386], atol=2e-3)
            assert np.allclose(np.abs(tr[0][6:8]), 1.109, atol=1e-3)


@pytest.mark.parametrize("case", util.decode_cases())
def test_oracle_reproduces_reference_golden(case):
    g = util.load_case(case)
    name = str(g["code"])
    arr = util.code_arrays(name)
    code = util.oracle_code(name)
    seeds = g["seeds"]
    a, b, acc = hostlib.gen_keys(seeds, arr["n"], float(g["qber"]))
    assert np.allclose(acc, g["acc_qber"])
    ab, bb = unpack_bits(a, arr["n"]), unpack_bits(b, arr["n"])
    it, fl, bits = cpu.qkd_ldpc_batch(code, int(g["alg"]), ab, bb, acc, max_iter=int(g["max_iter"]),
                                      primary=float(g["primary"]), secondary=float(g["secondary"]))
    assert (it == g["iters"]).all()
    assert (fl == g["flags"]).all()
    assert (bits == util.unpack(g["words"], arr["n"])).all()   # bit-exact words, converged or not
