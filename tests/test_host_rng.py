"""CPU: the C++ host's trial input generation against keys produced by the compiled reference (golden/rng.npz)."""
import numpy as np
import pytest

import util  # noqa: F401
from qkd_ldpc_v_b200 import hostlib

G = np.load(util.GOLDEN + "/rng.npz")
KEYS = sorted({k.split(".")[0] for k in G.files})


@pytest.mark.parametrize("key", KEYS)
def test_keys_match_reference(key):
    n = int(key.split("_n")[1])
    sim = int(key.split("_")[0][1:])
    seeds = hostlib.trial_seeds(sim, len(G[key + ".seeds"]))
    assert (seeds == G[key + ".seeds"]).all()
    a, b, acc = hostlib.gen_keys(seeds, n, float(G[key + ".qber"][0]))
    ga, gb = G[key + ".alice"], G[key + ".bob"]
    assert (a.view(np.uint8)[:, : ga.shape[1]] == ga).all()
    assert (b.view(np.uint8)[:, : gb.shape[1]] == gb).all()
    assert acc == G[key + ".acc"][0]
    # exactly floor(n*q) flips (inject_errors)
    flips = np.unpackbits((a ^ b).view(np.uint8), axis=1).sum(axis=1)
    assert (flips == int(n * float(G[key + ".qber"][0]))).all()


def _py_xoshiro_first_outputs(seed, count):
    """Independent pure-Python xoshiro256++ seeded by SplitMix64 (the published algorithms)."""
    M = (1 << 64) - 1
    rotl = lambda x, k: ((x << k) | (x >> (64 - k))) & M  # noqa: E731
    x, st = seed, []
    for _ in range(4):
        x = (x + 0x9E3779B97F4A7C15) & M
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        st.append(z ^ (z >> 31))
    assert seed != 1 or st[0] == 0x910A2DEC89025CC1      # SplitMix64(1): widely published first output
    out = []
    for _ in range(count):
        out.append((rotl((st[0] + st[3]) & M, 23) + st[0]) & M)
        t = (st[1] << 17) & M
        st[2] ^= st[0]; st[3] ^= st[1]; st[1] ^= st[2]; st[0] ^= st[3]; st[2] ^= t; st[3] = rotl(st[3], 45)  # noqa: E702
    return out


@pytest.mark.parametrize("seed", [0, 1, 777, 10012025, 2**63 + 12345])
def test_xoshiro_engine_against_independent_implementation(seed):
    """uniform_int_distribution<size_t>(0, SIZE_MAX) returns the raw 64-bit draw, so trial_seeds() exposes the
    engine's output stream. (The Xoshiro-cpp sources are not under /root/reference; see DESIGN.md 'RNG pin'.)"""
    got = [int(v) for v in hostlib.trial_seeds(seed, 8)]
    assert got == _py_xoshiro_first_outputs(seed, 8)
