"""Shared helpers of the test-suite: golden codes / vectors and packing."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_codes = None


def codes():
    global _codes
    if _codes is None:
        _codes = np.load(os.path.join(GOLDEN, "codes.npz"))
    return _codes


def code_arrays(name):
    c = codes()
    n, m, nnz, reg, fmt = (int(x) for x in c[f"{name}.meta"])
    return dict(n=n, m=m, nnz=nnz, is_regular=bool(reg), fmt=fmt, row_ptr=c[f"{name}.row_ptr"],
                col_idx=c[f"{name}.col_idx"], col_ptr=c[f"{name}.col_ptr"], row_idx=c[f"{name}.row_idx"],
                untp=c[f"{name}.untp"] if f"{name}.untp" in c.files else None, file=str(c[f"{name}.file"]))


def oracle_code(name):
    from oracle import cpu
    a = code_arrays(name)
    return cpu.Code(a["n"], a["m"], a["row_ptr"], a["col_idx"], a["col_ptr"], a["row_idx"])


def decode_cases():
    return sorted(f[len("decode_"):-4] for f in os.listdir(GOLDEN) if f.startswith("decode_") and f.endswith(".npz"))


def load_case(case):
    return np.load(os.path.join(GOLDEN, f"decode_{case}.npz"))


def unpack(words_u8, n):
    """golden words are np.packbits(..., bitorder='little') over uint8."""
    return np.unpackbits(words_u8, axis=-1, bitorder="little")[..., :n]
