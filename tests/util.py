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


def write_alist(path, name):
    """Writes golden code `name` as an alist file (format 1 of the reference)."""
    a = code_arrays(name)
    n, m = a["n"], a["m"]
    cw = np.diff(a["col_ptr"])
    rw = np.diff(a["row_ptr"])
    with open(path, "w") as f:
        f.write(f"{n} {m}\n{cw.max()} {rw.max()}\n")
        f.write(" ".join(map(str, cw)) + "\n" + " ".join(map(str, rw)) + "\n")
        for i in range(n):
            f.write(" ".join(str(x + 1) for x in a["row_idx"][a["col_ptr"][i]:a["col_ptr"][i + 1]]) + "\n")
        for j in range(m):
            f.write(" ".join(str(x + 1) for x in a["col_idx"][a["row_ptr"][j]:a["row_ptr"][j + 1]]) + "\n")


def write_sparse2(path, name, with_untp=True):
    """Writes golden code `name` in the "sparse_2" format (format 3) plus its .untp side-car when the code has one."""
    a = code_arrays(name)
    n, m = a["n"], a["m"]
    with open(path, "w") as f:
        f.write(f"{n} {m}\n")
        for j in range(m):
            f.write(" ".join(map(str, a["col_idx"][a["row_ptr"][j]:a["row_ptr"][j + 1]])) + "\n")
        for i in range(n):
            f.write(" ".join(map(str, a["row_idx"][a["col_ptr"][i]:a["col_ptr"][i + 1]])) + "\n")
    if with_untp and a["untp"] is not None and a["untp"].size:
        with open(os.path.splitext(path)[0] + ".untp", "w") as f:
            f.write(" ".join(map(str, a["untp"])) + "\n")
