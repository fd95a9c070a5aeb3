"""Dump a code of tests/golden/codes.npz as [n, m, nnz, row_ptr, col_idx] int32 for oc_layout_model: dump_code.py I80 /tmp/I80.bin"""
import sys
import numpy as np

d = np.load(__file__.rsplit("/tools/", 1)[0] + "/tests/golden/codes.npz")
name, out = sys.argv[1], sys.argv[2]
rp, ci = d[name + ".row_ptr"].astype(np.int32), d[name + ".col_idx"].astype(np.int32)
n = len(d[name + ".col_ptr"]) - 1
with open(out, "wb") as f:
    np.array([n, len(rp) - 1, len(ci)], np.int32).tofile(f)
    rp.tofile(f)
    ci.tofile(f)
