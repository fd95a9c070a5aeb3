"""CPU: the C-ABI library loads and exports every symbol include/qkdldpc.h declares (no compute without a GPU)."""
import os
import re

import pytest

import util  # noqa: F401
from qkd_ldpc_v_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "qkdldpc.h")).read()
    return sorted(set(re.findall(r"QKDLDPC_API[^;(]*?\b(qkdldpc_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_cabi.SYMBOLS)


def test_library_exports_every_symbol(built):
    L = _cabi.lib()
    for name in header_symbols():
        assert hasattr(L, name), name
    assert L.qkdldpc_version() == 200
    assert L.qkdldpc_tally_len(100) == 105
    assert isinstance(L.qkdldpc_last_error(), bytes)


def test_struct_sizes_match_header(tmp_path):
    """The ctypes mirrors against the header itself: a C program compiled with gcc prints sizeof / offsetof."""
    import ctypes as C
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "qkdldpc.h"\nint main(void) {\n'
                   'printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(qkdldpc_params), sizeof(qkdldpc_options), sizeof(qkdldpc_combination),\n'
                   'sizeof(qkdldpc_info), offsetof(qkdldpc_options, copy_chunks), offsetof(qkdldpc_info, last_precision),\n'
                   'offsetof(qkdldpc_params, message_precision)); return 0; }\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got == [C.sizeof(_cabi.Params), C.sizeof(_cabi.Options), C.sizeof(_cabi.Combination), C.sizeof(_cabi.Info),
                   _cabi.Options.copy_chunks.offset, _cabi.Info.last_precision.offset, _cabi.Params.message_precision.offset]


def test_no_cpu_fallback(built):
    """Without a device the compute entry points must fail loudly (status -2), never fall back."""
    L = _cabi.lib()
    if L.qkdldpc_device_count() > 0:
        pytest.skip("a GPU is present")
    import numpy as np
    import qkd_ldpc_v_b200 as q
    a = util.code_arrays("N6")
    with pytest.raises(_cabi.QkdLdpcError) as e:
        q.LdpcCode(a["n"], a["m"], a["row_ptr"], a["col_idx"])
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_invalid_graphs_are_rejected(built):
    import ctypes as C
    import numpy as np
    L = _cabi.lib()
    h = C.c_void_p()
    rp = np.array([0, 2, 4], np.int32)
    bad = np.array([1, 0, 0, 1], np.int32)          # row 0 not ascending (quirk Q1)
    rc = L.qkdldpc_code_create(C.byref(h), 2, 2, 4, rp.ctypes.data, bad.ctypes.data, 0, None)
    assert rc == -1 and b"ascending" in L.qkdldpc_last_error()
    oob = np.array([0, 5, 0, 1], np.int32)
    rc = L.qkdldpc_code_create(C.byref(h), 2, 2, 4, rp.ctypes.data, oob.ctypes.data, 0, None)
    assert rc == -1 and b"out of range" in L.qkdldpc_last_error()


def test_onchip_tables_build_for_every_golden_code(built):
    """qkdldpc_code_create builds (and self-checks: every edge once, every index in range) the index tables of the
    on-chip kernel BEFORE it touches the device, so the table builder -- including the conflict-aware grouping -- is
    exercised here without a GPU: on this box the call must get as far as "no CUDA device" (-2), never -4 (table
    self-check) or -1."""
    import ctypes as C
    L = _cabi.lib()
    if L.qkdldpc_device_count() > 0:
        pytest.skip("a GPU is present")
    for name in ("N6", "N7", "N100", "N10s1", "K1_3", "K1_4", "K1_5", "K1_hi", "A79", "A82", "I80", "I65", "I50"):
        a = util.code_arrays(name)
        h = C.c_void_p()
        rc = L.qkdldpc_code_create(C.byref(h), a["n"], a["m"], a["nnz"], a["row_ptr"].ctypes.data, a["col_idx"].ctypes.data, 0, None)
        assert rc == -2, (name, rc, L.qkdldpc_last_error())
