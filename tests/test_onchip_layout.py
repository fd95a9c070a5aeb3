"""CPU: the storage layout of the float32 on-chip min-sum kernel (qkd_ldpc_v_b200/csrc/onchip_layout.hpp) through the host-only
C-ABI entry qkdldpc_onchip_layout_model: every table entry is checked against the graph inside the call (the kernel does no
bounds tests), and the shared-memory bank model of the result is held to bounds -- a regression of the conflict-aware lane /
edge-order choice would show here without a GPU."""
import ctypes as C

import numpy as np
import pytest

import util
from qkd_ldpc_v_b200 import _cabi

CODES = ("N6", "N7", "N100", "N10s1", "K1_3", "K1_4", "K1_5", "K1_hi", "A79", "A82", "I80", "I65", "I50", "L100k")


def model(name):
    a = util.code_arrays(name)
    out = np.zeros(12, np.int64)
    rc = _cabi.lib().qkdldpc_onchip_layout_model(a["n"], a["m"], a["nnz"], np.ascontiguousarray(a["row_ptr"], np.int32).ctypes.data,
                                                 np.ascontiguousarray(a["col_idx"], np.int32).ctypes.data, out.ctypes.data)
    assert rc == 0, (name, rc, _cabi.lib().qkdldpc_last_error())
    return out


@pytest.mark.parametrize("name", CODES)
def test_layout_tables_are_sound(built, name):
    out = model(name)
    a = util.code_arrays(name)
    dc_max = int(np.diff(a["row_ptr"]).max())
    eligible = dc_max <= 64 and a["n"] <= 65000
    assert bool(out[0]) == eligible, (name, out)
    if eligible:
        assert out[1] >= out[2] > 0 and out[3] >= out[4] > 0
        assert bool(out[5]) == (a["m"] + int((np.diff(a["row_ptr"]) > 32).sum()) <= 2048)
    # the float32 kernel's 8-byte records: 27 edges per record, rows of 28..51 edges own two
    eligible8 = dc_max <= 51 and a["n"] <= 65000
    assert bool(out[6]) == eligible8, (name, out)
    if eligible8:
        assert out[7] >= out[8] > 0 and out[9] >= out[10] > 0
        assert out[10] * 2 == out[4] or not eligible   # half the wavefronts when conflict-free: 16 lanes per 128 bytes instead of 8
        assert bool(out[11]) == (a["m"] + int((np.diff(a["row_ptr"]) > 27).sum()) <= 2048)


@pytest.mark.parametrize("name,cn_bar,vn_bar", [("I80", 1.30, 1.75), ("A79", 1.25, 1.18), ("A82", 1.25, 1.18), ("I65", 1.35, 1.9), ("I50", 1.65, 1.9)])
def test_bank_model_bounds(built, name, cn_bar, vn_bar):
    """n = 10240 codes: check-phase gathers within 15-30 % of conflict-free (2.1x with natural order; the R = 0.5 code's short rows --
    10 to 12 edges, 160 check groups -- level less well: 1.55x), record gathers within 13 % (regular alist codes) / 65-90 % (irregular
    codes whose bits of degree 17 ... 66 meet a different octet of rows at every step)."""
    out = model(name)
    cn, vn = out[1] / out[2], out[3] / out[4]
    print(f"\n{name}: check-phase gathers {cn:.3f} x conflict-free, variable-phase record gathers {vn:.3f} x")
    assert cn <= cn_bar and vn <= vn_bar
    # 8-byte records: twice as many lanes per wavefront collide more often, but the absolute count must drop well below the
    # 16-byte format's (I80: 12 665 -> 7 292 wavefronts per iteration, A79: 5 765 -> 3 619)
    print(f"{name}: 8-byte records {out[9]} wavefronts ({out[9] / out[10]:.3f} x conflict-free) against {out[3]} with 16-byte records")
    assert out[6] and out[7] / out[8] <= cn_bar + 0.05 and out[9] <= 0.68 * out[3]
