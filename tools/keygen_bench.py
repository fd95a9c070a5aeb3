"""Times the reference-compatible device key generator (qkdldpc_generate_trial_inputs_device) on n = 10240."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402

import qkd_ldpc_v_b200 as q  # noqa: E402
from qkd_ldpc_v_b200 import hostlib  # noqa: E402

arr = util.code_arrays("A79")
words = (arr["n"] + 31) // 32
code = q.LdpcCode(arr["n"], arr["m"], arr["row_ptr"], arr["col_idx"], device=0)
F = 65536
seeds = hostlib.trial_seeds(1, F)
da = torch.zeros((F, words), dtype=torch.int32, device="cuda")
db = torch.zeros_like(da)
for rep in range(3):
    t0 = time.perf_counter()
    code.generate_trial_inputs_device(seeds, 0.02, da.data_ptr(), db.data_ptr())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print("device keygen: %.2f ms for %d frames of n=%d -> %.2f M frames/s" % (dt * 1e3, F, arr["n"], F / dt / 1e6))
