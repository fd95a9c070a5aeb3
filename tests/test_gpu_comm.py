"""GPU: the tally all-reduce behind the C ABI (csrc/comm.cu; SURVEY.md 8e / 8 b5: the handle owns the NCCL communicator).
On one GPU a one-rank communicator exercises the run-time binding of NCCL and both entry points; with two or more GPUs
(gpurun --gpus N) the in-process multi-device mode is checked against the host sum, and qkdldpc_sim --gpus 2 against the
reference executable."""
import threading

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def q(built):
    import qkd_ldpc_v_b200 as q
    return q


def _code(q, device=0):
    a = util.code_arrays("K1_5")
    return q.LdpcCode(a["n"], a["m"], a["row_ptr"], a["col_idx"], device=device)


def test_single_rank_communicator(q):
    import torch
    from qkd_ldpc_v_b200 import _cabi, hostlib
    assert _cabi.lib().qkdldpc_comm_nccl_version() >= 21800
    with _code(q) as code:
        assert code.comm_size() == 0
        with pytest.raises(_cabi.QkdLdpcError):
            code.tally_allreduce(np.arange(5, dtype=np.uint64))          # no communicator yet
        code.comm_init_rank(q.LdpcCode.comm_unique_id(), 1, 0)
        assert code.comm_size() == 1
        t = np.arange(105, dtype=np.uint64) * np.uint64(1 << 40)
        assert (code.tally_allreduce(t.copy()) == t).all()
        # device vector, enqueued behind a decode on the handle's stream
        arr = util.code_arrays("K1_5")
        a, b, acc = hostlib.gen_keys(hostlib.trial_seeds(5, 300), arr["n"], 0.02)
        r = code.QKD_LDPC_batch(a, b, acc, (0.75, 0.0), q.DecoderConfig(decoding_algorithm=2))
        d = torch.from_numpy(r.tally.astype(np.int64)).cuda()
        torch.cuda.synchronize()
        code.tally_allreduce_device(d.data_ptr(), d.numel())
        torch.cuda.synchronize()
        assert (d.cpu().numpy().astype(np.uint64) == r.tally).all()


def test_two_device_allreduce_equals_host_sum(q):
    import torch
    from qkd_ldpc_v_b200 import hostlib
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    arr = util.code_arrays("K1_5")
    seeds = hostlib.trial_seeds(99, 1001)
    cfg = q.DecoderConfig(decoding_algorithm=2, max_iterations=40)
    codes = [_code(q, d) for d in range(2)]
    q.LdpcCode.comm_init_all(codes)
    assert codes[0].comm_size() == codes[1].comm_size() == 2
    parts, sums = [None, None], [None, None]

    def work(d):
        lo, hi = 1001 * d // 2, 1001 * (d + 1) // 2
        r = codes[d].run_trials(seeds[lo:hi], 0.021, (0.75, 0.0), cfg, want_bits=False)
        parts[d] = r.tally.copy()
        sums[d] = codes[d].tally_allreduce(r.tally.copy())

    th = [threading.Thread(target=work, args=(d,)) for d in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    whole = codes[0].run_trials(seeds, 0.021, (0.75, 0.0), cfg, want_bits=False).tally
    assert (sums[0] == parts[0] + parts[1]).all() and (sums[1] == sums[0]).all() and (sums[0] == whole).all()
    for c in codes:
        c.close()
