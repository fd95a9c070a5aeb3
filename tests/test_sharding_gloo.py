"""CPU, world_size 2 over gloo: the N>1 path's host logic -- shard ranges, tally all-reduce, statistics from the
reduced tallies equal to the single-process statistics (exact: integer sums). The per-rank decode is done by the
CPU oracle here (test infrastructure); on the GPU box bench.py runs the same plumbing over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cpu
    from qkd_ldpc_v_b200 import hostlib, sharding, unpack_bits
    arr = util.code_arrays("K1_5")
    oc = util.oracle_code("K1_5")
    total = 301                                   # not divisible by the world size on purpose
    seeds = hostlib.trial_seeds(4711, total)
    lo, hi = sharding.shard_range(total, rank, world)
    a, b, acc = hostlib.gen_keys(seeds[lo:hi], arr["n"], 0.02)
    it, fl, _ = cpu.qkd_ldpc_batch(oc, 2, unpack_bits(a, arr["n"]), unpack_bits(b, arr["n"]), acc, max_iter=30,
                                   primary=0.75, threads=2, want_bits=False)
    mine = sharding.tally_from_results(it, fl, 30)
    reduced = sharding.allreduce_tally(mine)
    dist.barrier()
    q.put((rank, lo, hi, reduced.tolist()))
    dist.destroy_process_group()


def test_two_rank_tally_allreduce_equals_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    assert got[0][1] == 0 and got[0][2] == got[1][1] and got[1][2] == 301       # contiguous cover
    assert got[0][3] == got[1][3]                                               # both ranks hold the same sum

    from oracle import cpu
    from qkd_ldpc_v_b200 import hostlib, sharding, stats_from_tally, unpack_bits
    arr = util.code_arrays("K1_5")
    oc = util.oracle_code("K1_5")
    seeds = hostlib.trial_seeds(4711, 301)
    a, b, acc = hostlib.gen_keys(seeds, arr["n"], 0.02)
    it, fl, _ = cpu.qkd_ldpc_batch(oc, 2, unpack_bits(a, arr["n"]), unpack_bits(b, arr["n"]), acc, max_iter=30,
                                   primary=0.75, want_bits=False)
    single = sharding.tally_from_results(it, fl, 30)
    assert single.tolist() == got[0][3]
    # statistics exactly as process_trials_results computes them (simulation.cpp:594-624)
    st = stats_from_tally(np.array(got[0][3], np.uint64), 301)
    ok = (fl & 1) != 0
    assert st["iter_success_min"] == it[ok].min() and st["iter_success_max"] == it[ok].max()
    assert abs(st["iter_success_mean"] - it[ok].mean()) < 1e-12
    assert abs(st["iter_success_std"] - it[ok].std()) < 1e-9
    assert st["ratio_trials_success_ldpc"] == ((fl & 3) == 3).sum() / 301


@pytest.mark.parametrize("n,world", [(10, 3), (1, 4), (0, 2), (1000003, 8)])
def test_shard_ranges_cover_exactly(n, world):
    from qkd_ldpc_v_b200 import sharding
    r = [sharding.shard_range(n, k, world) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == n
    assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
    assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
