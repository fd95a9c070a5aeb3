#!/usr/bin/env python
"""Headline benchmark: decoded key Gbit/s of the batched LDPC syndrome decoder (BASELINE.json metric).

Headline workload (config.workload): `config 10k NMSA.json`'s operating family at the north-star point -- the irregular
n=10240, m=2048 (R=0.8, E=60430) code `matrices_2/(N=10240,M=2048,R=0.8).mtrx`, normalized min-sum alpha=0.7,
QBER 3 %, 100 iterations max, message clamp 100, float32 messages, synthetic keys with run_trial's distribution.
One "step" = one pass of the hot path over one batch of --frames frames per GPU.

The same JSON line carries, under "workloads", one entry per further BASELINE.json workload (value, e2e, roofline, mean
iterations, FER): the converging n=10k NMSA points, SPA / SPA-lin-approx at their FER=0.01 point, AOMSA under the
default precision policy (float64 state), and the n=102400 code with NMSA and SPA.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload ID] [--scaling weak|strong]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (golden code, algorithm, primary, secondary, qber, description)
    "I80_nmsa_q030": ("I80", 2, 0.70, 0.0, 0.030, "n=10240 m=2048 irregular R=0.8 (E=60430), NMSA alpha=0.70, QBER 3%"),
    "I80_nmsa_q015": ("I80", 2, 0.70, 0.0, 0.015, "n=10240 m=2048 irregular R=0.8 (E=60430), NMSA alpha=0.70, QBER 1.5%"),
    "A79_nmsa_q020": ("A79", 2, 0.71, 0.0, 0.020, "n=10240 m=2201 alist R=0.79 (E=40960), NMSA alpha=0.71, QBER 2%"),
    "A79_nmsa_q030": ("A79", 2, 0.71, 0.0, 0.030, "n=10240 m=2201 alist R=0.79 (E=40960), NMSA alpha=0.71, QBER 3%"),
    "A82_spa_q0162": ("A82", 0, 0.0, 0.0, 0.0162, "n=10240 m=1801 alist R=0.82 (E=40960), SPA, QBER 1.62%"),
    "A82_spalin_q0162": ("A82", 1, 0.0, 0.0, 0.0162, "n=10240 m=1801 alist R=0.82, SPA-lin-approx, QBER 1.62%"),
    "L100k_nmsa_q060": ("L100k", 2, 0.72, 0.0, 0.06, "n=102400 m=52301 alist R=0.49 (E=307200), NMSA alpha=0.72, QBER 6%"),
    "L100k_spa_q084": ("L100k", 0, 0.0, 0.0, 0.084, "n=102400 m=52301 alist R=0.49 (E=307200), SPA, QBER 8.4% (config 100k.json as shipped)"),
    "A82_omsa_q0154": ("A82", 3, 0.81, 0.0, 0.0154, "n=10240 m=1801 alist R=0.82, OMSA beta=0.81, QBER 1.54%"),
    "A82_anmsa_q0161": ("A82", 4, 0.80, 0.71, 0.0161, "n=10240 m=1801 alist R=0.82, ANMSA alpha=0.8 nu=0.71, QBER 1.61%"),
    "A82_aomsa_q0161": ("A82", 5, 0.68, 1.25, 0.0161, "n=10240 m=1801 alist R=0.82, AOMSA beta=0.68 sigma=1.25, QBER 1.61%"),
    "I80_spa_q015": ("I80", 0, 0.0, 0.0, 0.015, "n=10240 m=2048 irregular R=0.8 (E=60430), SPA, QBER 1.5% (4 E + 4 n bytes do not fit an SM: streaming path)"),
    "I80_aomsa_q015": ("I80", 5, 0.70, 0.99, 0.015, "n=10240 m=2048 irregular R=0.8 (E=60430), AOMSA beta=0.70 sigma=0.99, QBER 1.5% (ADAPTIVE R.json family)"),
}
HEADLINE = "I80_nmsa_q030"
# further workloads of the default run: (id, frames per GPU per step) -- sized so that each takes a few seconds
SECONDARY = [("A79_nmsa_q020", 65536), ("I80_nmsa_q015", 65536), ("A82_spa_q0162", 32768), ("A82_spalin_q0162", 32768),
             ("A82_aomsa_q0161", 32768), ("L100k_nmsa_q060", 4096), ("L100k_spa_q084", 1024), ("I80_spa_q015", 32768)]
MAX_ITER, THRESHOLD = 100, 100.0


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    p.add_argument("--frames", type=int, default=0, help="frames per GPU per step (0 = 65536; n=102400 codes 4096)")
    p.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                   help="weak: --frames per GPU; strong: --total-frames split over the GPUs")
    p.add_argument("--total-frames", type=int, default=262144, help="strong scaling: frames per step over all GPUs")
    p.add_argument("--precision", type=int, default=0, choices=[0, 32, 64],
                   help="message precision: 0 = the library's policy (float64 state for OMSA / ANMSA / AOMSA and SPA on n > 65536)")
    p.add_argument("--pool-slots", type=int, default=0)
    p.add_argument("--frames-per-lane", type=int, default=0)
    p.add_argument("--path", type=int, default=0, choices=[0, 1, 2],
                   help="decoder path: 0 auto (on-chip kernels when eligible), 1 streaming (messages in HBM), 2 on-chip")
    p.add_argument("--onchip-threads", type=int, default=0)
    p.add_argument("--vn-items", type=int, default=0,
                   help="streaming path, narrow variable-node buckets: items per warp (0 = auto, 1 = one item per warp)")
    p.add_argument("--vn-ctas", type=int, default=0, choices=[0, 1, 3, 4, 5, 6],
                   help="resident CTAs per SM of the dv <= 4 float32 variable-node kernel (0 = auto)")
    p.add_argument("--record-bytes", type=int, default=0, choices=[0, 8, 16],
                   help="float32 on-chip min-sum: record format (0 auto: 8-byte records when every row has at most 51 edges)")
    p.add_argument("--copy-chunks", type=int, default=0, help="pieces a host batch is cut into for copy / compute overlap (0 auto, 1 none)")
    p.add_argument("--no-compaction", action="store_true", help="streaming path: do not compact the tail of a draining batch")
    p.add_argument("--cpu-frames", type=int, default=0, help="frames of the CPU-baseline sample (0 = auto)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-secondary", action="store_true", help="only the selected workload, no `workloads` array")
    return p.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t_begin = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        """Start of the timed region. nvidia-smi needs about a second to deliver its first sample, so the process is
        started before the warm-up; only samples taken after this mark are used."""
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = getattr(self, "t_end", time.time())
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        window = "timed region of the headline workload"
        rows = [r for t, r in self.rows if (self.t_begin is None or t >= self.t_begin) and t <= t_end + 0.05]
        if not rows and self.rows:
            # a timed region shorter than the sampling period: the GPU ran the same steps during the warm-up just
            # before it, so the last samples before the end of the region stand in (and the window says so)
            rows = [r for t, r in self.rows if t <= t_end][-3:]
            window = "warm-up + timed region (timed region shorter than the 100 ms sampling period)"
        self.rows = rows
        self.window = window
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(len(r) >= 8 and r[4 + k].lower() == "active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": self.window, "reasons": reasons}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_db():
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ---- the reference's CPU decoder (oracle/_ref = its unmodified sources; oracle port only if that did not travel) --------

def reference_keys(n, n_frames, qber, seed):
    """Alice / Bob frames ([F][n] int32) with run_trial's distribution. With oracle/_ref: the REFERENCE's own generator
    (fill_random_bits + inject_errors through its RNG), nothing of this repo's host library."""
    from oracle import ref
    if ref.available():
        seeds = ref.trial_seeds(seed, n_frames)
        a = np.zeros((n_frames, n), np.int32)
        b = np.zeros((n_frames, n), np.int32)
        acc = 0.0
        for f in range(n_frames):
            a[f], b[f], acc = ref.gen_keys(int(seeds[f]), n, qber)
        return a, b, acc
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2, (n_frames, n), dtype=np.int32)
    b = a.copy()
    n_err = int(n * qber)
    for f in range(n_frames):
        b[f, rng.choice(n, n_err, replace=False)] ^= 1
    return a, b, n_err / n


def cpu_reference_run(wl, n_frames, threads, seed=20261018):
    """`n_frames` frames of the workload through the reference's own QKD_LDPC on `threads` host threads (static block
    partition like BS::thread_pool::detach_loop). Returns (seconds, kind, iterations, flags)."""
    import util
    name, alg, pri, sec, qber, _ = WORKLOADS[wl]
    arr = util.code_arrays(name)
    ai, bi, acc = reference_keys(arr["n"], n_frames, qber, seed)
    from oracle import ref
    if ref.available():
        import tempfile
        # the compiled reference reads matrices from text files: re-emit the golden CSR in its "sparse_2" format
        path = os.path.join(tempfile.mkdtemp(), "code.mtrx")
        with open(path, "w") as f:
            f.write(f"{arr['n']} {arr['m']}\n")
            for j in range(arr["m"]):
                f.write(" ".join(map(str, arr["col_idx"][arr["row_ptr"][j]:arr["row_ptr"][j + 1]])) + "\n")
            for i in range(arr["n"]):
                f.write(" ".join(map(str, arr["row_idx"][arr["col_ptr"][i]:arr["col_ptr"][i + 1]])) + "\n")
        m = ref.RefMatrix(path, 3)
        ref.set_cfg(alg, MAX_ITER, True, THRESHOLD)
        t0 = time.perf_counter()
        it, fl = m.qkd_ldpc_batch(ai, bi, acc, pri, sec, threads=threads)
        dt = time.perf_counter() - t0
        return dt, "reference", it, fl
    from oracle import cpu
    oc = util.oracle_code(name)
    t0 = time.perf_counter()
    it, fl, _ = cpu.qkd_ldpc_batch(oc, alg, ai.astype(np.uint8), bi.astype(np.uint8), acc, max_iter=MAX_ITER, primary=pri, secondary=sec,
                                   thr=THRESHOLD, threads=threads, want_bits=False)
    dt = time.perf_counter() - t0
    return dt, "port", it, fl


def auto_cpu_frames(wl, cores):
    name, alg, *_ = WORKLOADS[wl]
    import util
    arr = util.code_arrays(name)
    # ~21 ns/edge/iter (min-sum) or ~68 (SPA) per core for the reference; aim at ~15 s of CPU work
    per_frame = arr["nnz"] * MAX_ITER * (68e-9 if alg == 0 else 30e-9)
    return int(max(cores, min(4096, cores * max(1, round(15.0 / per_frame)))))


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation on this box's host cores (rank 0 only). Nothing of this
    repo's libraries is loaded on this arm: keys come from the reference's generator, the decoder is oracle/_ref."""
    if rank != 0:
        return
    import util
    name, alg, pri, sec, qber, desc = WORKLOADS[args.workload]
    arr = util.code_arrays(name)
    cores = os.cpu_count() or 1
    frames = args.cpu_frames or max(cores, auto_cpu_frames(args.workload, cores) // 4)
    for _ in range(args.warmup):
        cpu_reference_run(args.workload, max(cores, frames // 8), cores)
    t, kind, it = 0.0, "port", None
    for _ in range(args.steps):
        dt, kind, it, fl = cpu_reference_run(args.workload, frames, cores)
        t += dt
    gbit = arr["n"] * frames * args.steps / t / 1e9
    out = {
        "impl": "reference", "metric": "decoded key Gbit/s", "value": gbit, "unit": "Gbit/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "workload_id": args.workload, "frames_per_step": frames, "max_iterations": MAX_ITER,
                   "threshold": THRESHOLD, "mean_iterations": float(np.mean(it))},
        "cpu_baseline": {"value": gbit, "unit": "Gbit/s", "cores": cores, "kind": kind,
                         "sample": f"{frames} frames per step x {args.steps} steps, decode only, keys from the reference's generator"},
        "e2e": {"value": gbit, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ---- one workload on the GPU(s) ---------------------------------------------------------------------------------------------

class Ctx:
    pass


def measure(ctx, args, wl, F, steps, warmup, sampler=None, want_e2e=True):
    """Decodes `steps` batches of F frames per GPU of workload `wl`. Returns the fields of its bench entry."""
    import torch
    import torch.distributed as dist

    import qkd_ldpc_v_b200 as q
    import util
    dev, stream, rank, world, local_rank = ctx.dev, ctx.stream, ctx.rank, ctx.world, ctx.local_rank
    name, alg, pri, sec, qber, desc = WORKLOADS[wl]
    arr = util.code_arrays(name)
    n, m, nnz = arr["n"], arr["m"], arr["nnz"]
    words = (n + 31) // 32
    code = q.LdpcCode(n, m, arr["row_ptr"], arr["col_idx"], device=local_rank, pool_slots=args.pool_slots,
                      frames_per_lane_f32=args.frames_per_lane, decoder_path=args.path, onchip_threads=args.onchip_threads,
                      tail_compaction=-1 if args.no_compaction else 0, copy_chunks=args.copy_chunks,
                      onchip_record_bytes=args.record_bytes, vn_items_per_warp=args.vn_items, vn_ctas_per_sm=args.vn_ctas)
    code.set_stream(stream.cuda_stream)
    if world > 1:
        # the handle owns the NCCL communicator of the tally all-reduce (include/qkdldpc.h): rank 0 draws the unique id,
        # torch.distributed only carries its 128 bytes to the other ranks
        ident = torch.from_numpy(q.LdpcCode.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(dev)
        dist.broadcast(ident, 0)
        code.comm_init_rank(ident.cpu().numpy(), world, rank)
    cfg = q.DecoderConfig(decoding_algorithm=alg, max_iterations=MAX_ITER, enable_msg_llr_threshold=True,
                          msg_llr_threshold=THRESHOLD, message_precision=args.precision)

    # inputs resident in HBM: synthetic keys generated on the device, a different seed per rank (frames shard)
    d_alice = torch.empty((F, words), dtype=torch.int32, device=dev)
    d_bob = torch.empty((F, words), dtype=torch.int32, device=dev)
    acc_q = code.generate_keys_device(F, qber, 1234567 + 7919 * rank, d_alice.data_ptr(), d_bob.data_ptr())
    d_qber = torch.tensor([acc_q], dtype=torch.float64, device=dev)
    d_iters = torch.empty(F, dtype=torch.int32, device=dev)
    d_flags = torch.empty(F, dtype=torch.uint8, device=dev)
    d_bits = torch.empty((F, words), dtype=torch.int32, device=dev)
    tl = q.tally_len(MAX_ITER)
    d_tally = torch.zeros(tl, dtype=torch.int64, device=dev)

    def step():
        code.decode_batch_device(d_alice.data_ptr(), d_bob.data_ptr(), d_qber.data_ptr(), F, (pri, sec), cfg,
                                 d_out_bits=d_bits.data_ptr(), d_out_iters=d_iters.data_ptr(),
                                 d_out_flags=d_flags.data_ptr(), d_tally=d_tally.data_ptr())
        if world > 1:   # the only collective of the path: FER / iteration tallies (K5), on the decoder's stream
            code.tally_allreduce_device(d_tally.data_ptr(), tl)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    # ---- timed region: device-resident inputs, CUDA events on the launching stream; per-kernel profiling OFF (the streaming
    # ---- path replays its captured step graph, as it does for any caller) ----------------------------------------------------
    l0 = code.info()["kernel_launches"]
    barrier()
    if sampler:
        sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    batch_ms = 0.0
    for _ in range(steps):
        step()
        batch_ms += code.info()["last_batch_ms"]
    e1.record(stream)
    barrier()
    if sampler:
        sampler.mark_end()
    inf = code.info()
    launches = inf["kernel_launches"] - l0
    elapsed_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    tally = d_tally.cpu().numpy().astype(np.uint64)
    frames_total = F * world
    iters_executed = int(tally[3])           # of the LAST step, all ranks
    value = n * frames_total * steps / (elapsed_ms * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (algorithmic bytes: SURVEY.md 8d) ---------------------------
    prec = inf["last_precision"]
    sz = 4 if prec == 32 else 8
    it_rank = iters_executed / world         # frame-iterations one rank executed in one step
    cn_bytes, vn_bytes = 2 * sz * nnz * it_rank, (2 * sz * nnz + sz * n) * it_rank   # per step, all launches
    peak, peak_src = measured_peak()
    onchip = inf["last_path"] == 2
    db = ncu_db()
    if onchip:
        # ONE persistent kernel per step; the belief-propagation state never leaves shared memory, so the algorithmic
        # bytes (what a message-streaming decoder has to move) are not HBM traffic here: frac > 1 is the design's point,
        # and the kernel's own ceiling is on the SM -- `onchip` carries the shared-memory / issue figures of its ncu capture.
        k_ms = batch_ms / steps
        k_bytes = cn_bytes + vn_bytes
        ach_k = k_bytes / (k_ms * 1e-3) / 1e9
        kname = ("onchip_minsum_kernel" if prec == 32 else "onchip_minsum64_kernel") if alg >= 2 else "onchip_spa_kernel"
        ent = db.get(kname, {})
        tr = ent.get("dram_bytes_per_frame")
        # phase split of the min-sum kernels: a SECOND, untimed pass in which every CTA clocks its check / variable phases
        phases = None
        if kname in ("onchip_minsum_kernel", "onchip_minsum64_kernel"):
            code.set_profiling(True)
            step()
            pi = code.info()
            code.set_profiling(False)
            if pi["last_cn_ms"] > 0:
                phases = {"check_ms": pi["last_cn_ms"], "variable_ms": pi["last_vn_ms"], "batch_ms": pi["last_batch_ms"],
                          "how": "share of the CTAs' SM clocks spent in each phase x the batch time (second, untimed pass)"}
        roofline = {
            "bound": "hbm", "kernel": f"{kname}<ALG={alg}>", "achieved": ach_k, "peak": peak, "unit": "GB/s",
            "frac": ach_k / peak, "traffic": (tr * F if tr is not None else None), "peak_source": peak_src,
            "bytes_per_launch": k_bytes, "ms_per_launch": k_ms, "launches_per_step": 1,
            "whole_step_frac": (k_bytes / (elapsed_ms / steps * 1e-3) / 1e9) / peak,
            "onchip": ent.get("onchip"), "phases": phases,
            "note": "algorithmic bytes = (16*E + 4*N) per frame-iteration in float32, twice that in float64 (SURVEY.md 8d); this "
                    "kernel keeps them on chip "
                    + (("(4N + 16M bytes of shared memory per frame)" if prec == 32 else "(8N + 24M bytes of shared memory per frame)")
                       if alg >= 2 else "(4N + 4E bytes of shared memory per frame)")
                    + ", DRAM traffic is the packed keys only -- see `traffic`; its limit is on-SM -- see `onchip`",
        }
    else:
        # per-kernel times: a SECOND, untimed pass with every launch bracketed by CUDA events (this switches the step graph off)
        code.set_profiling(True)
        step()
        pi = code.info()
        code.set_profiling(False)
        prof_tally = d_tally.cpu().numpy().astype(np.uint64)
        it_prof = int(prof_tally[3]) / world
        cnb, vnb = 2 * sz * nnz * it_prof, (2 * sz * nnz + sz * n) * it_prof
        kern = {"cn": (cnb, pi["last_cn_ms"]), "vn": (vnb, pi["last_vn_ms"])}
        dom = max(kern, key=lambda k: kern[k][1])
        ach = {k: (v[0] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else 0.0) for k, v in kern.items()}
        launches_per_step = launches / steps
        n_cn_launches = max(1.0, (launches_per_step - 2) / 3)
        per_code = db.get(name, db) if sz == 4 and alg >= 2 else {}     # captures exist for the float32 min-sum kernels
        tr = per_code.get({"cn": "cn_kernel", "vn": "vn_kernel"}[dom], {}).get("dram_bytes_per_frame_iter")
        vn_name = "vn_kernel_ell_loop" if (pi.get("last_vn_items_per_warp") or 0) > 1 else "vn_kernel_ell"
        roofline = {
            "bound": "hbm", "kernel": {"cn": f"cn_kernel<{'float' if sz == 4 else 'double'},ALG={alg}>", "vn": vn_name}[dom],
            "achieved": ach[dom], "peak": peak, "unit": "GB/s", "frac": ach[dom] / peak,
            "traffic": (tr * it_rank / n_cn_launches if tr is not None and name in ("I80", "L100k") else None),
            "peak_source": peak_src,
            "bytes_per_launch": kern[dom][0] / n_cn_launches, "ms_per_launch": kern[dom][1] / n_cn_launches,
            "launches_per_step": n_cn_launches,
            "both_kernels": {k: {"achieved_gbs": ach[k], "frac": ach[k] / peak, "ms_per_step": kern[k][1]} for k in kern},
            "sched_ms_per_step": pi["last_sched_ms"],
            "per_kernel_times": "second untimed pass with per-launch CUDA events (step graph off); `value` is timed with the graph on",
            "whole_step_frac": ((cn_bytes + vn_bytes) / (elapsed_ms / steps * 1e-3) / 1e9) / peak,
        }

    # ---- e2e: the same metric through the host-buffer C-ABI call, pinned host memory, copies inside ---------------
    e2e = None
    if want_e2e and not args.no_e2e:
        h_alice = torch.empty((F, words), dtype=torch.int32).pin_memory()
        h_bob = torch.empty((F, words), dtype=torch.int32).pin_memory()
        h_alice.copy_(d_alice); h_bob.copy_(d_bob)  # noqa: E702
        torch.cuda.synchronize()
        a_np = h_alice.numpy().view(np.uint32)
        b_np = h_bob.numpy().view(np.uint32)
        # results land in pinned host memory too (a caller that decodes batch after batch reuses its buffers)
        h_bits = torch.empty((F, words), dtype=torch.int32).pin_memory()
        h_iters = torch.empty(F, dtype=torch.int32).pin_memory()
        h_flags = torch.empty(F, dtype=torch.uint8).pin_memory()
        outbuf = (h_bits.numpy().view(np.uint32), h_iters.numpy(), h_flags.numpy())
        code.QKD_LDPC_batch(a_np, b_np, acc_q, (pri, sec), cfg, out=outbuf)          # warm the staging buffers
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            r = code.QKD_LDPC_batch(a_np, b_np, acc_q, (pri, sec), cfg, out=outbuf)
            if world > 1:
                code.tally_allreduce(r.tally)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert (h_iters.numpy() == d_iters.cpu().numpy()).all(), "host-buffer call and device-buffer call disagree"
        e2e = {"value": n * frames_total * steps / dt / 1e9, "unit": "Gbit/s",
               "h2d_bytes_per_step": int(2 * F * words * 4 + 8), "d2h_bytes_per_step": int(F * words * 4 + F * 5 + tl * 8),
               "api": "qkdldpc_decode_batch (host buffers, pinned); copies overlap the decoding piece by piece on the on-chip paths"}
    stats = q.stats_from_tally(tally, frames_total)
    code.close()
    del d_alice, d_bob, d_bits
    torch.cuda.empty_cache()
    return {
        "workload_id": wl, "workload": desc, "n": n, "value": value, "unit": "Gbit/s", "ms_per_step": elapsed_ms / steps, "steps": steps,
        "warmup": warmup, "dtype": "f32" if prec == 32 else "f64", "frames_per_step_per_gpu": F,
        "accurate_qber": acc_q, "mean_iterations_executed": iters_executed / frames_total, "fer": stats["FER"],
        "decoder_path": (("on-chip min-sum" if alg >= 2 else "on-chip sum-product") + " (frame state in shared memory)")
                        if onchip else "streaming (messages in HBM)",
        "onchip_threads": inf.get("onchip_threads") if onchip else None,
        "onchip_record_bytes": (inf.get("onchip_record_bytes") or None) if onchip and alg >= 2 and prec == 32 else None,
        "frames_per_tile": inf["frames_per_tile"], "pool_tiles": inf["pool_tiles"], "pool_bytes": inf["pool_bytes"],
        "streaming": None if onchip else {"steps_per_poll": inf.get("last_steps_per_poll"),
                                          "vn_items_per_warp": inf.get("last_vn_items_per_warp"),
                                          "tail_compactions": inf.get("tail_compactions")},
        "l2_policy": ("inputs larger than L2: %.0f MB of packed keys in + decisions out per step, read once; decoder "
                      "state is in shared memory" % (3 * F * words * 4 / 1e6)) if onchip else
                     ("inputs larger than L2 (message pool %.1f GB >> 126 MB)" % (inf["pool_bytes"] / 1e9)),
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "tally_len": int(tl),
    }


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # Libraries (NCCL's version banner, ...) must not write to stdout: the contract is ONE JSON line from rank 0.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    ctx = Ctx()
    ctx.dev = torch.device("cuda", local_rank)
    ctx.rank, ctx.world, ctx.local_rank = rank, world, local_rank
    if world > 1:
        dist.init_process_group("nccl", device_id=ctx.dev)
    ctx.stream = torch.cuda.Stream(device=ctx.dev)      # a real stream: the decoder replays CUDA graphs on it
    torch.cuda.set_stream(ctx.stream)

    def frames_for(wl, default):
        if args.scaling == "strong":
            tot = args.total_frames if not WORKLOADS[wl][0].startswith("L100k") else max(world, args.total_frames // 16)
            return max(1, tot // world)
        return default

    big = WORKLOADS[args.workload][0].startswith("L100k")
    F = frames_for(args.workload, args.frames or (4096 if big else 65536))
    sampler = ClockSampler(local_rank)
    sampler.start()
    head = measure(ctx, args, args.workload, F, args.steps, args.warmup, sampler)
    clocks = sampler.stop()

    workloads = []
    if not args.no_secondary and args.workload == HEADLINE:
        for wl, f in SECONDARY:
            workloads.append(measure(ctx, args, wl, frames_for(wl, f), max(1, min(args.steps, 2)), max(1, min(args.warmup, 2))))

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference's own decoder on this box's host cores ---------
    cpu_baseline = None
    n = head["n"]
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cf = args.cpu_frames or auto_cpu_frames(args.workload, cores)
        dt, kind, it, fl = cpu_reference_run(args.workload, cf, cores)
        cpu_baseline = {"value": n * cf / dt / 1e9, "unit": "Gbit/s", "cores": cores, "kind": kind,
                        "sample": f"{cf} frames of the same workload (keys from the reference's generator), decode only, {dt:.1f} s",
                        "mean_iterations": float(np.mean(it)), "fer": float(1.0 - np.mean((fl & 3) == 3))}
        # and on ONE host thread (how the reference's authors run their throughput configs)
        cf1 = max(4, cf // (2 * cores))
        dt1, _, _, _ = cpu_reference_run(args.workload, cf1, 1)
        cpu_baseline["single_thread"] = {"value": n * cf1 / dt1 / 1e9, "unit": "Gbit/s", "sample": f"{cf1} frames, {dt1:.1f} s"}

    if rank == 0:
        out = {
            "metric": "decoded key Gbit/s", "value": head["value"], "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": head["dtype"], "data": "synthetic",
            "config": {"workload": head["workload"], "workload_id": args.workload, "frames_per_step_per_gpu": F,
                       "max_iterations": MAX_ITER, "threshold": THRESHOLD, "accurate_qber": head["accurate_qber"],
                       "frames_per_tile": head["frames_per_tile"], "pool_tiles": head["pool_tiles"], "pool_bytes": head["pool_bytes"],
                       "decoder_path": head["decoder_path"], "onchip_threads": head["onchip_threads"],
                       "onchip_record_bytes": head["onchip_record_bytes"], "l2_policy": head["l2_policy"],
                       "mean_iterations_executed": head["mean_iterations_executed"], "fer": head["fer"],
                       "parallelism": f"frames sharded over {world} GPU(s), tally all-reduce only (ncclAllReduce of {head['tally_len']} x u64 inside the library)"},
            "roofline": head["roofline"], "cpu_baseline": cpu_baseline, "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
            "clocks": clocks, "workloads": workloads,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
