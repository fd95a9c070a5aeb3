#!/usr/bin/env python
"""Headline benchmark: decoded key Gbit/s of the batched LDPC syndrome decoder (BASELINE.json metric).

Workload (config.workload): `config 10k NMSA.json`'s operating family at the north-star point -- the irregular
n=10240, m=2048 (R=0.8, E=60430) code `matrices_2/(N=10240,M=2048,R=0.8).mtrx`, normalized min-sum alpha=0.7,
QBER 3 %, 100 iterations max, message clamp 100, float32 messages, synthetic keys with run_trial's distribution.
One "step" = one pass of the hot path over one batch of --frames frames per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
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
    "I80_aomsa_q015": ("I80", 5, 0.70, 0.99, 0.015, "n=10240 m=2048 irregular R=0.8 (E=60430), AOMSA beta=0.70 sigma=0.99, QBER 1.5% (ADAPTIVE R.json family)"),
}
MAX_ITER, THRESHOLD = 100, 100.0


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default="I80_nmsa_q030", choices=sorted(WORKLOADS))
    p.add_argument("--frames", type=int, default=32768, help="frames per GPU per step")
    p.add_argument("--precision", type=int, default=0, choices=[0, 32, 64],
                   help="message precision: 0 = the library's policy (float64 state for OMSA / ANMSA / AOMSA and SPA on n > 65536)")
    p.add_argument("--pool-slots", type=int, default=0)
    p.add_argument("--frames-per-lane", type=int, default=0)
    p.add_argument("--path", type=int, default=0, choices=[0, 1, 2],
                   help="decoder path: 0 auto (on-chip kernels when eligible), 1 streaming (messages in HBM), 2 on-chip")
    p.add_argument("--onchip-threads", type=int, default=0)
    p.add_argument("--no-compaction", action="store_true", help="streaming path: do not compact the tail of a draining batch")
    p.add_argument("--cpu-frames", type=int, default=0, help="frames of the CPU-baseline sample (0 = auto)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
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

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.time()
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        window = "timed region"
        rows = [r for t, r in self.rows if self.t_begin is None or t >= self.t_begin]
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


def cpu_reference_run(wl, n_frames, threads, seed=20261018):
    """The reference's own CPU decoder (oracle/_ref: unmodified sources) or, if that library did not travel, our C
    port of it, on `n_frames` frames of the workload with `threads` host threads. Returns (seconds, kind, iters)."""
    import util
    from qkd_ldpc_v_b200 import hostlib, unpack_bits
    name, alg, pri, sec, qber, _ = WORKLOADS[wl]
    arr = util.code_arrays(name)
    seeds = hostlib.trial_seeds(seed, n_frames)
    a, b, acc = hostlib.gen_keys(seeds, arr["n"], qber)
    ab, bb = unpack_bits(a, arr["n"]), unpack_bits(b, arr["n"])
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
        ai, bi = ab.astype(np.int32), bb.astype(np.int32)
        t0 = time.perf_counter()
        it, fl = m.qkd_ldpc_batch(ai, bi, acc, pri, sec, threads=threads)
        dt = time.perf_counter() - t0
        return dt, "reference", it, fl
    from oracle import cpu
    oc = util.oracle_code(name)
    t0 = time.perf_counter()
    it, fl, _ = cpu.qkd_ldpc_batch(oc, alg, ab, bb, acc, max_iter=MAX_ITER, primary=pri, secondary=sec,
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
    """--impl reference: the reference's CPU implementation on this box's host cores (rank 0 only)."""
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
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "frames_per_step": frames, "max_iterations": MAX_ITER, "threshold": THRESHOLD,
                   "mean_iterations": float(np.mean(it))},
        "cpu_baseline": {"value": gbit, "unit": "Gbit/s", "cores": cores, "kind": kind,
                         "sample": f"{frames} frames per step x {args.steps} steps, decode only, reference RNG keys"},
        "e2e": {"value": gbit, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


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

    import qkd_ldpc_v_b200 as q
    import util

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    name, alg, pri, sec, qber, desc = WORKLOADS[args.workload]
    arr = util.code_arrays(name)
    n, m, nnz = arr["n"], arr["m"], arr["nnz"]
    words = (n + 31) // 32
    F = args.frames
    code = q.LdpcCode(n, m, arr["row_ptr"], arr["col_idx"], device=local_rank, pool_slots=args.pool_slots,
                      frames_per_lane_f32=args.frames_per_lane, decoder_path=args.path, onchip_threads=args.onchip_threads,
                      tail_compaction=-1 if args.no_compaction else 0)
    stream = torch.cuda.Stream(device=dev)      # a real stream: the decoder replays CUDA graphs on it
    torch.cuda.set_stream(stream)
    code.set_stream(stream.cuda_stream)
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
        if world > 1:   # the only collective of the path: FER / iteration tallies (K5)
            dist.all_reduce(d_tally)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    # ---- timed region: device-resident inputs, CUDA events on the launching stream, per-kernel events on ----------
    code.set_profiling(True)
    l0 = code.info()["kernel_launches"]
    barrier()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    cn_ms = vn_ms = sc_ms = batch_ms = 0.0
    for _ in range(args.steps):
        step()
        inf = code.info()
        cn_ms += inf["last_cn_ms"]; vn_ms += inf["last_vn_ms"]; sc_ms += inf["last_sched_ms"]  # noqa: E702
        batch_ms += inf["last_batch_ms"]
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    inf = code.info()
    launches = inf["kernel_launches"] - l0
    elapsed_ms = e0.elapsed_time(e1)
    code.set_profiling(False)
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    tally = d_tally.cpu().numpy().astype(np.uint64)
    frames_total = F * world
    iters_executed = int(tally[3])           # of the LAST step, all ranks
    value = n * frames_total * args.steps / (elapsed_ms * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (algorithmic bytes: SURVEY.md 8d) ---------------------------
    prec = inf["last_precision"]
    sz = 4 if prec == 32 else 8
    it_rank = iters_executed / world         # frame-iterations one rank executed in one step
    cn_bytes, vn_bytes = 2 * sz * nnz * it_rank, (2 * sz * nnz + sz * n) * it_rank   # per step, all launches
    peak, peak_src = measured_peak()
    onchip = inf["last_path"] == 2
    traffic_db = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic_db = json.load(f)
    except Exception:
        pass
    if onchip:
        # ONE persistent kernel per step; the belief-propagation state never leaves shared memory, so the algorithmic
        # bytes (what a message-streaming decoder has to move) are not HBM traffic here: frac > 1 is the design's point.
        k_ms = batch_ms / args.steps
        k_bytes = cn_bytes + vn_bytes
        ach_k = k_bytes / (k_ms * 1e-3) / 1e9
        kname = ("onchip_minsum_kernel" if prec == 32 else "onchip_minsum64_kernel") if alg >= 2 else "onchip_spa_kernel"
        tr = traffic_db.get(kname, {}).get("dram_bytes_per_frame")
        roofline = {
            "bound": "hbm", "kernel": f"{kname}<ALG={alg}>", "achieved": ach_k, "peak": peak, "unit": "GB/s",
            "frac": ach_k / peak, "traffic": (tr * F if tr is not None else None), "peak_source": peak_src,
            "bytes_per_launch": k_bytes, "ms_per_launch": k_ms, "launches_per_step": 1,
            "whole_step_frac": (k_bytes / (elapsed_ms / args.steps * 1e-3) / 1e9) / peak,
            "note": "algorithmic bytes = (16*E + 4*N) per frame-iteration (SURVEY.md 8d); this kernel keeps them on chip "
                    + (("(4N + 16M bytes of shared memory per frame)" if prec == 32 else "(8N + 24M bytes of shared memory per frame)")
                       if alg >= 2 else "(4N + 4E bytes of shared memory per frame)")
                    + ", DRAM traffic is the packed keys only -- see `traffic`",
        }
    else:
        kern = {"cn": (cn_bytes, cn_ms / args.steps), "vn": (vn_bytes, vn_ms / args.steps)}
        dom = max(kern, key=lambda k: kern[k][1])
        ach = {k: (v[0] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else 0.0) for k, v in kern.items()}
        launches_per_step = launches / args.steps
        n_cn_launches = max(1.0, (launches_per_step - 2) / 3)
        tr = traffic_db.get({"cn": "cn_kernel", "vn": "vn_kernel"}[dom], {}).get("dram_bytes_per_frame_iter")
        roofline = {
            "bound": "hbm", "kernel": {"cn": f"cn_kernel<{'float' if sz == 4 else 'double'},ALG={alg}>", "vn": "vn_kernel"}[dom],
            "achieved": ach[dom], "peak": peak, "unit": "GB/s", "frac": ach[dom] / peak,
            "traffic": (tr * it_rank / n_cn_launches if tr is not None and sz == 4 else None),
            "peak_source": peak_src,
            "bytes_per_launch": kern[dom][0] / n_cn_launches, "ms_per_launch": kern[dom][1] / n_cn_launches,
            "launches_per_step": n_cn_launches,
            "both_kernels": {k: {"achieved_gbs": ach[k], "frac": ach[k] / peak, "ms_per_step": kern[k][1]} for k in kern},
            "sched_ms_per_step": sc_ms / args.steps,
            "whole_step_frac": ((cn_bytes + vn_bytes) / (elapsed_ms / args.steps * 1e-3) / 1e9) / peak,
        }

    # ---- e2e: the same metric through the host-buffer C-ABI call, pinned host memory, copies inside ---------------
    e2e = None
    if not args.no_e2e:
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
        for _ in range(args.steps):
            r = code.QKD_LDPC_batch(a_np, b_np, acc_q, (pri, sec), cfg, out=outbuf)
            if world > 1:
                tt = torch.from_numpy(r.tally.astype(np.int64)).to(dev)
                dist.all_reduce(tt)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": n * frames_total * args.steps / dt / 1e9, "unit": "Gbit/s",
               "h2d_bytes_per_step": int(2 * F * words * 4 + 8), "d2h_bytes_per_step": int(F * words * 4 + F * 5 + tl * 8),
               "api": "qkdldpc_decode_batch (host buffers, pinned)"}

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference's own decoder on this box's host cores ---------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cf = args.cpu_frames or auto_cpu_frames(args.workload, cores)
        dt, kind, it, fl = cpu_reference_run(args.workload, cf, cores)
        cpu_baseline = {"value": n * cf / dt / 1e9, "unit": "Gbit/s", "cores": cores, "kind": kind,
                        "sample": f"{cf} frames of the same workload (reference RNG keys), decode only, {dt:.1f} s",
                        "mean_iterations": float(np.mean(it)), "fer": float(1.0 - np.mean((fl & 3) == 3))}
        # and on ONE host thread (how the reference's authors run their throughput configs)
        cf1 = max(4, cf // (2 * cores))
        dt1, _, _, _ = cpu_reference_run(args.workload, cf1, 1)
        cpu_baseline["single_thread"] = {"value": n * cf1 / dt1 / 1e9, "unit": "Gbit/s", "sample": f"{cf1} frames, {dt1:.1f} s"}

    if rank == 0:
        stats = q.stats_from_tally(tally, frames_total)
        out = {
            "metric": "decoded key Gbit/s", "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if prec == 32 else "f64",
            "data": "synthetic",
            "config": {"workload": desc, "workload_id": args.workload, "frames_per_step_per_gpu": F,
                       "max_iterations": MAX_ITER, "threshold": THRESHOLD, "accurate_qber": acc_q,
                       "frames_per_tile": inf["frames_per_tile"], "pool_tiles": inf["pool_tiles"],
                       "pool_bytes": inf["pool_bytes"],
                       "decoder_path": (("on-chip min-sum" if alg >= 2 else "on-chip sum-product") + " (frame state in shared memory)")
                                       if onchip else "streaming (messages in HBM)",
                       "onchip_threads": inf.get("onchip_threads") if onchip else None,
                       "l2_policy": ("inputs larger than L2: %.0f MB of packed keys in + decisions out per step, read once; decoder "
                                     "state is in shared memory" % (3 * F * words * 4 / 1e6)) if onchip else
                                    ("inputs larger than L2 (message pool %.1f GB >> 126 MB)" % (inf["pool_bytes"] / 1e9)),
                       "mean_iterations_executed": iters_executed / frames_total, "fer": stats["FER"],
                       "parallelism": f"frames sharded over {world} GPU(s), tally all-reduce only"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
