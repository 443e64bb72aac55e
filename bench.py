#!/usr/bin/env python3
"""RawBoost throughput benchmark: augmented utterances/sec on 64600-sample 16 kHz synthetic waveforms.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--algo 5] [--batch 4096] [--impl b200|reference]

One "step" = one pass of the hot path (process_Rawboost_feature, fused on the device) over one batch of
synthetic utterances. N > 1 is launched by torchrun, one rank per GPU: utterances are sharded by index with no
data-path collective (weak scaling: --batch utterances PER GPU); torch.distributed (NCCL) carries only the barrier
and the MAX reduction of the step time. Rank 0 prints ONE JSON line.

* ``value``  : whole-job utterances/s with inputs and the pre-drawn plan bank resident in HBM (CUDA events).
* ``e2e``    : the same metric through the host-buffer C ABI (``rb_process_host``): every step draws the plans on the
               host (the reference's numpy calls, parallel over host cores), copies waveforms + plans host->device
               from pinned/pageable host memory, runs the kernels and copies the result back.
* ``roofline``: the dominant kernel (FIR bank) -- algorithmic FLOPs of the actual drawn taps / its device time
               (events on its own stream, recorded by the library) against the FP32-pipe rate measured in this run
               by a register-resident FFMA2 chain; plus the whole step's algorithmic bytes against measured HBM.
* ``cpu_baseline`` (N=1, rank 0) and ``--impl reference``: the CPU oracle port of the reference's numpy/scipy path
               (the Python reference cannot travel to the GPU box), one process per host core, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "augmented utterances/sec (64600-sample, 16 kHz)"
UNIT = "utt/s"


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, one process per core
# ---------------------------------------------------------------------------------------------------------
_CPU_CACHE = {}


def _cpu_task(job):
    """Run the reference path (oracle port) on a fixed set of utterances; returns seconds spent inside it."""
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    from oracle import rawboost_oracle as orc  # CPU baseline leg: the only place bench.py executes oracle/
    algo, length, indices = job
    args = orc.make_args()
    key = (length, tuple(indices))
    if key not in _CPU_CACHE:
        _CPU_CACHE.clear()
        _CPU_CACHE[key] = [orc.synth_utterance(u, length, bool(u % 2)) for u in indices]
    waves = _CPU_CACHE[key]
    t0 = time.perf_counter()
    for u, x in zip(indices, waves):
        np.random.seed(orc.seed_for(u))
        orc.process(x, 16000, args, algo)
    return time.perf_counter() - t0


class CpuArm:
    def __init__(self, algo: int, length: int, per_core: int, cores: int):
        self.algo, self.length, self.per_core, self.cores = algo, length, per_core, cores
        for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ.setdefault(k, "1")
        self.pool = mp.get_context("fork").Pool(cores)
        self.jobs = [(algo, length, list(range(w * per_core, (w + 1) * per_core))) for w in range(cores)]

    def step(self) -> float:
        t0 = time.perf_counter()
        self.pool.map(_cpu_task, self.jobs, chunksize=1)
        return time.perf_counter() - t0

    @property
    def utterances_per_step(self) -> int:
        return self.per_core * self.cores

    def close(self):
        self.pool.close()
        self.pool.join()


def run_cpu_arm(algo, length, per_core, cores, steps, warmup):
    arm = CpuArm(algo, length, per_core, cores)
    try:
        for _ in range(warmup):
            arm.step()
        times = [arm.step() for _ in range(steps)]
    finally:
        arm.close()
    total = sum(times)
    return arm.utterances_per_step * steps / total, 1e3 * total / steps, arm.utterances_per_step


def workload_name(algo, batch, length):
    return f"RawBoost algo={algo} (main.py:258-298 default args), {batch} synthetic {length}-sample 16 kHz utterances per GPU"


def reference_arm(a):
    """--impl reference: the reference's CPU implementation of the path, all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    per_core = a.cpu_per_core
    value, ms, n = run_cpu_arm(a.algo, a.length, per_core, cores, a.steps, a.warmup)
    sample = (f"{n} utterances per step ({per_core} per core x {cores} cores) of the same seeded workload; "
              f"oracle port of datautils/RawBoost.py (numpy/scipy), BLAS threads pinned to 1")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a.algo, a.batch, a.length), "algo": a.algo, "utt_len": a.length, "sample_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampling of SM clocks and throttle reasons during the timed region."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.gpu_index = gpu_index
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [c.strip() for c in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power), "samples": len(sm),
                "reasons": sorted(reasons)}


def traffic_for(kernel, algo, batch):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from the committed
    ``ncu --set full`` capture of the same configuration (profiles/traffic.json), or None when there is none."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        table = json.load(f)
    return table.get(f"{kernel}:algo{algo}:b{batch}")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def gpu_arm(a):
    from scl_deepfake_audio_detection_b200 import sharding, workload
    rank, local_rank, world = sharding.env_rank_world()
    cores = host_cores()
    workers = max(1, cores // max(1, world))
    native = a.planner in ("native", "device")  # the resident plan bank is drawn by the native host planner in both modes
    pool = workload.PlanPool(1 if native else workers)  # forked before CUDA is initialised; workers only run numpy/scipy

    import ctypes as C
    import torch
    from scl_deepfake_audio_detection_b200.engine import Engine

    torch.cuda.set_device(local_rank)
    if world > 1:
        sharding.init_process_group("nccl")
    dev = torch.device("cuda", local_rank)
    eng = Engine(local_rank)
    lib = eng.lib
    args = workload.default_args()
    B, L, algo = a.batch, a.length, a.algo
    lo = rank * B  # weak scaling: every rank owns B utterances, global indices [rank*B, (rank+1)*B)
    seeds = [workload.seed_for(u) for u in range(lo, lo + B)]
    lengths = [L] * B

    # ---- synthetic inputs (pinned host copy) and the plan bank ------------------------------------------
    planners = []
    if native:
        from scl_deepfake_audio_detection_b200.native_planner import NativePlanner
        planners = [NativePlanner(threads=workers, pinned=True) for _ in range(3)]  # [0]: resident bank, [1],[2]: e2e double buffer
        t0 = time.perf_counter()
        bp = planners[0].draw(lengths, workload.SAMPLE_RATE, args, algo, seeds=seeds)
        t_plan = time.perf_counter() - t0
        # guard: the native planner must reproduce the numpy draws (integers and float32 taps) on a sample
        ref = pool.draw_batch(lengths[:4], workload.SAMPLE_RATE, args, algo, seeds[:4], ld=bp.ld)
        for name in ("lnl_tap_off", "lnl_taps", "isd_off", "isd_idx", "isd_fr"):
            r = getattr(ref, name)
            if r is not None:
                assert np.array_equal(r, getattr(bp, name)[:r.shape[0]]), f"native planner diverges from numpy on {name}"
    else:
        t0 = time.perf_counter()
        bp = pool.draw_batch(lengths, workload.SAMPLE_RATE, args, algo, seeds)
        t_plan = time.perf_counter() - t0
    ld = bp.ld
    x_host = torch.empty((B, ld), dtype=torch.float32).pin_memory()
    workload.synth_batch(lo, B, L, ld, out=x_host.numpy())
    y_host = torch.empty((B, ld), dtype=torch.float32).pin_memory()
    x = x_host.to(dev)
    ln = torch.from_numpy(bp.lengths).to(dev)
    dp = eng.upload_plan(bp)
    y = torch.zeros_like(x)
    torch.cuda.synchronize()

    # ---- FP32-pipe rate of this GPU, measured here (roofline denominator of the FIR kernel) --------------
    sink = torch.zeros(4, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def probe(packed):
        best = 0.0
        flops = C.c_double(0.0)
        for i in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.rb_probe_fp32(packed, 2000, sink.data_ptr(), C.byref(flops), C.c_void_p(stream))
            e1.record()
            e1.synchronize()
            assert rc == 0, rc
            if i:
                best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        return best

    fp32_ffma2 = probe(1)
    fp32_ffma = probe(0)
    fp32_peak = max(fp32_ffma2, fp32_ffma)
    hbm_peak, hbm_src = measured_peaks()

    # ---- device-resident timing --------------------------------------------------------------------------
    for _ in range(a.warmup):
        eng.process(algo, x, ln, dp, out=y)
    torch.cuda.synchronize()
    lib.rb_profile_read(None, None, 1)
    lib.rb_profile_enable(1)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    sharding.barrier()
    torch.cuda.synchronize()
    launches0 = lib.rb_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record()
    for i in range(a.steps):
        eng.process(algo, x, ln, dp, out=y)
        ev[i + 1].record()
    torch.cuda.synchronize()
    sharding.barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)]
    launches = int(lib.rb_launch_count() - launches0)
    clocks = sampler.stop() if sampler else None
    lib.rb_profile_enable(0)
    fir_ms, fir_n = C.c_double(0.0), C.c_uint64(0)
    lib.rb_profile_read(C.byref(fir_ms), C.byref(fir_n), 1)
    ms_per_step = sharding.max_over_ranks(total_ms / a.steps, dev)
    value = world * B / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer C ABI: plan draw (host, parallel) + H2D + kernels + D2H every step ---
    def e2e_run_native(steps, warm):
        """Plan drawing for step i+1 runs on the planner's host threads (no GIL) while step i is on the GPU."""
        import threading
        slots = planners[1:]
        box = {}

        def draw_into(k):
            box[k] = slots[k % 2].draw(lengths, workload.SAMPLE_RATE, args, algo, seeds=seeds, ld=ld)

        def run_steps(n):
            th = threading.Thread(target=draw_into, args=(0,))
            th.start()
            for k in range(n):
                th.join()
                plan_k = box.pop(k)
                if k + 1 < n:
                    th = threading.Thread(target=draw_into, args=(k + 1,))
                    th.start()
                eng.process_host(algo, x_host.numpy(), plan_k, out=y_host.numpy())

        run_steps(warm)
        sharding.barrier()
        t0 = time.perf_counter()
        run_steps(steps)
        dt = time.perf_counter() - t0
        sharding.barrier()
        return dt / steps

    seeds_np = np.asarray(seeds, dtype=np.uint32)

    def e2e_run_device(steps, warm):
        """Seeds in, results out: plans are drawn on the device inside the pipelined host-buffer call."""
        for _ in range(warm):
            eng.process_host_seeded(algo, x_host.numpy(), bp.lengths, seeds_np, workload.SAMPLE_RATE, args, out=y_host.numpy())
        sharding.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            eng.process_host_seeded(algo, x_host.numpy(), bp.lengths, seeds_np, workload.SAMPLE_RATE, args, out=y_host.numpy())
        dt = time.perf_counter() - t0
        sharding.barrier()
        return dt / steps

    y_host2 = torch.empty((B, ld), dtype=torch.float32).pin_memory() if a.planner == "device" and not a.no_e2e else None
    lengths_i32 = np.ascontiguousarray(bp.lengths, dtype=np.int32)

    def e2e_run_stream(steps, warm):
        """The streaming form a production loader uses: rb_submit_host_seeded / rb_ctx_wait with two calls in flight and two
        result buffers -- step k's results are collected while step k+1 is already being copied in. Every step still moves
        its waveforms host -> device and its results device -> host inside the timed region."""
        bufs = (y_host.numpy(), y_host2.numpy())

        def run(n):
            prev = None
            for k in range(n):
                t = eng.submit_host_seeded(algo, x_host.numpy(), lengths_i32, seeds_np, workload.SAMPLE_RATE, args, out=bufs[k % 2])
                if prev is not None:
                    eng.wait_host(prev)
                prev = t
            eng.wait_host(prev)

        run(warm)
        sharding.barrier()
        t0 = time.perf_counter()
        run(steps)
        dt = time.perf_counter() - t0
        sharding.barrier()
        return dt / steps

    def e2e_run(steps, warm):
        if a.planner == "device":
            return e2e_run_device(steps, warm)
        if native:
            return e2e_run_native(steps, warm)
        pending = pool.pool.map_async(workload._draw_chunk, _plan_jobs(lengths, args, algo, seeds, workers)) if pool.pool else None

        def next_plan():
            nonlocal pending
            if pending is None:
                return pool.draw_batch(lengths, workload.SAMPLE_RATE, args, algo, seeds, ld=ld)
            from scl_deepfake_audio_detection_b200 import plans as _pl
            got = _pl.pack([p for chunk in pending.get() for p in chunk], ld=ld)
            pending = pool.pool.map_async(workload._draw_chunk, _plan_jobs(lengths, args, algo, seeds, workers))
            return got

        for _ in range(warm):
            eng.process_host(algo, x_host.numpy(), next_plan(), out=y_host.numpy())
        sharding.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            eng.process_host(algo, x_host.numpy(), next_plan(), out=y_host.numpy())
        dt = time.perf_counter() - t0
        sharding.barrier()
        if pending is not None:
            pending.get()
        return dt / steps

    e2e_steps = max(1, min(a.steps, a.e2e_steps))
    if a.host_chunk:
        eng.set_host_chunk(a.host_chunk)
    if a.no_e2e:
        e2e_s, sync_s, copy_s, h2d, d2h, check = float("nan"), float("nan"), float("nan"), 0, 0, None
    else:
        sync_s = sharding.max_over_ranks(e2e_run(e2e_steps, 2), dev)   # one blocking call per step
        h2d, d2h = eng.last_host_traffic()
        e2e_check = y_host.numpy()[:, :L].copy() if a.planner == "device" else None
        e2e_s = sync_s
        if a.planner == "device":
            e2e_s = sharding.max_over_ranks(e2e_run_stream(e2e_steps, 2), dev)
            assert np.array_equal(y_host.numpy()[:, :L], e2e_check) and np.array_equal(y_host2.numpy()[:, :L], e2e_check), \
                "streamed results differ from the blocking call's"
        # copy + kernels only (plans pre-drawn), for the breakdown
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.process_host(algo, x_host.numpy(), bp, out=y_host.numpy())
        copy_s = sharding.max_over_ranks((time.perf_counter() - t0) / e2e_steps, dev)
        check = float(np.abs(y_host.numpy()[:, :L]).max())
        if e2e_check is not None:  # device-drawn plans must reproduce the host-drawn plan bank bit for bit
            assert np.array_equal(e2e_check, y_host.numpy()[:, :L]), "device-planned e2e result differs from the host-planned one"

    # ---- CPU baseline on this box's host cores (N=1 only) -----------------------------------------------
    cpu = None
    pool.close()
    if world == 1 and not a.no_cpu:
        v, ms, n = run_cpu_arm(algo, L, a.cpu_per_core, cores, 2, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} utterances per step ({a.cpu_per_core} per core x {cores} cores), 2 timed steps, oracle port of the "
                         f"reference numpy/scipy path (draw + apply), BLAS threads pinned to 1"}

    if rank == 0:
        flops_step = bp.fir_flops()
        bytes_step = bp.io_bytes()
        fir_avg_ms = fir_ms.value / max(1, fir_n.value)
        fir_per_step = fir_n.value / a.steps
        achieved_tf = flops_step / max(1e-9, fir_avg_ms * fir_per_step * 1e-3) / 1e12 if fir_n.value else None
        step_s = total_ms / a.steps * 1e-3
        step_hbm = {"bound": "hbm", "achieved": bytes_step / step_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": bytes_step / step_s / 1e9 / hbm_peak, "peak_source": hbm_src, "algorithmic_bytes_per_step": bytes_step}
        if fir_n.value:  # FP32-pipe bound: the FIR-bank kernel (with its fused per-utterance tail) is the step
            roofline = {
                "bound": "fp32", "kernel": "fir_bank_kernel", "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / fp32_peak, "traffic": traffic_for("fir_bank_kernel", algo, B),
                "peak_source": "FFMA2 register-resident chain measured in this run (rb_probe_fp32); scalar FFMA chain gave %.1f; "
                               "MEASURED_PEAKS.json carries HBM and bf16-tensor peaks only and north_star puts this path on the "
                               "non-tensor FP32 pipe" % fp32_ffma,
                "kernel_ms_per_launch": fir_avg_ms, "kernel_launches_per_step": fir_per_step,
                "kernel_share_of_step": (fir_avg_ms * fir_per_step) / (total_ms / a.steps),
                "algorithmic_flops_per_step": flops_step, "step_hbm": step_hbm}
        else:           # HBM bound (algo 2): one fused kernel per step, algorithmic bytes over the step time
            roofline = dict(step_hbm, kernel="isd_fused_kernel", traffic=traffic_for("isd_fused_kernel", algo, B),
                            kernel_ms_per_launch=total_ms / a.steps, kernel_launches_per_step=1.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(algo, B, L), "algo": algo, "batch_per_gpu": B, "global_batch": world * B, "utt_len": L,
                       "sharding": f"utterances by index over {world} GPU(s), no collective",
                       "l2": "inputs (%.2f GB per GPU) larger than L2; no explicit flush" % (x.numel() * 4 / 1e9),
                       "plan_bank": "drawn on the host (%s planner, reference RNG stream order), resident in HBM before timing" % ("numpy" if a.planner == "numpy" else "native"),
                       "step_ms_min_max": [min(step_ms), max(step_ms)], "plan_draw_s_setup": t_plan, "host_plan_workers": workers},
            "roofline": roofline,
            "e2e": {"value": world * B / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "includes": ("every step: H2D of its waveforms + seeds from pinned host memory, plan draw ON THE DEVICE (bit-exact "
                                 "replay of numpy's MT19937 stream), kernels, D2H of its results into pinned host memory; streaming "
                                 "API rb_submit_host_seeded / rb_ctx_wait with two calls in flight (results of step k collected "
                                 "while step k+1 is copied in); blocking_call_* = one rb_process_host_seeded call per step"
                                 if a.planner == "device" else
                                 "host plan draw (%s, %d host %s, overlapped with the previous step) + H2D + kernels + D2H"
                                 % (("native planner: bit-exact numpy MT19937 stream + float64 filter design", workers, "threads")
                                    if native else ("numpy/scipy", workers, "processes"))),
                    "planner": a.planner,
                    "plan_draw_s_per_batch": t_plan,
                    "ms_per_step": e2e_s * 1e3, "blocking_call_ms_per_step": sync_s * 1e3,
                    "blocking_call_value": world * B / sync_s, "copy_and_kernels_only_ms_per_step": copy_s * 1e3,
                    "copy_and_kernels_only_value": world * B / copy_s, "steps": e2e_steps, "result_peak_abs": check},  # peak |y| of the copied-back results (1.0 after normWav): a liveness check, not an error
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if a.no_e2e:  # diagnostic runs only: no end-to-end figure was taken
            line["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "skipped": "--no-e2e"}
        if cpu:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def _plan_jobs(lengths, args, algo, seeds, workers):
    n = len(lengths)
    step = max(1, (n + 4 * workers - 1) // (4 * workers))
    return [(list(lengths[i:i + step]), 16000, vars(args), algo, list(seeds[i:i + step])) for i in range(0, n, step)]


_REAL_STDOUT = None


def quiet_stdout():
    """Route everything libraries print to fd 1 (e.g. NCCL's version banner) to stderr; rank 0's JSON line goes to the real
    stdout through :func:`emit`, so stdout carries exactly one line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--algo", type=int, default=5, help="RawBoost algo (BASELINE config 3: algo 5 = LnL -> ISD)")
    ap.add_argument("--batch", type=int, default=4096, help="utterances per GPU per step (BASELINE config 3: 4096)")
    ap.add_argument("--length", type=int, default=64600)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--planner", choices=["device", "native", "numpy"], default="device",
                    help="plan drawing in the e2e leg: on the device from per-utterance seeds (default), the native host "
                         "re-implementation, or the numpy calls themselves")
    ap.add_argument("--host-chunk", type=int, default=0, help="utterances per pipeline chunk of the host-buffer entry (0 = 4 per SM)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    ap.add_argument("--cpu-per-core", type=int, default=8, help="utterances per host core per CPU-baseline step")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    quiet_stdout()
    if a.impl == "reference":
        reference_arm(a)
    else:
        gpu_arm(a)


if __name__ == "__main__":
    main()
