#!/usr/bin/env python
"""bench.py -- the headline measurement: GBM paths/sec, European call, ONE 2^30-path job on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path: pricing a European call (BASELINE.json configs[1]:
S0=100 K=100 r=0.05 sigma=0.2 T=1, single step) on 2^30 paths.  With N GPUs the SAME 2^30-path job is
sharded by path index (strong scaling, BASELINE.json's metric): rank g prices the chunks of its 64/N
reduction segments in ONE kernel launch that also folds those segments and stores them into every
rank's mailbox over NVLink; a one-warp final pass on a second stream runs the fixed tree.  The price is
bit-identical for every N (the line carries its hex digits).  `--scaling weak` prices 2^30 paths PER
GPU instead.  Rank 0 prints ONE JSON line.

  value    : whole-job paths/s, device-timed with CUDA events on the engine's own streams
             (mcb_pipeline_timer_*: from before the first pricing launch to after the last final pass,
             max over ranks); K jobs are submitted back to back, nothing to stage in HBM
             (the path has no input arrays: a path is a pure function of (seed, path id)).
  e2e      : the same metric through the public synchronous C-ABI call mcb_price_european, one call per
             step (host OptionData in, host mcb_result out through mapped pinned memory); at N > 1 it is
             rank 0 driving ONE engine over all N GPUs (mcb_engine_create_multi), the other ranks idle.
  roofline : dominant kernel european_job_kernel against the SM issue roofline (north_star: "FP32/SFU
             compute roofline" -- nothing here touches HBM or tensor cores), algorithmic
             thread-instructions per path from SURVEY.md 8(d); plus `roofline_trajectory`, the
             HBM-bound trajectory-store kernel (configs[2]) against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline : the UNMODIFIED reference CPU pricer (oracle/_ref, simulateOptionPriceCPU,
             inc/tool.cuh:104-130) on this box's host cores, bounded sample.  Reported, not a target.

`--impl reference` times only that CPU pricer (rank 0; other ranks exit 0).
oracle/ is used here ONLY as cpu_baseline / reference arm, never on the measured GPU path.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "GBM paths/sec (European call, 2^30 paths)"
UNIT = "paths/s"
PATHS_PER_GPU = 1 << 30
CFG = dict(S0=100.0, K=100.0, r=0.05, v=0.2, T=1.0)
SEED = 1234

# Algorithmic work per unit (restated in DESIGN.md "Rooflines")
INSTR_PER_EUROPEAN_PATH = 57          # 42 INT + 11 FP32 + 4 MUFU thread-instructions, canonical keying
EXECUTED_INSTR_PER_EUROPEAN_PATH = 52.3   # what the shipped SASS executes per path (ncu, profiles/)
IMAD_WIDE_PER_CLK_PER_SM = 26.7       # measured, profiles/pipe_microbench_r1.txt
WALK_IMAD_WIDE_PER_STEP = 4.75        # multi-step walk kernels (bullet, nested): 19 per Philox block of 4 steps
ISSUE_PER_CLK_PER_SM = 128            # 4 schedulers x 32 lanes
BYTES_PER_TRAJECTORY_STEP = 4         # one float stored per path-step
TRAJ_PATHS, TRAJ_STEPS = 1 << 20, 252


def workload_config(n_gpus: int, scaling: str):
    """The `config` object of BOTH arms (the reference arm times a bounded sample of this same workload and says so
    in cpu_baseline.sample): BASELINE.json configs[1], one job of 2^30 paths (strong) or 2^30 paths per GPU (weak)."""
    strong = scaling == "strong"
    n_total = PATHS_PER_GPU if strong else PATHS_PER_GPU * n_gpus
    return {"workload": "European call S0=100 K=100 r=0.05 sigma=0.2 T=1, single step, "
                        + ("ONE job of 2^30 paths sharded over the GPUs by path index" if strong else "2^30 paths per GPU")
                        + f" ({n_total} paths per step)",
            "paths_per_step": n_total, "n_gpus": n_gpus, "paths_per_gpu": n_total // max(n_gpus, 1),
            "l2": "n/a: the path reads no global memory (a path is a function of (seed, path id); 64 Ki chunk partials "
                  "of 8 B are written per job)",
            **CFG}


def bs_call(S0, K, T, r, v):
    d1 = (math.log(S0 / K) + (r + 0.5 * v * v) * T) / (v * math.sqrt(T))
    d2 = d1 - v * math.sqrt(T)
    phi = lambda x: 0.5 * math.erfc(-x / math.sqrt(2.0))
    return S0 * phi(d1) - K * math.exp(-r * T) * phi(d2)


def ncu_traffic(kernel):
    """DRAM bytes per launch (read + write) of `kernel` from the committed ncu --set full capture
    (profiles/r2_ncu_traffic.json, written by tools/ncu_traffic.py); None when not captured."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)[kernel]["traffic"]
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


# ------------------------------------------------------------------------------ clock sampler
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML (what nvidia-smi prints) every 20 ms."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, device_index: int):
        super().__init__(daemon=True)
        self.samples = []   # (t, sm_mhz, reasons_mask, phase)
        self.phase = "idle"
        self._halt = threading.Event()
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as exc:  # NVML missing: the JSON says so instead of inventing clocks
            self.err = repr(exc)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._halt.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.perf_counter(), mhz, mask, self.phase))
            except Exception:
                pass
            self._halt.wait(0.02)

    def stop(self):
        self._halt.set()

    def summary(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml")}
        timed = [s for s in self.samples if s[3] == "timed"]
        used = timed if len(timed) >= 3 else [s for s in self.samples if s[3] != "idle"]
        mhz = sorted(s[1] for s in used)
        mask = 0
        for s in used:
            mask |= s[2]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": self.max_mhz,
                "reasons": [name for bit, name in self.REASONS.items() if mask & bit],
                "samples": len(used), "window": "timed" if used is timed else "all load phases"}


# ----------------------------------------------------------------------------- reference arm
def reference_rate(threads: int, paths_per_thread: int, chunk: int = 1 << 20):
    """The unmodified reference CPU pricer on `threads` host threads (ctypes releases the GIL).
    Calls of 2^20 paths each: its single float accumulator is only valid to ~1e6-1e7 paths per
    call.  Returns (paths/s, seconds, mean price, kind)."""
    import ctypes as C
    import oracle

    o = oracle.option(N_PATHS=chunk, **CFG)
    if oracle.have_ref():
        ref = oracle.ref_cpu()
        kind = "reference"

        def work(out, i):
            done = C.c_uint64()
            out[i] = ref.ref_vanilla_cpu_chunked(C.byref(o), paths_per_thread, chunk, C.byref(done))
    else:  # oracle/_ref did not travel: fall back to the plain-C port (still CPU, still not the product)
        kind = "port"

        def work(out, i):
            s, _ = oracle.european(o, i * paths_per_thread, paths_per_thread, SEED, oracle.CALL)
            out[i] = oracle.price_from_sum(s, paths_per_thread, o.r, o.T)

    out = [0.0] * threads
    ts = [threading.Thread(target=work, args=(out, i)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    return threads * paths_per_thread / dt, dt, sum(out) / threads, kind


def reference_gpu_baseline(n_paths: int = 1 << 26):
    """The UNMODIFIED reference GPU wrappers (oracle/_ref/ref_gpu: inc/wrappers.cuh:33-93 compiled for
    sm_100 in the build container) timed on this GPU as a caller experiences them -- cudaMalloc of
    the XORWOW states, setup_kernel, pricing kernel, sync, copy, free.  A reported baseline like
    cpu_baseline; None when the executable did not travel."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_gpu")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe, str(n_paths), str(CFG["r"])], capture_output=True, text=True, timeout=120)
        price = [ln.split() for ln in out.stdout.splitlines() if ln.startswith("REFGPU ")][0]
        tm = [ln.split() for ln in out.stdout.splitlines() if ln.startswith("REFGPU_TIME")][0]
        n = int(tm[1])
        return {"kind": "reference", "what": "wrapper_gpu_option_vanilla / wrapper_gpu_bullet_option (100 steps), wall "
                "clock of the whole wrapper call", "n_paths": n, "european_paths_per_s": n / float(tm[2]),
                "bullet_path_steps_per_s": n * 100 / float(tm[3]), "european_price": float(price[2]),
                "bullet_price": float(price[3])}
    except Exception as exc:
        return {"error": repr(exc)}


def reference_bullet_rate(threads: int, paths_per_thread: int = 1 << 17):
    """The unmodified reference CPU bullet pricer (simulateBulletOptionPriceCPU, inc/tool.cuh:133-173),
    100 steps, on `threads` host threads: path-steps/s.  None without oracle/_ref."""
    import ctypes as C
    import oracle
    if not oracle.have_ref():
        return None
    ref = oracle.ref_cpu()
    o = oracle.option(N_PATHS=paths_per_thread, N_STEPS=100, B=120.0, P1=10, P2=50, **CFG)
    out = [0.0] * threads

    def work(i):
        out[i] = ref.ref_bullet_cpu(C.byref(o))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    return {"path_steps_per_s": threads * paths_per_thread * 100 / dt, "cores": threads, "price": sum(out) / threads,
            "sample": f"{threads} threads x {paths_per_thread} paths x 100 steps ({dt:.1f} s)"}


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    oracle.lib()
    threads = host_threads()
    # calibrate on one thread, then size a step so the whole run stays within ~2 minutes
    rate1, _, _, kind = reference_rate(1, 1 << 21)
    budget = min(1.5, 110.0 / max(1, args.steps + args.warmup))
    per_thread = max(1 << 20, int(rate1 * budget) >> 20 << 20)
    for _ in range(args.warmup):
        reference_rate(threads, per_thread)
    t_total, price = 0.0, 0.0
    for _ in range(args.steps):
        _, dt, p, kind = reference_rate(threads, per_thread)
        t_total += dt
        price += p
    paths = threads * per_thread
    value = paths * args.steps / t_total
    sample = f"{threads} threads x {per_thread} paths per step in calls of 2^20 (of the 2^30-path workload)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, args.scaling),
        "reference_note": "CPU simulateOptionPriceCPU (inc/tool.cuh:104-130), unmodified, on every host core: each step "
                          f"prices a bounded SAMPLE of the workload ({paths} paths; the rate metric does not depend on "
                          "the sample size)",
        "sample_paths_per_step": paths,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                         "value_1core": rate1},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "price": price / args.steps, "closed_form": bs_call(**CFG),
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------- own arm
def hexbits(x: float) -> str:
    return float(x).hex()


def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N>1 with torch.distributed.run (one process per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: this engine has no CPU fallback")
    torch.cuda.set_device(local)
    # stdout is reserved for the ONE JSON line: libraries that print there (NCCL's version banner does)
    # are sent to stderr at the file-descriptor level, the line goes out through the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")     # host-side barriers that keep the GPUs idle

    pkg = entry.load_package()
    import importlib
    sharded = importlib.import_module(entry.PKG_NAME + ".sharded")
    eng = pkg.Engine(local)
    pricer = sharded.ShardedPricer(eng, transport=args.transport)
    transport = pricer.transport
    hbm_gbs, sm_max_mhz, peak_src = measured_peaks()

    strong = args.scaling == "strong"
    n_total = PATHS_PER_GPU if strong else PATHS_PER_GPU * world
    opt = pkg.option(N_PATHS=PATHS_PER_GPU, **CFG)
    sampler = ClockSampler(local)
    sampler.start()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_over_ranks(x):
        if world == 1:
            return [x]
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = x
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def timed_jobs(n_paths, steps):
        """Device time (ms, max over ranks) of `steps` jobs of n_paths submitted back to back."""
        if transport == "peer":
            barrier()
            eng.pipeline_timer_start()
            for _ in range(steps):
                pricer.european_async(opt, n_paths, SEED, pkg.CALL)
            ms = eng.pipeline_timer_stop()
        else:
            # NCCL transport: kernels, the all-reduce and the events all on ONE torch stream
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                pricer.european_async(opt, n_paths, SEED, pkg.CALL)
            e1.record()
            torch.cuda.current_stream().synchronize()
            ms = float(e0.elapsed_time(e1))
        barrier()
        return max_over_ranks(ms)

    # ---- device-timed leg -------------------------------------------------------------
    torch.cuda.set_stream(pricer.stream)
    sampler.phase = "warmup"
    warm = max(args.warmup, 3)
    for _ in range(warm):
        pricer.european_async(opt, n_total, SEED, pkg.CALL)
    pricer.european_result()
    barrier()
    # Per-launch CUDA events bracket the pricing kernel.  With one GPU the jobs run back to back on one
    # stream and the events are recorded inside the timed region.  With several GPUs consecutive jobs
    # OVERLAP on two pricing streams (job e + 1 fills the SMs that job e's last wave leaves idle), so an
    # event pair would also time the wait for the previous job: there the per-launch duration is taken in
    # the one-job-at-a-time leg below and the roofline uses the timed region / jobs (an upper bound).
    overlapped = world > 1 and transport == "peer"
    eng.timing_read(pkg.KERNEL_EUROPEAN)
    eng.timing_enable(not overlapped)
    launches0 = eng.launch_count
    sampler.phase = "timed"
    ms_total = timed_jobs(n_total, args.steps)
    sampler.phase = "post"
    launches = eng.launch_count - launches0
    eng.timing_enable(False)
    kern_ms, kern_n = eng.timing_read(pkg.KERNEL_EUROPEAN)
    result = pricer.european_result()
    assert result.n_paths == n_total
    if not overlapped:
        assert 0.2 < kern_ms / max(ms_total, 1e-9) <= 1.0 + 1e-3 and kern_n == args.steps, \
            f"kernel events ({kern_ms:.3f} ms / {kern_n}) do not fit inside the timed region ({ms_total:.3f} ms)"
    value = n_total * args.steps / (ms_total * 1e-3)

    # one job at a time (submit, wait for the result, submit the next): what a latency-bound caller sees
    sampler.phase = "latency"
    barrier()
    eng.timing_enable(True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pricer.price_european(opt, n_total, SEED, pkg.CALL)
    single_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    eng.timing_enable(False)
    solo_ms, solo_n = eng.timing_read(pkg.KERNEL_EUROPEAN)
    rank_kernel_ms = gather_over_ranks(solo_ms / max(solo_n, 1))
    if overlapped:
        kern_ms, kern_n = ms_total, args.steps

    # ---- end-to-end leg: the public synchronous C-ABI call, host in / host out ----------------
    # N = 1: this engine.  N > 1: rank 0 drives ONE engine over all N GPUs (mcb_engine_create_multi,
    # the route a C caller of wrapper_gpu_option_vanilla takes); the other ranks wait on the CPU.
    sampler.phase = "e2e"
    barrier()
    e2e_value = e2e_s = None
    e2e_res = None
    multi_sweep = None
    if rank == 0:
        e2e_eng = eng if world == 1 else pkg.Engine(list(range(world)))
        for _ in range(3):
            e2e_eng.price_european(opt, n_total, SEED, pkg.CALL)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_res = e2e_eng.price_european(opt, n_total, SEED, pkg.CALL)
        e2e_s = time.perf_counter() - t0
        e2e_value = n_total * args.steps / e2e_s
        if world > 1 and not args.headline_only:
            # BASELINE configs[4] through the SAME single-process engine (synchronous C-ABI call, host arrays in, host
            # results out; the segments of the other shards cross NVLink as peer stores from segment_sets_kernel)
            import numpy as np
            K, V = np.meshgrid(np.linspace(60, 140, 32, dtype=np.float32), np.linspace(0.05, 0.8, 32, dtype=np.float32),
                               indexing="ij")
            k, v = K.ravel().copy(), V.ravel().copy()
            e2e_eng.price_sweep(opt, k, v, 1 << 26, SEED, pkg.CALL)
            t0 = time.perf_counter()
            for _ in range(3):
                sw = e2e_eng.price_sweep(opt, k, v, 1 << 26, SEED, pkg.CALL)
            dt = (time.perf_counter() - t0) / 3
            multi_sweep = {"ms": 1e3 * dt, "path_params_per_s": 1024 * (1 << 26) / dt,
                           "price_hex": hexbits(sw[16 * 32 + 6].price),
                           "api": f"mcb_price_sweep on mcb_engine_create_multi({world} GPUs), rank 0, synchronous"}
        if world > 1:
            e2e_eng.close()
    if world > 1:
        dist.barrier(group=cpu_group)
    # bytes: OptionData + Philox round keys + job arguments travel as kernel parameters (once per shard);
    # the result (40 B + its 8-byte sequence word) comes back through mapped pinned memory, written by the kernel
    h2d = (48 + 80) * world
    d2h = 48

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, args.scaling),
        "engine": {"rng": f"seed {SEED}, Philox4x32-10 keyed by (seed, path id)",
                   "parallelism": f"path-index shards x{world}, " + (
                       "one NCCL allreduce of 1 KiB" if transport == "nccl" else
                       "one pricing launch per GPU that also folds its segments and stores them into every "
                       "rank's mailbox over NVLink (CUDA IPC), final tree on a second stream"),
                   "transport": transport,
                   "jobs_in_flight": "up to 4 (back-to-back submits" + (", consecutive jobs overlap on two pricing streams)"
                                                                        if world > 1 else ")")},
        "gpu_launches": int(launches),
        "price": result.price, "std_error": result.std_error, "closed_form": bs_call(**CFG),
        "z_score": (result.price - bs_call(**CFG)) / result.std_error,
        "price_hex": hexbits(result.price), "sum_hex": hexbits(result.sum), "sumsq_hex": hexbits(result.sumsq),
        "single_job_latency": {"ms": single_ms, "paths_per_s": n_total / (single_ms * 1e-3),
                               "what": "submit one job, wait for its result on the host, then the next"},
        "rank_kernel_ms": rank_kernel_ms,
    }
    if rank == 0:
        line["e2e"] = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "api": "mcb_price_european" + ("" if world == 1 else f" on mcb_engine_create_multi({world} GPUs), rank 0"),
                       "ms_per_step": 1e3 * e2e_s / args.steps}
        line["e2e_price_bits_equal"] = bool(e2e_res.sum == result.sum and e2e_res.sumsq == result.sumsq
                                            and e2e_res.price == result.price)
    if kern_n:
        per_launch_s = kern_ms * 1e-3 / kern_n
        paths_per_launch = n_total / world  # this rank's shard
        achieved = INSTR_PER_EUROPEAN_PATH * paths_per_launch / per_launch_s / 1e12
        sms = eng.device_info().sm_count
        peak = sms * ISSUE_PER_CLK_PER_SM * sm_max_mhz * 1e6 / 1e12
        line["roofline"] = {
            "bound": "issue", "kernel": "european_kernel (shards of >= 2^26 paths; european_job_kernel below)",
            "achieved": achieved, "peak": peak, "unit": "Tinstr/s",
            "frac": achieved / peak, "traffic": ncu_traffic("european_kernel"),
            "traffic_note": "DRAM bytes per launch from ncu --set full (profiles/*_ncu_traffic.json); the kernel "
                            "reads no global memory, algorithmic bytes = 512 KiB of chunk partials (stay in L2)",
            "per_unit": f"{INSTR_PER_EUROPEAN_PATH} thread-instr/path (SURVEY 8(d))",
            "frac_executed": EXECUTED_INSTR_PER_EUROPEAN_PATH / INSTR_PER_EUROPEAN_PATH * achieved / peak,
            "per_unit_executed": f"{EXECUTED_INSTR_PER_EUROPEAN_PATH} thread-instr/path executed by the SASS (ncu)",
            "peak_how": f"{sms} SMs x {ISSUE_PER_CLK_PER_SM} thread-instr/clk x {sm_max_mhz:.0f} MHz ({peak_src} sm_max_mhz)",
            "kernel_ms": 1e3 * per_launch_s, "kernel_launches": kern_n,
            "kernel_ms_how": ("timed region / jobs: consecutive jobs overlap on two pricing streams, an upper bound of "
                              "the per-launch duration (rank_kernel_ms has the solo launches)") if overlapped else
                             "CUDA events around every launch inside the timed region",
            "kernel_paths_per_s": paths_per_launch / per_launch_s,
            "kernel_share_of_step": kern_ms / kern_n / (ms_total / args.steps),
        }

    # ---- N > 1 extras: the other scaling mode, and BASELINE configs[4] across the ranks ----
    if world > 1 and not args.headline_only:
        sampler.phase = "extra"
        other_n = PATHS_PER_GPU * world if strong else PATHS_PER_GPU
        for _ in range(2):
            pricer.european_async(opt, other_n, SEED, pkg.CALL)
        ms_other = timed_jobs(other_n, max(3, args.steps // 2))
        res_other = pricer.european_result()
        line["other_workloads"] = {("weak" if strong else "strong") + "_scaling_european": {
            "paths_per_step": other_n, "paths_per_s": other_n * max(3, args.steps // 2) / (ms_other * 1e-3),
            "ms_per_step": ms_other / max(3, args.steps // 2), "price": res_other.price,
            "std_error": res_other.std_error}}
        import numpy as np
        K, V = np.meshgrid(np.linspace(60, 140, 32, dtype=np.float32), np.linspace(0.05, 0.8, 32, dtype=np.float32),
                           indexing="ij")
        k, v = K.ravel().copy(), V.ravel().copy()
        for _ in range(2):
            pricer.sweep_async(opt, k, v, 1 << 26, SEED, pkg.CALL)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        s0.record()
        for _ in range(reps):
            pricer.sweep_async(opt, k, v, 1 << 26, SEED, pkg.CALL)
        s1.record()
        barrier()
        t_sweep = max_over_ranks(float(s0.elapsed_time(s1))) * 1e-3 / reps
        res = pricer._fetch(1024)
        if multi_sweep:
            line["other_workloads"]["sweep_1024x2^26_one_engine_e2e"] = multi_sweep
        line["other_workloads"]["sweep_1024x2^26_sharded"] = {
            "path_params_per_s": 1024 * (1 << 26) / t_sweep, "ms": 1e3 * t_sweep, "scaling": "strong",
            "collective": "one NCCL all-reduce of 1024 x 64 x 2 doubles (1 MiB)",
            "price_K100_v0.2ish": res[16 * 32 + 6].price, "price_hex": hexbits(res[16 * 32 + 6].price),
            "bound_per_gpu": "XU: 4.65e12 (path.set)/s"}

    # ---- other configs of BASELINE.json (rank 0, N = 1 only): bounded, each device-timed ----
    if world == 1 and not args.headline_only:
        sampler.phase = "extra"
        try:
            line["roofline_trajectory"], line["other_workloads"] = other_workloads(torch, pkg, eng, hbm_gbs, peak_src)
        except Exception as exc:  # never lose the headline line to a secondary leg
            line["other_workloads_error"] = repr(exc)

    sampler.stop()
    sampler.join(timeout=1.0)
    clocks = sampler.summary()
    line["clocks"] = clocks
    if clocks.get("sm_mhz") and "roofline" in line:
        rf = line["roofline"]
        rf["frac_at_sampled_clock"] = rf["achieved"] / (rf["peak"] * clocks["sm_mhz"] / sm_max_mhz)

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        threads = host_threads()
        rate1, _, _, kind = reference_rate(1, 1 << 22)
        per_thread = max(1 << 20, int(rate1 * 12.0) >> 20 << 20)   # ~12 s of CPU work per thread
        rate, dt, price, kind = reference_rate(threads, per_thread)
        line["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{threads} threads x {per_thread} paths ({dt:.1f} s) of the 2^30-path workload, "
                      f"calls of 2^20 paths", "value_1core": rate1, "price": price,
            "gpu_over_cpu": {"e2e_over_all_cores": e2e_value / rate, "e2e_over_1_core": e2e_value / rate1,
                             "note": "the all-core figure floats with the box's core count; the 1-core figure "
                                     "is the reference as shipped (single-threaded)"},
        }
        bullet_cpu = reference_bullet_rate(threads)
        if bullet_cpu:
            line["cpu_baseline"]["bullet"] = bullet_cpu

    # ---- the reference's own GPU wrappers on this same GPU (rank 0, N = 1 only; context, not a target) ----
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        ref_gpu = reference_gpu_baseline()
        if ref_gpu:
            line["reference_gpu_baseline"] = ref_gpu

    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    os.close(json_fd)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def other_workloads(torch, pkg, eng, hbm_gbs, peak_src):
    """configs[2] trajectories (HBM roofline), configs[3] nested MC, configs[4] sweep, bullet."""
    stream = torch.cuda.current_stream().cuda_stream
    out = {}

    def timed(fn, reps, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3 / reps

    # configs[0]: 1e6 paths (and hello.cu's own 1e5, hello.cu:14) through the synchronous public call
    # (host in / host out): call latency.  ONE launch, result through mapped pinned memory.
    for label, npaths in (("european_1e6_sync_call", 1_000_000), ("european_1e5_sync_call_hello_cu_size", 100_000)):
        o0 = pkg.option(N_PATHS=npaths, **CFG)
        for _ in range(500):
            r0 = eng.price_european(o0, 0, SEED, pkg.CALL)
        t0 = time.perf_counter()
        for _ in range(2000):
            r0 = eng.price_european(o0, 0, SEED, pkg.CALL)
        dt0 = (time.perf_counter() - t0) / 2000
        out[label] = {"us_per_call": 1e6 * dt0, "paths_per_s": npaths / dt0, "price": r0.price,
                      "std_error": r0.std_error, "closed_form": bs_call(**CFG),
                      "z_score": (r0.price - bs_call(**CFG)) / r0.std_error, "launches_per_call": 1,
                      "note": "through the Python binding (ctypes adds ~2 us per call); the same call from C is "
                              "c_abi_sync_call_us below"}
    # the same synchronous call from plain C (examples/price_c.c, built by build()): no interpreter in the loop
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "build", "price_c")
    if os.path.exists(exe):
        try:
            txt = subprocess.run([exe, str(torch.cuda.current_device())], capture_output=True, text=True, timeout=120).stdout
            lat = {}
            for line in txt.splitlines():
                if line.startswith("C_LAT"):
                    key, us = line.split("mcb_price_european(")[1].split(" paths")[0], float(line.split(": ")[1].split(" us")[0])
                    lat[key] = min(us, lat.get(key, us))     # (the small sizes are timed twice: the better window)
            if lat:
                out["c_abi_sync_call_us"] = {"paths_to_us": lat, "launches_per_call": 1,
                                             "what": "mcb_price_european from C99 (examples/price_c.c), host in / host out"}
        except Exception as exc:   # a missing example binary must not take the bench line down
            out["c_abi_sync_call_us"] = {"error": repr(exc)}

    # the optional second keying of SURVEY 8(d): path p <-> normal p & 3 of subsequence p >> 2 (four paths per Philox
    # block).  Reported separately; the headline above is the canonical keying.
    op = pkg.option(N_PATHS=PATHS_PER_GPU, **CFG)
    for _ in range(3):
        rp = eng.price_european_packed(op, PATHS_PER_GPU, SEED, pkg.CALL)
    eng.timing_read(pkg.KERNEL_EUROPEAN_PACKED)
    eng.timing_enable(True)
    t0 = time.perf_counter()
    for _ in range(10):
        rp = eng.price_european_packed(op, PATHS_PER_GPU, SEED, pkg.CALL)
    dtp = (time.perf_counter() - t0) / 10
    eng.timing_enable(False)
    pk_ms, pk_n = eng.timing_read(pkg.KERNEL_EUROPEAN_PACKED)
    out["european_2^30_packed_keying"] = {
        "paths_per_s": PATHS_PER_GPU / dtp, "ms_per_call": 1e3 * dtp, "kernel_ms": pk_ms / max(pk_n, 1),
        "price": rp.price, "std_error": rp.std_error, "z_score": (rp.price - bs_call(**CFG)) / rp.std_error,
        "bound": "XU: 3 MUFU per path -> 16/3 paths/clk/SM = 1.55e12 /s",
        "frac": PATHS_PER_GPU / (pk_ms / max(pk_n, 1) * 1e-3) / (148 * 16 / 3 * 1.965e9),
        "note": "synchronous public call mcb_price_european_packed (launch + sync included in paths_per_s)"}

    # configs[2]: 2^20 paths x 252 steps stored path-major to HBM (1.06 GB per launch > 126 MB L2)
    opt = pkg.option(N_STEPS=TRAJ_STEPS, N_PATHS=TRAJ_PATHS, B=120.0, **CFG)
    buf = torch.empty(TRAJ_PATHS * TRAJ_STEPS, dtype=torch.float32, device="cuda")
    eng.timing_read(pkg.KERNEL_TRAJECTORY)
    eng.timing_enable(True)
    t = timed(lambda: eng.trajectories_async(opt, 0, TRAJ_PATHS, SEED, buf.data_ptr(), None, stream), 50)
    eng.timing_enable(False)
    kms, kn = eng.timing_read(pkg.KERNEL_TRAJECTORY)
    nbytes = BYTES_PER_TRAJECTORY_STEP * TRAJ_PATHS * TRAJ_STEPS
    gbs = nbytes / (kms * 1e-3 / kn) / 1e9
    roofline_traj = {"bound": "hbm", "kernel": "trajectory_slab_kernel", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s",
                     "frac": gbs / hbm_gbs, "frac_of_8TBs_nominal": gbs / 8000.0,
                     "traffic": ncu_traffic("trajectory_slab_kernel"), "algorithmic_bytes": nbytes,
                     "peak_how": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})",
                     "per_unit": "4 B stored per path-step", "kernel_ms": kms / kn, "kernel_launches": kn,
                     "l2": "output 1.06 GB per launch, larger than L2"}
    out["trajectories_2^20x252"] = {"path_steps_per_s": TRAJ_PATHS * TRAJ_STEPS / t, "GB_per_s": nbytes / t / 1e9,
                                    "ms": 1e3 * t}
    # the reference's simulate_outer_trajectories (inc/trajectories.cuh:273-351) stores the barrier count next
    # to every price: 8 B per path-step
    cnt = torch.empty(TRAJ_PATHS * TRAJ_STEPS, dtype=torch.int32, device="cuda")
    eng.timing_read(pkg.KERNEL_TRAJECTORY)
    eng.timing_enable(True)
    t2 = timed(lambda: eng.trajectories_async(opt, 0, TRAJ_PATHS, SEED, buf.data_ptr(), cnt.data_ptr(), stream), 30)
    eng.timing_enable(False)
    kms2, kn2 = eng.timing_read(pkg.KERNEL_TRAJECTORY)
    gbs2 = 2 * nbytes / (kms2 * 1e-3 / kn2) / 1e9
    roofline_traj["with_counts"] = {"achieved": gbs2, "unit": "GB/s", "frac": gbs2 / hbm_gbs,
                                    "frac_of_8TBs_nominal": gbs2 / 8000.0, "algorithmic_bytes": 2 * nbytes,
                                    "per_unit": "8 B stored per path-step (float price + int barrier count)",
                                    "kernel_ms": kms2 / kn2, "kernel_launches": kn2}
    out["trajectories_with_counts_2^20x252"] = {"path_steps_per_s": TRAJ_PATHS * TRAJ_STEPS / t2,
                                                "GB_per_s": 2 * nbytes / t2 / 1e9, "ms": 1e3 * t2}
    del buf, cnt

    # pre-generated-normal pricing: 2^20 paths x 252 normals read from HBM (1.06 GB)
    import ctypes as C
    zbuf = torch.empty(TRAJ_PATHS * TRAJ_STEPS, dtype=torch.float32, device="cuda")
    pay = torch.empty(TRAJ_PATHS, dtype=torch.float32, device="cuda")
    lib = pkg.load_library()
    assert lib.mcb_generate_normals(eng._h, SEED, TRAJ_PATHS * TRAJ_STEPS, zbuf.data_ptr(), pkg.DEVICE) == 0
    pre = lambda: lib.mcb_price_from_normals(eng._h, C.byref(opt), zbuf.data_ptr(), TRAJ_PATHS, TRAJ_STEPS,
                                             pay.data_ptr(), pkg.DEVICE)
    for _ in range(3):
        assert pre() == 0
    t0 = time.perf_counter()
    for _ in range(20):
        pre()
    tp = (time.perf_counter() - t0) / 20
    out["pregenerated_normals_2^20x252"] = {"GB_per_s_read": nbytes / tp / 1e9, "ms": 1e3 * tp,
                                            "frac_of_hbm": nbytes / tp / 1e9 / hbm_gbs,
                                            "note": "synchronous call (launch + sync included), 4 B read per path-step"}
    del zbuf, pay

    # bullet option, 2^24 paths x 100 steps (hello.cu parameters, r as configs[0]); 16 384 CTAs = 18 waves
    ob = pkg.option(N_STEPS=100, N_PATHS=1 << 24, B=120.0, P1=10, P2=50, **CFG)
    seg = torch.zeros(2 * pkg.SEGMENTS, dtype=torch.float64, device="cuda")
    t = timed(lambda: eng.bullet_segments_async(ob, 1 << 24, SEED, 0, 0.0, 0, 0, 1, seg.data_ptr(), stream), 10)
    # fmaheavy: IMAD.WIDE per path-step at the measured 26.7 /clk/SM (profiles/pipe_microbench_r1.txt)
    walk_bound = 148 * IMAD_WIDE_PER_CLK_PER_SM * 1.965e9 / WALK_IMAD_WIDE_PER_STEP
    out["bullet_2^24x100"] = {"path_steps_per_s": (1 << 24) * 100 / t, "ms": 1e3 * t,
                              "bound": f"fmaheavy: {WALK_IMAD_WIDE_PER_STEP} IMAD.WIDE per path-step at "
                                       f"{IMAD_WIDE_PER_CLK_PER_SM}/clk/SM (measured) = {walk_bound:.3g} /s",
                              "frac": (1 << 24) * 100 / t / walk_bound,
                              "frac_of_issue_bound_2.0e12": (1 << 24) * 100 / t / 2.0e12}

    # configs[3]: nested MC 4096 outer x 4096 inner x 100 steps = 8.30e10 inner path-steps
    on = pkg.option(N_STEPS=100, N_PATHS=4096, N_PATHS_INNER=4096, B=120.0, P1=10, P2=50, **CFG)
    F = torch.empty(4096 * 100, dtype=torch.float32, device="cuda")
    Cn = torch.empty(4096 * 100, dtype=torch.int32, device="cuda")
    Pn = torch.empty(4096 * 100, dtype=torch.float32, device="cuda")
    t = timed(lambda: eng.nested_async(on, 0, 4096, 1234, 1235, pkg.DISCOUNT_COMPAT, F.data_ptr(), Pn.data_ptr(),
                                       Cn.data_ptr(), stream), 2, warm=1)
    inner_steps = 4096 * 4096 * sum(99 - k for k in range(100))
    # points whose barrier count already exceeds P2 are skipped (inc/nmc.cuh:53): the work actually done
    remaining = torch.arange(99, -1, -1, device="cuda", dtype=torch.float64)
    live = (Cn.view(4096, 100) <= 50).double()
    actual = float((live * remaining).sum()) * 4096
    out["nested_4096x4096x100"] = {"inner_path_steps_per_s": actual / t, "inner_path_steps": actual,
                                   "inner_path_steps_upper": inner_steps, "ms": 1e3 * t,
                                   "note": "points with count > P2 are skipped; 'upper' assumes none is",
                                   "bound": "fmaheavy, as bullet", "frac": actual / t / walk_bound,
                                   "mean_F": float(F.double().mean())}

    # configs[4]: 1024 parameter sets x 2^26 paths (whole job on this one GPU)
    import numpy as np
    K, V = np.meshgrid(np.linspace(60, 140, 32, dtype=np.float32), np.linspace(0.05, 0.8, 32, dtype=np.float32),
                       indexing="ij")
    os_ = pkg.option(**CFG)
    segs = torch.zeros(1024 * 2 * pkg.SEGMENTS, dtype=torch.float64, device="cuda")
    k, v = K.ravel().copy(), V.ravel().copy()
    t = timed(lambda: eng.sweep_segments_async(os_, k, v, 1 << 26, SEED, pkg.CALL, 0, 1, segs.data_ptr(), stream),
              3, warm=1)
    out["sweep_1024x2^26"] = {"path_params_per_s": 1024 * (1 << 26) / t, "ms": 1e3 * t,
                              "bound": "XU: one MUFU.EX2 per (path, set) -> 16/clk/SM = 4.65e12 /s",
                              "frac": 1024 * (1 << 26) / t / (148 * 16 * 1.965e9)}
    return roofline_traj, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default, BASELINE.json's metric): ONE 2^30-path job sharded over the GPUs; "
                         "weak: 2^30 paths per GPU")
    ap.add_argument("--transport", default="auto", choices=["auto", "nccl", "peer"],
                    help="N > 1: how the 64 partial-sum segments cross GPUs: direct NVLink stores into CUDA-IPC peer "
                         "mailboxes from inside the pricing kernel (peer), one NCCL all-reduce (nccl), or peer when "
                         "CUDA IPC connects on every rank and NCCL otherwise (auto, default); same bits either way")
    ap.add_argument("--headline-only", action="store_true", help="skip the other BASELINE configs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.steps < 1:
        raise SystemExit("--steps must be >= 1")
    return run_reference(args) if args.impl == "reference" else run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
