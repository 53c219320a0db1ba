#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric: "Hermite transform Gpts/s and
varf assembly ms; % of FP64 TC/HBM roofline").

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One JSON line on stdout (rank 0).  Headline step (workload c4 = BASELINE.json configs[3], the config BASELINE
quotes "at 1/2/4/8 GPUs"): forward + inverse 3-D Hermite transform on the 256^3 Gauss-Hermite grid at degree 255
per axis (cube set), operands resident in HBM; STRONG scaling -- the same grid at every N, outer grid axis
sharded over the ranks (hermipy_b200.dist.ShardedParityTransform: the exchange is the epilogue of a GEMM that
stores into the peers' buffers over NVLink); `value` = grid points consumed (forward) + produced (inverse) per
second.  At N > 1 the sharded coefficients of BOTH exchange modes are compared with the single-GPU plan's
(checks.c4_sharded_coef_rel_err) and the run exits non-zero above 1e-12.  The same line carries, under
"configs": c2 (BASELINE configs[1]: 2-D degree-200 transforms of a batch of 17 functions + the varf assembly
times of that config, replicas at N > 1), c5 (65 536 batched 1-D transforms, batch-sharded) and c3 (degree-400
varf).  "roofline": dominant kernel of the headline step; "e2e": the headline through the host-buffer API
(copies inside the timed region); "cpu_baseline": the reference's own CPU implementation timed on this box's
cores.  See DESIGN.md section "Measurement".
"""
import argparse
import json
import os
import signal
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C2 = dict(n=401, degree=200, batch=17, index_set="cube")
C5 = dict(n_rule=2001, degree=1000, batch=65536)
C4 = dict(n=256, degree=255)
C3 = dict(n_rule=801, degree=400)


def gauss_hermite(n):
    """Probabilists' Gauss-Hermite rule, weights summing to 1.  The rule is symmetric; scipy's eigenvalue solver
    leaves ulp-level asymmetry, which is averaged out so that x_i == -x_{n-1-i} and w_i == w_{n-1-i} bit for bit
    (the library then halves the Gram contraction of varf by parity)."""
    import scipy.special as sp
    x, w = sp.roots_hermitenorm(n)
    x, w = (x - x[::-1]) / 2, (w + w[::-1]) / 2
    return x, w / np.sqrt(2 * np.pi)


def c5_nodes():
    x, w = gauss_hermite(C5["n_rule"])
    keep = (w > 0) & (np.abs(x) <= 50)       # the reference formula is finite only here (SURVEY.md finding 4)
    return x[keep], w[keep]


def bistable(X, Y):
    return X ** 4 / 4 - X ** 2 / 2 + Y ** 4 / 4 - Y ** 2 / 2 + X * Y / 2


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def pause(self, on):
        """Stop / continue the nvidia-smi loop (our own child, by PID).  Its queries take driver locks for several
        milliseconds; one of them landing inside the ~35 ms end-to-end leg -- a host-driven copy pipeline -- moved
        that number between 4.9 and 3.9 Gpts/s from run to run.  The device-timed regions stay sampled."""
        if self.proc is not None and self.proc.poll() is None:
            try:
                os.kill(self.proc.pid, signal.SIGSTOP if on else signal.SIGCONT)
            except OSError:
                pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def ev_pair(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


class Flusher:
    """Write a buffer larger than the 126 MB L2 between timed iterations."""

    def __init__(self, torch):
        self.buf = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")

    def __call__(self):
        self.buf.zero_()


def time_steps(torch, fn, steps, warmup, flush, barrier):
    for _ in range(warmup):
        flush()
        fn()
    barrier()
    torch.cuda.synchronize()
    pairs = []
    for _ in range(steps):
        flush()
        e0, e1 = ev_pair(torch)
        e0.record()
        fn()
        e1.record()
        pairs.append((e0, e1))
    torch.cuda.synchronize()
    barrier()
    return sum(a.elapsed_time(b) for a, b in pairs)  # ms over the K steps


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def fp64_peak(torch):
    """FP64 tensor-core GEMM peak is not in MEASURED_PEAKS.json: measure cuBLAS DGEMM 8192^3 here
    (burst: best of 6; sustained: mean of a ~2 s back-to-back loop)."""
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    torch.matmul(a, b, out=c)
    best = 1e9
    for _ in range(6):
        e0, e1 = ev_pair(torch)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = ev_pair(torch)
    reps = 40
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record()
    torch.cuda.synchronize()
    flop = 2.0 * 8192 ** 3
    return flop / best * 1e-9, flop * reps / e0.elapsed_time(e1) * 1e-9


# ------------------------------------------------------------------------------------------------
# The reference's own CPU implementation (oracle/_ref = its unmodified C++ compiled where it lies).
def c4_function(x0, x1, x2):
    """the synthetic 3-D grid function of the headline (values on the tensor grid x0 x x1 x x2)"""
    return (np.exp(-(x0[:, None, None] ** 2 + x1[None, :, None] ** 2 + x2[None, None, :] ** 2) / 8)
            * (1 + 0.1 * x1[None, :, None] + 0.05 * x0[:, None, None] * x2[None, None, :]))


REF_SAMPLE = dict(n=32, degree=31, rows=3)     # config 4 is timed at n = 32, degree 31


def _ref_c4_job(args):
    """One worker: the reference's transform(), 3-D cube set, forward + inverse, on a `rows` x 32 x 32 slab of the
    32^3 Gauss-Hermite grid at degree 31 -- the reduced size BASELINE.md section 3 prescribes for this config.  The
    reference's loop (transform.cpp:115-146) does the same work for every (grid point, multi-index) pair, so the
    sample yields its pair rate; the headline size has 256^3 multi-indices per grid point.  (At the full degree one
    call spends ~30 s building the 16.8 M-entry multi-index list and string hash table, iterators.cpp:152-164,
    before the first grid point: no bounded sample at degree 255 fits a benchmark step.  The reduced sample favours
    the reference: its 262 kB coefficient vector stays in cache, the 134 MB one of the full size does not.)"""
    rows, seed = args
    from oracle import ref
    n, deg = REF_SAMPLE["n"], REF_SAMPLE["degree"]
    x, w = gauss_hermite(n)
    rng = np.random.default_rng(seed)
    lo = int(rng.integers(0, n - rows + 1))
    xs, ws = x[lo:lo + rows], w[lo:lo + rows]
    f = c4_function(xs, x, x).ravel()
    t0 = time.perf_counter()
    c = ref.transform(deg, f, [xs, x, x], [ws, w, w], True, index_set="cube")
    ref.transform(deg, c, [xs, x, x], [ws, w, w], False, index_set="cube")
    return time.perf_counter() - t0, 2 * rows * n * n * (deg + 1) ** 3      # seconds, (point, multi-index) pairs


def reference_c4_pass(cores, pool, seed, rows=None):
    """All host cores (the reference is single-threaded: one process per core), each on its own slab.  Returns
    (wall seconds, equivalent grid points of the HEADLINE size consumed + produced) -- pairs / 256^3."""
    rows = REF_SAMPLE["rows"] if rows is None else rows
    t0 = time.perf_counter()
    res = pool.map(_ref_c4_job, [(rows, seed + i) for i in range(cores)])
    return time.perf_counter() - t0, sum(r[1] for r in res) / float(C4["degree"] + 1) ** 3


def reference_sample_text(cores):
    return ("per step: %d processes (one per host core) x (forward + inverse reference transform(), 3-D cube set, on a "
            "%d x %d x %d slab of the %d^3 grid at degree %d = the reduced size of BASELINE.md section 3 for this config); "
            "the measured (grid point, multi-index) pair rate divided by the %d^3 multi-indices per grid point of the "
            "headline size -- EXTRAPOLATED, linear in pairs as the reference's loop is; at degree %d a single call "
            "spends ~30 s building its multi-index hash table before the first grid point"
            % (cores, REF_SAMPLE["rows"], REF_SAMPLE["n"], REF_SAMPLE["n"], REF_SAMPLE["n"], REF_SAMPLE["degree"],
               C4["degree"] + 1, C4["degree"]))


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on this box's host cores; same config, metric and
    unit as our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libhermite_ref.so missing"}))
        return
    cores = os.cpu_count() or 1
    with mp.get_context("spawn").Pool(cores) as pool:
        for i in range(max(args.warmup, 1)):
            reference_c4_pass(cores, pool, 100 * i, rows=1)
        t, pts = 0.0, 0
        for i in range(args.steps):
            dt, p = reference_c4_pass(cores, pool, 1000 + 100 * i)
            t += dt
            pts += p
    value = pts / t * 1e-9
    line = {
        "impl": "reference", "metric": "hermite_transform_throughput", "value": value, "unit": "Gpts/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "Gpts/s", "cores": cores, "kind": "reference",
                         "sample": reference_sample_text(cores)},
        "e2e": {"value": value, "unit": "Gpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(world):
    """identical in both arms (the driver compares them)"""
    return {"workload": "c4: 3-D Hermite transform forward+inverse, degree %d per axis (cube set), %d^3 "
                        "Gauss-Hermite grid, one function; the same grid at every N (outer grid axis sharded)"
                        % (C4["degree"], C4["n"]),
            "baseline_config": "configs[3] of BASELINE.json (configs[1], [2], [4] reported under 'configs')",
            "l2": "256 MB buffer written between timed iterations",
            "parallelism": "1 GPU" if world == 1 else "outer grid axis sharded over %d GPUs, one exchange per direction" % world}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    # stdout carries the JSON line and nothing else: anything a library prints on fd 1 (NCCL's banner and its
    # NCCL_DEBUG=INFO log go there) is sent to stderr instead, where the driver can still read it
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from hermipy_b200 import core, _lib
    from hermipy_b200.plan import Plan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = dict(torch=torch, dist=dist, core=core, lib=_lib.lib(), Plan=Plan, world=world, rank=rank, local=local,
               barrier=barrier, max_over_ranks=max_over_ranks, flush=Flusher(torch), K=args.steps, W=args.warmup,
               args=args)
    ctx["peaks"], ctx["peaks_src"] = measured_peaks()
    ctx["dgemm_burst"], ctx["dgemm_sustained"] = fp64_peak(torch)

    sampler = ClockSampler(local)
    sampler.start()
    ctx["sampler"] = sampler
    head = bench_c4(ctx)                      # the headline
    clocks = sampler.stop()                   # sampled over the headline's timed regions

    configs = {}
    if not args.skip_configs:
        configs["c2"] = bench_c2(ctx)
        configs["c5"] = bench_c5(torch, Plan, world, rank, ctx["flush"], barrier, max_over_ranks, args.steps, args.warmup,
                                 ctx["dgemm_burst"])
        if not args.skip_varf:
            torch.cuda.empty_cache()
            configs["c3"] = bench_c3(torch, Plan, world, rank, ctx["flush"], barrier, max_over_ranks, ctx["dgemm_burst"])

    cpu = None
    if world == 1 and not args.skip_cpu:      # rank 0, N = 1 only
        cpu = cpu_baseline()

    bad = [k for k, v in head["checks"].items() if v is not None and not (v <= 1e-12)]
    if rank == 0:
        line = {
            "metric": "hermite_transform_throughput", "value": head["value"], "unit": "Gpts/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world),
            "clocks": clocks, "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
            "roofline": head["roofline"], "cpu_baseline": cpu,
            "launch": head["launch"], "exchange": head["exchange"], "checks": head["checks"],
            "varf": configs.get("c2", {}).get("varf"), "configs": configs,
            "peaks": {"hbm_gbs": ctx["peaks"].get("hbm_gbs"), "hbm_source": ctx["peaks_src"],
                      "fp64_dgemm_tflops_burst": ctx["dgemm_burst"], "fp64_dgemm_tflops_sustained": ctx["dgemm_sustained"]},
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    if bad:
        sys.stderr.write("bench.py: parity checks above 1e-12: %s\n" % bad)
        sys.exit(3)


def pinned_affinity(torch):
    """One rank per GPU: the page-locked buffers and the thread that drives the copies should sit on the NUMA node the
    GPU hangs off (NVML knows which CPUs those are).  Returns (previous affinity, note)."""
    old = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(torch.cuda.current_device())
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)).encode())
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return old, "%d CPUs near the GPU (NVML)" % len(os.sched_getaffinity(0))
    except Exception as exc:                      # no NVML, no permission: run unbound
        return old, "unbound (%s)" % type(exc).__name__


def bench_c4(ctx):
    """HEADLINE.  3-D forward+inverse, 256 nodes per axis, degree 255 (cube box); one GPU: the plan; N > 1: outer grid
    axis sharded, every axis contraction split by parity, the exchange fused into a GEMM epilogue over NVLink."""
    torch, dist, core, lib, Plan = ctx["torch"], ctx["dist"], ctx["core"], ctx["lib"], ctx["Plan"]
    world, rank, K, W = ctx["world"], ctx["rank"], ctx["K"], ctx["W"]
    flush, barrier, max_over_ranks, peak = ctx["flush"], ctx["barrier"], ctx["max_over_ranks"], ctx["dgemm_burst"]
    x, w = gauss_hermite(C4["n"])
    n, P, deg = C4["n"], C4["degree"] + 1, C4["degree"]
    plan = Plan(deg, [x] * 3, [w] * 3, index_set="cube")
    g_host = c4_function(x, x, x)
    pts_per_step = 2 * n ** 3
    survey_flops = 2 * 3 * 2.0 * n ** 4          # 2 n^4 per axis and direction (sum-factorised)
    parity = bool(np.array_equal(x, -x[::-1]) and np.array_equal(w, w[::-1]))
    flops = survey_flops / 2 if parity else survey_flops      # executed: every axis contraction halved by parity
    checks, notes = {}, []
    F1 = plan.pack_grid(g_host[None])
    C1 = plan.empty_coef(1)
    G1 = plan.empty_grid(1)

    def step1():
        plan.forward(F1, C1)
        plan.backward(C1, G1)

    step1()                                       # first call per GEMM shape times candidate tile configurations
    torch.cuda.synchronize()
    sw_full = torch.from_numpy(np.sqrt(w[:, None, None] * w[None, :, None] * w[None, None, :]).ravel()).cuda()

    if world == 1:
        lib.hb_launch_count(1)
        step1()
        launches_per_step = int(lib.hb_launch_count(1))
        torch.cuda.synchronize()
        run, how = step1, "stream launches"
        try:   # a fixed launch sequence on fixed buffers: replay it as a CUDA graph (exactly the kernels counted above)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step1()
            run, how = graph.replay, "CUDA graph replay"
        except Exception as e:
            torch.cuda.synchronize()
            how = "stream launches (graph capture failed: %s)" % str(e).splitlines()[0][:60]
        total_ms = time_steps(torch, run, K, W, flush, barrier)
        run()
        checks["c4_roundtrip_rel_err"] = float((((G1 - F1) * sw_full).norm() / (F1 * sw_full).norm()).item())
        exchange = "none (one GPU)"
        # per-launch device times of the same step (events around every GEMM launch, recorded by the plan)
        plan.set_timing(True)
        acc = {}
        for _ in range(max(K, 3)):
            flush()
            plan.forward(F1, C1)
            for tag, ms, _ in plan.timings():
                acc.setdefault("fwd_axis%d" % (tag - 1), []).append(ms)
            plan.backward(C1, G1)
            for tag, ms, _ in plan.timings():
                acc.setdefault("bwd_axis%d" % (tag - 11), []).append(ms)
        plan.set_timing(False)
        stage_flops = flops / 6
        launches_info = {k: {"ms": float(np.mean(v)), "tflops_executed": stage_flops / np.mean(v) * 1e-9}
                         for k, v in acc.items()}
        gemm_ms = sum(float(np.mean(v)) for v in acc.values())
        # The six launches are two instantiations of one kernel template: five with the plain store epilogue
        # (dgemm_dmma_tma_kernel<64,128,...,EPI_STORE>), one -- the last forward stage -- with the look-up-table
        # scatter into the reference's enumeration order (<...,EPI_LUT>).  The DOMINANT kernel is the one with the
        # larger share of the step: the plain-store instantiation; achieved = flops per launch / its mean launch time.
        store_keys = [k for k in acc if k != "fwd_axis0"]
        store_ms = float(np.mean([np.mean(acc[k]) for k in store_keys]))
        lut_ms = float(np.mean(acc["fwd_axis0"])) if "fwd_axis0" in acc else None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("c4_transform_bytes_per_launch", {}).get("store")
        roofline = {"bound": "tensor", "kernel": "dgemm_dmma_tma_kernel<64,128,2,2,4,K-major,2 CTAs/SM,EPI_STORE> (the two "
                                                 "parity halves of one axis contraction per launch; %d of the 6 GEMM "
                                                 "launches of a step, %.0f %% of the step time)"
                                                 % (len(store_keys), 100 * store_ms * len(store_keys) / (total_ms / K)),
                    "achieved": stage_flops / store_ms * 1e-9, "peak": peak, "unit": "TFLOP/s",
                    "frac": stage_flops / store_ms * 1e-9 / peak, "traffic": traffic,
                    "flops_per_launch": stage_flops,
                    "frac_survey_equivalent": (2.0 if parity else 1.0) * stage_flops / store_ms * 1e-9 / peak,
                    "other_instantiation": {"kernel": "dgemm_dmma_tma_kernel<...,EPI_LUT> (last forward stage: scatter into the "
                                                      "reference's shell order)", "ms": lut_ms,
                                            "frac": None if lut_ms is None else stage_flops / lut_ms * 1e-9 / peak,
                                            "share_of_step": None if lut_ms is None else lut_ms / (total_ms / K)},
                    "flops_note": "executed flops: SURVEY 8d counts 2 n^4 per axis contraction; on the bit-symmetric "
                                  "Gauss-Hermite rule the library halves that by the parity of the Hermite functions "
                                  "(n^4 per launch); the survey-equivalent rate is twice the achieved one",
                    "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (burst; FP64 is not in %s); sustained %.1f"
                                   % (ctx["peaks_src"], ctx["dgemm_sustained"]),
                    "step_tflops_executed": flops * K / (total_ms * 1e-3) * 1e-12,
                    "step_frac": flops * K / (total_ms * 1e-3) * 1e-12 / peak,
                    "gemm_share_of_step": gemm_ms / (total_ms / K),
                    "launches": launches_info}
    else:
        from hermipy_b200.dist import ShardedParityTransform, plan_parity_tables
        plan.forward(F1, C1)                      # single-GPU coefficients, the reference's enumeration (every rank)
        tables = plan_parity_tables(plan)
        total_ms, exchange, how, launches_per_step, best = None, None, None, 0, None
        for mode in ("p2p", "nccl"):
            ok, this_ms = 1, None                 # every rank must take the same branch: agree after each attempt
            try:
                T = ShardedParityTransform([n] * 3, [P] * 3, tables, exchange=mode)
                x0 = T.local_x0()
                f = torch.zeros(T.local_grid_shape(), dtype=torch.float64, device="cuda")
                f[..., :n] = torch.from_numpy(np.ascontiguousarray(g_host[x0])).cuda()
                sw = torch.zeros_like(f)
                sw[..., :n] = torch.from_numpy(np.sqrt(w[x0][:, None, None] * w[None, :, None] * w[None, None, :])).cuda()
                c = T.forward(f)                  # first call per GEMM shape times candidate tile configurations
                gout = T.backward(c)
                lib.hb_launch_count(1)
                c = T.forward(f)
                gout = T.backward(c)
                this_launches = int(lib.hb_launch_count(1))
                # coefficients of this rank's block against the single-GPU plan, squared sums over all ranks
                valid, pos = T.set_positions(deg, "cube")
                want = C1[0][pos]
                sq = torch.stack([((c[valid] - want) ** 2).sum(), (want ** 2).sum(), c[~valid].abs().sum()])
                dist.all_reduce(sq)
                checks["c4_sharded_coef_rel_err_%s" % mode] = float((sq[0] / sq[1]).sqrt().item())
                checks["c4_sharded_pad_cells_abs_sum_%s" % mode] = float(sq[2].item())
                try:
                    graph, c, gout = T.capture(f)
                    run, this_how = graph.replay, "CUDA graph replay"
                except Exception as e:                      # capture unsupported: plain launches
                    notes.append("%s: graph capture failed (%s)" % (mode, str(e).splitlines()[0][:80]))
                    torch.cuda.synchronize()
                    state = {}

                    def run():
                        state["c"] = T.forward(f)
                        state["g"] = T.backward(state["c"])
                    run()
                    c, gout, this_how = state["c"], state["g"], "stream launches"
                this_ms = max_over_ranks(time_steps(torch, run, K, W, flush, barrier))
                run()
                torch.cuda.synchronize()
                # the timed (graph) path once more: coefficients and the global quadrature-weighted round trip
                sq = torch.stack([((c[valid] - want) ** 2).sum(), (want ** 2).sum(),
                                  (((gout - f) * sw) ** 2).sum(), ((f * sw) ** 2).sum()])
                dist.all_reduce(sq)
                checks["c4_sharded_coef_rel_err_%s_timed_path" % mode] = float((sq[0] / sq[1]).sqrt().item())
                checks["c4_roundtrip_rel_err_%s" % mode] = float((sq[2] / sq[3]).sqrt().item())
            except Exception as e:
                ok = 0
                notes.append("%s: unavailable (%s)" % (mode, str(e).splitlines()[0][:120]))
            flag = torch.tensor([ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1 and (total_ms is None or this_ms < total_ms):
                total_ms, how, launches_per_step, best = this_ms, this_how, this_launches, T
                exchange = {"p2p": "the GEMM before the exchange stores each row block straight into the destination "
                                   "rank's buffer over NVLink (symmetric memory), one device barrier per direction",
                            "nccl": "GEMM into a send buffer + one NCCL all_to_all_single per direction"}[mode]
        if total_ms is None:
            raise RuntimeError("c4 sharded transform failed: " + "; ".join(notes))
        checks["c4_sharded_coef_rel_err"] = max(v for k, v in checks.items() if k.startswith("c4_sharded_coef_rel_err_"))
        checks["c4_roundtrip_rel_err"] = max(v for k, v in checks.items() if k.startswith("c4_roundtrip_rel_err_"))
        roofline = {"bound": "tensor", "kernel": "dgemm_dmma_tma_kernel (whole sharded step, per GPU: 6 GEMM launches, "
                                                 "2 parity passes, 2 exchanges)",
                    "achieved": flops * K / (total_ms * 1e-3) * 1e-12 / world, "peak": peak, "unit": "TFLOP/s",
                    "frac": flops * K / (total_ms * 1e-3) * 1e-12 / world / peak, "traffic": None,
                    "flops_note": "executed (parity-halved) flops of the whole step / step time / GPUs",
                    "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (burst)"}
    value = pts_per_step * K / (total_ms * 1e-3) * 1e-9

    # ---------------- e2e: host buffers in, host buffers out, copies inside the timed region ----------------
    old_affinity, numa_note = pinned_affinity(torch)
    sampler = ctx["sampler"]
    if world == 1:
        Fh = torch.empty((1, n ** 3), dtype=torch.float64).pin_memory()
        Fh.copy_(torch.from_numpy(g_host.reshape(1, -1)))
        Fnp = Fh.numpy()
        Ch = torch.empty((1, plan.n_polys), dtype=torch.float64).pin_memory().numpy()
        Gh = torch.empty((1, n ** 3), dtype=torch.float64).pin_memory().numpy()
        nodes, weights = [x] * 3, [w] * 3

        def e2e_step():
            core.transform_batch(deg, Fnp, nodes, weights, True, index_set="cube", out=Ch)
            core.transform_batch(deg, Ch, nodes, weights, False, index_set="cube", out=Gh)
        api = "hermipy_b200.core.transform_batch (hb_transform_batch: host buffers, copies inside)"
        h2d = 8 * (n ** 3 + plan.n_polys)
        d2h = h2d
    else:
        T = best
        fh = torch.empty(T.local_grid_shape(), dtype=torch.float64).pin_memory()
        fh.zero_()
        fh[..., :n].copy_(torch.from_numpy(np.ascontiguousarray(g_host[T.local_x0()])))
        ch = torch.empty(T.local_coef_shape(), dtype=torch.float64).pin_memory()
        gh = torch.empty(T.local_grid_shape(), dtype=torch.float64).pin_memory()

        def e2e_step():
            T.forward_host(fh, ch)
            T.backward_host(ch, gh)
        api = "hermipy_b200.dist.ShardedParityTransform.forward_host / backward_host (each rank: its slab in, its coefficient block out)"
        h2d = world * 8 * (fh.numel() + ch.numel())
        d2h = h2d
    for _ in range(3):
        e2e_step()
    sampler.pause(True)
    try:
        barrier()
        t0 = time.perf_counter()
        e2e_reps = max(5, min(K, 20))
        for _ in range(e2e_reps):
            e2e_step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
    finally:
        sampler.pause(False)
    os.sched_setaffinity(0, old_affinity)
    if world == 1:
        checks["c4_e2e_vs_resident_coef_rel_err"] = float(np.linalg.norm(Ch[0] - C1[0, :plan.n_polys].cpu().numpy())
                                                          / np.linalg.norm(Ch[0]))
    e2e = {"value": pts_per_step * e2e_reps / e2e_s * 1e-9, "unit": "Gpts/s", "ms_per_step": e2e_s / e2e_reps * 1e3,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": api, "cpu_affinity": numa_note}
    return {"value": value, "ms": total_ms / K, "checks": checks, "roofline": roofline, "e2e": e2e,
            "gpu_launches": launches_per_step * K, "launch": how,
            "exchange": exchange + ("" if not notes else " [" + "; ".join(notes) + "]"),
            "tflops_executed": flops * K / (total_ms * 1e-3) * 1e-12, "parity_halved_contraction": parity,
            "survey_equivalent_tflops": survey_flops * K / (total_ms * 1e-3) * 1e-12}


def bench_c2(ctx):
    """BASELINE configs[1]: forward + inverse 2-D transforms of a batch of 17 functions on the 401 x 401 grid at
    degree 200 (cube set), per GPU (a single 2-D transform does not shard: replicas at N > 1), the same through the
    host-pointer C ABI, and the varf assembly of that config."""
    torch, core, lib, Plan = ctx["torch"], ctx["core"], ctx["lib"], ctx["Plan"]
    world, rank, K, W, args = ctx["world"], ctx["rank"], ctx["K"], ctx["W"], ctx["args"]
    flush, barrier, max_over_ranks = ctx["flush"], ctx["barrier"], ctx["max_over_ranks"]
    peaks, peaks_src, dgemm_burst = ctx["peaks"], ctx["peaks_src"], ctx["dgemm_burst"]
    x, w = gauss_hermite(C2["n"])
    plan = Plan(C2["degree"], [x, x], [w, w], index_set=C2["index_set"])
    B, n, P = C2["batch"], C2["n"], C2["degree"] + 1
    X, Y = np.meshgrid(x, x, indexing="ij")
    rng = np.random.default_rng(2 + rank)
    fns = [np.exp(-(X ** 2 + X * Y + Y ** 2) / 10)]
    for _ in range(B - 1):   # smooth synthetic functions: random low-order polynomial x Gaussian
        a = rng.standard_normal(6)
        fns.append((a[0] + a[1] * X + a[2] * Y + a[3] * X * Y + a[4] * X ** 2 + a[5] * Y ** 2) * np.exp(-(X ** 2 + Y ** 2) / 6))
    F_host = np.stack(fns)
    F = plan.pack_grid(F_host)
    Cf = plan.empty_coef(B)
    G = plan.empty_grid(B)

    def step():
        plan.forward(F, Cf)
        plan.backward(Cf, G)

    step()                       # first call per GEMM shape times candidate tile configurations (autotune)
    torch.cuda.synchronize()
    lib.hb_launch_count(1)
    step()
    launches_per_step = int(lib.hb_launch_count(1))
    torch.cuda.synchronize()
    run_step, how = step, "stream launches"
    try:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        run_step, how = graph.replay, "CUDA graph replay"
    except Exception as e:  # capture unsupported in this environment: plain launches
        torch.cuda.synchronize()
        how = "stream launches (graph capture failed: %s)" % str(e).splitlines()[0][:60]
    total_ms = max_over_ranks(time_steps(torch, run_step, K, W, flush, barrier))
    pts_per_step = 2 * B * n * n
    value = world * pts_per_step * K / (total_ms * 1e-3) * 1e-9

    plan.set_timing(True)
    acc, executed = {}, {}
    for _ in range(max(K, 3)):
        flush()
        plan.forward(F, Cf)
        for tag, ms, fl in plan.timings():
            acc.setdefault(("fwd", tag), []).append(ms)
            executed[("fwd", tag)] = fl
        plan.backward(Cf, G)
        for tag, ms, fl in plan.timings():
            acc.setdefault(("bwd", tag), []).append(ms)
            executed[("bwd", tag)] = fl
    plan.set_timing(False)
    # algorithmic flops per stage (2 n^2 P for the stage between the grid and the half-transformed
    # array, 2 n P^2 for the one next to the coefficients).  The inverse may run its two stages in either order
    # (last axis first, or axis 0 first): the stage that executed more flops is the 2 n^2 P one.
    big, small = 2.0 * B * n * n * P, 2.0 * B * P * n * P
    alg = {("fwd", 2): big, ("fwd", 1): small}
    first_is_big = executed.get(("bwd", 11), 1.0) >= executed.get(("bwd", 12), 0.0)
    alg[("bwd", 11)], alg[("bwd", 12)] = (big, small) if first_is_big else (small, big)
    # On the bit-symmetric Gauss-Hermite rule every axis contraction is halved by the parity of the Hermite functions
    # (round 2: at every size): the flops EXECUTED are half the SURVEY's count, and the roofline fraction is computed on
    # those (as for the headline); "tflops_algorithmic" is the survey-equivalent rate.
    parity = sum(executed.get(k, alg[k]) for k in alg) < 0.75 * sum(alg.values())
    exe = {k: (alg[k] / 2 if parity else alg[k]) for k in alg}
    launches_info = {"%s_tag%d" % k: {"ms": float(np.mean(v)), "tflops_algorithmic": alg[k] / np.mean(v) * 1e-9,
                                      "tflops_executed": exe[k] / np.mean(v) * 1e-9}
                     for k, v in acc.items() if k in alg}
    dom = max((k for k in acc if k in alg), key=lambda k: np.mean(acc[k]))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj.get("c2_transform_bytes_per_launch", {}).get("%s_tag%d" % dom, tj.get("c2_transform_dominant_bytes_per_launch"))
    roofline = {"bound": "tensor", "kernel": "dgemm_dmma_tma_kernel (%s stage, tag %d)" % dom,
                "achieved": exe[dom] / np.mean(acc[dom]) * 1e-9, "peak": dgemm_burst, "unit": "TFLOP/s",
                "frac": exe[dom] / np.mean(acc[dom]) * 1e-9 / dgemm_burst, "traffic": traffic,
                "parity_halved_contraction": bool(parity),
                "step_tflops_executed": sum(exe.values()) * K / (total_ms * 1e-3) * 1e-12,
                "step_frac": sum(exe.values()) * K / (total_ms * 1e-3) * 1e-12 / dgemm_burst,
                "step_tflops_algorithmic": sum(alg.values()) * K / (total_ms * 1e-3) * 1e-12,
                "step_frac_survey_equivalent": sum(alg.values()) * K / (total_ms * 1e-3) * 1e-12 / dgemm_burst,
                "launches": launches_info}

    # round-trip property at full size, in the quadrature-weighted L2 norm: un-weighted values at the outer nodes
    # are ill-conditioned by design (phi_200 ~ 1e82 there), for the reference as much as for this code
    step()
    sw = torch.from_numpy(np.sqrt(np.pad(np.outer(w, w), ((0, 0), (0, plan.grid_pitch - n))).ravel())).cuda()
    rt = float((((G - F) * sw).norm() / (F * sw).norm()).item())

    # ---------------- e2e: the host-pointer C ABI with pinned host buffers ----------------
    old_affinity, numa_note = pinned_affinity(torch)
    Fh = torch.empty((B, n * n), dtype=torch.float64).pin_memory()
    Fh.copy_(torch.from_numpy(F_host.reshape(B, -1)))
    Fnp = Fh.numpy()
    Ch = torch.empty((B, plan.n_polys), dtype=torch.float64).pin_memory().numpy()
    Gh = torch.empty((B, n * n), dtype=torch.float64).pin_memory().numpy()
    nodes, weights = [x, x], [w, w]

    def e2e_step():
        core.transform_batch(C2["degree"], Fnp, nodes, weights, True, index_set=C2["index_set"], out=Ch)
        core.transform_batch(C2["degree"], Ch, nodes, weights, False, index_set=C2["index_set"], out=Gh)

    for _ in range(3):                      # call 1 eager (tunes), call 2 captures the pipeline graph, call 3 replays
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_reps = max(10, min(3 * K, 30))  # ~1.1 ms each: enough of them that one host hiccup does not decide
    for _ in range(e2e_reps):
        e2e_step()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    n_polys = plan.n_polys
    e2e = {"value": world * pts_per_step * e2e_reps / e2e_s * 1e-9, "unit": "Gpts/s",
           "h2d_bytes_per_step": world * 8 * B * (n * n + n_polys), "d2h_bytes_per_step": world * 8 * B * (n * n + n_polys),
           "api": "hermipy_b200.core.transform_batch (hb_transform_batch: host buffers, copies inside)",
           "cpu_affinity": numa_note}
    # the same forward transforms through the OBJECT API a user of the reference calls -- Quad.transform_batch on
    # pageable NumPy arrays (rank 0 only; forward only: the reference has no batched eval)
    obj = None
    if rank == 0:
        import hermipy_b200 as hm
        quad = hm.Quad([x, x], [w, w])
        F_page = np.array(F_host.reshape(B, -1))          # ordinary (pageable) memory
        quad.transform_batch(F_page, C2["degree"], index_set=C2["index_set"])
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            series = quad.transform_batch(F_page, C2["degree"], index_set=C2["index_set"])
        dt = (time.perf_counter() - t0) / reps
        err = float(np.linalg.norm(series[0].coeffs - Ch[0]) / np.linalg.norm(Ch[0]))
        obj = {"value": B * n * n / dt * 1e-9, "unit": "Gpts/s (forward only)", "ms_per_call": dt * 1e3,
               "h2d_bytes_per_call": 8 * B * n * n, "d2h_bytes_per_call": 8 * B * n_polys,
               "api": "hermipy_b200.Quad.transform_batch (pageable NumPy input; factor evaluated and divided on the "
                      "device: hb_transform_weighted), list of Series out",
               "rel_diff_vs_core_transform_batch": err}
    os.sched_setaffinity(0, old_affinity)
    out = {"workload": "2-D forward+inverse, degree %d per axis (%s set), %d x %d grid, batch of %d functions per GPU"
                       % (C2["degree"], C2["index_set"], n, n, B),
           "scaling": "weak (replicas: a single 2-D transform does not shard)", "launch": how,
           "ms": total_ms / K, "gpts": value, "launches_per_step": launches_per_step, "roofline": roofline,
           "e2e": e2e, "e2e_object_api": obj, "roundtrip_rel_err": rt}
    if not args.skip_varf:
        out["varf"] = bench_c2_varf(ctx, plan, x, w, X, Y)
    if world == 1 and not args.skip_cpu:
        out["cpu"] = cpu_baseline_c2(F_host, x, w)
    return out


def bench_c2_varf(ctx, plan, x, w, X, Y):
    torch, Plan, args, K = ctx["torch"], ctx["Plan"], ctx["args"], ctx["K"]
    flush, barrier, max_over_ranks = ctx["flush"], ctx["barrier"], ctx["max_over_ranks"]
    peaks, dgemm_burst = ctx["peaks"], ctx["dgemm_burst"]
    n = C2["n"]
    varf = {}
    fV_host, fE_host = bistable(X, Y), np.exp(-3 * bistable(X, Y) / 10)
    vs = max(2, min(K, 5))
    for set_name in ("cube", "triangle"):
        vplan = plan if set_name == C2["index_set"] else Plan(C2["degree"], [x, x], [w, w], index_set=set_name)
        npol = vplan.n_polys
        out = torch.empty((npol, npol), dtype=torch.float64, device="cuda")
        fV, fE = vplan.pack_grid(fV_host[None])[0], vplan.pack_grid(fE_host[None])[0]
        ms_poly = max_over_ranks(time_steps(torch, lambda: vplan.varf(fV, "reference", out), vs, 1, flush, barrier)) / vs
        kept, path = vplan.varf_info()
        ref_poly = out.clone() if args.check else None
        # the same matrix as CSR straight from the band of retained triple products (no dense buffer, no zero-fill);
        # host-timed: the call returns nnz, i.e. it ends with a device -> host read
        vplan.varf_sparse(fV, fetch=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(vs):
            nnz_sp = vplan.varf_sparse(fV, fetch=False)
        ms_sparse = (time.perf_counter() - t0) / vs * 1e3
        ms_gram = max_over_ranks(time_steps(torch, lambda: vplan.varf(fE, "gram", out), vs, 1, flush, barrier)) / vs
        vplan.set_timing(True)
        vplan.varf(fE, "gram", out)
        tl = {tag: ms for tag, ms, _ in vplan.timings()}
        vplan.set_timing(False)
        # A is symmetric and only the m0 <= n0 half is computed and mirrored: HALF of 2*n*|I|^2;
        # |I| = P^2 (cube) or C(N+2,2) (triangle).  On a bit-symmetric node set the library also halves the
        # contraction over x0 by the parity of the Hermite functions (non-negative nodes only, against the
        # even / odd parts of f): the flops actually executed are ceil(n/2) |I|^2, and the roofline fraction
        # is computed on those.
        parity = bool(np.array_equal(x, -x[::-1]) and np.array_equal(w, w[::-1]))
        survey_flops2 = float(n) * float(npol) ** 2
        flops2 = float((n + 1) // 2 if parity else n) * float(npol) ** 2
        entry = {
            "matrix": "%d x %d dense (%.1f GB)" % (npol, npol, 8e-9 * npol ** 2),
            "assembly_ms_polynomial_f": ms_poly,
            "polynomial_f": {"mode": "reference algebra (thresholded sum over alpha: %d alpha kept, path %d = "
                                     "zero-fill + banded assembly)" % (kept, path),
                             "bound": "hbm", "achieved_gbs": 8.0 * npol ** 2 / ms_poly * 1e-6,
                             "peak_gbs": peaks.get("hbm_gbs"),
                             "frac": 8.0 * npol ** 2 / ms_poly * 1e-6 / peaks.get("hbm_gbs")},
            "assembly_ms_polynomial_f_sparse": ms_sparse,
            "polynomial_f_sparse": {"mode": "CSR emitted from the band (candidates -> radix sort -> indptr/indices/data), "
                                            "device-resident result; host-timed including the final nnz read",
                                    "nnz": int(nnz_sp), "csr_bytes": int(12 * nnz_sp + 8 * (npol + 1))},
            "assembly_ms_dense_f": ms_gram,
            "dense_f": {"mode": "Gram form, two-stage FP64 tensor-core contraction on 8x8-blocked pair tables "
                                "(persistent stage-2 kernel); symmetric half computed and mirrored; contraction "
                                "over x0 %s" % ("halved by parity (bit-symmetric nodes)" if parity else "over all nodes"),
                        "bound": "tensor", "stage2_ms": tl.get(22),
                        "algorithmic_flop_symmetric_half": survey_flops2, "executed_flop": flops2,
                        "achieved_tflops": flops2 / tl.get(22, np.nan) * 1e-9, "peak_tflops": dgemm_burst,
                        "frac": flops2 / tl.get(22, np.nan) * 1e-9 / dgemm_burst,
                        "survey_equivalent_tflops": survey_flops2 / tl.get(22, np.nan) * 1e-9,
                        "note": "bound by the scatter of misaligned 64-byte runs into the %.1f GB matrix (direct + "
                                "mirrored half): ~0.8 TB/s, the rate tools/scatter_probe.cu measures for that "
                                "pattern; see DESIGN.md 4.3" % (8e-9 * npol ** 2),
                        "scatter_tbs": 8e-12 * npol ** 2 / (tl.get(22, np.nan) * 1e-3)},
        }
        # consumer of the resident matrix: y = A x (hb_dev_matvec), HBM-bound at 8 |I|^2 bytes per product
        from hermipy_b200.resident import device_matvec
        try:
            mv_traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("matvec_c2_cube_bytes_per_launch")
        except (OSError, ValueError):
            mv_traffic = None
        xv = torch.randn(npol, dtype=torch.float64, device="cuda")
        yv = torch.empty(npol, dtype=torch.float64, device="cuda")
        ms_mv = max_over_ranks(time_steps(torch, lambda: device_matvec(out, xv, yv), 5, 2, flush, barrier)) / 5
        mv_err = float(((yv - out @ xv).norm() / yv.norm()).item())
        entry["matvec"] = {"kernel": "matvec_kernel (hb_dev_matvec)", "ms": ms_mv, "bound": "hbm",
                           "algorithmic_bytes": 8.0 * npol ** 2, "achieved_gbs": 8.0 * npol ** 2 / ms_mv * 1e-6,
                           "peak_gbs": peaks.get("hbm_gbs"),
                           "frac": 8.0 * npol ** 2 / ms_mv * 1e-6 / peaks.get("hbm_gbs"),
                           "rel_diff_vs_torch_mv": mv_err,
                           "traffic": (mv_traffic or {}).get("read") if set_name == "cube" else None}
        del xv, yv
        if args.check:
            vplan.varf(fV, "gram", out)
            entry["gram_vs_reference_algebra_rel_err_polynomial_f"] = float(((out - ref_poly).norm() / ref_poly.norm()).item())
            entry["symmetry_max_abs"] = float((out - out.T).abs().max().item())
        varf[set_name] = entry
        del out, ref_poly
    varf["assembly_ms_polynomial_f"] = varf[C2["index_set"]]["assembly_ms_polynomial_f"]
    varf["assembly_ms_dense_f"] = varf[C2["index_set"]]["assembly_ms_dense_f"]
    return varf



def bench_c5(torch, Plan, world, rank, flush, barrier, max_over_ranks, K, W, peak):
    """65536 x 1-D degree-1000 transforms, batch-sharded (no data-path collective), strong scaling."""
    from hermipy_b200.dist import split_range
    x, w = c5_nodes()
    n, P = len(x), C5["degree"] + 1
    lo, hi = split_range(C5["batch"], world, rank)
    Bl = hi - lo
    plan = Plan(C5["degree"], [x], [w])
    gen = torch.Generator(device="cuda").manual_seed(5 + rank)
    coef = torch.randn((Bl, plan.coef_stride), generator=gen, dtype=torch.float64, device="cuda")
    coef /= (1.0 + torch.arange(plan.coef_stride, device="cuda", dtype=torch.float64))
    f = plan.backward(coef)                       # rows = random series evaluated on the grid
    c = plan.empty_coef(Bl)
    g = plan.empty_grid(Bl)
    ms = max_over_ranks(time_steps(torch, lambda: (plan.forward(f, c), plan.backward(c, g)), K, W, flush, barrier)) / K
    # SURVEY 8d counts 2 B n P flops per direction; on a bit-symmetric node set the library halves that by the parity
    # of the Hermite functions (even / odd parts of f on the non-negative nodes): the roofline fraction is computed
    # on the flops actually executed
    parity = bool(np.array_equal(x, -x[::-1]) and np.array_equal(w, w[::-1]))
    survey_flops = 2 * 2.0 * C5["batch"] * n * P
    nh, Pe, Po = (n + 1) // 2, (P + 1) // 2, P // 2
    flops = 2 * 2.0 * C5["batch"] * nh * (Pe + Po) if parity else survey_flops
    return {"workload": "65536 x 1-D forward+inverse, degree 1000, %d nodes (w>0, |x|<=50 subset of the 2001-point rule)" % n,
            "scaling": "strong", "ms": ms, "gpts": 2.0 * C5["batch"] * n / ms * 1e-6, "tflops": flops / ms * 1e-9,
            "frac_of_dgemm_peak_per_gpu": flops / ms * 1e-9 / (peak * world),
            "parity_halved_contraction": parity, "survey_equivalent_tflops": survey_flops / ms * 1e-9,
            "collective": "none (batch shard)"}


def bench_c3(torch, Plan, world, rank, flush, barrier, max_over_ranks, peak):
    """BASELINE configs[2]: dense 2-D varf at degree 400 per axis, quadrature Gram form (two-stage FP64 tensor-core
    contraction) of a dense non-polynomial f.  One GPU: the triangle set (80 601^2 doubles = 52 GB).  N >= 2 GPUs:
    the cube set (160 801^2 = 207 GB), row-sharded, never gathered (each rank keeps its row block)."""
    from hermipy_b200.dist import split_range
    # the nodes of the 801-point rule the degree-400 basis reaches (|x| <= 45.1: 717 of them).  80 of their weights
    # underflow in double precision (w < 1e-308) while sqrt(w) phi_k is still O(1e-3) there: the plan gets sqrt(w) in
    # extended range, without which the rows above degree ~290 come out wrong
    from hermipy_b200.lib import gram_rule
    x, w, sw = gram_rule(C3["n_rule"], C3["degree"])
    n = len(x)
    set_name = "triangle" if world == 1 else "cube"
    plan = Plan(C3["degree"], [x, x], [w, w], index_set=set_name, sqrt_weights=[sw, sw])
    npol = plan.n_polys
    lo, hi = split_range(npol, world, rank)
    X, Y = np.meshgrid(x, x, indexing="ij")
    fE = plan.pack_grid(np.exp(-3 * bistable(X, Y) / 10)[None])[0]
    free = torch.cuda.mem_get_info()[0]
    need = 8.0 * (hi - lo) * npol
    if need > free - 6e9:
        return {"workload": "2-D varf degree 400 (%s)" % set_name, "skipped": "row block of %.0f GB does not fit (%.0f GB free)"
                % (need * 1e-9, free * 1e-9)}
    out = torch.empty((hi - lo, npol), dtype=torch.float64, device="cuda")
    reps = 2
    run = lambda: plan.varf(fE, "gram", out, row_begin=lo, row_end=hi)
    ms = max_over_ranks(time_steps(torch, run, reps, 1, flush, barrier)) / reps
    # a full matrix is computed as its m0 <= n0 half and mirrored (flops n |I|^2); a row block computes every
    # owned entry (2 n rows |I|)
    parity = bool(np.array_equal(x, -x[::-1]) and np.array_equal(w, w[::-1]))
    nk = (n + 1) // 2 if parity else n           # nodes actually contracted (parity of the Hermite functions)
    survey_flops = float(n) * float(npol) ** 2 if world == 1 else 2.0 * n * float(npol) ** 2
    flops = survey_flops * nk / n
    finite = bool(torch.isfinite(out[:: max(1, (hi - lo) // 64)]).all().item())
    sym = None
    if world == 1:
        k = 4096
        sym = float((out[:k, :k] - out[:k, :k].T).abs().max().item())
    res = {"workload": "2-D Galerkin varf assembly, degree %d per axis, %s set: %d x %d dense (%.0f GB)%s, Gram form of a "
                       "dense f on %d x %d nodes (the part of the %d-point rule the basis reaches; sqrt(w) in extended range)"
                       % (C3["degree"], set_name, npol, npol, 8e-9 * npol ** 2,
                          "" if world == 1 else ", row-sharded over %d GPUs" % world, n, n, C3["n_rule"]),
           "scaling": "strong", "assembly_ms": ms, "algorithmic_tflop": survey_flops * 1e-12,
           "executed_tflop": flops * 1e-12, "parity_halved_contraction": parity,
           "achieved_tflops": flops / ms * 1e-9, "frac_of_dgemm_peak_per_gpu": flops / ms * 1e-9 / (peak * world),
           "survey_equivalent_tflops": survey_flops / ms * 1e-9,
           "finite": finite, "symmetry_max_abs_leading_4096": sym, "collective": "none (row blocks stay on their rank)"}
    del out
    torch.cuda.empty_cache()
    return res


def cpu_baseline():
    """The reference's own code (oracle/_ref) on this box's host cores, bounded sample of the headline workload
    (~15-40 s), plus a sum-factorised NumPy/BLAS restatement of the same maths as the 'fair CPU' line."""
    import multiprocessing as mp
    from oracle import port, ref
    cores = os.cpu_count() or 1
    out = None
    if ref.available():
        with mp.get_context("spawn").Pool(cores) as pool:
            reference_c4_pass(cores, pool, 7, rows=1)
            t, pts = 0.0, 0.0
            for i in range(5):
                dt, p = reference_c4_pass(cores, pool, 1000 + 100 * i)
                t, pts = t + dt, pts + p
        out = {"value": pts / t * 1e-9, "unit": "Gpts/s", "cores": cores, "kind": "reference",
               "sample": reference_sample_text(cores).replace("per step: ", "") + "; %.1f s wall" % t}
    x, w = gauss_hermite(C4["n"])
    g = c4_function(x, x, x)
    t0 = time.perf_counter()
    c = port.transform_sumfac([C4["degree"]] * 3, g[None], [x] * 3, [w] * 3, True)
    port.transform_sumfac([C4["degree"]] * 3, c, [x] * 3, [w] * 3, False)
    dt = time.perf_counter() - t0
    fair = {"value": 2 * C4["n"] ** 3 / dt * 1e-9, "unit": "Gpts/s", "cores": cores,
            "kind": "port (sum-factorised NumPy/BLAS, not the reference's algorithm)", "sample": "the full c4 step, once"}
    if out is None:
        return fair
    out["fair_cpu"] = fair
    return out


def _ref_slab_job(args):
    """One worker: the reference's own transform() on a slab of the C2 grid (rows of axis 0), forward and
    inverse.  The naive loop is linear in the number of grid points, so a slab is a faithful bounded sample."""
    rows, seed = args
    from oracle import ref
    x, w = gauss_hermite(C2["n"])
    rng = np.random.default_rng(seed)
    lo = int(rng.integers(0, C2["n"] - rows + 1))
    xs, ws = x[lo:lo + rows], w[lo:lo + rows]
    X, Y = np.meshgrid(xs, x, indexing="ij")
    f = np.exp(-(X ** 2 + X * Y + Y ** 2) / 10).ravel()
    t0 = time.perf_counter()
    c = ref.transform(C2["degree"], f, [xs, x], [ws, w], True, index_set=C2["index_set"])
    ref.transform(C2["degree"], c, [xs, x], [ws, w], False, index_set=C2["index_set"])
    return time.perf_counter() - t0, 2 * rows * C2["n"]


def cpu_baseline_c2(F_host, x, w):
    """c2 beside its GPU numbers: the reference's transform() on 3 x 401 slabs of the grid on every core, and the
    sum-factorised NumPy/BLAS port on the full step."""
    import multiprocessing as mp
    from oracle import port, ref
    cores = os.cpu_count() or 1
    out = {}
    if ref.available():
        rows = 3
        with mp.get_context("spawn").Pool(cores) as pool:
            pool.map(_ref_slab_job, [(1, i) for i in range(cores)])
            t0 = time.perf_counter()
            res = pool.map(_ref_slab_job, [(rows, 1000 + i) for i in range(cores)])
            t = time.perf_counter() - t0
        out["reference"] = {"value": sum(r[1] for r in res) / t * 1e-9, "unit": "Gpts/s", "cores": cores,
                            "sample": "%d processes x (forward + inverse reference transform() of a %d x %d slab of the "
                                      "%d x %d grid); %.1f s wall" % (cores, rows, C2["n"], C2["n"], C2["n"], t)}
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        c = port.transform_sumfac([C2["degree"]] * 2, F_host, [x, x], [w, w], True)
        port.transform_sumfac([C2["degree"]] * 2, c, [x, x], [w, w], False)
    dt = (time.perf_counter() - t0) / reps
    out["fair_cpu"] = {"value": 2 * F_host.shape[0] * C2["n"] ** 2 / dt * 1e-9, "unit": "Gpts/s", "cores": cores,
                       "kind": "port (sum-factorised NumPy/BLAS, not the reference's algorithm)", "sample": "the full c2 step"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-varf", action="store_true")
    ap.add_argument("--skip-configs", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--check", action="store_true", help="extra full-size consistency checks (slower)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
