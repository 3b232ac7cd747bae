#!/usr/bin/env python
"""bench.py -- MLMC samples/sec (SPDE sample + Darcy solve) on the 3D hex Darcy MLMC workload.

Workload (BASELINE.json configs[1]): MLMC.cpp's default problem -- cube_hex 16^3 -> 8^3 -> 4^3 on [0,2]^3
(N = 17152 / 2240 / 304 dofs), lognormal field, variance 1, correlation length 0.1, eff_perm QoI, the reference's
default Krylov stopping rule (rel 1e-6, abs 1e-12, 300 iterations; CreateMLMCParameterList.hpp:58-70) -- with a
fixed sample array {1000, 3000, 6000} (levels 0, 1, 2) = 1e4 samples per GPU per step.  One "step" is one
MLMC_Manager::InitRun over that array (src/MLMC_Manager.cpp:103-179): for every sample the noise draw, the SPDE
saddle solve(s), exp, the Darcy saddle solve(s), the QoI and the moment sums.

  python bench.py [--gpus N] [--steps K] [--warmup W]         product arm (CUDA, one rank per GPU under torchrun)
  python bench.py --impl reference ...                         reference arm: the CPU path on the host cores

One JSON line on stdout (rank 0).  See the task contract for the keys.
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

METRIC = "MLMC samples/sec per level (SPDE sample + Darcy solve)"
LEVEL_SAMPLES = [1000, 3000, 6000]   # levels 0 (fine) .. 2 (coarse); "Array number of samples"
REL, ABS, MAXIT = 1e-6, 1e-12, 300   # CreateMLMCParameterList.hpp:67-69
# The headline e2e number copies every vector the reference's managers hold in host memory back to the device, every call,
# every step.  PMC_E2E_REUSE=1 (reported next to it as e2e.reuse_device_results) lets chained calls read the library's own
# previous results from device memory instead (option "cache_results": no host->device copy at all on this workload).
E2E_REUSE_ALSO = os.environ.get("PMC_E2E_REUSE", "1") != "0"
E2E_PARTS = [int(x) for x in os.environ.get("PMC_E2E_PARTS", "1,1,1").split(",")]   # sub-batches per level on the host-buffer path


def build_problem():
    from common import hex_problem
    return hex_problem(16, 3)


def stream_positions(p, samples, rank, world):
    """Absolute yarn5 positions: InitRun draws the coarsest level first (MLMC_Manager.cpp:110,140); within a level,
    rank r owns global samples [r*n, (r+1)*n) of the world*n realisations."""
    pos, base = {}, 0
    for lev in range(p["nlevels"] - 1, -1, -1):
        Ne = p["sampler"][lev].Ne
        pos[lev] = base + rank * samples[lev] * Ne
        base += world * samples[lev] * Ne
    return pos


# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=100):
        self.index = index
        self.period_ms = period_ms
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Summary of the samples that arrived in [t0, t1] (host clock: the timed region); nvidia-smi needs ~0.1 s before its
        first sample, so the sampler is started before the warm-up steps and the window is cut out afterwards.  If no sample
        falls into the window (a very short run) the nearest ones are used and the summary says so."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines, window = self.lines, "timed region"
        if t0 is not None and t1 is not None:
            inside = [l for l in self.lines if t0 <= l[0] <= t1]
            if inside:
                lines = inside
            elif self.lines:      # nearest samples (warm-up steps of the same workload run right before the region)
                lines = sorted(self.lines, key=lambda l: min(abs(l[0] - t0), abs(l[0] - t1)))[:3]
                window = "nearest samples (none fell into the timed region)"
        for _, l in lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(p, samples, threads, pos):
    """The CPU path (oracle port of the reference algorithm) on `threads` host threads; returns seconds."""
    from common import make_oracle
    o = cpu_reference_run.oracle
    if o is None:
        o = cpu_reference_run.oracle = make_oracle(p, True, REL, ABS, MAXIT)
    t0 = time.perf_counter()
    for lev in range(p["nlevels"] - 1, -1, -1):
        if samples[lev] > 0:
            o.mlmc_level(lev, samples[lev], pos[lev], nthreads=threads)
    return time.perf_counter() - t0


cpu_reference_run.oracle = None


def bounded_sample(p, threads, target_s):
    """A sample of the workload in the same 1:3:6 level proportions, sized for ~target_s of wall time (two calibration
    rounds: a tiny probe, then a ~2 s run, because the per-sample cost of the probe is dominated by set-up)."""
    smp = [max(1, threads // 4), max(1, 3 * threads // 4), max(1, 6 * threads // 4)]
    for goal in (1.0, 4.0, target_s):
        t = cpu_reference_run(p, smp, threads, stream_positions(p, smp, 0, 1))
        scale = max(1.0, goal / max(t, 1e-3))
        smp = [max(1, int(round(x * scale))) for x in smp]
    return smp


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p = build_problem()
    threads = os.cpu_count() or 1
    per_step_target = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    samples = bounded_sample(p, threads, per_step_target)
    pos = stream_positions(p, samples, 0, 1)
    for _ in range(args.warmup):
        cpu_reference_run(p, samples, threads, pos)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_run(p, samples, threads, pos)
    dt = time.perf_counter() - t0
    total = sum(samples) * args.steps
    v = total / dt
    sample_desc = (f"{samples[0]}/{samples[1]}/{samples[2]} samples on levels 0/1/2 per step (same 1:3:6 proportions as "
                   f"the 1000/3000/6000 workload), {threads} OpenMP threads over samples")
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(1),
           "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample_desc},
           "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "reference cannot be built here (needs MPI, MFEM, hypre, ParELAG, TRNG); this arm times the "
                   "oracle's C restatement of its per-sample path"}
    print(json.dumps(out), flush=True)


def workload_config(world):
    return {"workload": "MLMC.cpp SPDE sampler + Darcy, cube_hex 16^3/8^3/4^3 (N=17152/2240/304), lognormal, "
                        "corlen 0.1, variance 1, eff_perm QoI",
            "level_samples_per_gpu": LEVEL_SAMPLES, "samples_per_step": sum(LEVEL_SAMPLES) * world,
            "rel_tol": REL, "abs_tol": ABS, "max_iter": MAXIT,
            "l2_policy": "batched working set per level >> 126 MB L2 (level 0: ~2 GB of vectors per batch); no flush",
            "parallelism": f"samples sharded over {world} GPU(s), one ncclAllReduce of 3x9 sums per step through the C ABI "
                           "(pmc_allreduce_sums)"}


# ------------------------------------------------------------------------------------------------------------------
def pin_to_gpu_numa(local):
    """Bind this rank's host threads (and therefore its page-locked buffers, which are placed on first touch) to the CPU
    cores next to its GPU: 8 ranks x 3 level threads otherwise float over both sockets and half of the host<->device
    copies cross the socket interconnect.  The core set comes from NVML (nvmlDeviceGetCpuAffinity, what `nvidia-smi topo -m`
    prints), else from the GPU's sysfs numa_node.  Returns a note for the JSON line."""
    allowed = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (max(allowed) // 64) + 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1} & allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return f"rank bound to the {len(cpus)} CPUs NVML lists for GPU {local} (of {len(allowed)} allowed)"
        if cpus == allowed:
            return f"NVML lists all {len(allowed)} allowed CPUs for GPU {local}: nothing to bind"
    except Exception as e:
        nvml_err = type(e).__name__
    else:
        nvml_err = "empty mask"
    try:
        bdf = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if len(bdf.split(":")[0]) == 8:          # nvidia-smi prints an 8-digit PCI domain, sysfs uses 4
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return f"no CPU affinity information for the GPU (NVML: {nvml_err}; sysfs numa_node = -1)"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= allowed
        if not cpus:
            return f"NUMA node {node} has no allowed CPU"
        os.sched_setaffinity(0, cpus)
        return f"rank bound to the {len(cpus)} CPUs of NUMA node {node} (GPU {bdf})"
    except Exception as e:
        return f"not bound (NVML: {nvml_err}; sysfs: {type(e).__name__})"


# ------------------------------------------------------------------------------------------------------------------
SPE10_SAMPLES = [8, 32, 128, 512]      # realisations per GPU on levels 0 (fine) .. 3, one MLMC_Manager::InitRun


def run_spe10(args, rank, world, local, dist, dev):
    """BASELINE.json configs[4] as a bounded, driver-visible leg: the SPE10-size hierarchy (60 x 220 x 85 cells of
    20 x 10 x 2 ft, 4 levels, N = 4.5 M / 577 k / 75 k / 10 k, correlation length 100, the SPE10 boundary attributes --
    examples/SPE10/SPE10_MLMC.cpp with spe10_3D_parameters.xml) and one InitRun with SPE10_SAMPLES realisations per GPU
    and level (weak scaling), level after level; per level the device time of its launch (max over ranks), the
    algorithmic GB/s of the solver kernel, and the Darcy MINRES iterations per solve on 4 realisations."""
    import torch
    from common import spe10_problem, make_context
    t0 = time.perf_counter()
    p = spe10_problem(args.spe10_scale, 4)
    ctx = make_context(p, True, REL, 1e-14, 3000, device=local)
    setup_s = time.perf_counter() - t0
    nl = p["nlevels"]
    pos, base = {}, 0
    for lev in range(nl - 1, -1, -1):
        Ne = p["sampler"][lev].Ne
        pos[lev] = base + rank * SPE10_SAMPLES[lev] * Ne
        base += world * SPE10_SAMPLES[lev] * Ne
    per_level, total_ms, sums = {}, 0.0, np.zeros((nl, 9))
    for lev in range(nl - 1, -1, -1):
        ns = SPE10_SAMPLES[lev]
        ctx.mlmc_level_batch(lev, min(ns, 8), pos[lev])              # warm-up: workspace and program of the level
        if dist is not None:
            dist.barrier()
        ctx.reset_stats()
        _, _, its = ctx.mlmc_level_batch(lev, ns, pos[lev], sums=sums[lev])
        k = ctx.kernel_stats()["kernel"]
        ms = float(k["ms"])
        if dist is not None:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        # Darcy iterations per solve on this level (4 realisations of the level's own field)
        xi = ctx.sampler_sample_batch(lev, 4, pos[lev])
        kf, _, _ = ctx.sampler_eval_batch(lev, xi, want_embed=False)
        _, _, _, dits = ctx.darcy_solve_batch(lev, kf)
        per_level[f"level{lev}"] = {"N": int(p["darcy"][lev].N), "samples": ns * world, "ms": ms,
                                    "samples_per_s": ns * world / (ms * 1e-3),
                                    "kernel_gbs_per_gpu": k["algo_bytes"] / max(k["ms"] * 1e-3, 1e-12) / 1e9,
                                    "iterations_per_sample": its / ns, "darcy_its_per_solve": float(np.mean(dits))}
        total_ms += ms
    if dist is not None:
        t = torch.from_numpy(sums).to(dev)
        dist.all_reduce(t)
        sums = t.cpu().numpy()
    ctx.close()
    est = float((sums[:, 1] / (np.array(SPE10_SAMPLES) * world)).sum())
    return {"workload": "SPE10_MLMC.cpp geometry: %dx%dx%d hex cells (20x10x2 ft at scale 1), 4 levels, corlen 100, SPE10 BCs, "
                        "unit mass coefficient" % tuple(p["grid"]),
            "level_samples_per_gpu": SPE10_SAMPLES, "n_gpus": world, "rel_tol": REL, "value": sum(SPE10_SAMPLES) * world / (total_ms * 1e-3),
            "unit": "samples/s", "ms_per_initrun": total_ms, "host_setup_s": setup_s, "mlmc_estimate": est,
            "per_level": per_level, "timing": "CUDA events around each level's launch, max over ranks, levels run one after another"}


# ------------------------------------------------------------------------------------------------------------------
def run_product(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    affinity_note = pin_to_gpu_numa(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL announces its version on stdout when the first communicator is created; stdout carries the JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    from common import make_context
    from concurrent.futures import ThreadPoolExecutor
    p = build_problem()
    nl = p["nlevels"]
    # One handle (own stream, own workspace) per level: the level batches of an InitRun are independent, so they are
    # launched from one host thread each and their persistent kernels share the GPU (the finest level alone cannot
    # fill it with 1000 realisations = 250 tiles).
    ctxs = [make_context(p, True, REL, ABS, MAXIT, device=local)]
    ctxs += [ctxs[0].clone() for _ in range(nl - 1)]          # pmc_clone: same hierarchy, own stream and workspace
    pool = ThreadPoolExecutor(max_workers=nl)
    stream = torch.cuda.current_stream()
    if dist is not None:
        # the per-level sums are all-reduced by the library's own NCCL communicator, through the C ABI
        # (pmc_comm_unique_id / pmc_comm_init / pmc_allreduce_sums); torch.distributed only ships the 128-byte id and
        # provides the barrier and the max-over-ranks of the timings the bench contract asks for
        from parelagmc_b200 import capi
        box = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctxs[0].comm_init(world, rank, box[0])
    pos = stream_positions(p, LEVEL_SAMPLES, rank, world)
    dev = torch.device("cuda", local)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        """One InitRun on the device path: every level's fused loop (one kernel launch each), then one allreduce."""
        sums = np.zeros((nl, 9))
        futs = [pool.submit(ctxs[lev].mlmc_level_batch, lev, LEVEL_SAMPLES[lev], pos[lev], None, False, sums[lev])
                for lev in range(nl)]
        its = sum(f.result()[2] for f in futs)
        if dist is not None:
            ctxs[0].allreduce_sums(sums)      # ncclAllReduce of 3 x 9 doubles on the handle's stream
        return sums, its

    # ---- warm-up (the clock sampler starts here: nvidia-smi takes ~0.1 s to deliver its first sample) ----
    clocks = ClockSampler(local, args.clock_ms)
    if rank == 0 and not args.no_clocks:
        clocks.start()
    for _ in range(args.warmup):
        step()
    # ---- timed region: CUDA events on the launching stream; the library also brackets every launch of its
    #      persistent solver kernel (one per level batch, 3 per step) with events on the same stream ----
    for c in ctxs:
        c.reset_stats()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e0.record(stream)
    its_total = 0
    for _ in range(args.steps):
        sums, its = step()
        its_total += its
    e1.record(stream)
    barrier()
    t_host1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    stats = [c.kernel_stats() for c in ctxs]
    clk = clocks.stop(t_host0, t_host1) if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    total_samples = sum(LEVEL_SAMPLES) * world * args.steps
    value = total_samples / (ms * 1e-3)
    launches = int(sum(st["kernel"]["launches"] + st["kernel"]["other_launches"] for st in stats)) // max(1, args.steps)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    total_bytes = sum(x["kernel"]["algo_bytes"] for x in stats)
    n_launch = int(sum(x["kernel"]["launches"] for x in stats))
    # The three level batches of a step are three concurrent launches of the same kernel; their CUDA-event durations
    # overlap, so the honest denominator is the device time of the timed region itself (which also contains the
    # host-side gaps between steps): achieved is a lower bound.
    achieved = total_bytes / (ms * 1e-3) / 1e9
    # solo pass (outside the timed region): each level's launch alone on the GPU -> per-level throughput and the
    # bandwidth of the dominant launch without co-runners
    solo = {}
    for lev in range(nl):
        ctxs[lev].reset_stats()
        ctxs[lev].mlmc_level_batch(lev, LEVEL_SAMPLES[lev], pos[lev])
        k = ctxs[lev].kernel_stats()
        solo[lev] = k
    per_level = {f"level{lev}": LEVEL_SAMPLES[lev] * world / max(solo[lev]["kernel"]["ms"] * 1e-3, 1e-12) for lev in range(nl)}
    dom = max(range(nl), key=lambda l: solo[l]["kernel"]["algo_bytes"])
    st = solo[dom]
    kst = st["kernel"]
    classes = [n for n in st if n != "kernel"]
    # DRAM bytes of the dominant launch: counters need ncu, so they come from the committed capture of this very launch
    # (profiles/r2_traffic.json, tools/profile_tools.py) -- but only while the capture was taken from the kernel sources
    # this run is built from (hash of csrc/); a stale capture reports null instead of a number that no longer holds.
    traffic, traffic_note, frac_dram = None, None, None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from profile_tools import kernel_source_hash
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if tj.get("kernel_source_hash") == kernel_source_hash():
            traffic = float(tj["dram_bytes_read"] + tj["dram_bytes_write"])
            frac_dram = traffic / max(kst["ms"] * 1e-3, 1e-12) / 1e9 / peak
            traffic_note = (f"dram__bytes_read.sum + dram__bytes_write.sum of one level-{dom} launch, {tj['source']} (kernel source "
                            f"hash {tj['kernel_source_hash']} = this build); frac_dram = those bytes / this run's solo launch time / peak")
        else:
            traffic_note = (f"profiles/r2_traffic.json was captured from kernel sources {tj.get('kernel_source_hash')}, this build is "
                            f"{kernel_source_hash()}: stale, not reported")
    except Exception as e:
        traffic_note = f"no capture: {type(e).__name__}"
    roofline = {"bound": "hbm",
                "kernel": "k_run_program (tile-persistent solver: a whole level batch per launch; the 3 level batches of a "
                          "step run as 3 concurrent launches)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "frac_dram": frac_dram,
                "traffic_note": traffic_note,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "launches": n_launch, "algo_bytes_per_launch": total_bytes / max(1, n_launch),
                "avg_launch_us": 1e3 * sum(x["kernel"]["ms"] for x in stats) / max(1, n_launch),
                "definition": "algorithmic bytes of all launches in the timed region / CUDA-event time of the timed region "
                              "(launches overlap; lower bound).  Algorithmic bytes: 32-byte rows of the batched vectors an "
                              "operation reads and writes (DESIGN.md section 5); inside the Krylov loops only the realisations "
                              "of a tile that are still iterating are credited, and the coarsest-level Chebyshev iteration, "
                              "whose iterates live in shared memory, is credited its inputs and result once",
                "share_of_step": 1.0,
                f"solo_level{dom}_launch": {"achieved": kst["algo_bytes"] / max(kst["ms"] * 1e-3, 1e-12) / 1e9,
                                            "frac": kst["algo_bytes"] / max(kst["ms"] * 1e-3, 1e-12) / 1e9 / peak,
                                            "ms": kst["ms"], "algo_gb": kst["algo_bytes"] / 1e9,
                                            "class_cycle_share": {n: round(st[n]["cycle_share"], 4) for n in classes if st[n]["ops"]},
                                            "class_algo_gb": {n: round(st[n]["algo_bytes"] / 1e9, 3) for n in classes if st[n]["ops"]}},
                "timing": "CUDA events on the launching streams; in-kernel clock64 accounting for the class shares"}

    # ---- e2e: the same InitRun through the host-buffer plugin API (Sample / Eval / SolveFwd, batched) ----
    # The levels run concurrently (one handle and host thread each), so the host<->device copies of one level overlap the
    # solves of the others.  PMC_E2E_PARTS can cut a level's batch into sub-batches on cloned handles as well; measured
    # on this workload that only adds per-call overhead (122k -> 114k -> 82k samples/s for 1 / 2 / 3 parts of level 0).
    h2d = d2h = 0

    from parelagmc_b200.capi import pinned_empty
    Ne_l = [p["sampler"][l].Ne for l in range(nl)]
    jobs = []   # (level, first sample, count, handle, page-locked buffers of the vectors the managers hold)
    for lev in range(nl):
        parts = E2E_PARTS[lev] if lev < len(E2E_PARTS) else 1
        bounds = [LEVEL_SAMPLES[lev] * i // parts for i in range(parts + 1)]
        for i in range(parts):
            n = bounds[i + 1] - bounds[i]
            if n == 0:
                continue
            d = {"xi": pinned_empty((n, Ne_l[lev])), "s": pinned_empty((n, Ne_l[lev]))}
            if lev < nl - 1:
                d["sc"] = pinned_empty((n, Ne_l[lev + 1]))
                d["emb"] = pinned_empty((n, Ne_l[lev + 1]))
            jobs.append((lev, bounds[i], n, ctxs[lev] if i == 0 else ctxs[0].clone(), d))
    pool_e2e = ThreadPoolExecutor(max_workers=len(jobs))
    e2e_mode = {"reuse": False}

    def part_e2e(job):
        """One sub-batch of a level of InitRun through the host-buffer API (the reference managers' call sequence)."""
        lev, first, n, ctx, b = job
        p0 = pos[lev] + first * Ne_l[lev]
        hi = ho = 0
        xi = ctx.sampler_sample_batch(lev, n, p0, out=b["xi"])                # Sample(level, xi)
        ho += xi.nbytes
        # Every call writes its result to the caller's (page-locked) host vectors, as the reference's managers hold them,
        # and reads its input vectors from them (host->device copy inside the call).  In the secondary "reuse" pass a
        # vector that the library itself produced in the preceding call is not sent back to the device: the handle keeps
        # its device copy (option "cache_results").
        reuse = e2e_mode["reuse"]
        if lev == nl - 1:
            s, _, _ = ctx.sampler_eval_batch(lev, None if reuse else xi, xi_level=lev, want_embed=False, out_s=b["s"],
                                             nsamples=n)                       # Eval(level, xi, s)
            hi += 0 if reuse else xi.nbytes
            ho += s.nbytes
            q, c, _, _ = ctx.darcy_solve_batch(lev, None if reuse else s, nsamples=n)     # SolveFwd(level, s, q, c)
            hi += 0 if reuse else s.nbytes
            ho += q.nbytes
            y = q
        else:
            sc, emb, _ = ctx.sampler_eval_batch(lev + 1, None if reuse else xi, xi_level=lev, use_init=0, out_s=b["sc"],
                                                out_embed=b["emb"], nsamples=n)  # Eval(l+1, xi, s, init, false)
            hi += 0 if reuse else xi.nbytes
            ho += sc.nbytes + emb.nbytes
            qc, cc, _, _ = ctx.darcy_solve_batch(lev + 1, None if reuse else sc, nsamples=n)
            hi += 0 if reuse else sc.nbytes
            ho += qc.nbytes
            sf, _, _ = ctx.sampler_eval_batch(lev, None if reuse else xi, xi_level=lev, init_s=None if reuse else emb,
                                              init_level=lev + 1, use_init=1, want_embed=False, out_s=b["s"],
                                              nsamples=n)                      # Eval(l, xi, s, init, true)
            hi += 0 if reuse else xi.nbytes + emb.nbytes
            ho += sf.nbytes
            q, c, _, _ = ctx.darcy_solve_batch(lev, None if reuse else sf, nsamples=n)
            hi += 0 if reuse else sf.nbytes
            ho += q.nbytes
            y = q - qc
            c = c + cc
        row = [np.sum(y * y), np.sum(y), np.sum(np.abs(y)), np.sum(q * q), np.sum(q), np.sum(np.abs(q)),
               np.sum(c), np.sum(y ** 3), np.sum(y ** 4)]
        return lev, row, hi, ho

    def step_e2e():
        nonlocal h2d, d2h
        sums = np.zeros((nl, 9))
        res = [f.result() for f in [pool_e2e.submit(part_e2e, j) for j in jobs]]
        h2d = sum(r[2] for r in res)
        d2h = sum(r[3] for r in res)
        for lev, row, _, _ in res:
            sums[lev] += row
        if dist is not None:
            ctxs[0].allreduce_sums(sums)
        return sums

    e2e_steps = max(1, min(args.steps, 3))

    def time_e2e(reuse):
        nonlocal sums_e2e
        e2e_mode["reuse"] = reuse
        for j in jobs:
            j[3].set_option("cache_results", 1 if reuse else 0)
        step_e2e()
        barrier()
        e0.record(stream)
        for _ in range(e2e_steps):
            sums_e2e = step_e2e()
        e1.record(stream)
        barrier()
        ms_ = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms_], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item())
        return ms_, sum(LEVEL_SAMPLES) * world * e2e_steps / (ms_ * 1e-3), h2d, d2h

    sums_e2e = None
    reuse_rec = None
    if E2E_REUSE_ALSO:
        ms_r, v_r, h_r, d_r = time_e2e(True)
        reuse_rec = {"value": v_r, "h2d_bytes_per_step": int(h_r), "d2h_bytes_per_step": int(d_r),
                     "note": "chained calls pass NULL for vectors the library itself produced in the preceding call "
                             "(option cache_results): no host->device copy; NOT the headline"}
    ms_e, e2e_value, h2d, d2h = time_e2e(False)

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on a bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        smp = bounded_sample(p, threads, 12.0)
        tcpu = cpu_reference_run(p, smp, threads, stream_positions(p, smp, 0, 1))
        cpu = {"value": sum(smp) / tcpu, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"{smp[0]}/{smp[1]}/{smp[2]} samples on levels 0/1/2 (1:3:6 like the workload), "
                         f"{tcpu:.1f} s, {threads} OpenMP threads over samples, same tolerances"}
    # ---- configs[4]: SPE10-scale leg (bounded; its own hierarchy and handle) ----
    for j in jobs:
        if j[3] not in ctxs:
            j[3].close()
    for c in ctxs:
        c.close()
    jobs, ctxs = [], []
    spe10 = None
    if not args.no_spe10:
        try:
            spe10 = run_spe10(args, rank, world, local, dist, dev)
        except Exception as e:      # the headline line must not be lost to the secondary configuration
            spe10 = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        mean_y = sums[:, 1] / (np.array(LEVEL_SAMPLES) * world)
        out = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world),
               "per_level_samples_per_s_solo": per_level, "minres_iterations_per_step": its_total / args.steps,
               "mlmc_estimate": float(mean_y.sum()),
               "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
                       "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                       "h2d_gbs_per_gpu": h2d / max(ms_e * 1e-3 / e2e_steps, 1e-12) / 1e9,
                       "d2h_gbs_per_gpu": d2h / max(ms_e * 1e-3 / e2e_steps, 1e-12) / 1e9,
                       "reuse_device_results": reuse_rec, "cpu_affinity": affinity_note,
                       "api": "sampler_sample_batch / sampler_eval_batch / darcy_solve_batch with page-locked host buffers; levels "
                              f"concurrent (sub-batches per level: {E2E_PARTS})",
                       "mlmc_estimate": float((sums_e2e[:, 1] / (np.array(LEVEL_SAMPLES) * world)).sum())},
               "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clk, "config_spe10": spe10}
        print(json.dumps(out), flush=True)
    for j in jobs:
        if j[3] not in ctxs:
            j[3].close()
    for c in ctxs:
        c.close()
    pool.shutdown()
    pool_e2e.shutdown()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-clocks", action="store_true", help="do not poll nvidia-smi during the timed region")
    ap.add_argument("--no-spe10", action="store_true", help="skip the SPE10-scale leg (BASELINE configs[4])")
    ap.add_argument("--spe10-scale", type=float, default=1.0, help="shrink the SPE10 grid (1.0 = 60x220x85)")
    ap.add_argument("--clock-ms", type=int, default=20, help="nvidia-smi polling period (the default run times ~120 ms)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
