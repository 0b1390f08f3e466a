#!/usr/bin/env python
"""bench.py -- LsqDB assembly + lsqfit solve on N B200s (one process per GPU).

A "step" is one pass of the hot path over one batch of synthetic configurations:
    K1 neighbour lists -> K2/K3 basis evaluation into the device-resident design matrix
    -> K4 weighting -> K5 Gram (DMMA) / CholeskyQR passes (+ NCCL all-reduce of the Gram
    blocks when N > 1) -> coefficients, residual.
Workload at N=1 = BASELINE.json configs[1]: 2B (deg 14) + 3B (deg 11) bond-angle basis on
2000 rattled 64-atom Si cells, energies + forces + virials  (398 000 x 216 entries).
N > 1: weak scaling -- every rank holds its own 2000 configurations (different seeds);
the fit is over all N x 2000 configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--ncfg C]

`value`  : design-matrix entries / s of the whole step, inputs resident in HBM.
`e2e`    : same metric through the public API (LsqDB(...) + lsqfit(...)) with host buffers.
`--impl reference`: the restated reference CPU path (oracle, one task per (config, okey),
all host threads + LAPACK QR) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "cfg2: 2B(deg14)+3B(deg11) bond-angle Si basis, 2000 rattled 64-atom cells, E+F+V"
WEIGHTS = {"default": {"E": 100.0, "F": 1.0, "V": 1.0}}
E0 = -4.0


def make_basis():
    import ipfitting_jl_b200 as ipf
    r0 = ipf.rnn("Si")
    rcut = 2.5 * r0
    desc = ipf.BondAngleDesc(f"exp(- (r/{r0} - 1.0))", ipf.CosCut(rcut - 1.0, rcut))
    return ipf.NBodyBasis(ipf.nbpolys(2, desc, 14) + ipf.nbpolys(3, desc, 11))


def make_configs(ncfg, seed):
    from ipfitting_jl_b200 import synth
    return synth.rattled_configs("Si", 2, ncfg, seed=seed, calc=synth.StillingerWeber())


# ---------------------------------------------------------------- algorithmic flop model
def cluster_flops(blk):
    """FP64 operations (add = mul = 1, FMA = 2) the straightforward unordered-cluster algorithm
    (oracle/ipf_oracle.c: eval_cluster) spends on ONE cluster for all functions of a block:
    values + full gradients by prefix/suffix products, chain rule onto every leg, E/F/V accumulation."""
    nx = blk.N - 1
    npair = nx * (nx - 1) // 2
    nv = nx + npair
    f = npair * 5 + nv * max(blk.maxdeg - 1, 0) + (nx - 1) + nx * max(nx - 2, 0)
    nnz = int((blk.mono_exp[:, :nv] > 0).sum())
    nm = len(blk.mono_b)
    f += nm * (2 * nv + 1) + 4 * nnz
    f += blk.nb * (2 + nx * (5 + 3 + (nx - 1) * 14 + 6 + 18))
    return float(f)


class ClockSampler:
    def __init__(self, index=0):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.index = index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        # NVML in-process (spawning nvidia-smi every 200 ms takes driver locks that stall cudaMalloc/cudaFree)
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            while not self._stop.is_set():
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for nm, b in bits.items():
                    if r & b:
                        self.reasons.add(nm)
                self._stop.wait(0.05)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.2)

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except Exception:
                pass
        return self.index

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join(timeout=10)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------- reference arm (CPU)
def cpu_reference_step(basis, configs, nthreads):
    """One pass of the restated reference CPU path over `configs`: assembly with one task per
    (config, okey) (src/obsiter.jl:87-102) on all threads, then get_lsq_system + Householder QR."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ipf_oracle as O
    import lsq_oracle as LO
    from ipfitting_jl_b200.lsq_db import _alloc_rows
    nrows = _alloc_rows(configs)
    rows = {k: np.array([d.rows[k][0] + 1 if k in d.rows else 0 for d in configs], dtype=np.int64) for k in "EFV"}
    t0 = time.perf_counter()
    Psi = O.assemble(basis, configs, rows["E"], rows["F"], rows["V"], nrows, mode=1, nthreads=nthreads)
    t1 = time.perf_counter()
    c, kappa, rel, _ = LO.lsqfit_qr(Psi, configs, WEIGHTS, E0=E0)
    t2 = time.perf_counter()
    return nrows * len(basis), t1 - t0, t2 - t1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ipf_oracle as O
    basis = make_basis()
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1 to its workers)
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    nthreads = max(O.num_threads(), ncpu)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=nthreads)        # LAPACK QR of the reference solve
    except Exception:
        pass
    cal = make_configs(2, 999)
    ent, ta, ts = cpu_reference_step(basis, cal, nthreads)
    per_cfg = (ta + ts) / 2
    budget = 150.0 / max(1, args.steps + args.warmup)
    S = int(max(2, min(args.ncfg, budget / per_cfg)))
    configs = make_configs(S, 1)
    for _ in range(args.warmup):
        cpu_reference_step(basis, configs, nthreads)
    t = 0.0
    ent = 0
    for _ in range(args.steps):
        e, ta, ts = cpu_reference_step(basis, configs, nthreads)
        t += ta + ts
        ent += e
    val = ent / t
    line = {"impl": "reference", "metric": "lsqdb_entries_per_s", "value": val, "unit": "entries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{S} of 2000 configurations per step"},
            "cpu_baseline": {"value": val, "unit": "entries/s", "cores": nthreads, "kind": "port",
                             "sample": f"{S} configurations x {args.steps} steps, assembly (1 task per (config,okey), "
                                       f"{nthreads} threads) + LAPACK QR; restated reference CPU path (Julia not installable)"},
            "e2e": {"value": val, "unit": "entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------- our arm (GPU)
def run_ours(args):
    import torch
    import ipfitting_jl_b200 as ipf
    from ipfitting_jl_b200 import _capi, parallel
    from ipfitting_jl_b200.lsq import _fix_weights, _select, collect_observations, _Solver
    from ipfitting_jl_b200.lsq_db import LsqDB, flatten_configs, _alloc_rows

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    comm = None
    saved_stdout = None
    if world > 1:
        import torch.distributed as dist
        # stdout must carry exactly one JSON line: NCCL prints its version banner there, so fd 1 points
        # to stderr until the line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = parallel.Communicator()

    ctx = _capi.Context(local)
    lib = ctx.lib
    basis = make_basis()
    configs = make_configs(args.ncfg, seed=1 if os.environ.get("IPF_BENCH_SAME_SEED") else 1 + rank)
    nrows = _alloc_rows(configs)
    ncols = len(basis)
    entries_local = nrows * ncols
    dev_basis = _capi.DeviceBasis(ctx, basis)
    off, pos, cell, pbc, Z, rows = flatten_configs(configs)

    # ---- device-resident inputs for the `value` arm
    hcf = C.c_void_p()
    ctx.check(lib.ipf_configs_create(ctx.h, len(configs), _capi.ptr(off, C.c_int64), _capi.ptr(pos, C.c_double),
                                     _capi.ptr(cell, C.c_double), _capi.ptr(pbc, C.c_uint8), _capi.ptr(Z, C.c_int32),
                                     _capi.ptr(rows["E"], C.c_int64), _capi.ptr(rows["F"], C.c_int64),
                                     _capi.ptr(rows["V"], C.c_int64), C.byref(hcf)))
    Psi = _capi.DeviceMatrix(ctx, nrows, ncols)

    class _DB:      # minimal stand-in so that the host-side weight logic can be reused
        pass
    db = _DB(); db.configs = configs; db.nrows = nrows
    Vref = ipf.OneBody(E0)
    weights = _fix_weights(dict(WEIGHTS))
    Y, W, Icfg = collect_observations(db, weights, Vref)
    Irows, _ = _select(db, Y, W, Icfg, None, None)
    dY = torch.from_numpy(Y).cuda(); dW = torch.from_numpy(W).cuda(); dI = torch.from_numpy(Irows).cuda()
    torch.cuda.synchronize()

    def solve_dev():
        h = C.c_void_p()
        ctx.check(lib.ipf_solve_begin_dev(ctx.h, Psi.h, dY.data_ptr(), dW.data_ptr(), dI.data_ptr(), len(Irows),
                                          None, 0, None, C.byref(h)))
        try:
            done = C.c_int32(0)
            while not done.value:
                g, n1 = C.c_void_p(), C.c_int64()
                ctx.check(lib.ipf_solve_gram(ctx.h, h, C.byref(g), C.byref(n1)))
                if comm is not None:
                    comm.allreduce_device_f64(g.value, n1.value * n1.value, local)
                ctx.check(lib.ipf_solve_factor(ctx.h, h, C.byref(done)))
            c = np.empty(ncols)
            ssr, ssy = C.c_double(), C.c_double()
            ctx.check(lib.ipf_solve_finish(ctx.h, h, _capi.ptr(c, C.c_double), None, C.byref(ssr), C.byref(ssy)))
            npass, defect, shift = C.c_int32(), C.c_double(), C.c_double()
            lib.ipf_solve_info(h, C.byref(npass), C.byref(defect), C.byref(shift), None, None)
            solve_dev.info = {"passes": npass.value, "orth_defect_last_pass": defect.value, "shift": shift.value}
        finally:
            lib.ipf_solve_destroy(ctx.h, h)
        return c, ssr.value, ssy.value, npass.value

    def step_dev():
        ctx.check(lib.ipf_assemble_configs(ctx.h, dev_basis.h, hcf, Psi.h))
        return solve_dev()

    def barrier():
        if comm is not None:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()
    ctx.profile(True)
    ctx.profile_reset()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record(stream)
        t_asm = 0.0
        for _ in range(args.steps):
            ta = time.perf_counter()
            ctx.check(lib.ipf_assemble_configs(ctx.h, dev_basis.h, hcf, Psi.h))
            t_asm += time.perf_counter() - ta
            c_dev, ssr, ssy, npass = solve_dev()
        e1.record(stream)
        barrier()
    ms_dev = e0.elapsed_time(e1)
    launches = ctx.launches - l0
    prof = {k: ctx.profile_get(k) for k in ("neighbour_list", "pair_record", "site_deriv_2b", "site_deriv_3b", "gather_3b",
                                             "tsort_3b", "k4_build", "gram", "cholesky", "trsm", "residual")}
    clusters3 = ctx.profile_get("count:clusters_3b")[0]
    ordered3 = ctx.profile_get("count:ordered_pairs_3b")[0]
    pairs3 = ctx.profile_get("count:pairs_3b")[0]
    ctx.profile(False)
    if comm is not None:
        t = torch.tensor([ms_dev], device="cuda", dtype=torch.float64)
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev = float(t.item())
    entries_total = entries_local * world
    value = entries_total * args.steps / (ms_dev * 1e-3)

    # ---- e2e arm: the public API with host buffers
    def step_e2e():
        dbx = LsqDB("", basis, configs, ctx=ctx)
        IP, info = ipf.lsqfit(dbx, weights=dict(WEIGHTS), E0=E0, Vref=ipf.OneBody(E0), comm=comm)
        res = (info["c"], info["rel_rms"])
        dbx.delete()
        return res
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_e2e, rel_e2e = step_e2e()
    barrier()
    t_e2e = time.perf_counter() - t0
    if comm is not None:
        t = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_val = entries_total * args.steps / t_e2e
    h2d = off.nbytes + pos.nbytes + cell.nbytes + pbc.nbytes + 3 * rows["E"].nbytes + Y.nbytes + W.nbytes + Irows.nbytes
    d2h = 8 * ncols + 8 * ncols * ncols + 16

    if rank != 0:
        if comm is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---- rooflines: the three heavy kernels, the slowest one is reported as `roofline`
    blk3 = [b for b in basis.blocks if b.N == 3][0]
    D = blk3.maxdeg
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get("hbm_gbs", 6553.9))
    fp64_peak, dmma_peak = 33.9, 37.1      # TFLOP/s, tools/peaks.cu on this pool (profiles/r01_fp64_peaks.jsonl)
    # (a) moments kernel: FP64 operations of the factorised algorithm (FMA = 2, MUL = ADD = 1), DESIGN.md §kernels
    f_k = 15 + D + 7 * (D * (D + 1) // 2) + 2 * (D + 1)         # per ordered neighbour pair (j, k)
    n_single = int(np.sum(np.bincount(blk3.mono_b) == 1)); n_double = blk3.nb - n_single
    f_comb = 36 * n_double + 30 * n_single                      # per own leg: combine moments into all functions
    # per own leg and function: F[i] += g, F[j] -= g (6 adds), E += e (1), six Voigt products of the virial (12)
    f_scatter = 19 * blk3.nb
    msg, ng = prof["gather_3b"]
    resident = ng == 0            # configuration-resident kernel: moments + combine + in-SM scatter in one kernel
    flops3 = ordered3 * f_k + pairs3 * (f_comb + (f_scatter if resident else 0))
    ms3, n3 = prof["site_deriv_3b"]
    ach3 = flops3 / (ms3 * 1e-3) * 1e-12 if ms3 > 0 else 0.0
    naive3 = clusters3 * cluster_flops(blk3) / (ms3 * 1e-3) * 1e-12 if ms3 > 0 else 0.0
    traffic = {}
    tf = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tf):
        traffic = json.load(open(tf))
    tkey = "cfg3b" if resident else "moments3b"
    kname = (f"cfg3b_kernel<{D}, unit 0..7> (configuration-resident 3-body kernel: moments + combine + scatter into the "
             "E/F/V rows from shared memory; all units of a chunk in one launch, blockIdx.y = unit)" if resident else
             f"moments3b_kernel<{D}, unit 0..7> (3-body basis values + derivatives per ordered pair; 8 unit launches per batch)")
    r_mom = {"kernel": kname,
             "bound": "fp64", "achieved": ach3, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach3 / fp64_peak,
             "traffic": traffic.get(tkey, {}).get("dram_bytes_per_pair", 0) * pairs3 / max(1, n3) or None,
             "launches": n3, "avg_launch_ms": ms3 / max(1, n3), "algorithmic_flops_per_launch": flops3 / max(1, n3),
             "algorithmic_bytes_per_launch": (pairs3 * 8.0 * (8 + D + 2) + float(nrows) * blk3.nb * 8.0 * args.steps) / max(1, n3),
             "naive_algorithm_equivalent_tflops": naive3, "total_ms": ms3,
             "note": "one launch = all units of one chunk of configurations; flops = factorised algorithm (DESIGN.md section 4), "
                     "FMA = 2; algorithmic bytes = pair records read once + Psi rows written once"}
    # (b) gather kernel (scratch path only): HBM bytes = reverse scratch entry (24 B) per (pair, function) + the rows written
    bytes_g = pairs3 * blk3.nb * 24.0 + float(nrows) * blk3.nb * 8.0 * args.steps
    achg = bytes_g / (msg * 1e-3) * 1e-9 if msg > 0 else 0.0
    r_gat = {"kernel": "gather3b_kernel (scratch -> E/F/V rows)", "bound": "hbm", "achieved": achg, "peak": hbm_peak,
             "unit": "GB/s", "frac": achg / hbm_peak, "traffic": traffic.get("gather3b", {}).get("dram_bytes_per_config", 0) * len(configs) * args.steps / max(1, ng) or None, "launches": ng, "avg_launch_ms": msg / max(1, ng),
             "algorithmic_bytes_per_launch": bytes_g / max(1, ng), "total_ms": msg,
             "note": "reads only the reverse entries (24 B); own sums come pre-reduced from the producer"}
    # (c) Gram kernel: algorithmic flops M n (n+1) (symmetric half) on the FP64 tensor pipe
    gram_ms, gram_n = prof["gram"]
    M = len(Irows)
    gram_flops = float(M) * ncols * (ncols + 1) * gram_n
    achd = gram_flops / (gram_ms * 1e-3) * 1e-12 if gram_ms > 0 else 0.0
    r_gram = {"kernel": "gram_kernel (DMMA m8n8k4 f64, upper-triangular tiles)", "bound": "tensor", "achieved": achd,
              "peak": dmma_peak, "unit": "TFLOP/s", "frac": achd / dmma_peak, "traffic": None, "launches": gram_n,
              "avg_launch_ms": gram_ms / max(1, gram_n), "algorithmic_flops_per_launch": gram_flops / max(1, gram_n),
              "total_ms": gram_ms}
    # (d) re-orthogonalisation update A <- A L^-T (DMMA product with the explicit inverse): M n^2 flops (triangular)
    trsm_ms, trsm_n = prof["trsm"]
    trsm_flops = float(M) * ncols * ncols * trsm_n
    acht = trsm_flops / (trsm_ms * 1e-3) * 1e-12 if trsm_ms > 0 else 0.0
    r_trsm = {"kernel": "trinv_kernel + applyw_kernel (A <- A L^-T, DMMA m8n8k4 f64, in place)", "bound": "tensor",
              "achieved": acht, "peak": dmma_peak, "unit": "TFLOP/s", "frac": acht / dmma_peak, "traffic": None,
              "launches": trsm_n, "avg_launch_ms": trsm_ms / max(1, trsm_n),
              "algorithmic_flops_per_launch": trsm_flops / max(1, trsm_n), "total_ms": trsm_ms}
    cands = sorted([r for r in (r_mom, r_gat, r_gram, r_trsm) if r["launches"] > 0], key=lambda r: -r["total_ms"])
    roofline = cands[0]
    roofline["peak_source"] = ("HBM: MEASURED_PEAKS.json (driver); FP64 FMA 33.9 and DMMA 37.1 TFLOP/s: own microbenchmark "
                               "tools/peaks.cu (MEASURED_PEAKS.json has no FP64 entry), cuBLAS DGEMM 8192^3 = 35.5")
    kernels = {k: {"ms": v[0], "launches": v[1]} for k, v in prof.items()}

    # ---- CPU baseline (restated reference path) on a bounded sample, rank 0, N = 1 only
    cpu = None
    if world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ipf_oracle as O
        nth = O.num_threads()
        ent, ta, ts = cpu_reference_step(basis, configs[:2], nth)
        S = int(max(2, min(len(configs), 60.0 / ((ta + ts) / 2))))      # ~15-20 s of CPU work (threads scale with the task count)
        ent, ta, ts = cpu_reference_step(basis, configs[:S], nth)
        cpu = {"value": ent / (ta + ts), "unit": "entries/s", "cores": nth, "kind": "port",
               "sample": f"first {S} of {len(configs)} configurations: assembly {ta:.2f} s (1 task per (config,okey), E/F/V separately) "
                         f"+ LAPACK QR {ts:.3f} s; restated reference CPU path (Julia not installable)"}

    line = {"metric": "lsqdb_entries_per_s", "value": value, "unit": "entries/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "configs_per_gpu": len(configs), "rows_per_gpu": nrows, "cols": ncols,
                       "l2": "working set (Psi 0.69 GB + pair records) exceeds the 126 MB L2",
                       "solver": f"CholeskyQR, {npass} pass(es)", "parallelism": f"dp{world} (configs sharded, Gram all-reduce)"},
            "assemble_entries_per_s": entries_local * args.steps / t_asm if t_asm > 0 else None,
            "solve_ms": (ms_dev - 1e3 * t_asm) / args.steps,
            "e2e": {"value": e2e_val, "unit": "entries/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks.summary(), "roofline": roofline,
            "roofline_other": cands[1:], "kernels": kernels,
            "cpu_baseline": cpu, "fit": {"rel_rms": math.sqrt(ssr / ssy) if comm is None else rel_e2e,
                                         "coeff_diff_dev_vs_e2e": float(np.linalg.norm(c_dev - c_e2e) / np.linalg.norm(c_e2e)),
                                         "solve_info": getattr(solve_dev, "info", None)}}
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if comm is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ncfg", type=int, default=2000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
