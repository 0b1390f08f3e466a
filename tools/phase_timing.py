"""Wall-clock timing of the phases of one bench step (development aid)."""
import ctypes as C, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import _capi
from ipfitting_jl_b200.lsq import _fix_weights, _select, collect_observations
from ipfitting_jl_b200.lsq_db import flatten_configs, _alloc_rows

ncfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ctx = _capi.Context(0); lib = ctx.lib
basis = bench.make_basis(); configs = bench.make_configs(ncfg, 1)
nrows = _alloc_rows(configs); ncols = len(basis)
dev_basis = _capi.DeviceBasis(ctx, basis)
off, pos, cell, pbc, Z, rows = flatten_configs(configs)
hcf = C.c_void_p()
ctx.check(lib.ipf_configs_create(ctx.h, len(configs), _capi.ptr(off, C.c_int64), _capi.ptr(pos, C.c_double), _capi.ptr(cell, C.c_double), _capi.ptr(pbc, C.c_uint8), _capi.ptr(Z, C.c_int32), _capi.ptr(rows["E"], C.c_int64), _capi.ptr(rows["F"], C.c_int64), _capi.ptr(rows["V"], C.c_int64), C.byref(hcf)))
Psi = _capi.DeviceMatrix(ctx, nrows, ncols)
class _DB: pass
db = _DB(); db.configs = configs; db.nrows = nrows
Y, W, Icfg = collect_observations(db, _fix_weights(dict(bench.WEIGHTS)), ipf.OneBody(bench.E0))
Irows, _ = _select(db, Y, W, Icfg, None, None)
dY = torch.from_numpy(Y).cuda(); dW = torch.from_numpy(W).cuda(); dI = torch.from_numpy(Irows).cuda()
torch.cuda.synchronize()
for it in range(5):
    T = {}
    t = time.perf_counter(); ctx.check(lib.ipf_assemble_configs(ctx.h, dev_basis.h, hcf, Psi.h)); T["assemble"] = time.perf_counter() - t
    h = C.c_void_p()
    t = time.perf_counter(); ctx.check(lib.ipf_solve_begin_dev(ctx.h, Psi.h, dY.data_ptr(), dW.data_ptr(), dI.data_ptr(), len(Irows), None, 0, None, C.byref(h))); T["begin"] = time.perf_counter() - t
    done = C.c_int32(0); k = 0
    while not done.value:
        g, n1 = C.c_void_p(), C.c_int64()
        t = time.perf_counter(); ctx.check(lib.ipf_solve_gram(ctx.h, h, C.byref(g), C.byref(n1))); T[f"gram{k}"] = time.perf_counter() - t
        t = time.perf_counter(); ctx.check(lib.ipf_solve_factor(ctx.h, h, C.byref(done))); T[f"factor{k}"] = time.perf_counter() - t
        k += 1
    c = np.empty(ncols); ssr, ssy = C.c_double(), C.c_double()
    t = time.perf_counter(); ctx.check(lib.ipf_solve_finish(ctx.h, h, _capi.ptr(c, C.c_double), None, C.byref(ssr), C.byref(ssy))); T["finish"] = time.perf_counter() - t
    t = time.perf_counter(); lib.ipf_solve_destroy(ctx.h, h); T["destroy"] = time.perf_counter() - t
    print(it, " ".join(f"{k}={1e3*v:.2f}" for k, v in T.items()), "total=%.2f" % (1e3 * sum(T.values())))
print("--- profiling on")
ctx.profile(True)
for it in range(3):
    t0 = time.perf_counter()
    ctx.check(lib.ipf_assemble_configs(ctx.h, dev_basis.h, hcf, Psi.h)); ta = time.perf_counter() - t0
    h = C.c_void_p(); T = {}
    t = time.perf_counter(); ctx.check(lib.ipf_solve_begin_dev(ctx.h, Psi.h, dY.data_ptr(), dW.data_ptr(), dI.data_ptr(), len(Irows), None, 0, None, C.byref(h))); T["begin"] = time.perf_counter() - t
    done = C.c_int32(0); k = 0
    while not done.value:
        g, n1 = C.c_void_p(), C.c_int64()
        t = time.perf_counter(); ctx.check(lib.ipf_solve_gram(ctx.h, h, C.byref(g), C.byref(n1))); T[f"gram{k}"] = time.perf_counter() - t
        t = time.perf_counter(); ctx.check(lib.ipf_solve_factor(ctx.h, h, C.byref(done))); T[f"factor{k}"] = time.perf_counter() - t
        k += 1
    t = time.perf_counter(); ctx.check(lib.ipf_solve_finish(ctx.h, h, _capi.ptr(c, C.c_double), None, C.byref(ssr), C.byref(ssy))); T["finish"] = time.perf_counter() - t
    t = time.perf_counter(); lib.ipf_solve_destroy(ctx.h, h); T["destroy"] = time.perf_counter() - t
    print(it, "assemble=%.2f" % (1e3 * ta), " ".join(f"{k}={1e3*v:.2f}" for k, v in T.items()))
ctx.profile(False)
print("--- clock sampler on")
with bench.ClockSampler(0) as cs:
    for it in range(3):
        t0 = time.perf_counter()
        ctx.check(lib.ipf_assemble_configs(ctx.h, dev_basis.h, hcf, Psi.h)); ta = time.perf_counter() - t0
        h = C.c_void_p(); T = {}
        t = time.perf_counter(); ctx.check(lib.ipf_solve_begin_dev(ctx.h, Psi.h, dY.data_ptr(), dW.data_ptr(), dI.data_ptr(), len(Irows), None, 0, None, C.byref(h))); T["begin"] = time.perf_counter() - t
        done = C.c_int32(0); k = 0
        while not done.value:
            g, n1 = C.c_void_p(), C.c_int64()
            t = time.perf_counter(); ctx.check(lib.ipf_solve_gram(ctx.h, h, C.byref(g), C.byref(n1))); T[f"gram{k}"] = time.perf_counter() - t
            t = time.perf_counter(); ctx.check(lib.ipf_solve_factor(ctx.h, h, C.byref(done))); T[f"factor{k}"] = time.perf_counter() - t
            k += 1
        t = time.perf_counter(); ctx.check(lib.ipf_solve_finish(ctx.h, h, _capi.ptr(c, C.c_double), None, C.byref(ssr), C.byref(ssy))); T["finish"] = time.perf_counter() - t
        t = time.perf_counter(); lib.ipf_solve_destroy(ctx.h, h); T["destroy"] = time.perf_counter() - t
        print(it, "assemble=%.2f" % (1e3 * ta), " ".join(f"{k}={1e3*v:.2f}" for k, v in T.items()))
print(cs.summary())
