"""Wall-clock phases of one end-to-end step (LsqDB + lsqfit from host objects) -- development aid."""
import sys, os, time, functools, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import _capi, lsq_db, lsq

ncfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ctx = _capi.Context(0)
basis = bench.make_basis(); configs = bench.make_configs(range(ncfg), 1)
acc = collections.OrderedDict()

def timed(mod, name, label=None):
    f = getattr(mod, name)
    @functools.wraps(f)
    def g(*a, **k):
        t = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            acc[label or name] = acc.get(label or name, 0.0) + time.perf_counter() - t
    setattr(mod, name, g)

timed(lsq_db, "_alloc_rows"); timed(lsq_db, "flatten_configs")
timed(lsq_db.LsqDB, "_assemble", "LsqDB._assemble (flatten + ipf_assemble)")
timed(_capi, "DeviceBasis"); timed(_capi, "DeviceMatrix")
timed(lsq, "collect_observations"); timed(lsq, "_select"); timed(lsq, "_precon_blocks")
timed(lsq._Solver, "__init__", "solve_begin (H2D Y, W, I + build)")
timed(lsq._Solver, "gram"); timed(lsq._Solver, "factor"); timed(lsq._Solver, "finish"); timed(lsq._Solver, "close")
timed(np.linalg, "cond"); timed(lsq, "asm_fitinfo"); timed(lsq_db.LsqDB, "delete")

def step():
    db = lsq_db.LsqDB("", basis, configs, ctx=ctx)
    IP, info = lsq.lsqfit(db, weights=dict(bench.WEIGHTS), E0=bench.E0, Vref=ipf.OneBody(bench.E0))
    db.delete()
    return info

for _ in range(3):
    step()
acc.clear()
N = 5
t0 = time.perf_counter()
for _ in range(N):
    step()
tot = (time.perf_counter() - t0) / N
print("step %.2f ms" % (1e3 * tot))
for k, v in acc.items():
    print("  %-45s %7.2f ms" % (k, 1e3 * v / N))
