"""cProfile of one end-to-end step (LsqDB + lsqfit from host objects) -- development aid."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import _capi
from ipfitting_jl_b200.lsq_db import LsqDB
from ipfitting_jl_b200.lsq import lsqfit

ncfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ctx = _capi.Context(0)
basis = bench.make_basis(); configs = bench.make_configs(ncfg, 1)

def step():
    db = LsqDB("", basis, configs, ctx=ctx)
    IP, info = ipf.lsqfit(db, weights=dict(bench.WEIGHTS), E0=bench.E0, Vref=ipf.OneBody(bench.E0))
    db.delete()
    return info

for _ in range(3):
    t = time.perf_counter(); step(); print("step %.2f ms" % (1e3 * (time.perf_counter() - t)))
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
