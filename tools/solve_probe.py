"""Prints the per-pass orthogonality defects of the CholeskyQR solve on the bench workload (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200.lsq import _Solver, _fix_weights, _select, collect_observations
ncfg = int(sys.argv[1]) if len(sys.argv) > 1 else 400
basis = bench.make_basis(); configs = bench.make_configs(range(ncfg), 1)
db = ipf.LsqDB("", basis, configs)
Y, W, Icfg = collect_observations(db, _fix_weights(dict(bench.WEIGHTS)), ipf.OneBody(bench.E0))
Irows, _ = _select(db, Y, W, Icfg, None, None)
S = _Solver(db._context(), db.Psi_dev, Y, W, Irows, None, None)
k = 0
while True:
    S.gram(); done = S.factor(); print("pass", k, S.info()); k += 1
    if done: break
c, R, ssr, ssy = S.finish()
print("cond(R) %.3e rel_rms %.3e" % (np.linalg.cond(R), np.sqrt(ssr / ssy)))
sv = np.linalg.svd(R, compute_uv=False); print("sigma max/min", sv[0], sv[-1])
