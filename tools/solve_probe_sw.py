"""Per-pass diagnostics of the CholeskyQR solve on the SW 2B(14)+3B(14) system of tests/test_fit_gpu.py (cond ~ 3e16),
next to LAPACK's Householder QR of the same device-assembled matrix (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.linalg as sl
import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import synth
from ipfitting_jl_b200.lsq import _Solver, _fix_weights, _select, collect_observations
r0 = ipf.rnn("Si"); calc = synth.StillingerWeber()
data = synth.rattled_configs("Si", 2, 100, rmax=0.33 * r0, strain=0.0, seed=1, calc=calc, virial=False, mode="julip", configtype="set")
desc = ipf.BondAngleDesc(f"exp(- (r/{r0} - 1.0))", ipf.CosCut(calc.rcut - 1.0, calc.rcut))
B = ipf.nbpolys(2, desc, 14) + ipf.nbpolys(3, desc, 14)
db = ipf.LsqDB("", B, data)
Y, W, Icfg = collect_observations(db, _fix_weights({"default": {"E": 100.0, "F": 1.0}}), ipf.OneBody(0.0))
Irows, _ = _select(db, Y, W, Icfg, None, None)
S = _Solver(db._context(), db.Psi_dev, Y, W, Irows, None, None)
A0, y0 = S.system_to_host()
k = 0
while True:
    S.gram(); done = S.factor(); print("pass", k, S.info(), flush=True); k += 1
    if done: break
c, R, ssr, ssy = S.finish()
print("device: rel_rms (orthonormal basis) %.6e   |A0 c - y|/|y| %.6e  |c| %.3e" % (np.sqrt(ssr / ssy), np.linalg.norm(A0 @ c - y0) / np.linalg.norm(y0), np.linalg.norm(c)))
Q, Rl = np.linalg.qr(A0)
cl = sl.solve_triangular(Rl, Q.T @ y0)
print("lapack: |A0 c - y|/|y| %.6e  |c| %.3e cond %.3e" % (np.linalg.norm(A0 @ cl - y0) / np.linalg.norm(y0), np.linalg.norm(cl), np.linalg.cond(Rl)))
