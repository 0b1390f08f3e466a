"""Randomised check of the device solve against LAPACK over column counts, row counts and condition numbers
(development aid; run on a GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ipfitting_jl_b200 import _capi
from ipfitting_jl_b200.lsq import solve_device
ctx = _capi.Context(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
worst = 0.0
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 60
for case in range(ncase):
    n = int(rng.choice([1, 2, 3, 7, 31, 32, 33, 64, 127, 128, 129, 169, 170, 171, 255, 256, 257, 511, 805, 1023, 1200]))
    if case % 3 == 0:
        n = int(rng.integers(1, 900))
    M = n + int(rng.integers(0, 6 * n + 50))
    logc = rng.uniform(0, 10)
    U, _ = np.linalg.qr(rng.standard_normal((M, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    s = np.logspace(0, -logc, n) if n > 1 else np.ones(1)
    A = np.asfortranarray((U * s) @ V.T)
    x = rng.standard_normal(n)
    y = A @ x + 1e-4 * rng.standard_normal(M)
    w = rng.uniform(0.5, 2.0, M)
    Pm = _capi.DeviceMatrix(ctx, M, n)
    Pm.from_host(A)
    c, R, info = solve_device(ctx, Pm, y, w)
    Pm.close()
    Aw, yw = A * w[:, None], y * w
    c_ref = np.linalg.lstsq(Aw, yw, rcond=None)[0]
    r, r_ref = np.linalg.norm(Aw @ c - yw), np.linalg.norm(Aw @ c_ref - yw)
    e_res = abs(r - r_ref) / max(r_ref, 1e-300)
    e_pred = np.linalg.norm(Aw @ (c - c_ref)) / np.linalg.norm(yw)
    worst = max(worst, e_pred)
    flag = "" if (e_pred <= 1e-8 and e_res <= 1e-6) else "   <-- CHECK"
    print(f"n {n:5d} M {M:6d} cond 1e{logc:4.1f} passes {info['passes']} csne {info.get('csne_steps')}  resid diff {e_res:.1e}  pred diff {e_pred:.1e}{flag}", flush=True)
print("worst prediction difference %.2e" % worst)
