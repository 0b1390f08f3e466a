"""Randomised agreement of the three 3-body paths (configuration-resident, scratch + gather, table-driven), the two
2-body paths and the two 4-body paths (moment-factorised, table-driven) on mixed configuration sets -- development aid,
run on a GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import _capi, synth
from ipfitting_jl_b200.lsq_db import LsqDB
from ipfitting_jl_b200.prototypes import Dat

ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ctx = _capi.Context(0)
rng = np.random.default_rng(2024)
worst = 0.0
for case in range(ncase):
    sp = ["Si", "Cu", "W", "C"][rng.integers(4)]
    r0 = ipf.rnn(sp)
    D3 = int(rng.integers(2, 15)); D2 = int(rng.integers(2, 15))
    rc3 = float(rng.uniform(1.3, 2.1)) * r0; rc2 = float(rng.uniform(1.5, 2.6)) * r0
    tr = [f"exp(- (r/{r0} - 1.0))", f"({r0}/r)^3", f"exp(- 1.7 * (r/{r0} - 1.0))"][rng.integers(3)]
    B = ipf.nbpolys(2, ipf.BondAngleDesc(tr, ipf.CosCut(rc2 - 0.8, rc2)), D2) + \
        ipf.nbpolys(3, ipf.BondAngleDesc(tr, ipf.CosCut(rc3 - 0.6, rc3)), D3)
    D4 = 0
    if case % 2 == 1:          # every second case carries a 4-body block (the table-driven 4-body kernel is slow: small cutoff)
        D4 = int(rng.integers(1, 7))
        rc4 = float(rng.uniform(1.2, 1.7)) * r0
        B = B + ipf.nbpolys(4, ipf.BondAngleDesc(tr, ipf.CosCut(rc4 - 0.5, rc4)), D4)
    configs = []
    big = rng.random() < 0.15 and D4 == 0          # sometimes more than 256 configurations / more than 74 atoms
    for _ in range(int(rng.integers(260, 300)) if (big and D3 <= 6) else int(rng.integers(1, 7))):
        L = int(rng.integers(1, 3))
        d = synth.rattled_configs(sp, L, 1, seed=int(rng.integers(1 << 30)), calc=synth.PairPotential(r0),
                                  vacancies=int(rng.integers(0, 2)) if L > 1 else 0)[0]
        nat = len(d.at)
        cap = nat if (big and D3 > 6) else min(nat, 74)
        keep = np.sort(rng.choice(nat, size=int(rng.integers(1, cap + 1)), replace=False))
        pbc = tuple(bool(x) for x in rng.integers(0, 2, 3))
        at = ipf.Atoms(d.at.X[keep], d.at.cell, pbc, d.at.Z[keep])
        kw = {}
        if rng.random() < 0.8: kw["E"] = 0.0
        if rng.random() < 0.8: kw["F"] = np.zeros((len(keep), 3))
        if rng.random() < 0.8: kw["V"] = np.zeros((3, 3))
        if not kw: kw["E"] = 0.0
        configs.append(Dat(at, "r", **kw))
    out = {}
    for mode in (0, 2, 1):
        ctx.force_generic(mode)
        out[mode] = LsqDB("", B, configs, ctx=ctx).Psi
    ctx.force_generic(0)
    ref = out[1]
    # columns that vanish by symmetry (single-atom lattices) hold rounding noise only: floor the scale
    scale = np.maximum(np.abs(ref).max(axis=0), 1e-3 * np.abs(ref).max()) + 1e-300
    e0 = (np.abs(out[0] - ref) / scale).max(); e2 = (np.abs(out[2] - ref) / scale).max()
    worst = max(worst, e0, e2)
    flag = "" if max(e0, e2) < 1e-9 else "   <-- CHECK"
    print(f"case {case:3d} {sp:2s} D2={D2:2d} D3={D3:2d} D4={D4} ncfg={len(configs)} nat={[len(c.at) for c in configs][:8]} "
          f"rows={ref.shape[0]} resident-vs-table {e0:.2e} scratch-vs-table {e2:.2e}{flag}")
print("worst", worst)
sys.exit(0 if worst < 1e-9 else 1)
