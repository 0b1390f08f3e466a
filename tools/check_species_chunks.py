"""One-off check: species-resolved and environment-dependent runs across SEVERAL chunks (> 65 536 atoms) and mixed sizes
against the oracle (development aid; run on a GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import synth
from ipfitting_jl_b200.lsq_db import LsqDB, _alloc_rows
import ipf_oracle as O
from util import psi_errors, rows_for
desc = ipf.BondAngleDesc("exp(- (r/1.89 - 1.0))", ipf.CosCut(2.6, 3.4))
B = ipf.NBodyBasis(ipf.species_nbpolys(2, desc, 3, [14, 6]) + ipf.species_nbpolys(3, desc, 2, [14, 6])
                   + ipf.envpolys(3, desc, 2, 1, ipf.CosCut(2.0, 3.0)))
rng = np.random.default_rng(3)
configs = []
for k in range(150):
    L = [4, 4, 4, 3, 2, 1][k % 6]
    at = synth.rattle(synth.repeat(synth.zincblende(4.36, 14, 6), L), 0.1, rng, 0.02)
    configs.append(ipf.Dat(at, "sic", E=0.0, F=np.zeros((len(at), 3)), V=np.zeros(6)))
print("atoms", sum(len(d.at) for d in configs), "columns", len(B))
t = time.time(); db = LsqDB("", B, configs); Psi = db.Psi; print("gpu %.1f s" % (time.time() - t))
r = rows_for(configs)
t = time.time(); ref = O.assemble(B, configs, r["E"], r["F"], r["V"], db.nrows); print("oracle %.1f s" % (time.time() - t))
err = psi_errors(Psi, ref, configs)
print("max block-relative error %.2e" % err)
sys.exit(0 if err <= 1e-10 else 1)
