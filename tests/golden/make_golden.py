"""Generates the golden fixtures of tests/golden/ from the CPU oracle.

The reference holds no golden vectors for this path ("0 .xyz/.h5/.json
fixtures") and cannot be run here, so these vectors pin the ORACLE (and, through the
GPU tests, the CUDA path) against regressions; they are not reference outputs.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ipfitting_jl_b200 as ipf  # noqa: E402
from ipfitting_jl_b200 import synth  # noqa: E402
from ipfitting_jl_b200.lsq_db import _alloc_rows  # noqa: E402
import ipf_oracle as O  # noqa: E402
import lsq_oracle as LO  # noqa: E402
from util import rows_for, si_desc  # noqa: E402


def case_basis(name):
    if name == "si_2b3b":
        return ipf.NBodyBasis(ipf.nbpolys(2, si_desc(), 6) + ipf.nbpolys(3, si_desc(2.0), 4))
    if name == "si_4b":
        return ipf.NBodyBasis(ipf.nbpolys(3, si_desc(2.0), 3) + ipf.nbpolys(4, si_desc(1.8), 3))
    raise KeyError(name)


def case_configs(name):
    sw = synth.StillingerWeber()
    if name == "si_2b3b":
        return synth.rattled_configs("Si", 2, 2, seed=101, calc=sw) + synth.rattled_configs("Si", 1, 2, seed=102, calc=sw, configtype="small")
    if name == "si_4b":
        return synth.rattled_configs("Si", 1, 3, seed=103, calc=sw)
    raise KeyError(name)


CASES = ["si_2b3b", "si_4b"]
WEIGHTS = {"default": {"E": 30.0, "F": 1.0, "V": 0.5}}

if __name__ == "__main__":
    for name in CASES:
        B, configs = case_basis(name), case_configs(name)
        n = _alloc_rows(configs)
        r = rows_for(configs)
        Psi = O.assemble(B, configs, r["E"], r["F"], r["V"], n)
        c, kappa, rel, R = LO.lsqfit_qr(Psi, configs, WEIGHTS, E0=-4.0)
        at = configs[0].at
        I, J, S, Rv = O.neighbours(at.X, at.cell, at.pbc, B.rcut)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), Psi=Psi, c=c, kappa=kappa, rel_rms=rel,
                            nl_I=I, nl_J=J, nl_S=S, nl_R=Rv, X0=at.X, cell0=at.cell)
        print(name, Psi.shape, "cond", kappa, "rel_rms", rel, "pairs", len(I))
