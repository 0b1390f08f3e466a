"""A10: force preconditioner.  Properties of `test/test_precon.jl:89-110` (P symmetric, row sums = stab,
three smallest eigenvalues = stab) and parity of the preconditioned system / fit against the oracle."""
import numpy as np
import pytest
import scipy.linalg as sla

import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import synth
from ipfitting_jl_b200.lsq_db import LsqDB
import ipf_oracle as O
import lsq_oracle as LO
from util import rows_for, si_desc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("innerstab", [1.0, 0.1, 0.0])
def test_adjacency_precon_properties(ctx, innerstab):
    r0 = ipf.rnn("Si")
    at = synth.rattled_configs("Si", 2, 1, seed=3)[0].at
    stab = 1e-3
    iv = [(0.0, 1.2 * r0, 1.0), (1.2 * r0, 1.8 * r0, 0.3)]
    pre = ipf.AdjacancyPrecon(iv, stab=stab, innerstab=innerstab, ctx=ctx)
    P = pre(at)
    assert np.allclose(P, P.T, atol=1e-14)
    assert np.allclose(P.sum(axis=0), stab, atol=1e-11) and np.allclose(P.sum(axis=1), stab, atol=1e-11)
    ev = np.linalg.eigvalsh(P)
    assert np.allclose(ev[:3], stab, atol=1e-10)           # translation null space, shifted by stab
    assert np.allclose(P, LO.adjacency_precon(at, iv, stab, innerstab), rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("solver", ["chol", "sqrt"])
def test_preconditioned_system_and_fit(ctx, solver):
    r0 = ipf.rnn("Si")
    sw = synth.StillingerWeber()
    configs = synth.rattled_configs("Si", 2, 3, seed=31, calc=sw, configtype="bulk") + \
        synth.rattled_configs("Si", 1, 3, seed=32, calc=sw, configtype="small")
    B = ipf.nbpolys(2, si_desc(2.0), 4) + ipf.nbpolys(3, si_desc(1.7), 2)
    db = LsqDB("", B, configs, ctx=ctx)
    r = rows_for(configs)
    ref = O.assemble(db.basis, configs, r["E"], r["F"], r["V"], db.nrows)
    iv = [(0.0, 1.3 * r0, 1.0)]
    pre = ipf.AdjacancyPrecon(iv, stab=0.05, innerstab=0.1, ctx=ctx)
    w = {"default": {"E": 30.0, "F": 1.0, "V": 0.5}, "precon": {"default": pre, "_solver": solver}}
    A, y = ipf.get_lsq_system(db, Vref=ipf.OneBody(-4.0), weights=dict(w))
    A0, y0 = LO.forceprecon_system(ref, configs, {"default": w["default"]},
                                   lambda at: LO.adjacency_precon(at, iv, 0.05, 0.1), E0=-4.0, solver=solver)
    assert A.shape == A0.shape
    assert np.allclose(y, y0, rtol=1e-10, atol=1e-12 * np.abs(y0).max())
    scale = np.abs(A0).max(axis=0)
    assert (np.abs(A - A0).max(axis=0) <= 1e-9 * scale).all()
    IP, info = ipf.lsqfit(db, weights=dict(w), E0=-4.0, Vref=ipf.OneBody(-4.0))
    Q, R = sla.qr(A0, mode="economic")
    c_ref = sla.solve_triangular(R, Q.T @ y0)
    assert np.linalg.norm(info["c"] - c_ref) <= 1e-8 * np.linalg.norm(c_ref)


def test_precon_only_for_listed_configtypes(ctx):
    sw = synth.StillingerWeber()
    configs = synth.rattled_configs("Si", 1, 2, seed=41, calc=sw, configtype="a") + \
        synth.rattled_configs("Si", 1, 2, seed=42, calc=sw, configtype="b")
    B = ipf.nbpolys(2, si_desc(2.0), 3)
    db = LsqDB("", B, configs, ctx=ctx)
    pre = ipf.AdjacancyPrecon([(0.0, 3.0, 1.0)], stab=0.1, ctx=ctx)
    A0, y0 = ipf.get_lsq_system(db, Vref=None, weights={"default": {"E": 1.0, "F": 1.0, "V": 1.0}})
    A1, y1 = ipf.get_lsq_system(db, Vref=None, weights={"default": {"E": 1.0, "F": 1.0, "V": 1.0}, "precon": {"a": pre}})
    fa = np.concatenate([np.arange(*np.add.accumulate(configs[k].rows["F"])) for k in (0, 1)])
    fb = np.concatenate([np.arange(*np.add.accumulate(configs[k].rows["F"])) for k in (2, 3)])
    assert np.array_equal(A0[fb], A1[fb]) and not np.allclose(A0[fa], A1[fa])
    others = np.setdiff1d(np.arange(A0.shape[0]), fa)
    assert np.array_equal(A0[others], A1[others])
