"""CPU tests of the oracle itself (finite differences, invariances, neighbour symmetry)."""
import numpy as np
import pytest

import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import synth
import ipf_oracle as O
from util import si_desc


@pytest.fixture(scope="module")
def at():
    return synth.rattled_configs("Si", 1, 1, seed=3)[0].at


def _blocks():
    B = ipf.NBodyBasis(ipf.nbpolys(2, si_desc(), 6) + ipf.nbpolys(3, si_desc(2.0), 5) + ipf.nbpolys(4, si_desc(2.0), 3))
    return B.blocks


@pytest.mark.parametrize("blk", _blocks(), ids=lambda b: f"{b.N}B")
def test_forces_and_virial_match_finite_differences(at, blk):
    E, F, V = O.eval_config(blk, at)
    h = 1e-5
    for (k, d) in ((0, 0), (3, 1), (7, 2)):
        ap, am = at.copy(), at.copy()
        ap.X[k, d] += h
        am.X[k, d] -= h
        fd = -(O.eval_config(blk, ap)[0] - O.eval_config(blk, am)[0]) / (2 * h)
        assert np.allclose(fd, F[:, 3 * k + d], rtol=1e-6, atol=1e-7 * np.abs(F).max())
    for (p, q) in ((0, 0), (0, 1), (2, 1)):
        def strained(e):
            Fm = np.eye(3)
            Fm[q, p] += e
            a = at.copy()
            a.X = at.X @ Fm
            a.cell = at.cell @ Fm
            return a
        fd = -(O.eval_config(blk, strained(h))[0] - O.eval_config(blk, strained(-h))[0]) / (2 * h)
        assert np.allclose(fd, V[:, p, q], rtol=1e-6, atol=1e-7 * np.abs(V).max())
    # net force vanishes, virial is symmetric
    assert np.abs(F.reshape(blk.nb, -1, 3).sum(1)).max() < 1e-9 * max(1.0, np.abs(F).max())
    assert np.abs(V - V.transpose(0, 2, 1)).max() < 1e-12 * np.abs(V).max()


def test_invariances(at):
    blk = _blocks()[1]
    E0, F0, V0 = O.eval_config(blk, at)
    # translation by a lattice-incommensurate vector; permutation of atoms
    a = at.copy()
    a.X = a.X + np.array([0.37, -1.2, 2.9])
    assert np.allclose(O.eval_config(blk, a)[0], E0, rtol=1e-11)
    perm = np.random.default_rng(0).permutation(len(at))
    a = ipf.Atoms(at.X[perm], at.cell, at.pbc, at.Z[perm])
    E1, F1, _ = O.eval_config(blk, a)
    assert np.allclose(E1, E0, rtol=1e-11)
    assert np.allclose(F1.reshape(blk.nb, -1, 3), F0.reshape(blk.nb, -1, 3)[:, perm], rtol=1e-9, atol=1e-12)
    # rotation of positions and cell
    th = 0.7
    Q = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    a = ipf.Atoms(at.X @ Q.T, at.cell @ Q.T, at.pbc, at.Z)
    assert np.allclose(O.eval_config(blk, a)[0], E0, rtol=1e-11)


def test_neighbour_list_symmetric_and_sorted(at):
    I, J, S, R = O.neighbours(at.X, at.cell, at.pbc, 2.5 * ipf.rnn("Si"))
    key = list(zip(I.tolist(), J.tolist(), map(tuple, S.tolist())))
    assert key == sorted(key)
    fwd = {(i, j, s): tuple(r) for (i, j, s), r in zip(key, R.tolist())}
    for (i, j, s), r in fwd.items():
        rr = fwd[(j, i, tuple(-x for x in s))]
        assert rr == tuple(-x for x in r)


def test_reference_mode_equals_fused(at):
    """One task per (config, okey) as in src/obsiter.jl:87-102 gives the same matrix as the fused mode."""
    configs = synth.rattled_configs("Si", 1, 3, seed=9, calc=synth.StillingerWeber())
    from ipfitting_jl_b200.lsq_db import _alloc_rows
    from util import rows_for
    n = _alloc_rows(configs)
    B = ipf.NBodyBasis(ipf.nbpolys(2, si_desc(), 4) + ipf.nbpolys(3, si_desc(2.0), 3))
    r = rows_for(configs)
    A0 = O.assemble(B, configs, r["E"], r["F"], r["V"], n, mode=0)
    A1 = O.assemble(B, configs, r["E"], r["F"], r["V"], n, mode=1)
    assert np.array_equal(A0, A1)


def test_label_calculators_consistent():
    """SW / LJ labels: forces and virial are the derivatives of the energy."""
    for calc, sp in ((synth.StillingerWeber(), "Si"), (synth.PairPotential(ipf.rnn("Cu")), "Cu")):
        at = synth.rattled_configs(sp, 2, 1, seed=4)[0].at
        E, F, V = calc.efv(at)
        h = 1e-5
        for (k, d) in ((1, 0), (5, 2)):
            ap, am = at.copy(), at.copy()
            ap.X[k, d] += h
            am.X[k, d] -= h
            fd = -(calc.efv(ap)[0] - calc.efv(am)[0]) / (2 * h)
            assert abs(fd - F[k, d]) < 1e-6 * max(1.0, np.abs(F).max())
        Fm = np.eye(3)
        Fm[1, 0] += h
        ap = ipf.Atoms(at.X @ Fm, at.cell @ Fm, at.pbc, at.Z)
        Fm[1, 0] -= 2 * h
        am = ipf.Atoms(at.X @ Fm, at.cell @ Fm, at.pbc, at.Z)
        fd = -(calc.efv(ap)[0] - calc.efv(am)[0]) / (2 * h)
        assert abs(fd - V[0, 1]) < 1e-5 * max(1.0, np.abs(V).max())
        assert np.abs(F.sum(0)).max() < 1e-9


# ---------------------------------------------------------------- species-resolved blocks (JuLIP Atoms.Z)
def _sic(seed=3, L=1):
    at = synth.repeat(synth.zincblende(4.36, 14, 6), L)
    return synth.rattle(at, 0.15, np.random.default_rng(seed), 0.02)


def _sic_desc(rcut=4.6):
    return ipf.BondAngleDesc("exp(- (r/1.89 - 1.0))", ipf.CosCut(rcut - 1.0, rcut))


def test_species_blocks_sum_to_the_species_agnostic_basis():
    """Every cluster belongs to exactly one species tuple: the species-resolved functions whose orbit is contained in a
    symmetric orbit add up to the species-agnostic function."""
    at = _sic()
    for N, deg in ((2, 5), (3, 3), (4, 2)):
        B0 = ipf.NBodyBasis(ipf.nbpolys(N, _sic_desc(), deg))
        Bs = ipf.NBodyBasis(ipf.species_nbpolys(N, _sic_desc(), deg, [14, 6]))
        E0, F0, V0 = O.eval_config(B0.blocks[0], at)
        Et, Ft, Vt = np.zeros_like(E0), np.zeros_like(F0), np.zeros_like(V0)
        for blk in Bs.blocks:
            E, F, V = O.eval_config(blk, at)
            for k, p in enumerate(Bs.polys[blk.col0:blk.col0 + blk.nb]):
                for b, q in enumerate(B0.polys):
                    if set(p.orbit) <= set(q.orbit):
                        Et[b] += E[k]; Ft[b] += F[k]; Vt[b] += V[k]
        assert np.abs(Et - E0).max() <= 1e-12 * np.abs(E0).max()
        assert np.abs(Ft - F0).max() <= 1e-12 * np.abs(F0).max()
        assert np.abs(Vt - V0).max() <= 1e-12 * np.abs(V0).max()


def test_species_forces_match_finite_differences():
    at = _sic(seed=5)
    B = ipf.NBodyBasis(ipf.nbpolys(3, _sic_desc(), 3, (14, (6, 14))) + ipf.nbpolys(4, _sic_desc(3.6), 2, (6, (6, 14, 14))))
    h = 1e-5
    for blk in B.blocks:
        E, F, V = O.eval_config(blk, at)
        assert np.abs(E).max() > 0
        for (k, d) in ((0, 0), (5, 2)):
            ap, am = at.copy(), at.copy()
            ap.X[k, d] += h
            am.X[k, d] -= h
            fd = -(O.eval_config(blk, ap)[0] - O.eval_config(blk, am)[0]) / (2 * h)
            assert np.allclose(fd, F[:, 3 * k + d], rtol=1e-6, atol=1e-7 * np.abs(F).max())


# ---------------------------------------------------------------- environment-dependent blocks
def test_env_forces_match_finite_differences():
    """n_i^t-weighted functions (`basis.envpolys`): forces and virials are the derivatives of the energies, including the
    derivative of the smooth coordination number; t = 0 is the plain basis."""
    at = synth.rattled_configs("W", 2, 1, seed=9, vacancies=2)[0].at
    r0 = ipf.rnn("W")
    desc = ipf.BondAngleDesc(f"exp(- (r/{r0} - 1.0))", ipf.CosCut(1.5 * r0 - 0.6, 1.5 * r0))
    B = ipf.NBodyBasis(ipf.envpolys(3, desc, 3, 2, ipf.CosCut(1.1 * r0, 1.45 * r0)))
    assert [b.env_t for b in B.blocks] == [0, 1, 2]
    plain = O.eval_config(ipf.NBodyBasis(ipf.nbpolys(3, desc, 3)).blocks[0], at)
    E0, F0, V0 = O.eval_config(B.blocks[0], at)
    assert np.array_equal(E0, plain[0]) and np.array_equal(F0, plain[1])
    h = 1e-5
    for blk in B.blocks[1:]:
        E, F, V = O.eval_config(blk, at)
        assert np.abs(E).max() > 0 and not np.allclose(E, E0)
        for (k, d) in ((0, 0), (5, 2), (9, 1)):
            ap, am = at.copy(), at.copy()
            ap.X[k, d] += h
            am.X[k, d] -= h
            fd = -(O.eval_config(blk, ap)[0] - O.eval_config(blk, am)[0]) / (2 * h)
            assert np.allclose(fd, F[:, 3 * k + d], rtol=1e-6, atol=1e-7 * np.abs(F).max())
        for (p, q) in ((0, 0), (2, 1)):
            def strained(e):
                Fm = np.eye(3)
                Fm[q, p] += e
                a = at.copy()
                a.X = at.X @ Fm
                a.cell = at.cell @ Fm
                return a
            fd = -(O.eval_config(blk, strained(h))[0] - O.eval_config(blk, strained(-h))[0]) / (2 * h)
            assert np.allclose(fd, V[:, p, q], rtol=1e-6, atol=1e-7 * np.abs(V).max())
