"""Host-side logic (no GPU): basis tables, Dat / observation plumbing, row assignment,
weights, DB files, sharding."""
import json
import os

import numpy as np
import pytest

import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import parallel, synth
from ipfitting_jl_b200.lsq import _fix_weights, _select, collect_observations
from ipfitting_jl_b200.lsq_db import LsqDB, _alloc_rows, flatten_configs
import lsq_oracle as LO
from util import si_desc


def test_basis_sizes_match_survey():
    # 2B d14 -> 14, 3B d11 -> 202, 4B d8 -> 589 (Burnside counts)
    d = si_desc()
    assert len(ipf.nbpolys(2, d, 14)) == 14
    assert len(ipf.nbpolys(3, d, 11)) == 202
    assert len(ipf.nbpolys(4, d, 8)) == 589
    B = ipf.NBodyBasis(ipf.nbpolys(2, d, 14) + ipf.nbpolys(3, d, 11))
    assert len(B) == 216 and [b.N for b in B.blocks] == [2, 3] and B.blocks[1].col0 == 14
    # lower degree is a prefix; orbit members are permutations of each other
    assert ipf.nbpolys(3, d, 5) == ipf.nbpolys(3, d, 11)[:len(ipf.nbpolys(3, d, 5))]
    for p in ipf.nbpolys(4, d, 4):
        assert len({tuple(sorted(m[:3])) + tuple(sorted(m[3:])) for m in p.orbit}) == 1


def test_descriptor_parsing_and_roundtrip():
    r0 = ipf.rnn("Si")
    d = ipf.BondAngleDesc(f"exp(- (r/{r0} - 1.0))", ipf.CosCut(4.0, 5.0))
    assert d.kind == 0 and d.a == 1.0 and abs(d.r0 - r0) < 1e-15
    d2 = ipf.BondAngleDesc(f"exp(- 2.5 * (r/{r0} - 1))", ipf.CosCut(4.0, 5.0))
    assert d2.a == 2.5
    d3 = ipf.BondAngleDesc("(2.5/r)^3", ipf.CosCut(4.0, 5.0))
    assert d3.kind == 1 and d3.a == 3.0 and d3.r0 == 2.5
    with pytest.raises(ValueError):
        ipf.BondAngleDesc("sin(r)", ipf.CosCut(4.0, 5.0))
    B = ipf.NBodyBasis(ipf.nbpolys(2, d, 3) + ipf.nbpolys(3, d2, 2))
    assert ipf.NBodyBasis.read_dict(json.loads(json.dumps(B.write_dict()))) == B
    assert np.allclose(ipf.CosCut(4.0, 5.0)([3.0, 4.5, 5.0, 6.0]), [1.0, 0.5, 0.0, 0.0])


def test_dat_roundtrip_and_equality():
    """`test/test_dat.jl:16-31`."""
    at = synth.repeat(synth.bulk("Si"), 2)
    rng = np.random.default_rng(0)
    E, F, V = rng.random(), rng.random((len(at), 3)), rng.random((3, 3))
    V = V + V.T
    d = ipf.Dat(at, "test", E=E, F=F, V=V)
    assert ipf.energy(d) == E and np.array_equal(ipf.forces(d), F) and np.array_equal(ipf.virial(d), V)
    assert np.array_equal(d.D["V"], [V[0, 0], V[1, 1], V[2, 2], V[2, 1], V[2, 0], V[1, 0]])   # _IV = [1,5,9,6,3,2]
    d2 = ipf.Dat.read_dict(json.loads(json.dumps(d.write_dict())))
    assert d2 == d
    d2.rows["E"] = (5, 1)          # rows do not take part in ==
    assert d2 == d
    d2.D["E"][0] += 1.0
    assert d2 != d


def test_dat_reads_reference_style_rows():
    """`_info.json` stores the explicit 1-based index vector per observation (`src/lsq_db.jl:244-246`); a dict in the
    reference's layout (E: one row, F: nine consecutive rows) loads as blocks, and ours writes the same layout."""
    at = synth.rattled_configs("Si", 1, 1, seed=3)[0].at
    Dd = {"__id__": "IPFitting.Dat", "at": at.write_dict(), "configtype": "a",
          "D": {"E": [1.0], "F": [0.0] * 9}, "rows": {"E": [1], "F": list(range(2, 11))}, "info": {}}
    d = ipf.Dat.read_dict(Dd)
    assert d.rows["E"] == (0, 1) and d.rows["F"] == (1, 9)
    assert d.write_dict()["rows"] == {"E": [1], "F": list(range(2, 11))}
    Dd["rows"]["F"] = [2, 3, 5]
    with pytest.raises(ValueError):
        ipf.Dat.read_dict(Dd)


def test_row_assignment_follows_observation_order():
    """`_alloc_lsq_matrix` (`src/lsq_db.jl:237-250`) with arbitrary key order and missing observations."""
    at = synth.bulk("Si")
    z3 = np.zeros((len(at), 3))
    cfgs = [ipf.Dat(at, "a", E=1.0), ipf.Dat(at, "b", V=np.zeros(6), F=z3, E=0.0, okey_order=("V", "F", "E")),
            ipf.Dat(at, "c", F=z3)]
    n = _alloc_rows(cfgs)
    assert n == 1 + (6 + 24 + 1) + 24
    assert cfgs[0].rows == {"E": (0, 1)}
    assert cfgs[1].rows == {"V": (1, 6), "F": (7, 24), "E": (31, 1)}
    assert cfgs[2].rows == {"F": (32, 24)}
    assert [(k, i) for k, _, i in ipf.observations(cfgs)] == [("E", 0), ("V", 1), ("F", 1), ("E", 1), ("F", 2)]
    off, pos, cell, pbc, Z, rows = flatten_configs(cfgs)
    assert rows["E"].tolist() == [1, 32, 0] and rows["F"].tolist() == [0, 8, 33] and rows["V"].tolist() == [0, 2, 0]
    assert off.tolist() == [0, 8, 16, 24] and pos.shape == (24, 3) and cell.shape == (3, 3, 3)


class _FakeDB:
    def __init__(self, configs):
        self.configs = configs
        self.nrows = _alloc_rows(configs)


@pytest.mark.parametrize("weights", [None, {"default": {"E": 30.0, "F": 1.0}},
                                     {"default": {"E": 1.0, "F": 1.0, "V": 1.0}, "b": {"E": 5.0, "V": 2.0}, "ignore": ["c"]},
                                     {"default": {"E": 1.0, "F": 2.0, "V": 3.0}, "ignore": ["F"]}])
def test_observations_and_weights_match_reference_restatement(weights):
    """collect_observations / _get_weights / row selection (`src/lsq.jl:74-161,270-283`, quirks Q1, Q2, Q4)."""
    sw = synth.StillingerWeber()
    cfgs = (synth.rattled_configs("Si", 1, 2, seed=1, calc=sw, configtype="a")
            + synth.rattled_configs("Si", 1, 2, seed=2, calc=sw, configtype="b")
            + synth.rattled_configs("Si", 1, 2, seed=3, calc=sw, configtype="c"))
    cfgs[1].info["W"] = {"E": 7.0, "V": np.arange(1.0, 7.0)}         # overrides table AND weighthook, drops F
    cfgs[3].info["Vref"] = {"E": np.array([1.5]), "F": np.full(24, 0.25), "V": np.zeros(6)}
    db = _FakeDB(cfgs)
    w = _fix_weights(None if weights is None else dict(weights))
    Y, W, Icfg = collect_observations(db, w, ipf.OneBody(-4.0))
    Y0, W0, I0 = LO.collect_observations(cfgs, db.nrows, LO.fix_weights(weights), -4.0)
    assert np.array_equal(Y, Y0) and np.array_equal(W, W0) and np.array_equal(Icfg, I0)
    Irows, Icols = _select(db, Y, W, Icfg, None, [0, 1, 4])
    keep = (W0 != 0) & np.isin(I0, [0, 1, 4])
    assert np.array_equal(Irows, np.nonzero(keep)[0]) and Icols is None


def test_weights_errors():
    cfgs = synth.rattled_configs("Si", 1, 1, seed=1, calc=synth.StillingerWeber())
    cfgs[0].info["W"] = {"F": np.ones(24)}
    with pytest.raises(ValueError):
        collect_observations(_FakeDB(cfgs), _fix_weights(None), None)
    with pytest.raises(ValueError):
        ipf.lsqfit(_FakeDB(cfgs), weights={}, sigmas={})


def test_db_files_roundtrip(tmp_path):
    """`test/test_lsq_db.jl:46-75`: file names, reload gives identical basis / configs / Psi."""
    cfgs = synth.rattled_configs("Si", 1, 3, seed=5, calc=synth.StillingerWeber())
    B = ipf.nbpolys(2, si_desc(), 4)
    db = LsqDB(str(tmp_path / "db"), B, cfgs, _assemble=False)
    db._Psi_host = np.asfortranarray(np.random.default_rng(1).random((db.nrows, len(B))))
    db.flush()
    assert os.path.exists(str(tmp_path / "db_info.json"))
    assert os.path.exists(str(tmp_path / "db_kron.npy")) or os.path.exists(str(tmp_path / "db_kron.h5"))
    db2 = LsqDB(str(tmp_path / "db"))
    assert db2.basis == db.basis and db2.configs == db.configs
    assert np.array_equal(db2.Psi, db.Psi)
    assert all(a.rows == b.rows for a, b in zip(db.configs, db2.configs))
    db.flush()        # second save backs the old files up instead of overwriting (`_backupfile`)
    assert len(list(tmp_path.iterdir())) == 4
    info = json.load(open(str(tmp_path / "db_info.json")))
    assert info["version"] == 2 and info["configs"][0]["rows"]["E"] == [1]        # 1-based index vectors on disk
    assert info["configs"][0]["rows"]["F"] == list(range(2, 2 + 3 * len(cfgs[0])))
    # README-era argument order
    db3 = LsqDB("", cfgs, B, _assemble=False)
    assert db3.basis == db.basis


def test_filters():
    cfgs = synth.rattled_configs("Si", 1, 4, seed=5)
    cfgs[2].configtype = "x"
    B = ipf.nbpolys(2, si_desc(), 3) + ipf.nbpolys(3, si_desc(), 2)
    db = LsqDB("", B, cfgs, _assemble=False)
    assert ipf.filter_basis(db, lambda b: ipf.bodyorder(b) < 3).tolist() == [0, 1, 2]
    assert ipf.filter_configs(db, lambda d: ipf.configtype(d) == "x").tolist() == [2]


def test_lpt_sharding():
    """Descending-cost assignment (`sortperm(costs, rev=true)`, `src/tools.jl:35`)."""
    costs = [64, 16, 512, 64, 128, 16, 256, 64]
    sh = parallel.shard_configs(costs, 3)
    assert sorted(i for s in sh for i in s) == list(range(len(costs)))
    loads = [sum(costs[i] for i in s) for s in sh]
    assert max(loads) == 512 and min(loads) >= 256
    assert parallel.shard_configs(costs, 1) == [list(range(len(costs)))]
    assert parallel.shard_configs([], 2) == [[], []]


def test_read_xyz_round_trip(tmp_path):
    """`read_xyz` (`src/data.jl:129-243`): keys are matched case-insensitively, `config_type` drives include / exclude,
    missing observations stay missing; an extended-xyz round trip through `write_xyz` is exact."""
    from ipfitting_jl_b200 import data
    cfgs = (synth.rattled_configs("Si", 1, 3, seed=3, calc=synth.StillingerWeber(), configtype="bulk")
            + synth.rattled_configs("Si", 1, 2, seed=4, calc=synth.StillingerWeber(), configtype="liq", virial=False))
    fn = str(tmp_path / "train.xyz")
    data.write_xyz(fn, cfgs, energy_key="DFT_Energy")
    back = data.read_xyz(fn, verbose=False)                      # default key "dft_energy" matches "DFT_Energy"
    assert len(back) == 5 and all(a == b for a, b in zip(cfgs, back))
    assert list(back[0].D) == ["E", "F", "V"] and list(back[3].D) == ["E", "F"]
    assert [d.configtype for d in data.read_xyz(fn, verbose=False, exclude=["liq"])] == ["bulk"] * 3
    assert [d.configtype for d in data.read_xyz(fn, verbose=False, include=["liq"])] == ["liq"] * 2
    assert len(data.read_xyz(fn, verbose=False, index="1:4")) == 3
    assert "E" not in data.read_xyz(fn, verbose=False, energy_key="energy")[0].D
    assert "E" in data.read_xyz(fn, verbose=False, energy_key="energy", auto=True)[0].D     # closest key
    assert len(data.load_data(fn, verbose=False)) == 5
    with pytest.raises(ValueError):
        data.load_data(str(tmp_path / "train.jld2"))


def test_species_and_environment_bases_roundtrip():
    """Species-resolved and environment-dependent runs: block grouping, dict round trip (the `basis` entry of `_info.json`),
    symmetry only among equal-species legs, argument checks."""
    desc = si_desc(2.0)
    B = ipf.NBodyBasis(ipf.species_nbpolys(2, desc, 3, [14, 6]) + ipf.nbpolys(3, desc, 2, (14, (6, 14)))
                       + ipf.envpolys(3, desc, 2, 2, ipf.CosCut(2.0, 3.0)) + ipf.nbpolys(2, desc, 2))
    assert [(b.zc, b.zs, b.env_t) for b in B.blocks] == [(6, (6,), 0), (6, (14,), 0), (14, (6,), 0), (14, (14,), 0),
                                                          (14, (6, 14), 0), (0, (), 0), (0, (), 1), (0, (), 2), (0, (), 0)]
    assert ipf.NBodyBasis.read_dict(json.loads(json.dumps(B.write_dict()))) == B
    # two different neighbour species: no exchange symmetry (all monomials), equal species: the symmetric orbits
    assert len(ipf.nbpolys(3, desc, 4, (14, (6, 14)))) == 34 and len(ipf.nbpolys(3, desc, 4, (14, (6, 6)))) == len(ipf.nbpolys(3, desc, 4))
    with pytest.raises(ValueError):
        ipf.nbpolys(3, desc, 2, (14, (6,)))
    with pytest.raises(ValueError):
        ipf.envpolys(3, desc, 2, 1, ipf.CosCut(2.0, desc.rcut + 0.5))


def test_matrix_p_factor_solver_algebra():
    """`solver["P"]` as a full matrix (`src/lsq.jl:450-458,593`): the n x n post-processing of the factor of Psi equals
    the QR solve of Psi pinv(P)."""
    import scipy.linalg as sla
    from ipfitting_jl_b200 import lsq
    rng = np.random.default_rng(0)
    m, n = 300, 14
    A = rng.standard_normal((m, n)) * np.logspace(0, -4, n)
    y = rng.standard_normal(m)
    P = 2.1 * np.eye(n) - np.diag(np.ones(n - 1), 1) - np.diag(np.ones(n - 1), -1)
    B = np.linalg.pinv(P)
    Q, R = np.linalg.qr(A)
    c, extra = lsq._matrix_p_factor_solver(B)(R, Q.T @ y)
    Qr, Rr = sla.qr(A @ B, mode="economic")
    c_ref = B @ sla.solve_triangular(Rr, Qr.T @ y)
    assert np.linalg.norm(c - c_ref) <= 1e-10 * np.linalg.norm(c_ref)
    R2 = extra["R_factor"]
    assert np.all(np.diag(R2) > 0) and np.allclose(np.tril(R2, -1), 0)
    assert np.abs(np.abs(R2) - np.abs(Rr)).max() <= 1e-12 * np.abs(Rr).max()
    assert lsq._is_full_matrix(P) and not lsq._is_full_matrix(np.diag(np.arange(1.0, 4.0))) and not lsq._is_full_matrix(np.ones(3))


def test_lsqfit_host_logic_with_numpy_stand_in(monkeypatch):
    """The host side of `lsqfit` around the device solve -- full-matrix `solver["P"]`, `saveqr`, regulariser targets --
    with a numpy stand-in for `solve_device` (same signature and return values) on the oracle's Psi."""
    import ipf_oracle as O
    import lsq_oracle as LO
    from ipfitting_jl_b200 import lsq
    from util import rows_for
    cfgs = synth.rattled_configs("Si", 1, 6, seed=22, calc=synth.StillingerWeber(), configtype="small")
    B = ipf.nbpolys(2, si_desc(), 4)
    db = LsqDB("", B, cfgs, _assemble=False)
    r = rows_for(cfgs)
    ref = O.assemble(db.basis, cfgs, r["E"], r["F"], r["V"], db.nrows)
    db._Psi_host = np.asfortranarray(ref)
    monkeypatch.setattr(db, "_context", lambda: None)
    monkeypatch.setattr(type(db), "Psi_dev", property(lambda self: self._Psi_host))

    def stand_in(ctx, Psi, Y, W, Irows=None, Icols=None, Dinv=None, want_R=True, comm=None, precon=None,
                 factor_solver=None, reg=None, exchange="gram"):
        rows = np.arange(len(W)) if Irows is None else Irows
        A0 = (W[:, None] * Psi)[rows]
        y = (W * Y)[rows]
        if Icols is not None:
            A0 = A0[:, Icols]
        if reg is not None and reg[0] is not None:
            A0 = np.vstack([A0, reg[0]]); y = np.concatenate([y, reg[1]])
        A = A0 if Dinv is None else A0 * Dinv
        Q, R = np.linalg.qr(A)
        qty = Q.T @ y
        cs, extra = factor_solver(R, qty) if factor_solver is not None else (np.linalg.solve(R, qty), {})
        c = cs if Dinv is None else cs * Dinv
        info = {"rel_rms": np.linalg.norm(A0 @ c - y) / np.linalg.norm(y), "cond_R_estimate": -1.0}
        info.update(extra)
        return c, R, info

    monkeypatch.setattr(lsq, "solve_device", stand_in)
    n = len(B)
    P = 2.5 * np.eye(n) - np.diag(np.ones(n - 1), 1) - np.diag(np.ones(n - 1), -1)
    w = {"default": {"E": 40.0, "F": 1.0, "V": 0.5}}
    saveqr = {}
    solver = {"solver": "qr", "P": P.copy()}
    IP, info = ipf.lsqfit(db, weights=dict(w), E0=0.0, solver=solver, saveqr=saveqr)
    c_ref, kappa, rel_rms, R = LO.lsqfit_qr(ref, db.configs, w, E0=None, P=P)
    assert np.linalg.norm(info["c"] - c_ref) <= 1e-10 * np.linalg.norm(c_ref)
    assert abs(info["cond_R"] - kappa) <= 1e-9 * kappa and abs(info["rel_rms"] - rel_rms) <= 1e-9 * rel_rms
    assert np.abs(np.abs(saveqr["R"]) - np.abs(R)).max() <= 1e-10 * np.abs(R).max()
    assert "P" not in solver and "P" not in info                      # `delete!(solver, "P")`, src/lsq.jl:460-462
    # regulariser rows: targets appended to the saved right-hand side (`vcat(Y, Yreg)`)
    Preg, q = 0.1 * np.eye(n)[:3], np.array([0.5, -0.25, 0.0])
    saveqr = {}
    IP, info = ipf.lsqfit(db, weights=dict(w), E0=0.0, regularisers=[(Preg, q)], saveqr=saveqr)
    c_ref, _, _, _ = LO.lsqfit_qr(ref, db.configs, w, E0=None, regularisers=[(Preg, q)])
    assert np.linalg.norm(info["c"] - c_ref) <= 1e-10 * np.linalg.norm(c_ref)
    assert np.array_equal(saveqr["Y"][-3:], q)
