"""K2/K3 parity: LsqDB assembly through the C-ABI against the CPU oracle."""
import numpy as np
import pytest

import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import synth
from ipfitting_jl_b200.lsq_db import LsqDB
import ipf_oracle as O
from util import PSI_RTOL, psi_entrywise, psi_errors, rows_for, si_desc

pytestmark = pytest.mark.gpu


def _check(basis, configs, ctx):
    db = LsqDB("", basis, configs, ctx=ctx)
    rows = rows_for(configs)
    ref = O.assemble(db.basis, configs, rows["E"], rows["F"], rows["V"], db.nrows)
    Psi = db.Psi
    assert Psi.shape == ref.shape
    assert np.isfinite(Psi).all()
    err = psi_errors(Psi, ref, configs)
    assert err <= PSI_RTOL, f"max relative error {err:.3e}"
    return db, ref


def test_2b_deg14(ctx):
    configs = synth.rattled_configs("Si", 2, 6, seed=1, calc=synth.StillingerWeber())
    B = ipf.nbpolys(2, si_desc(), 14)
    db, ref = _check(B, configs, ctx)
    assert db.Psi.shape == (6 * 199, 14)


def test_3b(ctx):
    configs = synth.rattled_configs("Si", 2, 3, seed=2, calc=synth.StillingerWeber())
    B = ipf.nbpolys(2, si_desc(), 6) + ipf.nbpolys(3, si_desc(2.0), 6)
    _check(B, configs, ctx)


def test_4b(ctx):
    configs = synth.rattled_configs("Si", 1, 3, seed=3, calc=synth.StillingerWeber())
    B = ipf.nbpolys(3, si_desc(2.0), 4) + ipf.nbpolys(4, si_desc(1.8), 4)
    _check(B, configs, ctx)


def test_5b(ctx):
    configs = synth.rattled_configs("Si", 1, 2, seed=4, calc=synth.StillingerWeber())
    B = ipf.nbpolys(5, si_desc(1.3, 0.5), 3)
    _check(B, configs, ctx)


def test_missing_observations_and_key_order(ctx):
    """A config may hold only E (`test/test_lsq_db.jl:17-18`); key order is free (`src/obsiter.jl:43`)."""
    base = synth.rattled_configs("Si", 1, 4, seed=5, calc=synth.StillingerWeber())
    cfgs = [ipf.Dat(base[0].at, "a", E=base[0].D["E"]),
            ipf.Dat(base[1].at, "b", E=base[1].D["E"], F=base[1].D["F"], V=base[1].D["V"], okey_order=("V", "F", "E")),
            ipf.Dat(base[2].at, "c", F=base[2].D["F"]),
            ipf.Dat(base[3].at, "d", V=base[3].D["V"], E=base[3].D["E"], okey_order=("V", "E"))]
    B = ipf.nbpolys(2, si_desc(), 5) + ipf.nbpolys(3, si_desc(2.0), 3)
    db, _ = _check(B, cfgs, ctx)
    assert cfgs[1].rows["V"][0] < cfgs[1].rows["F"][0] < cfgs[1].rows["E"][0]


def test_mixed_sizes(ctx):
    a = synth.rattled_configs("Si", 1, 2, seed=6, calc=synth.StillingerWeber())
    b = synth.rattled_configs("Si", 2, 2, seed=7, calc=synth.StillingerWeber())
    c = synth.rattled_configs("W", 2, 2, seed=8, calc=synth.PairPotential(ipf.rnn("W")), vacancies=2)
    B = ipf.nbpolys(2, si_desc(), 8) + ipf.nbpolys(3, si_desc(2.0), 4)
    _check(B, [a[0], b[0], c[0], a[1], c[1], b[1]], ctx)


def test_empty(ctx):
    B = ipf.nbpolys(2, si_desc(), 3)
    db = LsqDB("", B, [], ctx=ctx)
    assert db.Psi.shape == (0, 3)


@pytest.mark.parametrize("path", ["cfg", "scratch"])
@pytest.mark.parametrize("D", [2, 3, 5, 8, 11, 14])
def test_3b_specialised_path(ctx, D, path):
    """Canonical nbpolys(3, desc, D) goes through the moment-factorised kernels: the configuration-resident
    kernel (basis_3b_cfg.cu, F accumulators in shared memory) or the HBM-scratch + gather path (basis_3b.cu)."""
    configs = (synth.rattled_configs("Si", 1, 2, seed=30 + D, calc=synth.StillingerWeber())
               + synth.rattled_configs("Si", 2, 1, seed=40 + D, calc=synth.StillingerWeber()))
    B = ipf.nbpolys(3, si_desc(2.5 if D <= 11 else 2.0), D)
    ctx.profile(True); ctx.profile_reset()
    ctx.force_generic(2 if path == "scratch" else 0)
    try:
        _check(B, configs, ctx)
        assert ctx.profile_get("site_deriv_3b")[1] > 0, "specialised 3-body path was not taken"
        if D <= 11:      # D = 14: 64 atoms x 63 functions of the widest unit do not fit in shared memory -> scratch path
            assert (ctx.profile_get("gather_3b")[1] > 0) == (path == "scratch")
    finally:
        ctx.force_generic(0)
        ctx.profile(False)


def test_3b_paths_agree_and_are_reproducible(ctx):
    """Config-resident and scratch kernels sum in different (fixed) orders: equal to rounding; each is bit-reproducible.
    Mixed sizes: single atoms, dimers (one pass, few rows) and 64-atom cells (several passes per configuration)."""
    configs = (synth.rattled_configs("Si", 2, 3, seed=61, calc=synth.StillingerWeber())
               + synth.rattled_configs("Si", 1, 2, seed=62, calc=synth.StillingerWeber(), vacancies=3))
    B = ipf.nbpolys(3, si_desc(), 9)
    a = LsqDB("", B, configs, ctx=ctx).Psi
    b = LsqDB("", B, configs, ctx=ctx).Psi
    assert np.array_equal(a, b)
    ctx.force_generic(2)
    try:
        s = LsqDB("", B, configs, ctx=ctx).Psi
    finally:
        ctx.force_generic(0)
    scale = np.abs(s).max(axis=0)
    assert np.all(np.abs(a - s) <= 1e-12 * scale)


def test_3b_specialised_equals_table_driven(ctx):
    configs = synth.rattled_configs("Si", 2, 3, seed=50, calc=synth.StillingerWeber())
    B = ipf.nbpolys(2, si_desc(), 4) + ipf.nbpolys(3, si_desc(), 7)
    fast = LsqDB("", B, configs, ctx=ctx).Psi
    ctx.force_generic(True)
    try:
        slow = LsqDB("", B, configs, ctx=ctx).Psi
    finally:
        ctx.force_generic(False)
    scale = np.abs(slow).max(axis=0)
    assert (np.abs(fast - slow).max(axis=0) <= 1e-12 * scale).all()


def test_non_canonical_3b_uses_table_driven_path(ctx):
    """A subset / reordering of the 3-body functions is not the canonical basis: generic kernels."""
    configs = synth.rattled_configs("Si", 1, 2, seed=51, calc=synth.StillingerWeber())
    B = ipf.nbpolys(3, si_desc(2.0), 5)[::2]
    ctx.profile(True); ctx.profile_reset()
    _check(B, configs, ctx)
    assert ctx.profile_get("gather_3b")[1] == 0
    ctx.profile(False)


def test_2b_specialised_equals_table_driven(ctx):
    configs = synth.rattled_configs("Si", 2, 2, seed=60, calc=synth.StillingerWeber()) + \
        synth.rattled_configs("W", 2, 2, seed=61, calc=synth.PairPotential(ipf.rnn("W")), vacancies=2)
    B = ipf.nbpolys(2, si_desc(), 14)
    fast = LsqDB("", B, configs, ctx=ctx).Psi
    ctx.force_generic(True)
    try:
        slow = LsqDB("", B, configs, ctx=ctx).Psi
    finally:
        ctx.force_generic(False)
    scale = np.abs(slow).max(axis=0)
    assert (np.abs(fast - slow).max(axis=0) <= 1e-13 * scale).all()


def test_equal_sized_chunks_do_not_share_pair_records(ctx):
    """Two device chunks (> 65536 atoms in total) with the SAME pair count but different geometries: the second
    chunk must get its own pair records (the record cache is keyed on the neighbour-list identity, not its size)."""
    rng = np.random.Generator(np.random.Philox(key=77))
    base = synth.repeat(synth.bulk("Si"), 2)                       # 64 atoms
    X = synth.rattle(base, 0.02, rng)                               # small rattles: neighbour shells stay separated,
    Y = synth.rattle(base, 0.02, rng)                               # so both have the same number of pairs
    from ipfitting_jl_b200.prototypes import Dat
    mk = lambda at: Dat(at, "rand", E=0.0, F=np.zeros((64, 3)), V=np.zeros((3, 3)))
    B = ipf.nbpolys(2, si_desc(), 4) + ipf.nbpolys(3, si_desc(2.0), 4)
    ref = LsqDB("", B, [mk(X), mk(Y)], ctx=ctx).Psi
    assert not np.allclose(ref[:199], ref[199:])
    big = LsqDB("", B, [mk(X) for _ in range(1024)] + [mk(Y) for _ in range(1024)], ctx=ctx).Psi
    assert np.array_equal(big[:199], ref[:199])
    assert np.array_equal(big[1023 * 199:1024 * 199], ref[:199])
    assert np.array_equal(big[1024 * 199:1025 * 199], ref[199:])
    assert np.array_equal(big[2047 * 199:], ref[199:])


def test_3b_odd_sized_configurations_only(ctx):
    """A chunk whose largest configuration has an odd atom count (shared-memory carve-up of the resident kernel) and
    a cell that is periodic along one axis only."""
    configs = synth.rattled_configs("Si", 2, 2, seed=71, calc=synth.StillingerWeber(), vacancies=3)
    for d in configs:
        while len(d.at) % 2 == 0:                      # force odd sizes
            keep = np.arange(len(d.at) - 1)
            d.at = ipf.Atoms(d.at.X[keep], d.at.cell, d.at.pbc, d.at.Z[keep])
            d.D["F"] = d.D["F"][:3 * len(keep)]
    iso = synth.rattled_configs("Si", 1, 1, seed=72, calc=synth.StillingerWeber())[0]
    iso.at = ipf.Atoms(iso.at.X[:7] + 50.0, iso.at.cell * 40.0, (False, True, False), iso.at.Z[:7])
    iso.D["F"] = iso.D["F"][:21]
    B = ipf.nbpolys(3, si_desc(2.0), 6)
    assert all(len(d.at) % 2 == 1 for d in configs + [iso])
    _check(B, configs + [iso], ctx)


@pytest.mark.parametrize("species, transform", [("W", "(2.74/r)^3"), ("Cu", "exp(- 2.0 * (r/2.55 - 1.0))"), ("C", (1, 2.5, 1.54))])
def test_other_transforms_and_lattices(ctx, species, transform):
    """Inverse-power and scaled Morse transforms (`README.md:34-37`), bcc / fcc / diamond neighbourhoods, through the
    fused 2-body and the resident 3-body kernels."""
    r0 = ipf.rnn(species)
    d2 = ipf.BondAngleDesc(transform, ipf.CosCut(2.2 * r0 - 0.7, 2.2 * r0))
    d3 = ipf.BondAngleDesc(transform, ipf.CosCut(1.8 * r0 - 0.5, 1.8 * r0))
    configs = (synth.rattled_configs(species, 2, 2, seed=81, calc=synth.PairPotential(r0), vacancies=1)
               + synth.rattled_configs(species, 1, 2, seed=82, calc=synth.PairPotential(r0)))
    B = ipf.nbpolys(2, d2, 9) + ipf.nbpolys(3, d3, 6)
    ctx.profile(True); ctx.profile_reset()
    try:
        _check(B, configs, ctx)
        assert ctx.profile_get("site_deriv_2b")[1] > 0 and ctx.profile_get("site_deriv_3b")[1] > 0
    finally:
        ctx.profile(False)


def _entrywise(Psi, ref):
    """Largest entry-wise relative error over the entries that are not cancellation residue (|ref| above 1e-6 of the
    column's largest entry of the same observation block is not assumed here: plain |a - b| / |b| where |b| > 0)."""
    nz = np.abs(ref) > 1e-8 * np.abs(ref).max(axis=0, keepdims=True)
    return float((np.abs(Psi - ref)[nz] / np.abs(ref)[nz]).max())


@pytest.mark.parametrize("D", [1, 2, 3, 4, 6, 8])
def test_4b_specialised_path(ctx, D):
    """nbpolys(4, desc, D) goes through the moment-factorised 4-body kernel (basis_4b.cu): small cells with many
    periodic images of the same atoms, vacancies (odd sizes) and an isolated cluster."""
    configs = (synth.rattled_configs("Si", 1, 2, seed=90 + D, calc=synth.StillingerWeber())
               + synth.rattled_configs("Si", 1, 1, seed=95 + D, calc=synth.StillingerWeber(), vacancies=3))
    iso = synth.rattled_configs("Si", 1, 1, seed=72, calc=synth.StillingerWeber())[0]
    iso.at = ipf.Atoms(iso.at.X[:6] + 50.0, iso.at.cell * 40.0, (False, False, False), iso.at.Z[:6])
    iso.D["F"] = iso.D["F"][:18]
    B = ipf.nbpolys(4, si_desc(1.9 if D <= 6 else 1.7), D)
    ctx.profile(True); ctx.profile_reset()
    try:
        db, ref = _check(B, configs + [iso], ctx)
        assert ctx.profile_get("site_deriv_4b")[1] > 0, "specialised 4-body path was not taken"
        assert _entrywise(db.Psi, ref) < 1e-7
    finally:
        ctx.profile(False)


def test_4b_deg8_64_atom_cells(ctx):
    """The north-star block: nbpolys(4, desc, 8) (589 functions) on rattled 64-atom Si cells at the README cutoff
    (2.5 r0, ~34 neighbours per atom) against the oracle (`README.md:41-47`)."""
    configs = synth.rattled_configs("Si", 2, 2, seed=11, calc=synth.StillingerWeber())
    B = ipf.nbpolys(4, si_desc(), 8)
    assert len(B) == 589
    ctx.profile(True); ctx.profile_reset()
    try:
        db, ref = _check(B, configs, ctx)
        assert ctx.profile_get("site_deriv_4b")[1] > 0
        print("4-body deg 8, 64 atoms: block-scaled error", psi_errors(db.Psi, ref, configs),
              "entry-wise (entries above 1e-8 of the column scale)", _entrywise(db.Psi, ref))
    finally:
        ctx.profile(False)


def test_4b_specialised_equals_table_driven_and_is_reproducible(ctx):
    configs = synth.rattled_configs("Si", 1, 3, seed=52, calc=synth.StillingerWeber()) + \
        synth.rattled_configs("W", 2, 1, seed=53, calc=synth.PairPotential(ipf.rnn("W")), vacancies=2)
    B = ipf.nbpolys(3, si_desc(2.0), 4) + ipf.nbpolys(4, si_desc(1.6), 5)
    fast = LsqDB("", B, configs, ctx=ctx).Psi
    again = LsqDB("", B, configs, ctx=ctx).Psi
    assert np.array_equal(fast, again)
    ctx.force_generic(True)
    try:
        slow = LsqDB("", B, configs, ctx=ctx).Psi
    finally:
        ctx.force_generic(False)
    scale = np.abs(slow).max(axis=0)
    assert (np.abs(fast - slow).max(axis=0) <= 1e-11 * scale).all()


def test_4b_subset_of_orbits_uses_the_specialised_path(ctx):
    """Any set of complete orbits (here every other function of the canonical basis, reordered) is served by the
    factorised kernel: its tables are built per function, not per canonical basis."""
    configs = synth.rattled_configs("Si", 1, 2, seed=54, calc=synth.StillingerWeber())
    full = ipf.nbpolys(4, si_desc(1.6), 5)
    B = full[1::2][::-1]
    ctx.profile(True); ctx.profile_reset()
    try:
        _check(B, configs, ctx)
        assert ctx.profile_get("site_deriv_4b")[1] > 0
    finally:
        ctx.profile(False)


def test_entrywise_error_report(ctx):
    """north_star states the 1e-10 bar entry-wise; the parity tests apply it per (column, observation type) block because
    forces of a nearly perfect crystal cancel (any summation order has absolute error eps x sum |terms|).  This test
    reports and bounds the entry-wise figure on a 2B + 3B + 4B basis: every entry that is not a cancellation zero
    (|ref| > 1e-6 of its column's largest entry) agrees to 1e-9 entry-wise, 99.9 % of them to 1e-10."""
    configs = synth.rattled_configs("Si", 2, 3, seed=21, calc=synth.StillingerWeber(), min_dist=0.75)
    B = ipf.nbpolys(2, si_desc(), 14) + ipf.nbpolys(3, si_desc(2.0), 8) + ipf.nbpolys(4, si_desc(1.8), 6)
    db, ref = _check(B, configs, ctx)
    emax, e999, share = psi_entrywise(db.Psi, ref)
    print(f"entry-wise relative error: max {emax:.2e}, 99.9 % quantile {e999:.2e}, entries within 1e-10: {100 * share:.3f} %")
    assert emax <= 1e-9 and e999 <= 1e-10 and share >= 0.9999


def test_species_resolved_blocks(ctx):
    """Two-species cells (zincblende Si/C, JuLIP `Atoms.Z`, `src/prototypes.jl:19-25`): species-resolved 2-, 3- and 4-body
    runs (one per centre species and multiset of neighbour species, polynomials symmetric in equal-species legs only)
    against the oracle, next to species-agnostic runs of the specialised kernels in the same basis."""
    desc = ipf.BondAngleDesc("exp(- (r/1.89 - 1.0))", ipf.CosCut(3.6, 4.6))
    desc4 = ipf.BondAngleDesc("exp(- (r/1.89 - 1.0))", ipf.CosCut(2.6, 3.6))
    rng = np.random.default_rng(11)
    configs = []
    for L in (1, 2, 1):
        at = synth.rattle(synth.repeat(synth.zincblende(4.36, 14, 6), L), 0.12, rng, 0.02)
        nat = len(at)
        configs.append(ipf.Dat(at, "sic", E=0.0, F=np.zeros((nat, 3)), V=np.zeros(6)))
    B = (ipf.species_nbpolys(2, desc, 6, [14, 6]) + ipf.nbpolys(3, desc, 4)
         + ipf.species_nbpolys(3, desc, 3, [14, 6]) + ipf.nbpolys(4, desc4, 2, (6, (6, 14, 14)))
         + ipf.nbpolys(4, desc4, 2, (14, (6, 6, 6))) + ipf.nbpolys(2, desc, 4))
    db, ref = _check(B, configs, ctx)
    assert np.abs(ref).max(axis=0).min() > 0          # every column sees clusters of its species tuple
    # a basis with a species-resolved run needs Z: the specialised (species-agnostic) kernels do not
    assert [b.zc for b in db.basis.blocks][:2] == [6, 6]


def test_environment_dependent_blocks(ctx):
    """BASELINE config 4 in miniature (`README.md:65`, `W5Bdeg12env`): W bcc cells with vacancies, environment-dependent
    3-, 4- and 5-body runs (site function times n_i^t, t = 0..2) against the oracle, mixed with plain runs."""
    r0 = ipf.rnn("W")
    d3 = ipf.BondAngleDesc(f"exp(- (r/{r0} - 1.0))", ipf.CosCut(1.5 * r0 - 0.6, 1.5 * r0))
    d5 = ipf.BondAngleDesc(f"exp(- (r/{r0} - 1.0))", ipf.CosCut(1.25 * r0 - 0.4, 1.25 * r0))
    env = ipf.CosCut(1.05 * r0, 1.2 * r0)
    configs = synth.rattled_configs("W", 2, 3, seed=31, calc=synth.PairPotential(r0), vacancies=2)
    B = (ipf.nbpolys(2, d3, 5) + ipf.envpolys(3, d3, 3, 2, ipf.CosCut(1.1 * r0, 1.45 * r0))
         + ipf.envpolys(4, d5, 2, 1, env) + ipf.envpolys(5, d5, 2, 1, env))
    db, ref = _check(B, configs, ctx)
    assert [b.env_t for b in db.basis.blocks] == [0, 0, 1, 2, 0, 1, 0, 1]
    assert np.abs(ref).max(axis=0).min() > 0
