"""K1 parity: the CUDA linked-cell neighbour list is bit-exact against the oracle's
brute-force enumeration (indices, image shifts AND bond vectors), through the C-ABI."""
import numpy as np
import pytest

import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import _capi, synth
import ipf_oracle as O

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.Generator(np.random.Philox(key=11))
    r0 = ipf.rnn("Si")
    out = []
    # 8-atom cell: rcut > L, many images incl. self-images
    out.append(("si8", synth.rattle(synth.bulk("Si"), 0.1, rng, 0.02), 2.5 * r0))
    out.append(("si64", synth.rattle(synth.repeat(synth.bulk("Si"), 2), 0.3, rng, 0.02), 2.5 * r0))
    out.append(("si64_perfect", synth.repeat(synth.bulk("Si"), 2), 2.5 * r0))
    out.append(("si216", synth.rattle(synth.repeat(synth.bulk("Si"), 3), 0.2, rng, 0.02), 2.0 * r0))
    # strongly triclinic
    at = synth.rattle(synth.repeat(synth.bulk("Si"), 2), 0.2, rng, 0.0)
    Fm = np.array([[1.0, 0.3, -0.2], [0.0, 1.0, 0.25], [0.0, 0.0, 1.0]])
    out.append(("triclinic", ipf.Atoms(at.X @ Fm, at.cell @ Fm, at.pbc, at.Z), 2.5 * r0))
    # unwrapped positions (atoms far outside the cell)
    at = synth.rattle(synth.repeat(synth.bulk("Si"), 2), 0.2, rng, 0.01)
    X = at.X + rng.integers(-3, 4, size=at.X.shape) @ at.cell
    out.append(("unwrapped", ipf.Atoms(X, at.cell, at.pbc, at.Z), 2.5 * r0))
    # mixed / no pbc
    at = synth.rattle(synth.repeat(synth.bulk("W"), 3), 0.1, rng, 0.01)
    out.append(("slab", ipf.Atoms(at.X, at.cell, (True, True, False), at.Z), 6.0))
    out.append(("cluster", ipf.Atoms(at.X, at.cell, (False, False, False), at.Z), 6.0))
    # single atom, periodic: only self images
    out.append(("single", ipf.Atoms([[0.1, 0.2, 0.3]], np.eye(3) * 3.0), 6.5))
    return out


@pytest.mark.parametrize("name,at,rcut", _cases(), ids=[c[0] for c in _cases()])
def test_neighbour_list_bit_exact(ctx, name, at, rcut):
    I0, J0, S0, R0 = O.neighbours(at.X, at.cell, at.pbc, rcut)
    I1, J1, S1, R1 = _capi.neighbours(ctx, at.X, at.cell, at.pbc, rcut)
    assert len(I0) > 0
    assert len(I1) == len(I0)
    assert np.array_equal(I1, I0)
    assert np.array_equal(J1, J0)
    assert np.array_equal(S1, S0)
    assert np.array_equal(R1.view(np.uint64), R0.view(np.uint64))    # bit-exact bond vectors


def test_neighbour_list_empty(ctx):
    at = ipf.Atoms([[0.0, 0.0, 0.0], [10.0, 0.0, 0.0]], np.eye(3) * 50.0, (False, False, False))
    I, J, S, R = _capi.neighbours(ctx, at.X, at.cell, at.pbc, 3.0)
    assert len(I) == 0


def test_bad_arguments(ctx):
    with pytest.raises(_capi.IPFError):
        _capi.neighbours(ctx, np.zeros((2, 3)), np.zeros((3, 3)), (True, True, True), 3.0)   # singular cell
