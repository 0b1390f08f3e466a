"""numpy prototype of the factorised 4-body own-leg algorithm (csrc/basis_4b.cu), checked against the CPU oracle.
Development aid: documents the formulas and the table construction the CUDA path uses."""
import sys, os, itertools, math
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import synth
import ipf_oracle as O


def mults(c):
    """multi-indices nu with |nu| = c and multinomial coefficients c!/nu!"""
    out = []
    for n0 in range(c + 1):
        for n1 in range(c - n0 + 1):
            n2 = c - n0 - n1
            out.append(((n0, n1, n2), math.factorial(c) // (math.factorial(n0) * math.factorial(n1) * math.factorial(n2))))
    return out


def eval_site_4b(blk, nbrs):
    """nbrs: list of dict(R, r, Rh, x, dx, fc, dfc, j).  Returns per own leg: g[nb,3], e[nb]."""
    n = len(nbrs)
    D = blk.maxdeg_total
    G = np.zeros((n, blk.nb, 3)); Eo = np.zeros((n, blk.nb))
    for jj, L in enumerate(nbrs):
        Rj = L["Rh"]
        ax = np.array([1.0, 0, 0]) if abs(Rj[0]) < 0.8 else np.array([0, 1.0, 0])
        e1 = np.cross(Rj, ax); e1 /= np.linalg.norm(e1); e2 = np.cross(Rj, e1)
        # moments
        M = {}; Z2 = {}; Z2q = {}
        for kk, K in enumerate(nbrs):
            if kk == jj:
                continue
            u = np.array([Rj @ K["Rh"], e1 @ K["Rh"], e2 @ K["Rh"]])
            for e in range(D + 1):
                xe = K["fc"] * K["x"] ** e
                for m0 in range(D + 1 - e):
                    for m1 in range(D + 1 - e - m0):
                        for m2 in range(D + 1 - e - m0 - m1):
                            M[(e, m0, m1, m2)] = M.get((e, m0, m1, m2), 0.0) + xe * u[0] ** m0 * u[1] ** m1 * u[2] ** m2
                for c in range(D + 1 - e):
                    z = K["fc"] ** 2 * K["x"] ** e * u[0] ** c
                    Z2[(e, c)] = Z2.get((e, c), 0.0) + z
                    Z2q[(e, c)] = Z2q.get((e, c), np.zeros(2)) + z * u[1:]
        xj = L["x"]
        for b in range(blk.nb):
            S0 = S1 = 0.0; T = np.zeros(2)
            for m in blk.orbits[b]:
                a1, a2, a3, c12, c13, c23 = m
                s = 0.0; t = np.zeros(2)
                for nu, kap in mults(c23):
                    Bv = kap * M[(a3, c13 + nu[0], nu[1], nu[2])]
                    s += M[(a2, c12 + nu[0], nu[1], nu[2])] * Bv
                    if c12 > 0:
                        t[0] += M[(a2, c12 - 1 + nu[0], nu[1] + 1, nu[2])] * Bv
                        t[1] += M[(a2, c12 - 1 + nu[0], nu[1], nu[2] + 1)] * Bv
                s -= Z2[(a2 + a3, c12 + c13)]
                if c12 > 0:
                    t -= Z2q[(a2 + a3, c12 + c13 - 1)]
                    t *= c12
                S0 += 0.5 * xj ** a1 * s
                if a1 > 0:
                    S1 += 0.5 * a1 * xj ** (a1 - 1) * s
                T += xj ** a1 * t
            radial = L["dfc"] * S0 + L["fc"] * L["dx"] * S1
            G[jj, b] = radial * Rj + (L["fc"] / L["r"]) * (T[0] * e1 + T[1] * e2)
            Eo[jj, b] = L["fc"] * S0 / 3.0
    return G, Eo


def main():
    deg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    r0 = ipf.rnn("Si")
    desc = ipf.BondAngleDesc(f"exp(- (r/{r0} - 1.0))", ipf.CosCut(1.8 * r0 - 1.0, 1.8 * r0))
    B = ipf.as_basis(ipf.nbpolys(4, desc, deg))
    blk = B.blocks[0]
    blk.orbits = [p.orbit for p in B.polys]
    blk.maxdeg_total = deg
    cfg = synth.rattled_configs("Si", 1, 1, seed=3, calc=synth.StillingerWeber())[0]
    at = cfg.at
    Eref, Fref, Vref = O.eval_config(blk, at)
    I, J, S, R = O.neighbours(at.X, at.cell, at.pbc, desc.rcut)
    nat = len(at)
    E = np.zeros(blk.nb); F = np.zeros((blk.nb, nat, 3)); V = np.zeros((blk.nb, 3, 3))
    for i in range(nat):
        idx = np.where(I == i)[0]
        nbrs = []
        for p in idx:
            r = np.linalg.norm(R[p]); x = float(desc.x(r)); fc = float(desc.cutoff(r))
            dx = -desc.a / desc.r0 * x
            rc1, rc2 = desc.cutoff.rc1, desc.cutoff.rc2
            dfc = 0.0 if (r <= rc1 or r >= rc2) else -0.5 * np.pi / (rc2 - rc1) * np.sin(np.pi * (r - rc1) / (rc2 - rc1))
            nbrs.append(dict(R=R[p], r=r, Rh=R[p] / r, x=x, dx=dx, fc=fc, dfc=dfc, j=J[p]))
        G, Eo = eval_site_4b(blk, nbrs)
        for jj, L in enumerate(nbrs):
            E += Eo[jj]
            F[:, L["j"], :] -= G[jj]
            F[:, i, :] += G[jj]
            V -= G[jj][:, :, None] * L["R"][None, None, :]
    sc = lambda a: np.abs(a).max()
    print("nb", blk.nb, "E err", sc(E - Eref) / sc(Eref), "F err", sc(F.reshape(blk.nb, -1) - Fref) / sc(Fref),
          "V err", sc(V - Vref) / sc(Vref))


if __name__ == "__main__":
    main()
