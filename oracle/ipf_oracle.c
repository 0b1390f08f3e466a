/* ipf_oracle.c -- CPU restatement of the LsqDB assembly hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's shared object.
 *
 * PARITY UNPINNED: the reference (IPFitting.jl v0.10.3) delegates the
 * arithmetic of this path to un-vendored Julia packages (JuLIP.jl >= 0.9.3,
 * NeighbourLists.jl, NBodyIPs.jl; no Manifest.toml => no pinned versions), has
 * no golden vectors, and Julia is not installable here.  This file restates
 * the published algorithm and is anchored on the reference's own call sites:
 *
 *   eval_obs(:E|:F|:V, B, dat) = energy|forces|virial(B, dat.at)
 *                                        src/datatypes.jl:28,40,65
 *   vec_obs F: [3 nat] atom-major xyz    src/datatypes.jl:37-50
 *   vec_obs V: V[[1,5,9,6,3,2]] of the column-major 3x3 = (xx,yy,zz,zy,zx,yx)
 *                                        src/datatypes.jl:57,62,69-76
 *   db.Psi[irows, :] = vec_obs(...)      src/lsq_db.jl:268-278
 *   one task per (config, okey), descending cost, dynamic queue
 *                                        src/obsiter.jl:87-102, src/tools.jl:35-56
 *
 * JuLIP conventions restated (site-energy form, as recalled; test_fit2b.jl:84-87
 * pins "no 1/2 on ordered pairs"):
 *   E_b  = sum_i V_b(i)
 *   F    : for each site i, neighbour a with dV = dV_b(i)/dR_ia (R_ia = X_a - X_i):
 *            F[a] -= dV ;  F[i] += dV
 *   Vir  = - sum_i sum_a dV (x) R_ia       (3x3, Vir[p][q] = -sum dV_p R_q)
 *
 * Algorithmic choices are deliberately different from the CUDA path: brute-force
 * neighbour enumeration over all (j, S) in lexicographic order (no cell list),
 * unordered-cluster loops with the full gradient per cluster (the GPU uses an
 * ordered "own-leg" decomposition), naive power tables.
 *
 * Build:  make -C oracle      (gcc -O2 -fopenmp -ffp-contract=off)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define IPFO_MAX_VARS 10
#define IPFO_MAX_DEG 20
#define IPFO_MAX_LEGS 4

/* ------------------------------------------------------------------ */
/* Neighbour enumeration                                              */
/* ------------------------------------------------------------------ */

static void inv3(const double* C, double* Ci) {
    /* C row-major 3x3; Ci = inverse */
    double a = C[0], b = C[1], c = C[2], d = C[3], e = C[4], f = C[5], g = C[6], h = C[7], k = C[8];
    double det = a * (e * k - f * h) - b * (d * k - f * g) + c * (d * h - e * g);
    double id = 1.0 / det;
    Ci[0] = (e * k - f * h) * id; Ci[1] = (c * h - b * k) * id; Ci[2] = (b * f - c * e) * id;
    Ci[3] = (f * g - d * k) * id; Ci[4] = (a * k - c * g) * id; Ci[5] = (c * d - a * f) * id;
    Ci[6] = (d * h - e * g) * id; Ci[7] = (b * g - a * h) * id; Ci[8] = (a * e - b * d) * id;
}

/* The bond vector, evaluated with individually rounded operations in this exact
 * order (the CUDA kernel uses __dadd_rn/__dmul_rn in the same order):
 *   sh_d = ((Sx*C[0][d] + Sy*C[1][d]) + Sz*C[2][d]);  R_d = (X_j,d - X_i,d) + sh_d
 *   r2   = (Rx*Rx + Ry*Ry) + Rz*Rz
 * (this translation unit is compiled with -ffp-contract=off). */
static inline double bond(const double* pos, const double* C, int i, int j,
                          int sx, int sy, int sz, double* R) {
    for (int d = 0; d < 3; ++d) {
        double sh = ((double)sx * C[0 + d] + (double)sy * C[3 + d]) + (double)sz * C[6 + d];
        R[d] = (pos[3 * j + d] - pos[3 * i + d]) + sh;
    }
    return (R[0] * R[0] + R[1] * R[1]) + R[2] * R[2];
}

/* All (i, j, S) with r2 < rcut^2 and not (i==j, S==0), lexicographic in
 * (i, j, Sx, Sy, Sz).  Returns the total count; writes at most `cap` entries.
 * I, J: int32[cap]; S: int32[3*cap]; R: double[3*cap] (may be NULL). */
int64_t ipfo_neighbours(int32_t nat, const double* pos, const double* cell,
                        const uint8_t* pbc, double rcut, int64_t cap,
                        int32_t* I, int32_t* J, int32_t* S, double* Rout) {
    double Ci[9];
    inv3(cell, Ci);
    /* fractional coordinates s = X * C^{-1} (row vector convention) */
    int nmax[3];
    for (int a = 0; a < 3; ++a) {
        if (!pbc[a]) { nmax[a] = 0; continue; }
        /* height of the cell along direction a = 1 / |column a of C^{-1}| */
        double n2 = Ci[0 + a] * Ci[0 + a] + Ci[3 + a] * Ci[3 + a] + Ci[6 + a] * Ci[6 + a];
        double h = 1.0 / sqrt(n2);
        double smin = 1e300, smax = -1e300;
        for (int i = 0; i < nat; ++i) {
            double s = pos[3 * i] * Ci[0 + a] + pos[3 * i + 1] * Ci[3 + a] + pos[3 * i + 2] * Ci[6 + a];
            if (s < smin) smin = s;
            if (s > smax) smax = s;
        }
        nmax[a] = (int)floor(rcut / h + (smax - smin)) + 2;
    }
    const double rc2 = rcut * rcut;
    int64_t n = 0;
    for (int i = 0; i < nat; ++i)
        for (int j = 0; j < nat; ++j)
            for (int sx = -nmax[0]; sx <= nmax[0]; ++sx)
                for (int sy = -nmax[1]; sy <= nmax[1]; ++sy)
                    for (int sz = -nmax[2]; sz <= nmax[2]; ++sz) {
                        if (i == j && sx == 0 && sy == 0 && sz == 0) continue;
                        double R[3];
                        double r2 = bond(pos, cell, i, j, sx, sy, sz, R);
                        if (r2 < rc2) {
                            if (n < cap) {
                                I[n] = i; J[n] = j;
                                S[3 * n] = sx; S[3 * n + 1] = sy; S[3 * n + 2] = sz;
                                if (Rout) { Rout[3 * n] = R[0]; Rout[3 * n + 1] = R[1]; Rout[3 * n + 2] = R[2]; }
                            }
                            ++n;
                        }
                    }
    return n;
}

/* ------------------------------------------------------------------ */
/* Descriptor                                                          */
/* ------------------------------------------------------------------ */
typedef struct {
    int transform;          /* 0: x = exp(-a (r/r0 - 1));  1: x = (r0/r)^a */
    double a, r0, rc1, rc2;
} desc_t;

static const double IPFO_PI = 3.14159265358979323846;

static void desc_eval(const desc_t* D, double r, double* x, double* dx, double* fc, double* dfc) {
    if (D->transform == 0) {
        *x = exp(-D->a * (r / D->r0 - 1.0));
        *dx = -D->a / D->r0 * (*x);
    } else {
        *x = pow(D->r0 / r, D->a);
        *dx = -D->a * (*x) / r;
    }
    if (r <= D->rc1) { *fc = 1.0; *dfc = 0.0; }
    else if (r >= D->rc2) { *fc = 0.0; *dfc = 0.0; }
    else {
        double w = D->rc2 - D->rc1;
        double s = (r - D->rc1) / w;
        *fc = 0.5 * (1.0 + cos(IPFO_PI * s));
        *dfc = -0.5 * IPFO_PI / w * sin(IPFO_PI * s);
    }
}

/* ------------------------------------------------------------------ */
/* Per-config evaluation of one basis block                            */
/* ------------------------------------------------------------------ */
typedef struct {
    int N, nb, nmono;
    const int32_t* mono_b;      /* [nmono] */
    const uint8_t* mono_exp;    /* [nmono][IPFO_MAX_VARS] */
    desc_t D;
    /* species-resolved block (JuLIP Atoms.Z): the functions act on the clusters whose centre has species zc and whose
     * N-1 neighbours have exactly the (sorted) species zs, the legs assigned to the variable slots in species order
     * (legs of equal species in enumeration order: the polynomial is symmetric in them).  zc = 0: every cluster. */
    int zc, zs[IPFO_MAX_LEGS];
    /* environment-dependent block (NBodyIPs EnvIP as recalled; README.md:65 names a 5-body `env` database): the site
     * function is multiplied by n_i^env_t with n_i = sum_j fenv(r_ij), fenv = CosCut(env_rc1, env_rc2).  env_t = 0: none. */
    int env_t;
    double env_rc1, env_rc2;
} block_t;

typedef struct {
    int j;            /* neighbour atom index */
    double R[3], r, Rh[3], x, dx, fc, dfc;
} nbr_t;

/* E[nb], F[nb][3*nat] (basis-major), V[nb][9] (row-major 3x3) are ACCUMULATED.
 * what: bit0 = E, bit1 = F, bit2 = V. */
static double g_unused_scale = 1.0;
static void eval_cluster_s(const block_t* B, int i, const nbr_t** leg, int what,
                           int nat, double* E, double* F, double* V, double* scratch, double scale);
static void eval_cluster(const block_t* B, int i, const nbr_t** leg, int what,
                         int nat, double* E, double* F, double* V, double* scratch) {
    (void)g_unused_scale;
    eval_cluster_s(B, i, leg, what, nat, E, F, V, scratch, 1.0);
}
/* scale: common factor of the cluster's contribution (n_i^t of an environment-dependent block) */
static void eval_cluster_s(const block_t* B, int i, const nbr_t** leg, int what,
                           int nat, double* E, double* F, double* V, double* scratch, double scale) {
    const int nx = B->N - 1;
    const int npair = nx * (nx - 1) / 2;
    const int nv = nx + npair;
    double var[IPFO_MAX_VARS];
    int pa[6], pc[6];
    for (int a = 0; a < nx; ++a) var[a] = leg[a]->x;
    {
        int k = 0;
        for (int a = 0; a < nx; ++a)
            for (int c = a + 1; c < nx; ++c) {
                pa[k] = a; pc[k] = c;
                var[nx + k] = leg[a]->Rh[0] * leg[c]->Rh[0] + leg[a]->Rh[1] * leg[c]->Rh[1] +
                              leg[a]->Rh[2] * leg[c]->Rh[2];
                ++k;
            }
    }
    /* power tables */
    double pw[IPFO_MAX_VARS][IPFO_MAX_DEG + 1];
    for (int v = 0; v < nv; ++v) {
        pw[v][0] = 1.0;
        for (int e = 1; e <= IPFO_MAX_DEG; ++e) pw[v][e] = pw[v][e - 1] * var[v];
    }
    double FC = 1.0, FCwo[IPFO_MAX_LEGS];
    for (int a = 0; a < nx; ++a) FC *= leg[a]->fc;
    for (int a = 0; a < nx; ++a) {
        double p = 1.0;
        for (int c = 0; c < nx; ++c) if (c != a) p *= leg[c]->fc;
        FCwo[a] = p * scale;
    }
    FC *= scale;
    /* P[b], dP[b][v] */
    double* P = scratch;
    double* dP = scratch + B->nb;
    memset(scratch, 0, sizeof(double) * (size_t)B->nb * (1 + IPFO_MAX_VARS));
    for (int m = 0; m < B->nmono; ++m) {
        const uint8_t* ex = B->mono_exp + (size_t)m * IPFO_MAX_VARS;
        const int b = B->mono_b[m];
        /* prefix/suffix products give every leave-one-out product in O(nv) */
        double pre[IPFO_MAX_VARS + 1], suf[IPFO_MAX_VARS + 1];
        pre[0] = 1.0;
        for (int v = 0; v < nv; ++v) pre[v + 1] = pre[v] * pw[v][ex[v]];
        P[b] += pre[nv];
        if (what & 6) {
            suf[nv] = 1.0;
            for (int v = nv - 1; v >= 0; --v) suf[v] = suf[v + 1] * pw[v][ex[v]];
            for (int v = 0; v < nv; ++v) {
                if (ex[v] == 0) continue;
                double d = (double)ex[v] * pw[v][ex[v] - 1];
                dP[(size_t)b * IPFO_MAX_VARS + v] += d * (pre[v] * suf[v + 1]);
            }
        }
    }
    for (int b = 0; b < B->nb; ++b) {
        if (what & 1) E[b] += FC * P[b];
        if (!(what & 6)) continue;
        const double* g = dP + (size_t)b * IPFO_MAX_VARS;
        for (int a = 0; a < nx; ++a) {
            const nbr_t* L = leg[a];
            double radial = L->dfc * FCwo[a] * P[b] + FC * g[a] * L->dx;
            double dV[3] = { radial * L->Rh[0], radial * L->Rh[1], radial * L->Rh[2] };
            for (int k = 0; k < npair; ++k) {
                int o;
                if (pa[k] == a) o = pc[k]; else if (pc[k] == a) o = pa[k]; else continue;
                double w = var[nx + k];
                double coef = FC * g[nx + k] / L->r;
                for (int d = 0; d < 3; ++d) dV[d] += coef * (leg[o]->Rh[d] - w * L->Rh[d]);
            }
            if (what & 2) {
                double* Fb = F + (size_t)b * 3 * nat;
                for (int d = 0; d < 3; ++d) { Fb[3 * L->j + d] -= dV[d]; Fb[3 * i + d] += dV[d]; }
            }
            if (what & 4) {
                double* Vb = V + (size_t)b * 9;
                for (int p = 0; p < 3; ++p)
                    for (int q = 0; q < 3; ++q) Vb[3 * p + q] -= dV[p] * L->R[q];
            }
        }
    }
}

/* species rule for one unordered cluster: returns without contribution when the neighbour species do not match */
static void eval_cluster_z(const block_t* B, const int32_t* Z, int i, const nbr_t** leg, int what,
                           int nat, double* E, double* F, double* V, double* scratch, double scale) {
    if (B->zc == 0 || Z == NULL) { eval_cluster_s(B, i, leg, what, nat, E, F, V, scratch, scale); return; }
    const int nx = B->N - 1;
    const nbr_t* srt[IPFO_MAX_LEGS];
    for (int a = 0; a < nx; ++a) srt[a] = leg[a];
    for (int a = 1; a < nx; ++a) {                 /* stable insertion sort by species */
        const nbr_t* t = srt[a];
        int c = a - 1;
        while (c >= 0 && Z[srt[c]->j] > Z[t->j]) { srt[c + 1] = srt[c]; --c; }
        srt[c + 1] = t;
    }
    for (int a = 0; a < nx; ++a) if (Z[srt[a]->j] != B->zs[a]) return;
    eval_cluster_s(B, i, srt, what, nat, E, F, V, scratch, scale);
}

/* all clusters of site i (neighbours base[0..n)) */
static void site_clusters(const block_t* B, const int32_t* Z, int i, const nbr_t* base, int n, int what, int nat,
                          double* E, double* F, double* V, double* scratch, double scale) {
    const int nx = B->N - 1;
    const nbr_t* leg[IPFO_MAX_LEGS];
    if (nx == 1) {
        for (int a = 0; a < n; ++a) { leg[0] = base + a; eval_cluster_z(B, Z, i, leg, what, nat, E, F, V, scratch, scale); }
    } else if (nx == 2) {
        for (int a = 0; a < n; ++a)
            for (int c = a + 1; c < n; ++c) {
                leg[0] = base + a; leg[1] = base + c;
                eval_cluster_z(B, Z, i, leg, what, nat, E, F, V, scratch, scale);
            }
    } else if (nx == 3) {
        for (int a = 0; a < n; ++a)
            for (int c = a + 1; c < n; ++c)
                for (int e = c + 1; e < n; ++e) {
                    leg[0] = base + a; leg[1] = base + c; leg[2] = base + e;
                    eval_cluster_z(B, Z, i, leg, what, nat, E, F, V, scratch, scale);
                }
    } else if (nx == 4) {
        for (int a = 0; a < n; ++a)
            for (int c = a + 1; c < n; ++c)
                for (int e = c + 1; e < n; ++e)
                    for (int g = e + 1; g < n; ++g) {
                        leg[0] = base + a; leg[1] = base + c; leg[2] = base + e; leg[3] = base + g;
                        eval_cluster_z(B, Z, i, leg, what, nat, E, F, V, scratch, scale);
                    }
    }
}

static void eval_config_block_par(const block_t* B, int nat, const double* pos, const double* cell,
                                  const uint8_t* pbc, const int32_t* Z, int what, double* E, double* F, double* V, int par) {
    const double rcut = B->D.rc2;
    int64_t cap = ipfo_neighbours(nat, pos, cell, pbc, rcut, 0, NULL, NULL, NULL, NULL);
    int32_t* I = (int32_t*)malloc(sizeof(int32_t) * (size_t)(cap + 1));
    int32_t* J = (int32_t*)malloc(sizeof(int32_t) * (size_t)(cap + 1));
    int32_t* S = (int32_t*)malloc(sizeof(int32_t) * 3 * (size_t)(cap + 1));
    double* R = (double*)malloc(sizeof(double) * 3 * (size_t)(cap + 1));
    ipfo_neighbours(nat, pos, cell, pbc, rcut, cap, I, J, S, R);
    nbr_t* nb = (nbr_t*)malloc(sizeof(nbr_t) * (size_t)(cap + 1));
    for (int64_t p = 0; p < cap; ++p) {
        nbr_t* q = nb + p;
        q->j = J[p];
        double r2 = (R[3 * p] * R[3 * p] + R[3 * p + 1] * R[3 * p + 1]) + R[3 * p + 2] * R[3 * p + 2];
        q->r = sqrt(r2);
        for (int d = 0; d < 3; ++d) { q->R[d] = R[3 * p + d]; q->Rh[d] = R[3 * p + d] / q->r; }
        desc_eval(&B->D, q->r, &q->x, &q->dx, &q->fc, &q->dfc);
    }
    const int nx = B->N - 1;
    /* first pair of every site */
    int64_t* soff = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nat + 1));
    {
        int64_t q = 0;
        for (int i = 0; i < nat; ++i) { soff[i] = q; while (q < cap && I[q] == i) ++q; }
        soff[nat] = q;
    }
    /* par != 0: the sites are shared out over the OpenMP threads (dynamic), every thread accumulates into private
     * E/F/V copies that are added up in thread order at the end (bench.py's CPU arms on a small sample of
     * configurations: keeps all host threads busy; the sums differ from the serial order by rounding only) */
#pragma omp parallel if (par)
    {
    double *Et = E, *Ft = F, *Vt = V;
    int priv = 0;
#ifdef _OPENMP
    if (par && omp_get_num_threads() > 1) {
        priv = 1;
        Et = (double*)calloc((size_t)B->nb, sizeof(double));
        Ft = (double*)calloc((size_t)B->nb * 3 * (size_t)nat + 1, sizeof(double));
        Vt = (double*)calloc((size_t)B->nb * 9, sizeof(double));
    }
#endif
    double* scratch = (double*)malloc(sizeof(double) * (size_t)B->nb * (1 + IPFO_MAX_VARS));
#pragma omp for schedule(dynamic, 1)
    for (int i = 0; i < nat; ++i) {
        double *E = Et, *F = Ft, *V = Vt;
        if (B->zc != 0 && Z != NULL && Z[i] != B->zc) continue;
        const int64_t p0 = soff[i], p1 = soff[i + 1];
        const int n = (int)(p1 - p0);
        const nbr_t* base = nb + p0;
        if (B->env_t <= 0) {
            site_clusters(B, Z, i, base, n, what, nat, E, F, V, scratch, 1.0);
        } else {
            /* environment factor n_i^t: V0_b(i) from an energy-only pass, then the scaled clusters and the derivative of
             * the factor, dV = t n^(t-1) fenv'(r_ij) Rhat_ij V0_b(i) for every neighbour inside the environment cutoff */
            double ni = 0.0;
            const double we = B->env_rc2 - B->env_rc1;
            for (int a = 0; a < n; ++a) {
                const double r = base[a].r;
                if (r <= B->env_rc1) ni += 1.0;
                else if (r < B->env_rc2) ni += 0.5 * (1.0 + cos(IPFO_PI * (r - B->env_rc1) / we));
            }
            double nt = 1.0, ntm1 = 1.0;
            for (int k = 0; k < B->env_t; ++k) { ntm1 = nt; nt *= ni; }
            ntm1 *= (double)B->env_t;
            double* V0 = (double*)calloc((size_t)B->nb, sizeof(double));
            site_clusters(B, Z, i, base, n, 1, nat, V0, NULL, NULL, scratch, 1.0);
            site_clusters(B, Z, i, base, n, what, nat, E, F, V, scratch, nt);
            if (what & 6) {
                for (int a = 0; a < n; ++a) {
                    const nbr_t* L = base + a;
                    if (!(L->r > B->env_rc1 && L->r < B->env_rc2)) continue;
                    const double dfe = -0.5 * IPFO_PI / we * sin(IPFO_PI * (L->r - B->env_rc1) / we);
                    for (int b = 0; b < B->nb; ++b) {
                        const double c = ntm1 * dfe * V0[b];
                        const double dV[3] = { c * L->Rh[0], c * L->Rh[1], c * L->Rh[2] };
                        if (what & 2) {
                            double* Fb = F + (size_t)b * 3 * nat;
                            for (int d = 0; d < 3; ++d) { Fb[3 * L->j + d] -= dV[d]; Fb[3 * i + d] += dV[d]; }
                        }
                        if (what & 4) {
                            double* Vb = V + (size_t)b * 9;
                            for (int p = 0; p < 3; ++p)
                                for (int q = 0; q < 3; ++q) Vb[3 * p + q] -= dV[p] * L->R[q];
                        }
                    }
                }
            }
            free(V0);
        }
    }
    free(scratch);
    if (priv) {
#pragma omp critical(ipfo_site_reduce)
        {
            for (int b = 0; b < B->nb; ++b) E[b] += Et[b];
            for (size_t k = 0; k < (size_t)B->nb * 3 * (size_t)nat; ++k) F[k] += Ft[k];
            for (int k = 0; k < B->nb * 9; ++k) V[k] += Vt[k];
        }
        free(Et); free(Ft); free(Vt);
    }
    }
    free(soff); free(nb); free(R); free(S); free(J); free(I);
}

static void eval_config_block(const block_t* B, int nat, const double* pos, const double* cell,
                              const uint8_t* pbc, const int32_t* Z, int what, double* E, double* F, double* V) {
    eval_config_block_par(B, nat, pos, cell, pbc, Z, what, E, F, V, 0);
}

/* ------------------------------------------------------------------ */
/* Assembly of one basis block into the column-major LSQ matrix        */
/* ------------------------------------------------------------------ */
/* rowE/rowF/rowV: 1-based first row of each observation block, 0 = absent
 * (rows come from the host's _alloc_lsq_matrix, src/lsq_db.jl:237-250).
 * Psi: column-major, leading dimension ld (= nrows).
 * mode 0: fused E+F+V per config (one task per config).
 * mode 1: reference-like: one task per (config, okey), E/F/V evaluated
 *         separately (neighbour list + basis recomputed each time), tasks
 *         taken from a shared queue in descending natoms order
 *         (src/obsiter.jl:87-102, src/tools.jl:35-56).
 * mode 2: the tasks of mode 1 one after the other, the SITES of each task shared out over the threads
 *         (for timing a small sample of configurations with every host thread busy). */
/* Z: atomic numbers of all atoms (NULL with zc = 0); zc, zs[N-1]: species tuple of the block (zc = 0: none) */
int32_t ipfo_assemble_block_z(int32_t N, int32_t nb, int32_t nmono, const int32_t* mono_b,
                              const uint8_t* mono_exp, int32_t transform, double a, double r0,
                              double rc1, double rc2, int64_t ncfg, const int64_t* atom_off,
                              const double* pos, const double* cell, const uint8_t* pbc,
                              const int64_t* rowE, const int64_t* rowF, const int64_t* rowV,
                              int64_t col0, double* Psi, int64_t ld, int32_t mode, int32_t nthreads,
                              const int32_t* Zall, int32_t zc, const int32_t* zs,
                              int32_t env_t, double env_rc1, double env_rc2) {
    if (N < 2 || N > 5) return -1;
    block_t B;
    B.N = N; B.nb = nb; B.nmono = nmono; B.mono_b = mono_b; B.mono_exp = mono_exp;
    B.D.transform = transform; B.D.a = a; B.D.r0 = r0; B.D.rc1 = rc1; B.D.rc2 = rc2;
    B.zc = (Zall != NULL && zs != NULL) ? zc : 0;
    for (int k = 0; k < IPFO_MAX_LEGS; ++k) B.zs[k] = (B.zc && k < N - 1) ? zs[k] : 0;
    B.env_t = env_t; B.env_rc1 = env_rc1; B.env_rc2 = env_rc2;
    if (env_t > 0 && !(env_rc1 > 0.0 && env_rc2 > env_rc1 && env_rc2 <= rc2)) return -1;

    /* task list */
    const int tasks_per_cfg = mode ? 3 : 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const int64_t ntask = ncfg * tasks_per_cfg;
    int64_t* order = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ntask + 1));
    for (int64_t t = 0; t < ntask; ++t) order[t] = t;
    /* stable counting sort of the tasks by descending natoms (tools.jl:35) */
    {
        int64_t maxn = 0;
        for (int64_t c = 0; c < ncfg; ++c) { int64_t n = atom_off[c + 1] - atom_off[c]; if (n > maxn) maxn = n; }
        int64_t* cnt = (int64_t*)calloc((size_t)(maxn + 2), sizeof(int64_t));
        for (int64_t t = 0; t < ntask; ++t) { int64_t c = t / tasks_per_cfg; cnt[maxn - (atom_off[c + 1] - atom_off[c])]++; }
        int64_t acc = 0;
        for (int64_t k = 0; k <= maxn; ++k) { int64_t v = cnt[k]; cnt[k] = acc; acc += v; }
        for (int64_t t = 0; t < ntask; ++t) { int64_t c = t / tasks_per_cfg; order[cnt[maxn - (atom_off[c + 1] - atom_off[c])]++] = t; }
        free(cnt);
    }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    if (mode == 2) {
        for (int64_t slot = 0; slot < ntask; ++slot) {
            const int64_t t = order[slot];
            const int64_t c = t / tasks_per_cfg;
            int what = 1 << (int)(t % 3);
            if (!rowE[c]) what &= ~1;
            if (!rowF[c]) what &= ~2;
            if (!rowV[c]) what &= ~4;
            if (!what) continue;
            const int nat = (int)(atom_off[c + 1] - atom_off[c]);
            double* E = (double*)calloc((size_t)nb, sizeof(double));
            double* F = (double*)calloc((size_t)nb * 3 * (size_t)nat + 1, sizeof(double));
            double* V = (double*)calloc((size_t)nb * 9, sizeof(double));
            eval_config_block_par(&B, nat, pos + 3 * atom_off[c], cell + 9 * c, pbc + 3 * c,
                                  B.zc ? Zall + atom_off[c] : NULL, what, E, F, V, 1);
            for (int b = 0; b < nb; ++b) {
                double* col = Psi + (size_t)(col0 + b) * (size_t)ld;
                if (what & 1) col[rowE[c] - 1] = E[b];
                if (what & 2)
                    for (int k = 0; k < 3 * nat; ++k) col[rowF[c] - 1 + k] = F[(size_t)b * 3 * nat + k];
                if (what & 4) {
                    const double* Vb = V + (size_t)b * 9;
                    double* o = col + (rowV[c] - 1);
                    o[0] = Vb[0]; o[1] = Vb[4]; o[2] = Vb[8];
                    o[3] = Vb[3 * 2 + 1]; o[4] = Vb[3 * 2 + 0]; o[5] = Vb[3 * 1 + 0];
                }
            }
            free(V); free(F); free(E);
        }
        free(order);
        return 0;
    }
    int64_t next = 0;
#pragma omp parallel
    {
        for (;;) {
            int64_t slot;
#pragma omp atomic capture
            slot = next++;
            if (slot >= ntask) break;
            const int64_t t = order[slot];
            const int64_t c = t / tasks_per_cfg;
            int what = mode ? (1 << (int)(t % 3)) : 7;
            if (!rowE[c]) what &= ~1;
            if (!rowF[c]) what &= ~2;
            if (!rowV[c]) what &= ~4;
            if (!what) continue;
            const int nat = (int)(atom_off[c + 1] - atom_off[c]);
            double* E = (double*)calloc((size_t)nb, sizeof(double));
            double* F = (double*)calloc((size_t)nb * 3 * (size_t)nat + 1, sizeof(double));
            double* V = (double*)calloc((size_t)nb * 9, sizeof(double));
            eval_config_block(&B, nat, pos + 3 * atom_off[c], cell + 9 * c, pbc + 3 * c,
                              B.zc ? Zall + atom_off[c] : NULL, what, E, F, V);
            /* block placement: Psi[irows, col0 + b]; the reference takes a global lock
             * here (src/lsq_db.jl:274-276); rows of different tasks are disjoint. */
#pragma omp critical(ipfo_psi_write)
            {
                for (int b = 0; b < nb; ++b) {
                    double* col = Psi + (size_t)(col0 + b) * (size_t)ld;
                    if (what & 1) col[rowE[c] - 1] = E[b];
                    if (what & 2)
                        for (int k = 0; k < 3 * nat; ++k) col[rowF[c] - 1 + k] = F[(size_t)b * 3 * nat + k];
                    if (what & 4) {
                        /* Julia V[[1,5,9,6,3,2]] on a column-major 3x3: (1,1),(2,2),(3,3),(3,2),(3,1),(2,1) */
                        const double* Vb = V + (size_t)b * 9;
                        double* o = col + (rowV[c] - 1);
                        o[0] = Vb[0]; o[1] = Vb[4]; o[2] = Vb[8];
                        o[3] = Vb[3 * 2 + 1]; o[4] = Vb[3 * 2 + 0]; o[5] = Vb[3 * 1 + 0];
                    }
                }
            }
            free(V); free(F); free(E);
        }
    }
    free(order);
    return 0;
}

int32_t ipfo_assemble_block(int32_t N, int32_t nb, int32_t nmono, const int32_t* mono_b,
                            const uint8_t* mono_exp, int32_t transform, double a, double r0,
                            double rc1, double rc2, int64_t ncfg, const int64_t* atom_off,
                            const double* pos, const double* cell, const uint8_t* pbc,
                            const int64_t* rowE, const int64_t* rowF, const int64_t* rowV,
                            int64_t col0, double* Psi, int64_t ld, int32_t mode, int32_t nthreads) {
    return ipfo_assemble_block_z(N, nb, nmono, mono_b, mono_exp, transform, a, r0, rc1, rc2, ncfg, atom_off, pos, cell,
                                 pbc, rowE, rowF, rowV, col0, Psi, ld, mode, nthreads, NULL, 0, NULL, 0, 0.0, 0.0);
}

/* Single-config convenience for tests: evaluates one block, returns dense
 * E[nb], F[nb][3 nat], V[nb][9]. */
int32_t ipfo_eval_config_z(int32_t N, int32_t nb, int32_t nmono, const int32_t* mono_b,
                           const uint8_t* mono_exp, int32_t transform, double a, double r0,
                           double rc1, double rc2, int32_t nat, const double* pos,
                           const double* cell, const uint8_t* pbc, double* E, double* F, double* V,
                           const int32_t* Z, int32_t zc, const int32_t* zs,
                           int32_t env_t, double env_rc1, double env_rc2) {
    block_t B;
    B.N = N; B.nb = nb; B.nmono = nmono; B.mono_b = mono_b; B.mono_exp = mono_exp;
    B.D.transform = transform; B.D.a = a; B.D.r0 = r0; B.D.rc1 = rc1; B.D.rc2 = rc2;
    B.zc = (Z != NULL && zs != NULL) ? zc : 0;
    for (int k = 0; k < IPFO_MAX_LEGS; ++k) B.zs[k] = (B.zc && k < N - 1) ? zs[k] : 0;
    B.env_t = env_t; B.env_rc1 = env_rc1; B.env_rc2 = env_rc2;
    if (env_t > 0 && !(env_rc1 > 0.0 && env_rc2 > env_rc1 && env_rc2 <= rc2)) return -1;
    memset(E, 0, sizeof(double) * (size_t)nb);
    memset(F, 0, sizeof(double) * (size_t)nb * 3 * (size_t)nat);
    memset(V, 0, sizeof(double) * (size_t)nb * 9);
    eval_config_block(&B, nat, pos, cell, pbc, B.zc ? Z : NULL, 7, E, F, V);
    return 0;
}

int32_t ipfo_eval_config(int32_t N, int32_t nb, int32_t nmono, const int32_t* mono_b,
                         const uint8_t* mono_exp, int32_t transform, double a, double r0,
                         double rc1, double rc2, int32_t nat, const double* pos,
                         const double* cell, const uint8_t* pbc, double* E, double* F, double* V) {
    return ipfo_eval_config_z(N, nb, nmono, mono_b, mono_exp, transform, a, r0, rc1, rc2, nat, pos, cell, pbc, E, F, V,
                              NULL, 0, NULL, 0, 0.0, 0.0);
}

int32_t ipfo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
