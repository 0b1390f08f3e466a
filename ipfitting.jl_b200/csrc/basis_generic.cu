// basis_generic.cu -- K2/K3 (table-driven): basis evaluation with analytic
// derivatives for any body-order 2..5 / any monomial table, and the deterministic
// gather into the E / F / V rows of the column-major LSQ matrix.
//
// Replaces energy/forces/virial(B::IPBasis, at) (reference call sites
// src/datatypes.jl:28,40,65) + vec_obs (src/datatypes.jl:26,42-50,69-76) +
// the block placement db.Psi[irows,:] = ... (src/lsq_db.jl:268-278).
//
// Decomposition ("own-leg"): one thread per ORDERED neighbour pair p = (i, j).
// With position 1 of every cluster pinned to the thread's own neighbour j, the
// thread loops over all unordered sets K of the other N-2 neighbours of i and
// accumulates, per basis function b,
//     S0 = sum_K FC' P_b            S1 = sum_K FC' dP_b/dx_1
//     T  = sum_K FC' sum_c dP_b/dw_1c (Rhat_c - w_1c Rhat_j)          FC' = prod_{k in K} fc_k
// so that   g_p = dV_b(i)/dR_ij = (fc_j' S0 + fc_j x_j' S1) Rhat_j + fc_j/r_j T
//           e_p = fc_j S0 / (N-1)       (each cluster is visited once per leg)
// No atomics: (g_p, e_p) go to an L2-resident scratch tile owned by the CTA and
// are gathered in a fixed order:
//     E_b      = sum_p e_p
//     F_b[a]   = sum_{p in out(a)} (g_p - g_rev(p))          rev(i,j,S) = (j,i,-S)
//     Vir_b    = - sum_p g_p (x) R_p
#include "ipf_common.cuh"

namespace {

constexpr int REC = 8;   // head of a pair record: Rh[3], 1/r, x, dx/dr, fc, dfc/dr; then x^0..x^maxdeg (stride rs)

// thread per pair; the 32 records of a warp are staged in shared memory and written as one contiguous run
// (records are AoS: a per-thread store would touch 32 sectors per instruction)
constexpr int PR_THREADS = 128;
constexpr int PR_MAXRS = 32;

__global__ void __launch_bounds__(PR_THREADS)
pair_record_kernel(int64_t npairs, const double* __restrict__ pR, int transform, double a,
                   double r0, double rc1, double rc2, int maxdeg, int rs, double* __restrict__ rec) {
    __shared__ double tile[PR_THREADS / 32][32 * PR_MAXRS + 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t p0 = ((int64_t)blockIdx.x * PR_THREADS) + warp * 32;      // first pair of the warp
    if (p0 >= npairs) return;
    const int cnt = (int)min((int64_t)32, npairs - p0);
    const int rsp = rs | 1;                 // odd stride in shared memory: conflict-free per-lane records
    if (lane < cnt) {
        const int64_t p = p0 + lane;
        const double Rx = pR[3 * p], Ry = pR[3 * p + 1], Rz = pR[3 * p + 2];
        const double r = ipf_bond_length(Rx, Ry, Rz);
        double x, dx, fc, dfc;
        ipf_descriptor(r, transform, a, r0, rc1, rc2, x, dx, fc, dfc);
        double* o = tile[warp] + lane * rsp;
        o[0] = Rx / r; o[1] = Ry / r; o[2] = Rz / r; o[3] = 1.0 / r;
        o[4] = x; o[5] = dx; o[6] = fc; o[7] = dfc;
        double v = 1.0;
        for (int e = 0; e <= maxdeg; ++e) { o[REC + e] = v; v *= x; }
        for (int e = REC + maxdeg + 1; e < rs; ++e) o[e] = 0.0;
    }
    __syncwarp();
    double* out = rec + (int64_t)rs * p0;
    const double* t = tile[warp];
    for (int i = lane; i < cnt * rs; i += 32) {
        const int q = i / rs, e = i - q * rs;
        out[i] = t[q * rsp + e];
    }
}

__device__ __forceinline__ double ipow(double w, int e) {
    double v = 1.0;
    for (int k = 0; k < e; ++k) v *= w;
    return v;
}

struct GenArgs {
    int nb, unit, nunits, maxpairs;
    int64_t nitems, ncfg;
    const int32_t* mono_start;
    const uint8_t* mono_exp;
    const int64_t* atom_off;
    const int64_t* cfg_pair_off;   // [ncfg+1]
    const int64_t* site_off;
    const int32_t* pi;
    const int32_t* prev;
    const double* pR;
    const double* rec;
    int rs;                        // record stride in doubles
    int64_t npairs;
    const int64_t* rowE;
    const int64_t* rowF;
    const int64_t* rowV;
    double* Psi;
    int64_t ld, col0;
    double* scratch;       // [gridDim][unit][maxpairs][4]
    int* counter;
    // species-resolved block (zc > 0): atomic numbers of the chunk's atoms, neighbour atom of every pair, the sorted
    // neighbour species and the exponent tables with the own-leg slot moved to position 0 (capi.cu)
    const int32_t* Z;
    const int32_t* pj;
    int zc, zs[4];
    const uint8_t* mono_exp_var[4];
    // environment-dependent block (env_t > 0): site function times n_i^env_t, n_i = sum_j CosCut(env_rc1, env_rc2)(r_ij)
    int env_t;
    double env_rc1, env_rc2;
};

constexpr int GEN_THREADS = 256;
constexpr int MAX_UNIT_MONO = 1024;

// One cluster: own leg p + NX-1 others (chunk-global pair indices in oth[]).
// Accumulates S0, S1, T for basis function with monomials [m0, m1).
template <int NX>
__device__ __forceinline__ void cluster_accumulate(const GenArgs& A, int64_t p, const int64_t* oth,
                                                   const double* __restrict__ ownrec, const uint8_t* __restrict__ sexp,
                                                   int m0, int m1, double& S0, double& S1, double* T) {
    constexpr int NO = NX - 1;
    constexpr int NPAIR = NX * (NX - 1) / 2;
    double orec_rh[NO > 0 ? NO : 1][3];
    double FCp = 1.0;
#pragma unroll
    for (int c = 0; c < NO; ++c) {
        const double* r = A.rec + (int64_t)A.rs * oth[c];
        orec_rh[c][0] = r[0]; orec_rh[c][1] = r[1]; orec_rh[c][2] = r[2];
        FCp *= r[6];
    }
    // w variables in lexicographic pair order (0,1),(0,2),..,(1,2),..  position 0 = own leg
    double w[NPAIR > 0 ? NPAIR : 1];
    {
        int k = 0;
#pragma unroll
        for (int a = 0; a < NX; ++a)
#pragma unroll
            for (int c = a + 1; c < NX; ++c) {
                const double* ra = (a == 0) ? ownrec : orec_rh[a - 1];
                const double* rc = orec_rh[c - 1];
                w[k++] = ra[0] * rc[0] + ra[1] * rc[1] + ra[2] * rc[2];
            }
    }
    double P = 0.0, dPx = 0.0;
    double dPw[NO > 0 ? NO : 1];
#pragma unroll
    for (int c = 0; c < NO; ++c) dPw[c] = 0.0;
    for (int m = m0; m < m1; ++m) {
        const uint8_t* ex = sexp + m * IPF_MAX_VARS;
        const int e0 = ex[0];
        const double* ownp = A.rec + (int64_t)A.rs * p + REC;
        const double Aown = ownp[e0];
        const double dAown = e0 > 0 ? (double)e0 * ownp[e0 - 1] : 0.0;
        double rest = 1.0;
#pragma unroll
        for (int c = 0; c < NO; ++c) rest *= A.rec[(int64_t)A.rs * oth[c] + REC + ex[1 + c]];
        // w pairs not touching the own leg
#pragma unroll
        for (int k = NO; k < NPAIR; ++k) rest *= ipow(w[k], ex[NX + k]);
        double wv[NO > 0 ? NO : 1], dwv[NO > 0 ? NO : 1];
        double W0 = 1.0;
#pragma unroll
        for (int c = 0; c < NO; ++c) {
            const int e = ex[NX + c];
            const double wm1 = e > 0 ? ipow(w[c], e - 1) : 0.0;
            wv[c] = e > 0 ? wm1 * w[c] : 1.0;
            dwv[c] = (double)e * wm1;
            W0 *= wv[c];
        }
        const double base = rest * W0;
        P += Aown * base;
        dPx += dAown * base;
#pragma unroll
        for (int c = 0; c < NO; ++c) {
            double o = Aown * rest * dwv[c];
#pragma unroll
            for (int c2 = 0; c2 < NO; ++c2) if (c2 != c) o *= wv[c2];
            dPw[c] += o;
        }
    }
    S0 += FCp * P;
    S1 += FCp * dPx;
#pragma unroll
    for (int c = 0; c < NO; ++c) {
        const double g = FCp * dPw[c];
#pragma unroll
        for (int d = 0; d < 3; ++d) T[d] += g * (orec_rh[c][d] - w[c] * ownrec[d]);
    }
}

template <int NX>
__global__ void __launch_bounds__(GEN_THREADS) site_deriv_generic_kernel(GenArgs A) {
    __shared__ int s_item;
    __shared__ int s_mstart[64 + 1];
    __shared__ uint8_t s_exp[MAX_UNIT_MONO * IPF_MAX_VARS];
    const int tid = threadIdx.x;
    double* G = A.scratch + (size_t)blockIdx.x * A.unit * A.maxpairs * 4;
    for (;;) {
        if (tid == 0) s_item = atomicAdd(A.counter, 1);
        __syncthreads();
        const int64_t item = s_item;
        if (item >= A.nitems) break;
        const int64_t c = item / A.nunits;
        const int u = (int)(item % A.nunits);
        const int b0 = u * A.unit;
        const int nbu = min(A.nb - b0, A.unit);
        const int mbase = A.mono_start[b0];
        for (int k = tid; k <= nbu; k += blockDim.x) s_mstart[k] = A.mono_start[b0 + k] - mbase;
        const int nm = A.mono_start[b0 + nbu] - mbase;
        for (int k = tid; k < nm * IPF_MAX_VARS; k += blockDim.x) s_exp[k] = A.mono_exp[(size_t)mbase * IPF_MAX_VARS + k];
        __syncthreads();
        const int64_t P0 = A.cfg_pair_off[c], P1 = A.cfg_pair_off[c + 1];
        const int npc = (int)(P1 - P0);
        const int64_t a0 = A.atom_off[c];
        const int nat = (int)(A.atom_off[c + 1] - a0);
        // ---- phase A: per ordered pair, per basis function ----
        for (int pl = tid; pl < npc; pl += blockDim.x) {
            const int64_t p = P0 + pl;
            double own[REC];
#pragma unroll
            for (int k = 0; k < REC; ++k) own[k] = A.rec[(int64_t)A.rs * p + k];
            const int64_t s = a0 + A.pi[p];
            const int64_t o0 = A.site_off[s], o1 = A.site_off[s + 1];
            // species rule: the own leg takes the first slot of its species, the other legs (sorted by species) must
            // match the remaining slots; the exponent table is the one with that slot moved to position 0
            bool active = true;
            const uint8_t* etab = s_exp;
            int rem[3] = {0, 0, 0};
            if (A.zc > 0) {
                const int zj = A.Z[a0 + A.pj[p]];
                int s0 = -1;
#pragma unroll
                for (int t = 0; t < NX; ++t) if (s0 < 0 && A.zs[t] == zj) s0 = t;
                active = A.Z[s] == A.zc && s0 >= 0;
                if (active) {
                    etab = A.mono_exp_var[s0] + (size_t)mbase * IPF_MAX_VARS;
                    int k = 0;
#pragma unroll
                    for (int t = 0; t < NX; ++t) if (t != s0) { if (k < 3) rem[k] = A.zs[t]; ++k; }
                }
            }
            for (int bl = 0; bl < nbu; ++bl) {
                const int m0 = s_mstart[bl], m1 = s_mstart[bl + 1];
                double S0 = 0.0, S1 = 0.0, T[3] = {0.0, 0.0, 0.0};
                int64_t oth[3] = {0, 0, 0};
                if (!active) {
                } else if (NX == 1) {
                    cluster_accumulate<NX>(A, p, oth, own, etab, m0, m1, S0, S1, T);
                } else if (NX == 2) {
                    for (int64_t k = o0; k < o1; ++k) {
                        if (k == p) continue;
                        if (A.zc > 0 && A.Z[a0 + A.pj[k]] != rem[0]) continue;
                        oth[0] = k;
                        cluster_accumulate<NX>(A, p, oth, own, etab, m0, m1, S0, S1, T);
                    }
                } else if (NX == 3) {
                    for (int64_t k = o0; k < o1; ++k) {
                        if (k == p) continue;
                        for (int64_t l = k + 1; l < o1; ++l) {
                            if (l == p) continue;
                            oth[0] = k; oth[1] = l;
                            if (A.zc > 0) {
                                const int zk = A.Z[a0 + A.pj[k]], zl = A.Z[a0 + A.pj[l]];
                                if (zk > zl) { oth[0] = l; oth[1] = k; }          // stable: equal species keep the index order
                                if (min(zk, zl) != rem[0] || max(zk, zl) != rem[1]) continue;
                            }
                            cluster_accumulate<NX>(A, p, oth, own, etab, m0, m1, S0, S1, T);
                        }
                    }
                } else {
                    for (int64_t k = o0; k < o1; ++k) {
                        if (k == p) continue;
                        for (int64_t l = k + 1; l < o1; ++l) {
                            if (l == p) continue;
                            for (int64_t q = l + 1; q < o1; ++q) {
                                if (q == p) continue;
                                oth[0] = k; oth[1] = l; oth[2] = q;
                                if (A.zc > 0) {
                                    int z0 = A.Z[a0 + A.pj[k]], z1 = A.Z[a0 + A.pj[l]], z2 = A.Z[a0 + A.pj[q]];
                                    // stable three-element insertion sort by species
                                    if (z0 > z1) { int t = z0; z0 = z1; z1 = t; int64_t u = oth[0]; oth[0] = oth[1]; oth[1] = u; }
                                    if (z1 > z2) {
                                        int t = z1; z1 = z2; z2 = t; int64_t u = oth[1]; oth[1] = oth[2]; oth[2] = u;
                                        if (z0 > z1) { t = z0; z0 = z1; z1 = t; u = oth[0]; oth[0] = oth[1]; oth[1] = u; }
                                    }
                                    if (z0 != rem[0] || z1 != rem[1] || z2 != rem[2]) continue;
                                }
                                cluster_accumulate<NX>(A, p, oth, own, etab, m0, m1, S0, S1, T);
                            }
                        }
                    }
                }
                const double radial = own[7] * S0 + own[6] * own[5] * S1;
                const double tang = own[6] * own[3];
                double* g = G + ((size_t)bl * A.maxpairs + pl) * 4;
                g[0] = radial * own[0] + tang * T[0];
                g[1] = radial * own[1] + tang * T[1];
                g[2] = radial * own[2] + tang * T[2];
                g[3] = own[6] * S0 / (double)NX;
            }
        }
        __syncthreads();
        if (A.env_t > 0) {
            // ---- phase A2: environment factor.  Per (function, centre atom): n_i and V0 = sum of the energy shares of
            // its pairs; then g_p <- n^t g_p + t n^(t-1) fenv'(r_p) V0 Rhat_p, e_p <- n^t e_p (fixed order)
            constexpr double kPiD = 3.14159265358979323846;
            const double we = A.env_rc2 - A.env_rc1;
            for (int idx = tid; idx < nbu * nat; idx += blockDim.x) {
                const int bl = idx / nat, at = idx - bl * nat;
                const int64_t s = a0 + at;
                const int64_t q0 = A.site_off[s], q1 = A.site_off[s + 1];
                double* Gb = G + (size_t)bl * A.maxpairs * 4;
                double ni = 0.0, V0 = 0.0;
                for (int64_t p = q0; p < q1; ++p) {
                    const double r = 1.0 / A.rec[(int64_t)A.rs * p + 3];
                    if (r <= A.env_rc1) ni += 1.0;
                    else if (r < A.env_rc2) ni += 0.5 * (1.0 + cos(kPiD * (r - A.env_rc1) / we));
                    V0 += Gb[(p - P0) * 4 + 3];
                }
                double nt = 1.0, ntm1 = 1.0;
                for (int k = 0; k < A.env_t; ++k) { ntm1 = nt; nt *= ni; }
                ntm1 *= (double)A.env_t;
                for (int64_t p = q0; p < q1; ++p) {
                    const double* rc = A.rec + (int64_t)A.rs * p;
                    const double r = 1.0 / rc[3];
                    double c = 0.0;
                    if (r > A.env_rc1 && r < A.env_rc2) c = ntm1 * (-0.5 * kPiD / we * sin(kPiD * (r - A.env_rc1) / we)) * V0;
                    double* g = Gb + (p - P0) * 4;
                    g[0] = nt * g[0] + c * rc[0];
                    g[1] = nt * g[1] + c * rc[1];
                    g[2] = nt * g[2] + c * rc[2];
                    g[3] = nt * g[3];
                }
            }
            __syncthreads();
        }
        // ---- phase B: deterministic gather into Psi ----
        const int64_t rE = A.rowE[c], rF = A.rowF[c], rV = A.rowV[c];
        if (rF) {
            const int nrow = 3 * nat;
            for (int idx = tid; idx < nbu * nrow; idx += blockDim.x) {
                const int bl = idx / nrow, rd = idx - bl * nrow;
                const int at = rd / 3, d = rd - 3 * at;
                const int64_t s = a0 + at;
                const double* Gb = G + (size_t)bl * A.maxpairs * 4;
                double acc = 0.0;
                for (int64_t p = A.site_off[s]; p < A.site_off[s + 1]; ++p)
                    acc += Gb[(p - P0) * 4 + d] - Gb[((int64_t)A.prev[p] - P0) * 4 + d];
                A.Psi[(size_t)(A.col0 + b0 + bl) * A.ld + (rF - 1) + rd] = acc;
            }
        }
        {
            // warp per (bl, comp): comp 0 = E, 1..6 = Voigt (xx,yy,zz,zy,zx,yx)
            const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
            for (int t = warp; t < nbu * 7; t += nwarp) {
                const int bl = t / 7, comp = t - 7 * bl;
                if (comp == 0 ? !rE : !rV) continue;
                const double* Gb = G + (size_t)bl * A.maxpairs * 4;
                double acc = 0.0;
                if (comp == 0) {
                    for (int pl = lane; pl < npc; pl += 32) acc += Gb[(size_t)pl * 4 + 3];
                } else {
                    const int pp = (comp == 1) ? 0 : (comp == 2) ? 1 : (comp == 3) ? 2 : (comp == 4) ? 2 : (comp == 5) ? 2 : 1;
                    const int qq = (comp == 1) ? 0 : (comp == 2) ? 1 : (comp == 3) ? 2 : (comp == 4) ? 1 : (comp == 5) ? 0 : 0;
                    for (int pl = lane; pl < npc; pl += 32) acc -= Gb[(size_t)pl * 4 + pp] * A.pR[3 * (P0 + pl) + qq];
                }
                acc = warp_sum(acc);
                if (lane == 0) {
                    const int64_t row = (comp == 0) ? (rE - 1) : (rV - 1 + comp - 1);
                    A.Psi[(size_t)(A.col0 + b0 + bl) * A.ld + row] = acc;
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace

// pair records of one descriptor in ctx scratch slot 1: per pair `rs` doubles =
// [Rh(3), 1/r, x, dx, fc, dfc, x^0 .. x^maxdeg, pad]; rs is even (16-byte aligned records)
int ipf_pair_records(ipf_ctx* ctx, const BlockDev& blk, const NeighList& nl, const double** rec_out, int* rs_out) {
    const int64_t npairs = nl.npairs;
    // reuse the records of the previous block when it had the same descriptor on the same neighbour list
    // (e.g. nbpolys(2, D, 14) followed by nbpolys(3, D, 11)) and at least this block's degree
    if (ctx->rec_nl == nl.serial && nl.serial >= 0 && ctx->rec_npairs == npairs && ctx->rec_transform == blk.transform &&
        ctx->rec_a == blk.a && ctx->rec_r0 == blk.r0 && ctx->rec_rc1 == blk.rc1 && ctx->rec_rc2 == blk.rc2 &&
        ctx->rec_maxdeg == blk.maxdeg && ctx->rec_ptr) {
        *rec_out = ctx->rec_ptr;
        *rs_out = ctx->rec_rs;
        return IPF_OK;
    }
    const int rs = (REC + blk.maxdeg + 1 + 1) & ~1;
    void* rec = nullptr;
    IPF_CHECK(ipf_scratch(ctx, 1, sizeof(double) * rs * (size_t)(npairs + 1), &rec));
    if (npairs > 0) {
        ProfScope prof_(ctx, "pair_record");
        pair_record_kernel<<<(unsigned)((npairs + PR_THREADS - 1) / PR_THREADS), PR_THREADS, 0, ctx->stream>>>(
            npairs, nl.pR.p, blk.transform, blk.a, blk.r0, blk.rc1, blk.rc2, blk.maxdeg, rs, (double*)rec);
        ctx->launches++;
    }
    *rec_out = (const double*)rec;
    *rs_out = rs;
    ctx->rec_nl = nl.serial; ctx->rec_npairs = npairs; ctx->rec_transform = blk.transform;
    ctx->rec_a = blk.a; ctx->rec_r0 = blk.r0; ctx->rec_rc1 = blk.rc1; ctx->rec_rc2 = blk.rc2;
    ctx->rec_maxdeg = blk.maxdeg; ctx->rec_ptr = (const double*)rec; ctx->rec_rs = rs;
    return IPF_OK;
}

int ipf_assemble_block_generic(ipf_ctx* ctx, const BlockDev& blk, const ChunkDev& ch, const NeighList& nl,
                               ipf_matrix* Psi) {
    cudaStream_t st = ctx->stream;
    if (ch.ncfg == 0 || blk.nb == 0) return IPF_OK;
    const int64_t npairs = nl.npairs;
    // pair records + power table for this descriptor
    const double* rec = nullptr;
    int rs = 0;
    void *scr = nullptr, *cnt = nullptr;
    IPF_CHECK(ipf_pair_records(ctx, blk, nl, &rec, &rs));
    // unit size: bounded by the shared-memory monomial table and by 64 functions
    int unit = 64;
    {
        // largest unit such that any run of `unit` consecutive functions has <= MAX_UNIT_MONO monomials
        for (;;) {
            bool ok = true;
            for (int b0 = 0; b0 < blk.nb && ok; b0 += unit) {
                const int b1 = std::min(blk.nb, b0 + unit);
                if (blk.h_mono_start[b1] - blk.h_mono_start[b0] > MAX_UNIT_MONO) ok = false;
            }
            if (ok || unit == 1) break;
            unit /= 2;
        }
    }
    // keep units small enough that there are a few work items per SM
    while (unit > 4 && (int64_t)((blk.nb + unit - 1) / unit) * ch.ncfg < 4LL * ctx->sm_count) unit /= 2;
    const int nunits = (blk.nb + unit - 1) / unit;
    const int grid = (int)std::min<int64_t>((int64_t)ch.ncfg * nunits, 2LL * ctx->sm_count);
    const int maxpairs = std::max(1, nl.maxpairs_cfg);
    IPF_CHECK(ipf_scratch(ctx, 3, sizeof(double) * 4 * (size_t)grid * unit * maxpairs, &scr));
    IPF_CHECK(ipf_scratch(ctx, 5, 256, &cnt));
    IPF_CUDA(ctx, cudaMemsetAsync(cnt, 0, sizeof(int), st));
    GenArgs A;
    A.nb = blk.nb; A.unit = unit; A.nunits = nunits; A.maxpairs = maxpairs;
    A.nitems = (int64_t)ch.ncfg * nunits; A.ncfg = ch.ncfg;
    A.mono_start = blk.mono_start; A.mono_exp = blk.mono_exp;
    A.atom_off = ch.atom_off.p; A.cfg_pair_off = nl.cfg_pair_off.p; A.site_off = nl.site_off.p;
    A.pi = nl.pi.p; A.prev = nl.prev.p; A.pR = nl.pR.p; A.rec = rec; A.rs = rs;
    A.npairs = npairs; A.rowE = ch.rowE.p; A.rowF = ch.rowF.p; A.rowV = ch.rowV.p;
    A.Psi = Psi->d; A.ld = Psi->ld; A.col0 = blk.col0; A.scratch = (double*)scr; A.counter = (int*)cnt;
    A.Z = ch.Z.p; A.pj = nl.pj.p; A.zc = blk.zc;
    for (int t = 0; t < 4; ++t) { A.zs[t] = blk.zs[t]; A.mono_exp_var[t] = blk.mono_exp_var[t]; }
    A.env_t = blk.env_t; A.env_rc1 = blk.env_rc1; A.env_rc2 = blk.env_rc2;
    if (ctx->profiling) {
        // algorithmic work counter: unordered clusters sum_i C(n_i, N-1) (roofline numerator)
        static const char* cnames[6] = {"", "", "count:clusters_2b", "count:clusters_3b", "count:clusters_4b", "count:clusters_5b"};
        double tot = 0.0;
        const int nx = blk.N - 1;
        for (int64_t s = 0; s < nl.nsites; ++s) {
            const double n = (double)(nl.h_site_off[s + 1] - nl.h_site_off[s]);
            double c = 1.0;
            for (int k = 0; k < nx; ++k) c *= (n - k) / (double)(k + 1);
            if (n >= nx) tot += c;
        }
        ctx->prof[cnames[blk.N]].ms += tot;
        ctx->prof["count:pairs"].ms += (double)nl.npairs;
    }
    static const char* names[6] = {"", "", "site_deriv_2b", "site_deriv_3b", "site_deriv_4b", "site_deriv_5b"};
    {
    ProfScope prof_(ctx, names[blk.N >= 2 && blk.N <= 5 ? blk.N : 0]);
    switch (blk.N) {
        case 2: site_deriv_generic_kernel<1><<<grid, GEN_THREADS, 0, st>>>(A); break;
        case 3: site_deriv_generic_kernel<2><<<grid, GEN_THREADS, 0, st>>>(A); break;
        case 4: site_deriv_generic_kernel<3><<<grid, GEN_THREADS, 0, st>>>(A); break;
        case 5: site_deriv_generic_kernel<4><<<grid, GEN_THREADS, 0, st>>>(A); break;
        default: return ipf_set_error(ctx, IPF_EINVAL, "body-order %d not supported", blk.N);
    }
    }
    ctx->launches++;
    IPF_CUDA(ctx, cudaGetLastError());
    return IPF_OK;
}
