// basis_3b_mom.cu -- K3 fast path, variant "moments in HBM": canonical 3-body basis nbpolys(3, desc, D).
//
// Same factorisation as basis_3b.cu (own leg j, moments Z_c[e], Q_c[e] over the second leg k), but the
// producer stops after the k loop and stores the RAW MOMENTS of every ordered pair (NM = 210 doubles for
// D = 11, i.e. 1.7 KB instead of 4.8 KB of per-function derivatives), and the consumer turns them into
// basis functions on the fly:
//   moments3b_raw_kernel<D,U> : CTA = 128 ordered pairs x one unit of w-exponents; neighbour records staged
//                               in shared memory; epilogue = plain stores of the accumulators.
//   combine3b_kernel<D>       : CTA = one configuration, thread = basis function (a <= b, c).  For every
//                               atom and every neighbour entry p it visits the moment row of p (own leg) and
//                               of rev(p) (the same bond seen from the neighbour), evaluates
//                                   S0 = x^a Z_c[b] + x^b Z_c[a],  S1 = a x^(a-1) Z_c[b] + b x^(b-1) Z_c[a],
//                                   T  = c (x^a Q_c[b] + x^b Q_c[a]),
//                                   g  = (fc' S0 + fc x' S1) Rhat + fc/r (T1 e1 + T2 e2)
//                               and accumulates  F_b[a] += g_p - g_rev(p),  E_b += fc S0 / 2,  Vir_b -= g_p (x) R_p
//                               in registers, in a fixed order (deterministic, no atomics, no scratch reduction).
// HBM traffic per ordered pair: 1.7 KB written + 3.4 KB read (own + reverse row) instead of ~10 KB.
#include <cstdlib>

#include "ipf_common.cuh"

int ipf_pair_records(ipf_ctx* ctx, const BlockDev& blk, const NeighList& nl, const double** rec, int* rs);

namespace {

constexpr int REC = 8;
constexpr int A_THREADS = 128;
constexpr int MAXD = 14;
constexpr int A_MAXREC = 224;

template <int D>
struct Lay {
    static constexpr int RS = (REC + D + 2) & ~1;                 // pair record stride
    static constexpr int NZ = ((D + 1) * (D + 2) / 2 + 1) & ~1;   // scalar moments (padded to even)
    static constexpr int NQ = D * (D + 1) / 2;                    // vector moments (2 components each)
    static constexpr int NM = NZ + 2 * NQ;                        // doubles per moment row (even)
    __host__ __device__ static constexpr int zoff(int c) { int n = 0; for (int k = 0; k < c; ++k) n += D - k + 1; return n; }
    __host__ __device__ static constexpr int qoff(int c) { int n = 0; for (int k = 1; k < c; ++k) n += D - k + 1; return n; }
};

__device__ __forceinline__ double4 ld4(const double* p) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

template <int E>
__device__ __forceinline__ double ipow_c(double w) {
    if constexpr (E <= 0) return 1.0;
    else if constexpr (E == 1) return w;
    else {
        const double h = ipow_c<E / 2>(w);
        if constexpr (E % 2 == 0) return h * h;
        else return h * h * w;
    }
}

// orthonormal basis (e1, e2) of the plane perpendicular to the unit vector R (same code in producer and consumer)
__device__ __forceinline__ void plane_basis(double R0, double R1, double R2, double (&E1)[3], double (&E2)[3]) {
    const bool usex = fabs(R0) < 0.8;
    const double c0 = usex ? 0.0 : -R2, c1 = usex ? R2 : 0.0, c2 = usex ? -R1 : R0;
    const double inv = rsqrt(c0 * c0 + c1 * c1 + c2 * c2);
    E1[0] = c0 * inv; E1[1] = c1 * inv; E1[2] = c2 * inv;
    E2[0] = R1 * E1[2] - R2 * E1[1];
    E2[1] = R2 * E1[0] - R0 * E1[2];
    E2[2] = R0 * E1[1] - R1 * E1[0];
}

template <int D, int C0, int C1, int NC, int EMAX>
__device__ __forceinline__ void moment_loop(const double* __restrict__ rk, int n, int jself, double Rj0, double Rj1,
                                            double Rj2, const double (&E1)[3], const double (&E2)[3],
                                            double (&Z)[NC][EMAX + 1], double (&U)[NC][EMAX + 1][2]) {
    constexpr int RS = Lay<D>::RS;
    for (int kk = 0; kk < n; ++kk, rk += RS) {
        const double2 r01 = *reinterpret_cast<const double2*>(rk);
        const double2 r23 = *reinterpret_cast<const double2*>(rk + 2);
        const double2 a1 = *reinterpret_cast<const double2*>(rk + 6);      // fc, dfc
        double xe[EMAX + 2];
#pragma unroll
        for (int e = 0; e <= EMAX; e += 2) {
            const double2 v = *reinterpret_cast<const double2*>(rk + REC + e);
            xe[e] = v.x; xe[e + 1] = v.y;
        }
        const double w = Rj0 * r01.x + Rj1 * r01.y + Rj2 * r23.x;
        const double q1 = E1[0] * r01.x + E1[1] * r01.y + E1[2] * r23.x;
        const double q2 = E2[0] * r01.x + E2[1] * r01.y + E2[2] * r23.x;
        const double fck = (kk == jself) ? 0.0 : a1.x;
        double wl[NC];
        if constexpr (C0 >= 1) {
            wl[0] = fck * ipow_c<(C0 >= 1 ? C0 - 1 : 0)>(w);
#pragma unroll
            for (int ci = 1; ci < NC; ++ci) wl[ci] = wl[ci - 1] * w;
        } else {
            wl[0] = 0.0;
            if constexpr (NC > 1) {
                wl[1] = fck;
#pragma unroll
                for (int ci = 2; ci < NC; ++ci) wl[ci] = wl[ci - 1] * w;
            }
        }
#pragma unroll
        for (int e = 0; e <= EMAX; ++e) {
#pragma unroll
            for (int ci = 0; ci < NC; ++ci) {
                const int c = C0 + ci;
                if (e <= D - c) {
                    if (c == 0) {
                        Z[ci][e] = fma(xe[e], fck, Z[ci][e]);
                    } else {
                        const double t = xe[e] * wl[ci];
                        U[ci][e][0] = fma(t, q1, U[ci][e][0]);
                        U[ci][e][1] = fma(t, q2, U[ci][e][1]);
                        Z[ci][e] = fma(t, w, Z[ci][e]);
                    }
                }
            }
        }
    }
}

// units: consecutive c packed while the accumulators fit
template <int D>
struct UnitTable {
    int n = 0;
    int c0[MAXD + 1] = {}, c1[MAXD + 1] = {};
    constexpr UnitTable() {
        int c = 0;
        while (c <= D) {
            int acc = 0, e = c;
            while (e <= D) {
                const int add = (e == 0 ? 1 : 3) * (D - e + 1);
                if (acc > 0 && acc + add > 45) break;
                acc += add;
                ++e;
            }
            c0[n] = c; c1[n] = e - 1; ++n;
            c = e;
        }
    }
};

struct RawArgs {
    int64_t p0, np;
    const double* rec;
    const int64_t* pbeg;
    const int32_t* pcnt;
    double* M;                 // [np][NM]
};

template <int D, int U>
__global__ void __launch_bounds__(A_THREADS) moments3b_raw_kernel(RawArgs A) {
    extern __shared__ __align__(16) double sm[];
    __shared__ int64_t s_kbeg;
    __shared__ int s_nrec;
    constexpr UnitTable<D> UT{};
    constexpr int C0 = UT.c0[U], C1 = UT.c1[U];
    constexpr int RS = Lay<D>::RS;
    constexpr int NC = C1 - C0 + 1;
    constexpr int EMAX = D - C0;
    const int tid = threadIdx.x;
    const int64_t pl0 = (int64_t)blockIdx.x * A_THREADS;
    const int64_t pl = pl0 + tid;
    const bool live = pl < A.np;
    const int nrow = (int)min((int64_t)A_THREADS, A.np - pl0);
    const int64_t p = A.p0 + (live ? pl : pl0);
    const int64_t k0 = A.pbeg[p];
    const int n = live ? A.pcnt[p] : 0;
    if (tid == 0) s_kbeg = k0;
    if (tid == nrow - 1) s_nrec = (int)min((int64_t)(A_MAXREC + 1), k0 + A.pcnt[p] - A.pbeg[A.p0 + pl0]);
    __syncthreads();
    const int64_t kbeg = s_kbeg;
    const bool staged = s_nrec <= A_MAXREC;
    if (staged) {
        const double2* src = reinterpret_cast<const double2*>(A.rec + kbeg * RS);
        double2* dst = reinterpret_cast<double2*>(sm);
        const int n2 = s_nrec * (RS / 2);
        for (int i = tid; i < n2; i += A_THREADS) dst[i] = __ldg(src + i);
    }
    const double4 o0v = ld4(A.rec + (int64_t)RS * p);
    const double Rj0 = o0v.x, Rj1 = o0v.y, Rj2 = o0v.z;
    const int jself = (int)(p - k0);
    double Z[NC][EMAX + 1];
    double Q[NC][EMAX + 1][2];
#pragma unroll
    for (int ci = 0; ci < NC; ++ci)
#pragma unroll
        for (int e = 0; e <= EMAX; ++e) { Z[ci][e] = 0.0; Q[ci][e][0] = Q[ci][e][1] = 0.0; }
    double E1[3], E2[3];
    plane_basis(Rj0, Rj1, Rj2, E1, E2);
    __syncthreads();
    if (staged) moment_loop<D, C0, C1, NC, EMAX>(sm + (k0 - kbeg) * RS, n, jself, Rj0, Rj1, Rj2, E1, E2, Z, Q);
    else moment_loop<D, C0, C1, NC, EMAX>(A.rec + k0 * RS, n, jself, Rj0, Rj1, Rj2, E1, E2, Z, Q);
    if (!live) return;
    double* row = A.M + pl * Lay<D>::NM;
#pragma unroll
    for (int ci = 0; ci < NC; ++ci) {
        const int c = C0 + ci;
#pragma unroll
        for (int e = 0; e <= D - c; ++e) {
            row[Lay<D>::zoff(c) + e] = Z[ci][e];
            if (c >= 1)
                *reinterpret_cast<double2*>(row + Lay<D>::NZ + 2 * (Lay<D>::qoff(c) + e)) = make_double2(Q[ci][e][0], Q[ci][e][1]);
        }
    }
}

// ---- consumer ------------------------------------------------------------------------------------------
struct CombArgs {
    int64_t p0, c0;
    int nb, rs;
    const double* M;
    const double* rec;
    const int8_t* fdesc;        // [nb][4]: c, a, b, 0
    const int64_t* atom_off;
    const int64_t* site_off;
    const int32_t* prev;
    const double* pR;
    const int64_t* rowE;
    const int64_t* rowF;
    const int64_t* rowV;
    double* Psi;
    int64_t ld, col0;
};

constexpr int C_THREADS = 256;
constexpr int C_NV = 32;           // visits (moment rows) staged per batch
constexpr int C_CST = 32;          // doubles of per-visit constants
constexpr int C_NFT = 2;           // basis functions per thread (nb <= 512)

template <int D>
__global__ void __launch_bounds__(C_THREADS) combine3b_kernel(CombArgs A) {
    constexpr int NM = Lay<D>::NM;
    constexpr int NZ = Lay<D>::NZ;
    extern __shared__ __align__(16) double c_smem[];
    double (*rows)[NM] = reinterpret_cast<double (*)[NM]>(c_smem);
    double (*cst)[C_CST] = reinterpret_cast<double (*)[C_CST]>(c_smem + C_NV * NM);     // 0-2 Rhat, 3 fc'(r), 4 fc*x', 5 fc/r, 6-8 e1, 9-11 e2, 12 own(1)/rev(0),
                                            // 13 fc/2, 14-16 R_p, 17.. x^0..x^D
    const int cfg = (int)(A.c0 + blockIdx.x);
    const int tid = threadIdx.x;
    const int64_t a0 = A.atom_off[cfg];
    const int nat = (int)(A.atom_off[cfg + 1] - a0);
    const int64_t rE = A.rowE[cfg], rF = A.rowF[cfg], rV = A.rowV[cfg];
    // my functions
    int fb[C_NFT], zA[C_NFT], zB[C_NFT], qA[C_NFT], qB[C_NFT], ea[C_NFT], eb[C_NFT];
    double cc[C_NFT];
    bool flive[C_NFT], fdiag[C_NFT];
#pragma unroll
    for (int t = 0; t < C_NFT; ++t) {
        const int b = tid + t * C_THREADS;
        flive[t] = b < A.nb;
        fb[t] = b;
        const int c = flive[t] ? A.fdesc[4 * b] : 0, ia = flive[t] ? A.fdesc[4 * b + 1] : 0, ib = flive[t] ? A.fdesc[4 * b + 2] : 0;
        cc[t] = (double)c; ea[t] = ia; eb[t] = ib; fdiag[t] = ia == ib;
        zA[t] = Lay<D>::zoff(c) + ia; zB[t] = Lay<D>::zoff(c) + ib;
        qA[t] = c >= 1 ? NZ + 2 * (Lay<D>::qoff(c) + ia) : 0;
        qB[t] = c >= 1 ? NZ + 2 * (Lay<D>::qoff(c) + ib) : 0;
    }
    double Eacc[C_NFT], Vacc[C_NFT][6];
#pragma unroll
    for (int t = 0; t < C_NFT; ++t) { Eacc[t] = 0.0; for (int k = 0; k < 6; ++k) Vacc[t][k] = 0.0; }

    for (int at = 0; at < nat; ++at) {
        const int64_t q0 = A.site_off[a0 + at], q1 = A.site_off[a0 + at + 1];
        const int nv = 2 * (int)(q1 - q0);
        double F[C_NFT][3];
#pragma unroll
        for (int t = 0; t < C_NFT; ++t) F[t][0] = F[t][1] = F[t][2] = 0.0;
        for (int vb = 0; vb < nv; vb += C_NV) {
            const int nvis = min(C_NV, nv - vb);
            __syncthreads();
            // stage the moment rows of this batch of visits
            for (int idx = tid; idx < nvis * (NM / 2); idx += C_THREADS) {
                const int v = idx / (NM / 2), k = idx - v * (NM / 2);
                const int vv = vb + v;
                const int64_t p = q0 + (vv >> 1);
                const int64_t pp = (vv & 1) ? (int64_t)A.prev[p] : p;
                reinterpret_cast<double2*>(rows[v])[k] = __ldg(reinterpret_cast<const double2*>(A.M + (pp - A.p0) * NM) + k);
            }
            if (tid < nvis) {
                const int vv = vb + tid;
                const int64_t p = q0 + (vv >> 1);
                const bool own = (vv & 1) == 0;
                const int64_t pp = own ? p : (int64_t)A.prev[p];
                const double* r = A.rec + (int64_t)A.rs * pp;
                double* c = cst[tid];
                const double R0 = r[0], R1 = r[1], R2 = r[2];
                c[0] = R0; c[1] = R1; c[2] = R2;
                c[3] = r[7];                  // fc'
                c[4] = r[6] * r[5];           // fc x'
                c[5] = r[6] * r[3];           // fc / r
                double E1[3], E2[3];
                plane_basis(R0, R1, R2, E1, E2);
                c[6] = E1[0]; c[7] = E1[1]; c[8] = E1[2]; c[9] = E2[0]; c[10] = E2[1]; c[11] = E2[2];
                c[12] = own ? 1.0 : 0.0;
                c[13] = 0.5 * r[6];
                c[14] = A.pR[3 * p]; c[15] = A.pR[3 * p + 1]; c[16] = A.pR[3 * p + 2];
                for (int e = 0; e <= D; ++e) c[17 + e] = r[REC + e];
            }
            __syncthreads();
            for (int v = 0; v < nvis; ++v) {
                const double* row = rows[v];
                const double* c = cst[v];
                const bool own = c[12] != 0.0;
#pragma unroll
                for (int t = 0; t < C_NFT; ++t) {
                    if (!flive[t]) continue;
                    const double Xa = c[17 + ea[t]], Xb = c[17 + eb[t]];
                    const double dXa = ea[t] > 0 ? (double)ea[t] * c[16 + ea[t]] : 0.0;
                    const double dXb = eb[t] > 0 ? (double)eb[t] * c[16 + eb[t]] : 0.0;
                    const double Za = row[zA[t]], Zb = row[zB[t]];
                    double S0, S1, T1 = 0.0, T2 = 0.0;
                    if (fdiag[t]) {
                        S0 = Xa * Za; S1 = dXa * Za;
                        if (cc[t] != 0.0) { T1 = Xa * row[qA[t]]; T2 = Xa * row[qA[t] + 1]; }
                    } else {
                        S0 = fma(Xa, Zb, Xb * Za); S1 = fma(dXa, Zb, dXb * Za);
                        if (cc[t] != 0.0) {
                            T1 = fma(Xa, row[qB[t]], Xb * row[qA[t]]);
                            T2 = fma(Xa, row[qB[t] + 1], Xb * row[qA[t] + 1]);
                        }
                    }
                    const double radial = fma(c[3], S0, c[4] * S1);
                    const double tc = c[5] * cc[t];
                    T1 *= tc; T2 *= tc;
                    const double g0 = fma(radial, c[0], fma(T1, c[6], T2 * c[9]));
                    const double g1 = fma(radial, c[1], fma(T1, c[7], T2 * c[10]));
                    const double g2 = fma(radial, c[2], fma(T1, c[8], T2 * c[11]));
                    if (own) {
                        F[t][0] += g0; F[t][1] += g1; F[t][2] += g2;
                        Eacc[t] = fma(c[13], S0, Eacc[t]);
                        Vacc[t][0] -= g0 * c[14]; Vacc[t][1] -= g1 * c[15]; Vacc[t][2] -= g2 * c[16];
                        Vacc[t][3] -= g2 * c[15]; Vacc[t][4] -= g2 * c[14]; Vacc[t][5] -= g1 * c[14];
                    } else {
                        F[t][0] -= g0; F[t][1] -= g1; F[t][2] -= g2;
                    }
                }
            }
        }
        if (rF) {
#pragma unroll
            for (int t = 0; t < C_NFT; ++t)
                if (flive[t]) {
                    double* o = A.Psi + (size_t)(A.col0 + fb[t]) * A.ld + (rF - 1) + 3 * at;
                    o[0] = F[t][0]; o[1] = F[t][1]; o[2] = F[t][2];
                }
        }
    }
#pragma unroll
    for (int t = 0; t < C_NFT; ++t)
        if (flive[t]) {
            double* colp = A.Psi + (size_t)(A.col0 + fb[t]) * A.ld;
            if (rE) colp[rE - 1] = Eacc[t];
            if (rV) for (int k = 0; k < 6; ++k) colp[rV - 1 + k] = Vacc[t][k];
        }
}

template <int D, int U>
void launch_raw_from(ipf_ctx* ctx, const RawArgs& A) {
    constexpr UnitTable<D> UT{};
    if constexpr (U < UT.n) {
        const unsigned grid = (unsigned)((A.np + A_THREADS - 1) / A_THREADS);
        auto kern = moments3b_raw_kernel<D, U>;
        constexpr size_t smem = sizeof(double) * (size_t)A_MAXREC * Lay<D>::RS;
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_set = true;
        }
        kern<<<grid, A_THREADS, smem, ctx->stream>>>(A);
        ctx->launches++;
        launch_raw_from<D, U + 1>(ctx, A);
    }
}

template <int D>
void launch_pair(ipf_ctx* ctx, const RawArgs& R, const CombArgs& C, int ncfg, cudaStream_t st2) {
    {
        ProfScope prof_(ctx, "site_deriv_3b");
        launch_raw_from<D, 0>(ctx, R);
    }
    cudaEvent_t produced;
    cudaEventCreateWithFlags(&produced, cudaEventDisableTiming);
    cudaEventRecord(produced, ctx->stream);
    cudaStreamWaitEvent(st2, produced, 0);
    cudaEventDestroy(produced);
    {
        ProfScope prof_(ctx, "gather_3b", st2);
        constexpr size_t csmem = sizeof(double) * (size_t)C_NV * (Lay<D>::NM + C_CST);
        static bool cattr = false;
        if (!cattr) {
            cudaFuncSetAttribute(combine3b_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem);
            cattr = true;
        }
        combine3b_kernel<D><<<(unsigned)ncfg, C_THREADS, csmem, st2>>>(C);
        ctx->launches++;
    }
}

}  // namespace

int ipf_assemble_block_3b_mom(ipf_ctx* ctx, const BlockDev& blk, const ChunkDev& ch, const NeighList& nl,
                              ipf_matrix* Psi, bool* handled) {
    *handled = false;
    if (blk.N != 3 || blk.maxdeg > MAXD || blk.nb > C_NFT * C_THREADS) return IPF_OK;
    int D = 0;
    for (int m = 0; m < blk.nmono; ++m) {
        const uint8_t* e = &blk.h_mono_exp[(size_t)m * IPF_MAX_VARS];
        D = std::max(D, (int)e[0] + e[1] + e[2]);
    }
    if (D < 2 || D > MAXD) return IPF_OK;
    // canonical check + per-function descriptors (c, a, b)
    std::vector<int8_t> h_fdesc((size_t)4 * blk.nb, 0);
    int expect = 0;
    for (int deg = 1; deg <= D; ++deg)
        for (int a = 0; a <= deg; ++a)
            for (int b = a; a + b <= deg; ++b) {
                const int c = deg - a - b;
                if (expect >= blk.nb) return IPF_OK;
                const int m0 = blk.h_mono_start[expect], m1 = blk.h_mono_start[expect + 1];
                const int want = (a == b) ? 1 : 2;
                if (m1 - m0 != want) return IPF_OK;
                const uint8_t* e0 = &blk.h_mono_exp[(size_t)m0 * IPF_MAX_VARS];
                if (e0[0] != a || e0[1] != b || e0[2] != c) return IPF_OK;
                if (want == 2) {
                    const uint8_t* e1 = e0 + IPF_MAX_VARS;
                    if (e1[0] != b || e1[1] != a || e1[2] != c) return IPF_OK;
                }
                h_fdesc[4 * expect] = (int8_t)c; h_fdesc[4 * expect + 1] = (int8_t)a; h_fdesc[4 * expect + 2] = (int8_t)b;
                ++expect;
            }
    if (expect != blk.nb) return IPF_OK;
    *handled = true;
    if (ch.ncfg == 0 || nl.npairs == 0) return IPF_OK;

    cudaStream_t st = ctx->stream;
    const double* rec = nullptr;
    int rs = 0;
    IPF_CHECK(ipf_pair_records(ctx, blk, nl, &rec, &rs));
    ABuf<int8_t> fdesc;
    IPF_CHECK(fdesc.alloc(ctx, h_fdesc.size()));
    IPF_CUDA(ctx, cudaMemcpyAsync(fdesc.p, h_fdesc.data(), h_fdesc.size(), cudaMemcpyHostToDevice, st));
    IPF_CUDA(ctx, cudaStreamSynchronize(st));
    int NM = 0;
    switch (D) {
#define IPF_CASE(DD) case DD: NM = Lay<DD>::NM; break;
        IPF_CASE(2) IPF_CASE(3) IPF_CASE(4) IPF_CASE(5) IPF_CASE(6) IPF_CASE(7) IPF_CASE(8)
        IPF_CASE(9) IPF_CASE(10) IPF_CASE(11) IPF_CASE(12) IPF_CASE(13) IPF_CASE(14)
#undef IPF_CASE
    }
    if (rs != ((REC + D + 2) & ~1)) return ipf_set_error(ctx, IPF_EINVAL, "internal: record stride mismatch");
    // moment rows live in HBM, double buffered in batches of whole configurations
    const size_t m_budget = (size_t)3 << 30;
    const int64_t max_pairs_batch = std::max<int64_t>(nl.maxpairs_cfg, (int64_t)(m_budget / (8 * (size_t)NM)));
    const size_t m_bytes = (size_t)8 * NM * (size_t)std::min<int64_t>(max_pairs_batch, nl.npairs);
    void* Mbuf[2] = {nullptr, nullptr};
    IPF_CHECK(ipf_scratch(ctx, 3, m_bytes, &Mbuf[0]));
    IPF_CHECK(ipf_scratch(ctx, 4, m_bytes, &Mbuf[1]));
    if (ctx->profiling && !nl.h_site_off.empty()) {
        double tot = 0.0;
        for (int64_t s = 0; s < nl.nsites; ++s) {
            const double n = (double)(nl.h_site_off[s + 1] - nl.h_site_off[s]);
            tot += 0.5 * n * (n - 1.0);
        }
        ctx->prof["count:clusters_3b"].ms += tot;
        ctx->prof["count:ordered_pairs_3b"].ms += 2.0 * tot;
        ctx->prof["count:pairs_3b"].ms += (double)nl.npairs;
    }
    static const bool no_overlap = getenv("IPF_NO_OVERLAP") != nullptr;
    cudaStream_t st2 = no_overlap ? ctx->stream : ctx->stream2;
    cudaEvent_t consumed[2] = {nullptr, nullptr};
    int64_t c0 = 0;
    int ib = 0;
    while (c0 < ch.ncfg) {
        int64_t c1 = c0 + 1;
        while (c1 < ch.ncfg && nl.h_cfg_pair_off[c1 + 1] - nl.h_cfg_pair_off[c0] <= max_pairs_batch) ++c1;
        const int buf = ib & 1;
        if (consumed[buf]) {
            IPF_CUDA(ctx, cudaStreamWaitEvent(st, consumed[buf], 0));
            cudaEventDestroy(consumed[buf]);
            consumed[buf] = nullptr;
        }
        RawArgs R;
        R.p0 = nl.h_cfg_pair_off[c0]; R.np = nl.h_cfg_pair_off[c1] - R.p0; R.rec = rec;
        R.pbeg = nl.pbeg.p; R.pcnt = nl.pcnt.p; R.M = (double*)Mbuf[buf];
        CombArgs C;
        C.p0 = R.p0; C.c0 = c0; C.nb = blk.nb; C.rs = rs; C.M = (const double*)Mbuf[buf]; C.rec = rec; C.fdesc = fdesc.p;
        C.atom_off = ch.atom_off.p; C.site_off = nl.site_off.p; C.prev = nl.prev.p; C.pR = nl.pR.p;
        C.rowE = ch.rowE.p; C.rowF = ch.rowF.p; C.rowV = ch.rowV.p; C.Psi = Psi->d; C.ld = Psi->ld; C.col0 = blk.col0;
        if (R.np > 0) {
            switch (D) {
#define IPF_CASE(DD) case DD: launch_pair<DD>(ctx, R, C, (int)(c1 - c0), st2); break;
                IPF_CASE(2) IPF_CASE(3) IPF_CASE(4) IPF_CASE(5) IPF_CASE(6) IPF_CASE(7) IPF_CASE(8)
                IPF_CASE(9) IPF_CASE(10) IPF_CASE(11) IPF_CASE(12) IPF_CASE(13) IPF_CASE(14)
#undef IPF_CASE
            }
        }
        IPF_CUDA(ctx, cudaEventCreateWithFlags(&consumed[buf], cudaEventDisableTiming));
        IPF_CUDA(ctx, cudaEventRecord(consumed[buf], st2));
        c0 = c1;
        ++ib;
    }
    for (int b = 0; b < 2; ++b)
        if (consumed[b]) {
            IPF_CUDA(ctx, cudaStreamWaitEvent(st, consumed[b], 0));
            cudaEventDestroy(consumed[b]);
        }
    IPF_CUDA(ctx, cudaGetLastError());
    return IPF_OK;
}
