// basis_3b_shared.cuh -- pieces shared by the 3-body fast-path kernels (basis_3b.cu, basis_3b_cfg.cu):
// record layout, function counting, unit tables and the moment accumulation loop.
#pragma once
#include "ipf_common.cuh"

namespace {

constexpr int REC = 8;
constexpr int MAXD = 14;

// number of basis functions with w-exponent c (degree-D basis): pairs a <= b, a + b <= D - c, minus the constant
__host__ __device__ constexpr int nfun_c(int D, int c) {
    int n = 0;
    for (int a = 0; a <= (D - c) / 2; ++a) n += (D - c - a) - a + 1;
    return n - (c == 0 ? 1 : 0);
}
__host__ __device__ constexpr int scol_base(int D, int C0) {
    int n = 0;
    for (int c = 0; c < C0; ++c) n += nfun_c(D, c);
    return n;
}

__host__ __device__ constexpr int cmidx(int c, int a, int b) { return (c * (MAXD + 1) + a) * (MAXD + 1) + b; }

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

template <int D>
struct RecStride { static constexpr int value = (REC + D + 2) & ~1; };
// compact record staged by the config-resident kernel: [Rhat(3), fc, x^0..x^D]
template <int D>
struct CRecStride { static constexpr int value = (4 + D + 2) & ~1; };

// moment accumulation over the neighbours of the thread's centre atom; RK points at the record of
// the first neighbour (shared or global memory: the caller instantiates one copy per address space)
template <int D, int C0, int C1, int NC, int EMAX, bool COMPACT = false, int RSTRIDE = 0>
__device__ __forceinline__ void moment_loop(const double* __restrict__ rk, int n, int jself, double Rj0, double Rj1,
                                            double Rj2, const double (&E1)[3], const double (&E2)[3],
                                            double (&Z)[NC][EMAX + 1], double (&U)[NC][EMAX + 1][2]) {
    constexpr int RS = RSTRIDE ? RSTRIDE : (COMPACT ? CRecStride<D>::value : RecStride<D>::value);
    constexpr int OX = COMPACT ? 4 : REC;
    for (int kk = 0; kk < n; ++kk, rk += RS) {
        const double2 r01 = *reinterpret_cast<const double2*>(rk);
        const double2 r23 = *reinterpret_cast<const double2*>(rk + 2);
        double fc_k;
        if constexpr (COMPACT) fc_k = r23.y;
        else fc_k = reinterpret_cast<const double2*>(rk + 6)->x;      // (fc, dfc)
        double xe[EMAX + 2];
#pragma unroll
        for (int e = 0; e <= EMAX; e += 2) {
            const double2 v = *reinterpret_cast<const double2*>(rk + OX + e);
            xe[e] = v.x; xe[e + 1] = v.y;
        }
        const double w = Rj0 * r01.x + Rj1 * r01.y + Rj2 * r23.x;
        const double q1 = E1[0] * r01.x + E1[1] * r01.y + E1[2] * r23.x;      // in-plane components of Rhat_k
        const double q2 = E2[0] * r01.x + E2[1] * r01.y + E2[2] * r23.x;
        const double fck = (kk == jself) ? 0.0 : fc_k;      // the own leg is not its own partner
        // wl[ci] = fck w^(c-1), c = C0 + ci  (unused for c = 0)
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
                        const double t = xe[e] * wl[ci];          // fck x^e w^(c-1)
                        U[ci][e][0] = fma(t, q1, U[ci][e][0]);
                        U[ci][e][1] = fma(t, q2, U[ci][e][1]);
                        Z[ci][e] = fma(t, w, Z[ci][e]);
                    }
                }
            }
        }
    }
}

// unit tables: consecutive c packed while the accumulators fit (<= 44 doubles)
template <int D>
struct UnitTable {
    int n = 0;
    int c0[MAXD + 1] = {}, c1[MAXD + 1] = {};
    int base[MAXD + 2] = {};      // first scratch column of every unit (unit sizes padded to even), [n] = row length
    constexpr UnitTable() {
        int c = 0;
        int b = 0;
        while (c <= D) {
            int acc = 0, e = c;
            while (e <= D) {
                const int add = (e == 0 ? 1 : 3) * (D - e + 1);
                if (acc > 0 && acc + add > 36) break;
                acc += add;
                ++e;
            }
            c0[n] = c; c1[n] = e - 1;
            base[n] = b;
            int nf = 0;
            for (int cc = c; cc < e; ++cc) nf += nfun_c(D, cc);
            b += (nf + 1) & ~1;
            ++n;
            c = e;
        }
        base[n] = b;
    }
};

// The gather pass is shared by the 3-body scratch path (basis_3b.cu) and the 4-body path (basis_4b.cu): both producers
// write g_p to S[pair][scol][3] and the per-centre-atom own sums to O[site][slot][scol][4].
// ---- gather: S (reverse entries) + O (own sums) -> E / F / V rows ------------------
struct Gather3Args {
    int64_t p0;               // first pair of the batch
    int64_t c0;               // first config (chunk-local) of the batch
    int64_t s0;               // first site of the batch
    int nb;
    const double* S;          // [np][nb][3] (gstride = 0) or group-major [nb / 8][np][8][3] (gstride = 24 np)
    int64_t pstride, gstride; // doubles between consecutive pairs / between consecutive groups of 8 columns
    int cfg_aligned;          // 1: the producer's 32-row blocks start at the first pair of every configuration
    const double* O;          // [nsites][nslot][nb][4]
    int nslot;
    const int16_t* scol2col;  // [nb] scratch column -> local basis column
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

constexpr int G_THREADS = 256;
constexpr int G_WARPS = G_THREADS / 32;
constexpr int G_ACH = 64;         // atoms per shared-memory tile
constexpr int G_LDT = 3 * G_ACH + 1;

// CTA = (config, 32 scratch columns); lanes = columns, warps split the atoms.
__global__ void __launch_bounds__(G_THREADS) gather3b_kernel(Gather3Args A) {
    extern __shared__ double g_smem[];
    double* Fs = g_smem;                                            // [32][G_LDT]
    double (*part)[7][32] = reinterpret_cast<double (*)[7][32]>(g_smem + 32 * G_LDT);   // [G_WARPS][7][32]
    const int c = (int)(A.c0 + blockIdx.x);
    const int sc = blockIdx.y * 32 + (threadIdx.x & 31);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool live = sc < A.nb && A.scol2col[sc] >= 0;
    const int scl = sc < A.nb ? sc : 0;
    const int64_t a0 = A.atom_off[c];
    const int nat = (int)(A.atom_off[c + 1] - a0);
    const int64_t rE = A.rowE[c], rF = A.rowF[c], rV = A.rowV[c];
    const double* S3 = A.gstride ? A.S + (int64_t)(scl >> 3) * A.gstride + 3 * (scl & 7) : A.S + 3 * scl;
    const double4* O4 = reinterpret_cast<const double4*>(A.O) + scl;
    double e = 0.0, v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0, v4 = 0.0, v5 = 0.0;
    for (int ab = 0; ab < nat; ab += G_ACH) {
        const int na = min(G_ACH, nat - ab);
        for (int al = warp; al < na; al += G_WARPS) {
            const int64_t s = a0 + ab + al;
            const int64_t q0 = A.site_off[s], q1 = A.site_off[s + 1];
            double f0 = 0.0, f1 = 0.0, f2 = 0.0;
            if (q1 > q0) {
                // own sums: one slot per 32-row block of the producer that holds pairs of this atom
                const int64_t b0 = A.cfg_aligned ? A.site_off[a0] : A.p0;
                const int ns = (int)(((q1 - 1 - b0) >> 5) - ((q0 - b0) >> 5)) + 1;
                for (int sl = 0; sl < ns; ++sl) {
                    const double4 o = O4[((s - A.s0) * A.nslot + sl) * A.nb];
                    f0 += o.x; f1 += o.y; f2 += o.z; e += o.w;
                }
            }
#pragma unroll 4
            for (int64_t p = q0; p < q1; ++p) {
                const double* h = S3 + ((int64_t)A.prev[p] - A.p0) * A.pstride;
                const double hx = __ldg(h), hy = __ldg(h + 1), hz = __ldg(h + 2);
                const double r0 = A.pR[3 * p], r1 = A.pR[3 * p + 1], r2 = A.pR[3 * p + 2];
                f0 -= hx; f1 -= hy; f2 -= hz;
                v0 += hx * r0; v1 += hy * r1; v2 += hz * r2;      // xx, yy, zz
                v3 += hz * r1; v4 += hz * r0; v5 += hy * r0;      // zy, zx, yx
            }
            Fs[lane * G_LDT + 3 * al] = f0;
            Fs[lane * G_LDT + 3 * al + 1] = f1;
            Fs[lane * G_LDT + 3 * al + 2] = f2;
        }
        __syncthreads();
        if (rF) {
            const int nrow = 3 * na;
            for (int cl = 0; cl < 32; ++cl) {
                const int scc = blockIdx.y * 32 + cl;
                if (scc >= A.nb) break;
                if (A.scol2col[scc] < 0) continue;
                double* colp = A.Psi + (size_t)(A.col0 + A.scol2col[scc]) * A.ld + (rF - 1) + 3 * ab;
                for (int r = threadIdx.x; r < nrow; r += G_THREADS) colp[r] = Fs[cl * G_LDT + r];
            }
        }
        __syncthreads();
    }
    part[warp][0][lane] = e; part[warp][1][lane] = v0; part[warp][2][lane] = v1; part[warp][3][lane] = v2;
    part[warp][4][lane] = v3; part[warp][5][lane] = v4; part[warp][6][lane] = v5;
    __syncthreads();
    if (warp == 0 && live) {
        double* colp = A.Psi + (size_t)(A.col0 + A.scol2col[sc]) * A.ld;
#pragma unroll
        for (int comp = 0; comp < 7; ++comp) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < G_WARPS; ++w) t += part[w][comp][lane];
            if (comp == 0) { if (rE) colp[rE - 1] = t; }
            else if (rV) colp[rV - 1 + comp - 1] = t;
        }
    }
}

}  // namespace
