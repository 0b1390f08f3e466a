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

}  // namespace
