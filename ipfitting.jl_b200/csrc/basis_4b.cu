// basis_4b.cu -- K3 fast path for 4-body bond-angle bases (nbpolys(4, desc, D), D <= 8; any S3-closed set of
// monomial orbits in (x1, x2, x3; w12, w13, w23)).
//
// Own-leg decomposition as in the other kernels: the thread of ordered pair p = (i, j) owns
//     g_p = dV_b(i)/dR_ij    and    e_p = its third of V_b(i),
// but the double sum over the other two legs (k, l) is FACTORISED.  In the frame of the own leg
// (u_k = (w_jk, q1_k, q2_k): components of Rhat_k along Rhat_j and along an orthonormal basis (e1, e2) of the plane
// perpendicular to it) the coupling of the other two legs is w_kl = u_k . u_l, so
//     w_kl^c = sum_{|nu| = c} (c!/nu!) u_k^nu u_l^nu
// and for a monomial m = x_j^a1 x_k^a2 x_l^a3 w_jk^c12 w_jl^c13 w_kl^c23
//     S(m)  = sum_{k != l} fc_k fc_l x_k^a2 x_l^a3 w_jk^c12 w_jl^c13 w_kl^c23
//           = sum_nu kappa_nu M[a2, c12 e0 + nu] M[a3, c13 e0 + nu]  -  Z2[a2 + a3, c12 + c13]
//     T'(m) = in-plane gradient with respect to Rhat_j through w_jk only
//           = c12 ( sum_nu kappa_nu M[a2, (c12-1) e0 + nu + d_i] M[a3, c13 e0 + nu] - Z2q_i[a2 + a3, c12 + c13 - 1] ),  i = 1, 2
// with the per-own-leg MOMENTS  M[e, mu] = sum_{k != j} fc_k x_k^e u_k^mu  (495 for D = 8),
// Z2[e, c] = sum_k fc_k^2 x_k^e w^c,  Z2q_i[e, c] = sum_k fc_k^2 x_k^e w^c q_i  (the k = l terms).  Then, with the sums
// running over the distinct monomials of the orbit of basis function b,
//     S0 = 1/2 sum_m x_j^a1 S(m)      S1 = 1/2 sum_m a1 x_j^(a1-1) S(m)      T = sum_m x_j^a1 T'(m)
//     g_p = (fc_j' S0 + fc_j x_j' S1) Rhat_j + (fc_j / r_j)(T_1 e1 + T_2 e2)        e_p = fc_j S0 / 3.
// (sum_m T(m) = 2 sum_m T'(m) because the orbit is closed under the exchange of the two other legs.)
// Per own leg this is ~20k FMAs for the moments + ~35k for the combination instead of C(n-1, 2) clusters x 3002 monomials:
// two orders of magnitude fewer FP64 operations than the table-driven kernel (tools/proto4b.py is the numpy statement
// of exactly this algorithm, checked against the oracle).
//
// Mapping.  CTA = 32 consecutive ordered pairs (rows) x G4_NW worker warps; lane = row.
//   phase 1: worker w accumulates its share of the 612 moments of the 32 rows in registers over the k loop
//            (generated straight-line bodies, basis_4b_gen.inc) and stores them to shared memory M[moment][row];
//   phase 2: worker w evaluates its share of the basis functions from table-driven product lists (warp-uniform
//            indices, conflict-free [moment][row] reads), stages (g_p, e_p) of 4 functions at a time and flushes
//            g_p to the HBM scratch S[pair][scol][3] and the per-centre-atom own sums to O[site][slot][scol][4];
//   then gather3b_kernel (basis_3b_shared.cuh) turns S and O into the E / F / V rows in a fixed order (no atomics).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <array>
#include <map>

#include "basis_3b_shared.cuh"
#include "basis_4b_gen.inc"

int ipf_pair_records(ipf_ctx* ctx, const BlockDev& blk, const NeighList& nl, const double** rec, int* rs);

namespace {

constexpr int Q_NG = 4;                     // functions per staged output group
constexpr int Q_LD = 4 * Q_NG + 2;          // doubles per staged row: g(3) x 4, e x 4, pad 2
constexpr int Q_EOFF = 3 * Q_NG;
constexpr int Q_THREADS = 32 * G4_NW;
constexpr size_t Q_SMEM = sizeof(double) * ((size_t)G4_NROWS * 32 + (size_t)(G4_D + 1) * 32 + (size_t)G4_NW * 32 * Q_LD);

struct __align__(16) T4Term {               // one nu of one entry
    uint32_t offA, offB, offA1, offA2, offB1, offB2;      // byte offsets of moment rows in shared memory
    double kappa;
};
struct __align__(16) T4Ent {                // one class {m, m'} of monomials (m' = m with the two other legs exchanged)
    uint32_t term0;
    uint16_t nterm;
    uint8_t a1, flags;                      // flags: bit0 = raised-A products (c12 > 0), bit1 = raised-B products
    uint32_t offZ2, offZq1, offZq2;
    float cA, cB, wS;
};
static_assert(sizeof(T4Term) == 32 && sizeof(T4Ent) == 32, "table records are read as two 16-byte words");

struct Site4Args {
    int64_t p0, np;                 // batch pair range
    const double* rec;              // standard pair records (own leg: Rhat, 1/r, x, x', fc, fc')
    int rs;
    const double* rec4;             // compact neighbour records [Rhat(3), fc, fc x^0..x^D, pad]
    const int32_t* psite;
    const int64_t* pbeg;
    const int32_t* pcnt;
    const T4Term* term;
    const T4Ent* ent;
    const uint32_t* ent0;           // [nbs + 1] entries of every scratch column
    int wsc[G4_NW + 1];             // scratch-column range of every worker (multiples of Q_NG)
    int nb;                         // scratch columns (row length of S)
    double* S;
    double* O;
    int64_t s0;
    int nslot;
};

__global__ void __launch_bounds__(128) rec4_kernel(int64_t npairs, const double* __restrict__ rec, int rs,
                                                   double* __restrict__ rec4) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    const double* r = rec + (int64_t)rs * p;
    double* o = rec4 + (int64_t)G4_RS * p;
    const double x = r[4], fc = r[6];
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = fc;
    double v = fc;
#pragma unroll
    for (int e = 0; e <= G4_D; ++e) { o[4 + e] = v; v *= x; }
    for (int e = 4 + G4_D + 1; e < G4_RS; ++e) o[e] = 0.0;
}

__device__ __forceinline__ double q_column_sum(const double* __restrict__ stage, int r0, int r1, int lane) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int r = r0;
    for (; r + 3 < r1; r += 4) {
        a0 += stage[r * Q_LD + lane];
        a1 += stage[(r + 1) * Q_LD + lane];
        a2 += stage[(r + 2) * Q_LD + lane];
        a3 += stage[(r + 3) * Q_LD + lane];
    }
    for (; r < r1; ++r) a0 += stage[r * Q_LD + lane];
    return (a0 + a1) + (a2 + a3);
}

// write-out of the Q_NG staged functions of the warp's 32 rows: S rows (96 contiguous bytes per pair) and the
// per-centre-atom column sums O (same protocol as flush_group of basis_3b.cu, which gather3b_kernel reads)
__device__ __forceinline__ void q_flush(const Site4Args& A, const double* __restrict__ wstage, int64_t plw, int nrw,
                                        int sc0, unsigned pmask, int pout) {
    __syncwarp();
    const int lane = threadIdx.x & 31;
    constexpr int CH = 3 * Q_NG / 2;
    double2* S2 = reinterpret_cast<double2*>(A.S);
    const double2* st2 = reinterpret_cast<const double2*>(wstage);
    for (int q = lane; q < nrw * CH; q += 32) {
        const int row = q / CH, c = q - row * CH;
        S2[((plw + row) * A.nb + sc0) * 3 / 2 + c] = st2[row * (Q_LD / 2) + c];
    }
    const bool vlive = lane < 4 * Q_NG;
    const int f = lane < Q_EOFF ? lane / 3 : lane - Q_EOFF;
    const int d = lane < Q_EOFF ? lane - 3 * f : 3;
    unsigned m = pmask;
    while (m) {
        const int r0 = __ffs(m) - 1;
        m &= m - 1;
        const int r1 = m ? (__ffs(m) - 1) : nrw;
        const double acc = vlive ? q_column_sum(wstage, r0, r1, lane) : 0.0;
        const int orow = __shfl_sync(0xffffffffu, pout, r0);
        if (vlive) A.O[((int64_t)orow * A.nb + sc0 + f) * 4 + d] = acc;
    }
    __syncwarp();
}

template <bool HA, bool HB>
__device__ __forceinline__ void q_terms(const T4Term* __restrict__ t, int nterm, const char* __restrict__ Mb, double& s,
                                        double& a1, double& a2, double& b1, double& b2) {
    for (int i = 0; i < nterm; ++i, ++t) {
        const uint4 o0 = __ldg(reinterpret_cast<const uint4*>(t));
        const uint4 o1 = __ldg(reinterpret_cast<const uint4*>(t) + 1);
        const double kappa = __hiloint2double((int)o1.w, (int)o1.z);
        const double Av = *reinterpret_cast<const double*>(Mb + o0.x);
        const double Bk = kappa * *reinterpret_cast<const double*>(Mb + o0.y);
        s = fma(Av, Bk, s);
        if constexpr (HA) {
            a1 = fma(*reinterpret_cast<const double*>(Mb + o0.z), Bk, a1);
            a2 = fma(*reinterpret_cast<const double*>(Mb + o0.w), Bk, a2);
        }
        if constexpr (HB) {
            const double Ak = kappa * Av;
            b1 = fma(*reinterpret_cast<const double*>(Mb + o1.x), Ak, b1);
            b2 = fma(*reinterpret_cast<const double*>(Mb + o1.y), Ak, b2);
        }
    }
}

__global__ void __maxnreg__(168) site4b_kernel(Site4Args A) {
    extern __shared__ __align__(16) double sm[];
    double* M = sm;                                          // [G4_NROWS][32]
    double* xp = sm + (size_t)G4_NROWS * 32;                 // [G4_D + 1][32]  x_j^e
    double* stage = xp + (size_t)(G4_D + 1) * 32;            // [G4_NW][32][Q_LD]
    const int lane = threadIdx.x & 31, worker = threadIdx.x >> 5;
    const int64_t pl0 = (int64_t)blockIdx.x * 32;
    const int64_t pl = pl0 + lane;
    const bool live = pl < A.np;
    const int nrw = (int)min((int64_t)32, A.np - pl0);
    const int64_t p = A.p0 + (live ? pl : pl0);
    const int64_t k0 = A.pbeg[p];
    const int n = live ? A.pcnt[p] : 0;
    const int site = A.psite[p];
    const int site_prev = __shfl_up_sync(0xffffffffu, site, 1);
    const bool piece_start = live && (lane == 0 || site != site_prev);
    const unsigned pmask = __ballot_sync(0xffffffffu, piece_start);
    const int pout = (int)(site - A.s0) * A.nslot + (int)((pl >> 5) - ((k0 - A.p0) >> 5));
    const double* recp = A.rec + (int64_t)A.rs * p;
    const double4 o0v = ld4(recp);
    const double4 o1v = ld4(recp + 4);
    const double Rj0 = o0v.x, Rj1 = o0v.y, Rj2 = o0v.z, rinvj = o0v.w;
    const double xj = o1v.x, dxj = o1v.y, fcj = o1v.z, dfcj = o1v.w;
    double E1[3], E2[3];
    {
        const bool usex = fabs(Rj0) < 0.8;
        const double c0 = usex ? 0.0 : -Rj2, c1 = usex ? Rj2 : 0.0, c2 = usex ? -Rj1 : Rj0;
        const double inv = rsqrt(c0 * c0 + c1 * c1 + c2 * c2);
        E1[0] = c0 * inv; E1[1] = c1 * inv; E1[2] = c2 * inv;
        E2[0] = Rj1 * E1[2] - Rj2 * E1[1];
        E2[1] = Rj2 * E1[0] - Rj0 * E1[2];
        E2[2] = Rj0 * E1[1] - Rj1 * E1[0];
    }
    if (worker == 0) {
        double v = 1.0;
#pragma unroll
        for (int e = 0; e <= G4_D; ++e) { xp[e * 32 + lane] = v; v *= xj; }
    }
    // ---- phase 1: moments of the 32 rows
    g4_phase1(worker, A.rec4 + k0 * G4_RS, n, (int)(p - k0), Rj0, Rj1, Rj2, E1, E2, M + lane);
    __syncthreads();

    // ---- phase 2: basis functions of this worker, 4 scratch columns at a time
    const char* Mb = reinterpret_cast<const char*>(M + lane);
    double* wstage = stage + (size_t)worker * 32 * Q_LD;
    double* myrow = wstage + lane * Q_LD;
    const double tang = fcj * rinvj, fdx = fcj * dxj, third = fcj * (1.0 / 3.0);
    for (int sc = A.wsc[worker]; sc < A.wsc[worker + 1]; sc += Q_NG) {
#pragma unroll 1
        for (int f = 0; f < Q_NG; ++f) {
            const uint32_t e0 = __ldg(A.ent0 + sc + f), e1 = __ldg(A.ent0 + sc + f + 1);
            double S0 = 0.0, S1 = 0.0, T1 = 0.0, T2 = 0.0;
            for (uint32_t ei = e0; ei < e1; ++ei) {
                const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(A.ent + ei));
                const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(A.ent + ei) + 1);
                const int nterm = (int)(w0.y & 0xffffu), a1 = (int)((w0.y >> 16) & 0xffu), flags = (int)(w0.y >> 24);
                const T4Term* t = A.term + w0.x;
                double s = 0.0, ta1 = 0.0, ta2 = 0.0, tb1 = 0.0, tb2 = 0.0;
                if (flags == 3) q_terms<true, true>(t, nterm, Mb, s, ta1, ta2, tb1, tb2);
                else if (flags == 1) q_terms<true, false>(t, nterm, Mb, s, ta1, ta2, tb1, tb2);
                else q_terms<false, false>(t, nterm, Mb, s, ta1, ta2, tb1, tb2);
                s -= *reinterpret_cast<const double*>(Mb + w0.z);
                const double xa = xp[a1 * 32 + lane];
                const double xd = a1 > 0 ? (double)a1 * xp[(a1 - 1) * 32 + lane] : 0.0;
                const double wS = (double)__uint_as_float(w1.w);
                const double ws = wS * s;
                S0 = fma(xa, ws, S0);
                S1 = fma(xd, ws, S1);
                if (flags) {
                    const double z1 = *reinterpret_cast<const double*>(Mb + w0.w);
                    const double z2 = *reinterpret_cast<const double*>(Mb + w1.x);
                    const double cA = (double)__uint_as_float(w1.y), cB = (double)__uint_as_float(w1.z);
                    // c12 (tA - z) + c13 (tB - z)
                    const double u1 = fma(cA, ta1 - z1, cB * (tb1 - z1));
                    const double u2 = fma(cA, ta2 - z2, cB * (tb2 - z2));
                    T1 = fma(xa, u1, T1);
                    T2 = fma(xa, u2, T2);
                }
            }
            const double radial = fma(dfcj, S0, fdx * S1);
            T1 *= tang; T2 *= tang;
            myrow[3 * f] = fma(radial, Rj0, fma(T1, E1[0], T2 * E2[0]));
            myrow[3 * f + 1] = fma(radial, Rj1, fma(T1, E1[1], T2 * E2[1]));
            myrow[3 * f + 2] = fma(radial, Rj2, fma(T1, E1[2], T2 * E2[2]));
            myrow[Q_EOFF + f] = third * S0;
        }
        q_flush(A, wstage, pl0, nrw, sc, pmask, pout);
    }
}

// ---- host tables --------------------------------------------------------------------------------------------
struct Tables4 {
    int D = 0, nbs = 0;
    std::vector<T4Term> term;
    std::vector<T4Ent> ent;
    std::vector<uint32_t> ent0;
    std::vector<int16_t> s2c;
    int wsc[G4_NW + 1] = {};
    double flops_leg = 0.0;         // combine flops per own leg (FMA = 2)
    T4Term* d_term = nullptr;
    T4Ent* d_ent = nullptr;
    uint32_t* d_ent0 = nullptr;
    int16_t* d_s2c = nullptr;
};

struct Mono { int a[3], c12, c13, c23; };      // slots 1,2,3; pairs (1,2), (1,3), (2,3)

inline uint32_t row_off(int row) { return (uint32_t)row * 32u * 8u; }
inline int midx(int e, int m0, int m1, int m2) {
    constexpr int n1 = G4_D + 1;
    if (e < 0 || m0 < 0 || m1 < 0 || m2 < 0 || e >= n1 || m0 >= n1 || m1 >= n1 || m2 >= n1) return -1;
    return g4_midx[((e * n1 + m0) * n1 + m1) * n1 + m2];
}

// returns false when the block is not a set of S3-closed orbits of total degree <= G4_D
bool build_tables(const BlockDev& blk, Tables4& T) {
    const int nb = blk.nb;
    struct Cls { int a1; int a2, c12, a3, c13, c23; bool self; };
    std::vector<std::vector<Cls>> fcls(nb);
    std::vector<double> cost(nb, 0.0);
    int D = 0;
    for (int b = 0; b < nb; ++b) {
        const int m0 = blk.h_mono_start[b], m1 = blk.h_mono_start[b + 1];
        std::vector<std::array<int, 6>> mons;
        for (int m = m0; m < m1; ++m) {
            const uint8_t* e = &blk.h_mono_exp[(size_t)m * IPF_MAX_VARS];
            std::array<int, 6> v = {e[0], e[1], e[2], e[3], e[4], e[5]};
            const int deg = v[0] + v[1] + v[2] + v[3] + v[4] + v[5];
            if (deg < 1 || deg > G4_D) return false;
            D = std::max(D, deg);
            mons.push_back(v);
        }
        std::sort(mons.begin(), mons.end());
        if (std::adjacent_find(mons.begin(), mons.end()) != mons.end()) return false;
        auto has = [&](const std::array<int, 6>& v) { return std::binary_search(mons.begin(), mons.end(), v); };
        // closure under the two generators of S3 acting on the slots: (2 3) and (1 2)
        for (const auto& v : mons) {
            const std::array<int, 6> s23 = {v[0], v[2], v[1], v[4], v[3], v[5]};       // w12<->w13
            const std::array<int, 6> s12 = {v[1], v[0], v[2], v[3], v[5], v[4]};       // w13<->w23
            if (!has(s23) || !has(s12)) return false;
        }
        for (const auto& v : mons) {
            const std::array<int, 6> sw = {v[0], v[2], v[1], v[4], v[3], v[5]};
            const bool self = sw == v;
            if (!self && sw < v) continue;          // the class is emitted from its smaller member
            Cls c;
            c.a1 = v[0]; c.c23 = v[5]; c.self = self;
            // orientation: the A side carries a raised product whenever one of the two sides does
            const bool flip = !self && v[3] == 0 && v[4] > 0;
            if (!flip) { c.a2 = v[1]; c.c12 = v[3]; c.a3 = v[2]; c.c13 = v[4]; }
            else { c.a2 = v[2]; c.c12 = v[4]; c.a3 = v[1]; c.c13 = v[3]; }
            fcls[b].push_back(c);
            const int nt = (c.c23 + 1) * (c.c23 + 2) / 2;
            const bool ha = c.c12 > 0, hb = !self && c.c13 > 0;
            cost[b] += nt * (3.0 + (ha ? 4.0 : 0.0) + (hb ? 5.0 : 0.0)) + 14.0;
        }
        cost[b] += 25.0;
    }
    T.D = D;
    // functions -> workers (longest processing time first), each worker's list padded to a multiple of Q_NG
    std::vector<int> order(nb);
    for (int b = 0; b < nb; ++b) order[b] = b;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
    std::vector<std::vector<int>> wl(G4_NW);
    std::vector<double> wload(G4_NW, 0.0);
    for (int b : order) {
        int w = 0;
        for (int k = 1; k < G4_NW; ++k) if (wload[k] < wload[w]) w = k;
        wl[w].push_back(b);
        wload[w] += cost[b];
    }
    T.term.clear(); T.ent.clear(); T.ent0.clear(); T.s2c.clear();
    T.flops_leg = 0.0;
    int sc = 0;
    for (int w = 0; w < G4_NW; ++w) {
        T.wsc[w] = sc;
        std::sort(wl[w].begin(), wl[w].end());
        const int npad = (int)((wl[w].size() + Q_NG - 1) / Q_NG * Q_NG);
        for (int q = 0; q < npad; ++q) {
            T.ent0.push_back((uint32_t)T.ent.size());
            if (q >= (int)wl[w].size()) { T.s2c.push_back(-1); ++sc; continue; }
            const int b = wl[w][q];
            T.s2c.push_back((int16_t)b);
            ++sc;
            T.flops_leg += cost[b];
            for (const Cls& c : fcls[b]) {
                T4Ent E;
                memset(&E, 0, sizeof(E));
                E.term0 = (uint32_t)T.term.size();
                const bool ha = c.c12 > 0, hb = !c.self && c.c13 > 0;
                E.a1 = (uint8_t)c.a1;
                E.flags = (uint8_t)((ha ? 1 : 0) | (hb ? 2 : 0));
                if (hb && !ha) return false;      // excluded by the orientation rule
                E.cA = (float)c.c12; E.cB = hb ? (float)c.c13 : 0.f;
                E.wS = c.self ? 0.5f : 1.0f;
                const int zr = g4_z2idx[(c.a2 + c.a3) * (G4_D + 1) + c.c12 + c.c13];
                if (zr < 0) return false;
                E.offZ2 = row_off(zr);
                if (ha) {
                    const int q1 = g4_z2qidx[((0) * (G4_D + 1) + c.a2 + c.a3) * (G4_D + 1) + c.c12 + c.c13 - 1];
                    const int q2 = g4_z2qidx[((1) * (G4_D + 1) + c.a2 + c.a3) * (G4_D + 1) + c.c12 + c.c13 - 1];
                    if (q1 < 0 || q2 < 0) return false;
                    E.offZq1 = row_off(q1); E.offZq2 = row_off(q2);
                }
                int nt = 0;
                for (int n0 = 0; n0 <= c.c23; ++n0)
                    for (int n1 = 0; n1 <= c.c23 - n0; ++n1) {
                        const int n2 = c.c23 - n0 - n1;
                        double kap = 1.0;      // c23! / (n0! n1! n2!)
                        {
                            auto fact = [](int k) { double f = 1.0; for (int i = 2; i <= k; ++i) f *= i; return f; };
                            kap = fact(c.c23) / (fact(n0) * fact(n1) * fact(n2));
                        }
                        T4Term t;
                        memset(&t, 0, sizeof(t));
                        const int rA = midx(c.a2, c.c12 + n0, n1, n2), rB = midx(c.a3, c.c13 + n0, n1, n2);
                        if (rA < 0 || rB < 0) return false;
                        t.offA = row_off(rA); t.offB = row_off(rB);
                        if (ha) {
                            const int r1 = midx(c.a2, c.c12 - 1 + n0, n1 + 1, n2), r2 = midx(c.a2, c.c12 - 1 + n0, n1, n2 + 1);
                            if (r1 < 0 || r2 < 0) return false;
                            t.offA1 = row_off(r1); t.offA2 = row_off(r2);
                        }
                        if (hb) {
                            const int r1 = midx(c.a3, c.c13 - 1 + n0, n1 + 1, n2), r2 = midx(c.a3, c.c13 - 1 + n0, n1, n2 + 1);
                            if (r1 < 0 || r2 < 0) return false;
                            t.offB1 = row_off(r1); t.offB2 = row_off(r2);
                        }
                        t.kappa = kap;
                        T.term.push_back(t);
                        ++nt;
                    }
                E.nterm = (uint16_t)nt;
                T.ent.push_back(E);
            }
        }
    }
    T.wsc[G4_NW] = sc;
    T.ent0.push_back((uint32_t)T.ent.size());
    T.nbs = sc;
    return true;
}

void free_tables(void* q) {
    Tables4* T = (Tables4*)q;
    if (!T) return;
    if (T->d_term) cudaFree(T->d_term);
    if (T->d_ent) cudaFree(T->d_ent);
    if (T->d_ent0) cudaFree(T->d_ent0);
    if (T->d_s2c) cudaFree(T->d_s2c);
    delete T;
}

}  // namespace

void ipf_basis4b_forget(ipf_ctx* ctx, const ipf_basis* B) {
    for (const BlockDev& b : B->blocks) {
        auto jt = ctx->tabs4b.find(&b);
        if (jt == ctx->tabs4b.end()) continue;
        free_tables(jt->second);
        ctx->tabs4b.erase(jt);
    }
}

void ipf_basis4b_forget_ctx(ipf_ctx* ctx) {
    for (auto& kv : ctx->tabs4b) free_tables(kv.second);
    ctx->tabs4b.clear();
}

int ipf_assemble_block_4b(ipf_ctx* ctx, const BlockDev& blk, const ChunkDev& ch, const NeighList& nl,
                          ipf_matrix* Psi, bool* handled) {
    *handled = false;
    if (blk.N != 4) return IPF_OK;
    auto it = ctx->tabs4b.find(&blk);
    if (it == ctx->tabs4b.end()) {
        Tables4* T = new Tables4();
        if (!build_tables(blk, *T)) { delete T; T = nullptr; }
        if (T) {
            IPF_CUDA(ctx, cudaMalloc((void**)&T->d_term, sizeof(T4Term) * std::max<size_t>(1, T->term.size())));
            IPF_CUDA(ctx, cudaMalloc((void**)&T->d_ent, sizeof(T4Ent) * std::max<size_t>(1, T->ent.size())));
            IPF_CUDA(ctx, cudaMalloc((void**)&T->d_ent0, sizeof(uint32_t) * T->ent0.size()));
            IPF_CUDA(ctx, cudaMalloc((void**)&T->d_s2c, sizeof(int16_t) * std::max<size_t>(1, T->s2c.size())));
            IPF_CUDA(ctx, cudaMemcpy(T->d_term, T->term.data(), sizeof(T4Term) * T->term.size(), cudaMemcpyHostToDevice));
            IPF_CUDA(ctx, cudaMemcpy(T->d_ent, T->ent.data(), sizeof(T4Ent) * T->ent.size(), cudaMemcpyHostToDevice));
            IPF_CUDA(ctx, cudaMemcpy(T->d_ent0, T->ent0.data(), sizeof(uint32_t) * T->ent0.size(), cudaMemcpyHostToDevice));
            IPF_CUDA(ctx, cudaMemcpy(T->d_s2c, T->s2c.data(), sizeof(int16_t) * T->s2c.size(), cudaMemcpyHostToDevice));
        }
        it = ctx->tabs4b.emplace((const void*)&blk, (void*)T).first;
    }
    const Tables4* T = (const Tables4*)it->second;
    if (!T) return IPF_OK;
    if (ch.ncfg == 0 || nl.npairs == 0) { *handled = true; return IPF_OK; }

    cudaStream_t st = ctx->stream;
    const double* rec = nullptr;
    int rs = 0;
    IPF_CHECK(ipf_pair_records(ctx, blk, nl, &rec, &rs));
    void* rec4 = nullptr;
    IPF_CHECK(ipf_scratch(ctx, 8, sizeof(double) * G4_RS * (size_t)(nl.npairs + 1), &rec4));
    {
        ProfScope prof_(ctx, "pair_record4");
        rec4_kernel<<<(unsigned)((nl.npairs + 127) / 128), 128, 0, st>>>(nl.npairs, rec, rs, (double*)rec4);
        ctx->launches++;
        IPF_CUDA(ctx, cudaGetLastError());
    }
    const int nbs = T->nbs;
    static const size_t s_budget = getenv("IPF_4B_SCRATCH_GB") ? (size_t)atol(getenv("IPF_4B_SCRATCH_GB")) << 30 : (size_t)4 << 30;
    const int64_t max_pairs_batch = std::max<int64_t>(nl.maxpairs_cfg, (int64_t)(s_budget / (24 * (size_t)nbs)));
    const size_t s_bytes = (size_t)24 * nbs * (size_t)std::min<int64_t>(max_pairs_batch, nl.npairs);
    int64_t max_sites_batch = 0;
    {
        int64_t c0 = 0;
        while (c0 < ch.ncfg) {
            int64_t c1 = c0 + 1;
            while (c1 < ch.ncfg && nl.h_cfg_pair_off[c1 + 1] - nl.h_cfg_pair_off[c0] <= max_pairs_batch) ++c1;
            max_sites_batch = std::max(max_sites_batch, ch.h_atom_off[c1] - ch.h_atom_off[c0]);
            c0 = c1;
        }
    }
    const int nslot = (nl.max_nbrs + 31) / 32 + 1;
    const size_t o_bytes = (size_t)32 * nslot * nbs * (size_t)max_sites_batch;
    void* Sbuf[2] = {nullptr, nullptr};
    void* Obuf[2] = {nullptr, nullptr};
    IPF_CHECK(ipf_scratch(ctx, 3, s_bytes, &Sbuf[0]));
    IPF_CHECK(ipf_scratch(ctx, 4, s_bytes, &Sbuf[1]));
    IPF_CHECK(ipf_scratch(ctx, 5, o_bytes, &Obuf[0]));
    IPF_CHECK(ipf_scratch(ctx, 0, o_bytes, &Obuf[1]));
    if (ctx->profiling && !nl.h_site_off.empty()) {
        double legs = 0.0, lk = 0.0, cl = 0.0;
        for (int64_t s = 0; s < nl.nsites; ++s) {
            const double n = (double)(nl.h_site_off[s + 1] - nl.h_site_off[s]);
            legs += n; lk += n * (n - 1.0); cl += n * (n - 1.0) * (n - 2.0) / 6.0;
        }
        ctx->prof["count:clusters_4b"].ms += cl;
        ctx->prof["count:own_legs_4b"].ms += legs;
        ctx->prof["count:leg_partners_4b"].ms += lk;
        // factorised algorithm: 2 flops per moment and (own leg, partner) + 30 for the frame components and powers,
        // plus the combination (table cost) and 19 flops per (own leg, function) for the scatter sums
        ctx->prof["count:flops_4b"].ms += lk * (2.0 * G4_NROWS + 30.0) + legs * (T->flops_leg + 19.0 * blk.nb);
    }
    IPF_CUDA(ctx, cudaFuncSetAttribute(site4b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Q_SMEM));
    constexpr size_t g_bytes = sizeof(double) * (32 * G_LDT + G_WARPS * 7 * 32);
    IPF_CUDA(ctx, cudaFuncSetAttribute(gather3b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_bytes));
    static const bool no_overlap = getenv("IPF_NO_OVERLAP") != nullptr;
    cudaStream_t st2 = no_overlap ? ctx->stream : ctx->stream2;
    cudaEvent_t gathered[2] = {nullptr, nullptr};
    int64_t c0 = 0;
    int ib = 0;
    while (c0 < ch.ncfg) {
        int64_t c1 = c0 + 1;
        while (c1 < ch.ncfg && nl.h_cfg_pair_off[c1 + 1] - nl.h_cfg_pair_off[c0] <= max_pairs_batch) ++c1;
        const int buf = ib & 1;
        Site4Args A;
        A.p0 = nl.h_cfg_pair_off[c0]; A.np = nl.h_cfg_pair_off[c1] - A.p0;
        A.rec = rec; A.rs = rs; A.rec4 = (const double*)rec4;
        A.psite = nl.psite.p; A.pbeg = nl.pbeg.p; A.pcnt = nl.pcnt.p;
        A.term = T->d_term; A.ent = T->d_ent; A.ent0 = T->d_ent0;
        for (int w = 0; w <= G4_NW; ++w) A.wsc[w] = T->wsc[w];
        A.nb = nbs; A.S = (double*)Sbuf[buf]; A.O = (double*)Obuf[buf];
        A.s0 = ch.h_atom_off[c0]; A.nslot = nslot;
        if (gathered[buf]) {
            IPF_CUDA(ctx, cudaStreamWaitEvent(st, gathered[buf], 0));
            cudaEventDestroy(gathered[buf]);
            gathered[buf] = nullptr;
        }
        if (A.np > 0) {
            ProfScope prof_(ctx, "site_deriv_4b");
            site4b_kernel<<<(unsigned)((A.np + 31) / 32), Q_THREADS, Q_SMEM, st>>>(A);
            ctx->launches++;
            const cudaError_t le = cudaGetLastError();
            if (le != cudaSuccess) {
                cudaFuncAttributes fa;
                memset(&fa, 0, sizeof(fa));
                cudaFuncGetAttributes(&fa, site4b_kernel);
                return ipf_set_error(ctx, IPF_ECUDA, "%s:%d: site4b_kernel<<<%lld, %d, %zu>>>: %s (regs %d, static smem %zu, "
                                     "max dynamic smem %d, max threads %d, local %zu)", __FILE__, __LINE__,
                                     (long long)((A.np + 31) / 32), Q_THREADS, Q_SMEM, cudaGetErrorString(le), fa.numRegs,
                                     fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes, fa.maxThreadsPerBlock, fa.localSizeBytes);
            }
        }
        cudaEvent_t produced;
        IPF_CUDA(ctx, cudaEventCreateWithFlags(&produced, cudaEventDisableTiming));
        IPF_CUDA(ctx, cudaEventRecord(produced, st));
        IPF_CUDA(ctx, cudaStreamWaitEvent(st2, produced, 0));
        cudaEventDestroy(produced);
        Gather3Args G;
        G.p0 = A.p0; G.c0 = c0; G.s0 = A.s0; G.nb = nbs; G.S = (const double*)Sbuf[buf];
        G.O = (const double*)Obuf[buf]; G.nslot = nslot; G.scol2col = T->d_s2c;
        G.atom_off = ch.atom_off.p; G.site_off = nl.site_off.p; G.prev = nl.prev.p;
        G.pR = nl.pR.p; G.rowE = ch.rowE.p; G.rowF = ch.rowF.p; G.rowV = ch.rowV.p;
        G.Psi = Psi->d; G.ld = Psi->ld; G.col0 = blk.col0;
        {
            ProfScope prof_(ctx, "gather_4b", st2);
            dim3 grid((unsigned)(c1 - c0), (unsigned)((nbs + 31) / 32));
            gather3b_kernel<<<grid, G_THREADS, g_bytes, st2>>>(G);
            ctx->launches++;
            IPF_CUDA(ctx, cudaGetLastError());
        }
        IPF_CUDA(ctx, cudaEventCreateWithFlags(&gathered[buf], cudaEventDisableTiming));
        IPF_CUDA(ctx, cudaEventRecord(gathered[buf], st2));
        c0 = c1;
        ++ib;
    }
    for (int b = 0; b < 2; ++b)
        if (gathered[b]) {
            IPF_CUDA(ctx, cudaStreamWaitEvent(st, gathered[b], 0));
            cudaEventDestroy(gathered[b]);
        }
    IPF_CUDA(ctx, cudaGetLastError());
    *handled = true;
    return IPF_OK;
}
