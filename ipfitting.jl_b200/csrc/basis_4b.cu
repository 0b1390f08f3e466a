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
//   phase 2: worker w evaluates its share of the basis functions, one ENTRY (class {m, m'} of one function) at a time.
//            The moment rows are laid out so that the terms of one entry with equal nu0 are runs of consecutive rows
//            (block (e, m0), then s = m1 + m2, then m2): an entry is one 64-byte record (block rows of both sides,
//            prefetched one entry ahead) and a body specialised on (c23, raised sides) with every row offset and every
//            multinomial weight a compile-time constant -- no per-term table loads (the first version spent 60 % of its
//            stall samples waiting for them).  Reads are conflict-free ([moment][row], lane = row).  (g_p, e_p) of 4
//            functions at a time are staged and flushed: g_p to the HBM scratch S[pair][scol][3], the per-centre-atom
//            own sums to O[site][slot][scol][4];
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
constexpr int Q_GLD = 3 * Q_NG + 2;         // doubles per staged row of g: (gx, gy, gz) x 4, pad 2 (two-way conflicts at most)
constexpr int Q_ELD = 34;                   // stride of e[function][.]: the four e columns of a column-sum load hit distinct
                                            // banks (stride 32 made them collide: 8 wavefronts per load instead of 2, 17 %
                                            // of the kernel's shared-memory traffic, which is what bounds phase 2)
constexpr int Q_STAGE = 32 * Q_GLD + Q_NG * Q_ELD;  // doubles per worker: g rows, then e[function][row]
constexpr int Q_THREADS = 32 * G4_NW;
constexpr size_t Q_SMEM = sizeof(double) * ((size_t)G4_NROWS * 32 + (size_t)G4_NW * Q_STAGE);
static_assert(Q_NG == 4, "the flush maps 5 rows x 6 double2 onto a warp and 16 columns onto a half warp");

// one class {m, m'} of monomials of one basis function (m' = m with the two other legs exchanged), 16 words:
//   w[0..4]   rows of the blocks M[a2][c12 - 1 + t][.][.], t = 0..9 (16 bits each; t = 0 is unused when c12 = 0)
//   w[5..9]   the same for the B side, M[a3][c13 - 1 + t]
//   w[10]     row of Z2[a2 + a3][c12 + c13] | row of Z2q_1[a2 + a3][c12 + c13 - 1] << 16
//   w[11]     row of Z2q_2
//   w[12..14] c12, c13 (0 unless raised) and the weight of S (1/2 for a self-conjugate class): the HIGH words of
//             the doubles (their low words are zero)
//   w[15]     class | last << 5 | a1 << 8 | max(a1 - 1, 0) << 12   (class = 3 c23 + number of raised sides; last entry of
//             its function)
// (every word is read: a dead word of the prefetched record would be reused by the compiler as a temporary, and that
//  write-after-write dependency on the load in flight stalled every entry for a full memory latency)
struct __align__(16) T4Ent {
    uint32_t w[16];
};
static_assert(sizeof(T4Ent) == 64, "entry records are read as four 16-byte words");
constexpr uint32_t ENT_LAST = 1u << 5;
inline uint32_t dbl_hi(double v) { uint64_t u; memcpy(&u, &v, 8); return (uint32_t)(u >> 32); }     // v has a zero low word

struct Site4Args {
    int64_t p0, np;                 // batch pair range
    const double* rec;              // standard pair records (own leg: Rhat, 1/r, x, x', fc, fc')
    int rs;
    const double* rec4;             // compact neighbour records [Rhat(3), fc, fc x^0..x^D, pad]; record npairs_all is all zero
    int64_t npairs_all;             // pairs of the chunk
    const int32_t* psite;
    const int64_t* pbeg;
    const int32_t* pcnt;
    const T4Ent* ent;
    const int4* blk;                // row blocks of the batch: {first pair, first pair of its configuration, rows, -}
    int went[G4_NW + 1];            // entry range of every worker
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

// write-out of the Q_NG staged functions of the warp's 32 rows.  S is group-major, S[group of 8 columns][pair][8][3]:
// a flush writes one 96-byte half of 32 consecutive 192-byte records (5 rows x 6 double2 per warp instruction; the
// other half follows with the worker's next flush and merges in L2), and the gather reads whole 192-byte records
// (three full 64-byte DRAM bursts; 96-byte records cost it twice the algorithmic DRAM traffic).  The per-centre-atom
// column sums go to O[site][slot][scol][4] (the protocol gather3b_kernel reads): lane = (column 0..15, half), the two
// halves of the warp sum the even and the odd rows of a piece.
__device__ __forceinline__ void q_flush(const Site4Args& A, const double* __restrict__ wstage, int64_t plw, int nrw,
                                        int sc0, unsigned pmask, int pout) {
    __syncwarp();
    const int lane = threadIdx.x & 31;
    {
        double2* Sg = reinterpret_cast<double2*>(A.S + ((int64_t)(sc0 >> 3) * A.np + plw) * (6 * Q_NG) + (sc0 & 4) * 3);
        const double2* st2 = reinterpret_cast<const double2*>(wstage);
        const int rl = lane / 6, cl = lane - 6 * rl;        // lanes 30, 31 idle
        if (lane < 30) {
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const int row = 5 * i + rl;
                if (row < nrw) Sg[row * 12 + cl] = st2[row * (Q_GLD / 2) + cl];
            }
        }
    }
    const int col = lane & 15, half = lane >> 4;
    // column of the staged group: 0..11 = g of function col / 3, 12..15 = e of function col - 12
    const double* cp = col < 3 * Q_NG ? wstage + col : wstage + 32 * Q_GLD + (col - 3 * Q_NG) * Q_ELD;
    const int cstride = col < 3 * Q_NG ? Q_GLD : 1;
    const int f = col < 3 * Q_NG ? col / 3 : col - 3 * Q_NG;
    const int d = col < 3 * Q_NG ? col - 3 * f : 3;
    unsigned m = pmask;
    while (m) {
        const int r0 = __ffs(m) - 1;
        m &= m - 1;
        const int r1 = m ? (__ffs(m) - 1) : nrw;
        double a0 = 0.0, a1 = 0.0;
        int r = r0 + half;
        for (; r + 2 < r1; r += 4) { a0 += cp[r * cstride]; a1 += cp[(r + 2) * cstride]; }
        if (r < r1) a0 += cp[r * cstride];
        double acc = a0 + a1;
        acc += __shfl_down_sync(0xffffffffu, acc, 16);
        const int orow = __shfl_sync(0xffffffffu, pout, r0);
        if (half == 0) A.O[((int64_t)orow * A.nb + sc0 + f) * 4 + d] = acc;
    }
    __syncwarp();
}

__host__ __device__ constexpr double q_fact(int k) { return k <= 1 ? 1.0 : (double)k * q_fact(k - 1); }
__host__ __device__ constexpr int q_tri(int s) { return s * (s + 1) / 2; }

// byte offset (row << 8) of the 16-bit row number in the low / high half of a record word: one byte permute
template <int HALF>
__device__ __forceinline__ uint32_t q_off(uint32_t v) { return __byte_perm(v, 0u, HALF ? 0x4324u : 0x4104u); }

// row t of side SIDE (0 = A, 1 = B) of an entry record, as a byte offset
template <int SIDE, int T>
__device__ __forceinline__ uint32_t q_row(const uint32_t (&w)[16]) { return q_off<(T & 1)>(w[5 * SIDE + T / 2]); }

// accumulators of one entry; the sums alternate between two sets so that consecutive terms do not wait for each other
struct QAcc {
    double s[2], a1[2], a2[2], b1[2], b2[2];
};

// term nu = (N0, SD - N2, N2) of a run; the multinomial weight is a compile-time constant
template <int C, int N0, int N2, bool HA, bool HB>
__device__ __forceinline__ void q_term(const double* __restrict__ pA, const double* __restrict__ pB, const double* __restrict__ pRA,
                                       const double* __restrict__ pRB, double& ra, double& rb, QAcc& Q) {
    constexpr int SD = C - N0;
    constexpr int P = N2 & 1;
    constexpr double kappa = q_fact(C) / (q_fact(N0) * q_fact(SD - N2) * q_fact(N2));
    const double Av = pA[N2 * 32];
    const double Bk = kappa == 1.0 ? pB[N2 * 32] : kappa * pB[N2 * 32];
    Q.s[P] = fma(Av, Bk, Q.s[P]);
    if constexpr (HA) {
        const double ra1 = pRA[(N2 + 1) * 32];
        Q.a1[P] = fma(ra, Bk, Q.a1[P]);
        Q.a2[P] = fma(ra1, Bk, Q.a2[P]);
        ra = ra1;
    }
    if constexpr (HB) {
        const double Ak = kappa == 1.0 ? Av : kappa * Av;
        const double rb1 = pRB[(N2 + 1) * 32];
        Q.b1[P] = fma(rb, Ak, Q.b1[P]);
        Q.b2[P] = fma(rb1, Ak, Q.b2[P]);
        rb = rb1;
    }
    if constexpr (N2 < SD) q_term<C, N0, N2 + 1, HA, HB>(pA, pB, pRA, pRB, ra, rb, Q);
}

template <int C, int N0, bool HA, bool HB>
__device__ __forceinline__ void q_run(const uint32_t (&w)[16], const char* __restrict__ Mb, QAcc& Q) {
    constexpr int SD = C - N0;              // nu1 + nu2 of this run
    const double* pA = reinterpret_cast<const double*>(Mb + q_row<0, N0 + 1>(w)) + q_tri(SD) * 32;
    const double* pB = reinterpret_cast<const double*>(Mb + q_row<1, N0 + 1>(w)) + q_tri(SD) * 32;
    const double* pRA = nullptr;
    const double* pRB = nullptr;
    if constexpr (HA) pRA = reinterpret_cast<const double*>(Mb + q_row<0, N0>(w)) + q_tri(SD + 1) * 32;
    if constexpr (HB) pRB = reinterpret_cast<const double*>(Mb + q_row<1, N0>(w)) + q_tri(SD + 1) * 32;
    double ra = 0.0, rb = 0.0;
    if constexpr (HA) ra = pRA[0];
    if constexpr (HB) rb = pRB[0];
    q_term<C, N0, 0, HA, HB>(pA, pB, pRA, pRB, ra, rb, Q);
    if constexpr (N0 < C) q_run<C, N0 + 1, HA, HB>(w, Mb, Q);
}

// one entry of class CLS = 3 C + V (C = c23, V = number of raised sides; a raised side needs c12 >= 1, i.e. C <= D - 1,
// two raised sides C <= D - 2): the sum over nu (|nu| = C) and the entry's contribution to (S0, S1, T1, T2) of its function.
// Everything that depends on the class is resolved here: only the accumulators the class uses exist, the k = l rows are
// the starting values of the sums, and entries without raised products touch neither Z2q nor T.
template <int CLS>
__device__ __forceinline__ void q_entry(const uint32_t (&w)[16], const char* __restrict__ Mb, double x1, double x2, double x4,
                                        double& S0, double& S1, double& T1, double& T2) {
    constexpr int C = CLS / 3, V = CLS % 3;
    if constexpr (C + V <= G4_D) {
        const double zs = *reinterpret_cast<const double*>(Mb + q_off<0>(w[10]));
        double z1 = 0.0, z2 = 0.0;
        if constexpr (V >= 1) {
            z1 = *reinterpret_cast<const double*>(Mb + q_off<1>(w[10]));
            z2 = *reinterpret_cast<const double*>(Mb + q_off<0>(w[11]));
        }
        // x_j^a1 and a1 x_j^(a1-1) from three registers (x, x^2, x^4; a1 <= 8): the combination is bound by shared-memory
        // traffic, not by arithmetic -- the two table loads per entry this replaces were 10 % of it
        const uint32_t meta = w[15];
        double xm = (meta & 0x1000u) ? x1 : 1.0;
        if (meta & 0x2000u) xm *= x2;
        if (meta & 0x4000u) xm *= x4;
        const double xd = (double)(int)((meta >> 8) & 0xfu) * xm;
        const double xa = (meta & 0xf00u) ? xm * x1 : 1.0;
        QAcc Q;
        Q.s[0] = -zs; Q.s[1] = 0.0;
        Q.a1[0] = -z1; Q.a1[1] = 0.0; Q.a2[0] = -z2; Q.a2[1] = 0.0;
        Q.b1[0] = -z1; Q.b1[1] = 0.0; Q.b2[0] = -z2; Q.b2[1] = 0.0;
        q_run<C, 0, (V >= 1), (V >= 2)>(w, Mb, Q);
        const double s = C >= 1 ? Q.s[0] + Q.s[1] : Q.s[0];
        const double ws = __hiloint2double((int)w[14], 0) * s;
        S0 = fma(xa, ws, S0);
        S1 = fma(xd, ws, S1);
        if constexpr (V >= 1) {
            // c12 (tA - z) + c13 (tB - z)
            const double cA = __hiloint2double((int)w[12], 0);
            double u1 = cA * (C >= 1 ? Q.a1[0] + Q.a1[1] : Q.a1[0]);
            double u2 = cA * (C >= 1 ? Q.a2[0] + Q.a2[1] : Q.a2[0]);
            if constexpr (V >= 2) {
                const double cB = __hiloint2double((int)w[13], 0);
                u1 = fma(cB, C >= 1 ? Q.b1[0] + Q.b1[1] : Q.b1[0], u1);
                u2 = fma(cB, C >= 1 ? Q.b2[0] + Q.b2[1] : Q.b2[0], u2);
            }
            T1 = fma(xa, u1, T1);
            T2 = fma(xa, u2, T2);
        }
    }
}

__device__ __forceinline__ void q_dispatch(uint32_t cls, const uint32_t (&w)[16], const char* __restrict__ Mb, double x1,
                                           double x2, double x4, double& S0, double& S1, double& T1, double& T2) {
    static_assert(G4_D == 8, "one case per class 0 .. 3 (D + 1) - 1");
    // (a loop-driven body for the rare classes with c23 >= 5 -- 3 % of the entries, 70 % of the code of phase 2, which
    //  exceeds the 32 KB instruction cache -- was tried: the kernel got 7 % slower, 18 % with c23 >= 4)
#define Q_CASE(K) case K: q_entry<K>(w, Mb, x1, x2, x4, S0, S1, T1, T2); break;
    switch (cls) {
        Q_CASE(0) Q_CASE(1) Q_CASE(2) Q_CASE(3) Q_CASE(4) Q_CASE(5) Q_CASE(6) Q_CASE(7) Q_CASE(8)
        Q_CASE(9) Q_CASE(10) Q_CASE(11) Q_CASE(12) Q_CASE(13) Q_CASE(14) Q_CASE(15) Q_CASE(16) Q_CASE(17)
        Q_CASE(18) Q_CASE(19) Q_CASE(20) Q_CASE(21) Q_CASE(22) Q_CASE(23) Q_CASE(24) Q_CASE(25) Q_CASE(26)
        default: break;
    }
#undef Q_CASE
}

__device__ __forceinline__ void q_load_entry(const T4Ent* __restrict__ e, uint4 (&v)[4]) {
    const uint4* p = reinterpret_cast<const uint4*>(e);
    v[0] = __ldg(p); v[1] = __ldg(p + 1); v[2] = __ldg(p + 2); v[3] = __ldg(p + 3);
}

#ifdef IPF_4B_TIMING
// development aid (not in the product build): cycles per worker warp, summed over the CTAs
__device__ unsigned long long g4_tcyc[3][G4_NW];
#define G4_TICK(slot) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g4_tcyc[slot][threadIdx.x >> 5], (unsigned long long)(clock64() - t_begin)); } while (0)
#else
#define G4_TICK(slot) do { } while (0)
#endif

__global__ void __maxnreg__(168) site4b_kernel(Site4Args A) {
#ifdef IPF_4B_TIMING
    const long long t_begin = clock64();
#endif
    extern __shared__ __align__(16) double sm[];
    double* M = sm;                                          // [G4_NROWS][32]
    double* stage = sm + (size_t)G4_NROWS * 32;              // [G4_NW][Q_STAGE]
    const int lane = threadIdx.x & 31, worker = threadIdx.x >> 5;
    // row blocks start at the first pair of every configuration (the last block of a configuration is short): the
    // partial sums of a block never depend on how the configurations were batched or sharded over GPUs
    const int4 bk = __ldg(A.blk + blockIdx.x);
    const int64_t pl0 = (int64_t)bk.x - A.p0;
    const int64_t pl = pl0 + lane;
    const int nrw = bk.z;
    const bool live = lane < nrw;
    const int64_t p = A.p0 + (live ? pl : pl0);
    const int64_t k0 = A.pbeg[p];
    const int n = live ? A.pcnt[p] : 0;
    const int site = A.psite[p];
    const int site_prev = __shfl_up_sync(0xffffffffu, site, 1);
    const bool piece_start = live && (lane == 0 || site != site_prev);
    const unsigned pmask = __ballot_sync(0xffffffffu, piece_start);
    const int pout = (int)(site - A.s0) * A.nslot + (int)(((int64_t)(bk.x - bk.y) >> 5) - ((k0 - bk.y) >> 5));
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
    // ---- phase 1: moments of the 32 rows.  The neighbour records of the block's centre atoms (one contiguous run of
    // pairs) are staged in the output staging area, which is idle until phase 2: every worker walks the same records.
    {
        const int64_t kb0 = __shfl_sync(0xffffffffu, k0, 0);
        const int64_t kb1 = __shfl_sync(0xffffffffu, k0 + n, nrw - 1);
        const int nd2 = (int)(kb1 - kb0) * (G4_RS / 2);             // double2 words of the run
        if (nd2 + G4_RS / 2 <= G4_NW * Q_STAGE / 2) {
            const double2* src = reinterpret_cast<const double2*>(A.rec4 + kb0 * G4_RS);
            double2* dst = reinterpret_cast<double2*>(stage);
            for (int q = threadIdx.x; q < nd2; q += Q_THREADS) dst[q] = __ldg(src + q);
            if (threadIdx.x < G4_RS / 2) dst[nd2 + threadIdx.x] = make_double2(0.0, 0.0);      // the all-zero record
            __syncthreads();
            g4_phase1<true>(worker, stage + (k0 - kb0) * G4_RS, stage + 2 * (size_t)nd2, n, (int)(p - k0), Rj0, Rj1, Rj2, E1, E2,
                            M + lane);
        } else {
            g4_phase1<false>(worker, A.rec4 + k0 * G4_RS, A.rec4 + A.npairs_all * G4_RS, n, (int)(p - k0), Rj0, Rj1, Rj2, E1, E2,
                             M + lane);
        }
    }
    G4_TICK(0);
    __syncthreads();
    G4_TICK(1);

    // ---- phase 2: basis functions of this worker, entry by entry; 4 scratch columns per staged group
    const char* Mb = reinterpret_cast<const char*>(M + lane);
    const double xj2 = xj * xj, xj4 = xj2 * xj2;
    double* wstage = stage + (size_t)worker * Q_STAGE;
    double* myrow = wstage + lane * Q_GLD;
    double* mye = wstage + 32 * Q_GLD + lane;
    const double tang = fcj * rinvj, fdx = fcj * dxj, third = fcj * (1.0 / 3.0);
    int ei = A.went[worker];
    const int ei_end = A.went[worker + 1];
    int sc = A.wsc[worker], f = 0;
    uint4 nx[4];
    if (ei < ei_end) q_load_entry(A.ent + ei, nx);
    double S0 = 0.0, S1 = 0.0, T1 = 0.0, T2 = 0.0;
#pragma unroll 1
    for (; ei < ei_end; ++ei) {
        uint32_t w[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) { w[4 * q] = nx[q].x; w[4 * q + 1] = nx[q].y; w[4 * q + 2] = nx[q].z; w[4 * q + 3] = nx[q].w; }
        if (ei + 1 < ei_end) q_load_entry(A.ent + ei + 1, nx);        // the next record is in flight during this entry
        // the whole entry is specialised on its class (term sums and epilogue); padding entries have class 0, rows 0 and
        // weight 0
        q_dispatch(w[15] & 0x1fu, w, Mb, xj, xj2, xj4, S0, S1, T1, T2);
        if (w[15] & ENT_LAST) {
            const double radial = fma(dfcj, S0, fdx * S1);
            T1 *= tang; T2 *= tang;
            myrow[3 * f] = fma(radial, Rj0, fma(T1, E1[0], T2 * E2[0]));
            myrow[3 * f + 1] = fma(radial, Rj1, fma(T1, E1[1], T2 * E2[1]));
            myrow[3 * f + 2] = fma(radial, Rj2, fma(T1, E1[2], T2 * E2[2]));
            mye[f * Q_ELD] = third * S0;
            S0 = 0.0; S1 = 0.0; T1 = 0.0; T2 = 0.0;
            if (++f == Q_NG) {
                q_flush(A, wstage, pl0, nrw, sc, pmask, pout);
                sc += Q_NG;
                f = 0;
            }
        }
    }
    G4_TICK(2);
}

// ---- host tables --------------------------------------------------------------------------------------------
struct Tables4 {
    int D = 0, nbs = 0;
    std::vector<T4Ent> ent;
    std::vector<int16_t> s2c;
    int went[G4_NW + 1] = {};
    int wsc[G4_NW + 1] = {};
    double flops_leg = 0.0;         // combine flops per own leg (FMA = 2)
    T4Ent* d_ent = nullptr;
    int16_t* d_s2c = nullptr;
};

inline int block_row(int e, int m0) {
    constexpr int n1 = G4_D + 1;
    if (e < 0 || m0 < 0 || e >= n1 || m0 >= n1) return -1;
    return g4_base[e * n1 + m0];
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
        if (mons.empty()) return false;
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
            // per term: kappa B, S (3 flops); raised A: 2 FMA; raised B: kappa A + 2 FMA; per entry: 14
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
    T.ent.clear(); T.s2c.clear();
    T.flops_leg = 0.0;
    auto put16 = [](T4Ent& E, int word, int half, int v) { E.w[word] |= (uint32_t)(v & 0xffff) << (16 * half); };
    int sc = 0;
    for (int w = 0; w < G4_NW; ++w) {
        T.wsc[w] = sc;
        T.went[w] = (int)T.ent.size();
        std::sort(wl[w].begin(), wl[w].end());
        const int npad = (int)((wl[w].size() + Q_NG - 1) / Q_NG * Q_NG);
        for (int q = 0; q < npad; ++q) {
            ++sc;
            if (q >= (int)wl[w].size()) {
                // padding column: one entry of weight zero that closes the (never gathered) function
                T.s2c.push_back(-1);
                T4Ent E;
                memset(&E, 0, sizeof(E));
                E.w[15] = ENT_LAST;
                T.ent.push_back(E);
                continue;
            }
            const int b = wl[w][q];
            T.s2c.push_back((int16_t)b);
            T.flops_leg += cost[b];
            for (size_t ic = 0; ic < fcls[b].size(); ++ic) {
                const Cls& c = fcls[b][ic];
                T4Ent E;
                memset(&E, 0, sizeof(E));
                const bool ha = c.c12 > 0, hb = !c.self && c.c13 > 0;
                if (hb && !ha) return false;      // excluded by the orientation rule
                if (c.a2 + c.c12 + c.c23 > G4_D || c.a3 + c.c13 + c.c23 > G4_D) return false;
                for (int t = 0; t <= c.c23 + 1; ++t) {
                    // block rows of M[a2][c12 - 1 + t] and M[a3][c13 - 1 + t]; t = 0 is read by the raised products only
                    const int rA = block_row(c.a2, c.c12 - 1 + t), rB = block_row(c.a3, c.c13 - 1 + t);
                    if ((t >= 1 || ha) && rA < 0) return false;
                    if ((t >= 1 || hb) && rB < 0) return false;
                    put16(E, t / 2, t & 1, std::max(rA, 0));
                    put16(E, 5 + t / 2, t & 1, std::max(rB, 0));
                }
                const int zr = g4_z2idx[(c.a2 + c.a3) * (G4_D + 1) + c.c12 + c.c13];
                if (zr < 0) return false;
                put16(E, 10, 0, zr);
                if (ha) {
                    const int q1 = g4_z2qidx[((0) * (G4_D + 1) + c.a2 + c.a3) * (G4_D + 1) + c.c12 + c.c13 - 1];
                    const int q2 = g4_z2qidx[((1) * (G4_D + 1) + c.a2 + c.a3) * (G4_D + 1) + c.c12 + c.c13 - 1];
                    if (q1 < 0 || q2 < 0) return false;
                    put16(E, 10, 1, q1);
                    put16(E, 11, 0, q2);
                }
                const uint32_t cls = 3u * (uint32_t)c.c23 + (ha ? 1u : 0u) + (hb ? 1u : 0u);
                E.w[15] = cls | (ic + 1 == fcls[b].size() ? ENT_LAST : 0u) | (uint32_t)c.a1 << 8 | (uint32_t)std::max(c.a1 - 1, 0) << 12;
                E.w[12] = dbl_hi((double)c.c12);
                E.w[13] = dbl_hi(hb ? (double)c.c13 : 0.0);
                E.w[14] = dbl_hi(c.self ? 0.5 : 1.0);
                T.ent.push_back(E);
            }
        }
    }
    T.wsc[G4_NW] = sc;
    T.went[G4_NW] = (int)T.ent.size();
    while (sc % 8) { T.s2c.push_back(-1); ++sc; }      // whole groups of 8 scratch columns (never written, never gathered)
    T.nbs = sc;
    return true;
}

void free_tables(void* q) {
    Tables4* T = (Tables4*)q;
    if (!T) return;
    if (T->d_ent) cudaFree(T->d_ent);
    if (T->d_s2c) cudaFree(T->d_s2c);
    delete T;
}

}  // namespace

#ifdef IPF_4B_TIMING
extern "C" int32_t ipf_debug_4b_timing(double* out /* [3][G4_NW] */, int32_t reset) {
    unsigned long long h[3][G4_NW];
    if (cudaMemcpyFromSymbol(h, g4_tcyc, sizeof(h)) != cudaSuccess) return -1;
    for (int a = 0; a < 3; ++a) for (int w = 0; w < G4_NW; ++w) out[a * G4_NW + w] = (double)h[a][w];
    if (reset) { memset(h, 0, sizeof(h)); cudaMemcpyToSymbol(g4_tcyc, h, sizeof(h)); }
    return G4_NW;
}
#endif

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
            IPF_CUDA(ctx, cudaMalloc((void**)&T->d_ent, sizeof(T4Ent) * std::max<size_t>(1, T->ent.size())));
            IPF_CUDA(ctx, cudaMalloc((void**)&T->d_s2c, sizeof(int16_t) * std::max<size_t>(1, T->s2c.size())));
            IPF_CUDA(ctx, cudaMemcpy(T->d_ent, T->ent.data(), sizeof(T4Ent) * T->ent.size(), cudaMemcpyHostToDevice));
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
        IPF_CUDA(ctx, cudaMemsetAsync((double*)rec4 + (size_t)nl.npairs * G4_RS, 0, sizeof(double) * G4_RS, st));
        ctx->launches++;
        IPF_CUDA(ctx, cudaGetLastError());
    }
    const int nbs = T->nbs;
    static const size_t s_budget = getenv("IPF_4B_SCRATCH_GB") ? (size_t)atol(getenv("IPF_4B_SCRATCH_GB")) << 30 : (size_t)8 << 30;
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
    // row blocks of the chunk, aligned to the configurations
    std::vector<int4> h_blk;
    std::vector<int64_t> cfg_blk(ch.ncfg + 1, 0);
    for (int64_t c = 0; c < ch.ncfg; ++c) {
        const int64_t q0 = nl.h_cfg_pair_off[c], q1 = nl.h_cfg_pair_off[c + 1];
        cfg_blk[c] = (int64_t)h_blk.size();
        for (int64_t q = q0; q < q1; q += 32) h_blk.push_back(make_int4((int)q, (int)q0, (int)std::min<int64_t>(32, q1 - q), 0));
    }
    cfg_blk[ch.ncfg] = (int64_t)h_blk.size();
    if (nl.npairs > (int64_t)0x7fffffff) return ipf_set_error(ctx, IPF_EINVAL, "4-body path: more than 2^31 pairs in one chunk");
    void* d_blk = nullptr;
    IPF_CHECK(ipf_scratch(ctx, 10, sizeof(int4) * std::max<size_t>(1, h_blk.size()), &d_blk));
    IPF_CUDA(ctx, cudaMemcpyAsync(d_blk, h_blk.data(), sizeof(int4) * h_blk.size(), cudaMemcpyHostToDevice, st));
    IPF_CUDA(ctx, cudaStreamSynchronize(st));          // h_blk must outlive the copy
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
    // The gather follows its producer on the same stream.  Running it on the side stream next to the producer of the next
    // batch (IPF_4B_OVERLAP=1) is slower (cfg3 step 834 ms against 824 ms; with the producer's stream at a higher priority
    // 846 ms): a producer CTA needs a completely free SM, and the many small gather CTAs -- HBM bound, 75 % of the copy
    // bandwidth on their own -- spread over all SMs and keep the producer's CTAs waiting.
    static const bool overlap = getenv("IPF_4B_OVERLAP") != nullptr;
    cudaStream_t st2 = overlap ? ctx->stream2 : ctx->stream;
    cudaEvent_t gathered[2] = {nullptr, nullptr};
    int64_t c0 = 0;
    int ib = 0;
    while (c0 < ch.ncfg) {
        int64_t c1 = c0 + 1;
        while (c1 < ch.ncfg && nl.h_cfg_pair_off[c1 + 1] - nl.h_cfg_pair_off[c0] <= max_pairs_batch) ++c1;
        const int buf = ib & 1;
        Site4Args A;
        A.p0 = nl.h_cfg_pair_off[c0]; A.np = nl.h_cfg_pair_off[c1] - A.p0;
        A.rec = rec; A.rs = rs; A.rec4 = (const double*)rec4; A.npairs_all = nl.npairs;
        A.psite = nl.psite.p; A.pbeg = nl.pbeg.p; A.pcnt = nl.pcnt.p;
        A.ent = T->d_ent;
        for (int w = 0; w <= G4_NW; ++w) { A.wsc[w] = T->wsc[w]; A.went[w] = T->went[w]; }
        A.nb = nbs; A.S = (double*)Sbuf[buf]; A.O = (double*)Obuf[buf];
        A.s0 = ch.h_atom_off[c0]; A.nslot = nslot;
        if (gathered[buf]) {
            IPF_CUDA(ctx, cudaStreamWaitEvent(st, gathered[buf], 0));
            cudaEventDestroy(gathered[buf]);
            gathered[buf] = nullptr;
        }
        A.blk = (const int4*)d_blk + cfg_blk[c0];
        const unsigned nblocks = (unsigned)(cfg_blk[c1] - cfg_blk[c0]);
        if (nblocks > 0) {
            ProfScope prof_(ctx, "site_deriv_4b");
            site4b_kernel<<<nblocks, Q_THREADS, Q_SMEM, st>>>(A);
            ctx->launches++;
            const cudaError_t le = cudaGetLastError();
            if (le != cudaSuccess) {
                cudaFuncAttributes fa;
                memset(&fa, 0, sizeof(fa));
                cudaFuncGetAttributes(&fa, site4b_kernel);
                return ipf_set_error(ctx, IPF_ECUDA, "%s:%d: site4b_kernel<<<%lld, %d, %zu>>>: %s (regs %d, static smem %zu, "
                                     "max dynamic smem %d, max threads %d, local %zu)", __FILE__, __LINE__,
                                     (long long)nblocks, Q_THREADS, Q_SMEM, cudaGetErrorString(le), fa.numRegs,
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
        G.pstride = 6 * Q_NG; G.gstride = (int64_t)6 * Q_NG * A.np; G.cfg_aligned = 1;
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
