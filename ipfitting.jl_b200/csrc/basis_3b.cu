// basis_3b.cu -- K3 fast path: canonical 3-body bond-angle basis nbpolys(3, desc, D).
//
// Same own-leg decomposition as basis_generic.cu (one thread per ordered pair
// p = (i, j)), but the sum over the second leg k is factorised.  Every monomial
//     x_j^a x_k^b w_jk^c            (orbit partner: x_j^b x_k^a w_jk^c)
// is (thread constant) x (function of k only), so per own leg only the MOMENTS
//     Z_c[e] = sum_k fc_k x_k^e w_jk^c                         (scalar,  e + c <= D)
//     Q_c[e] = sum_k fc_k x_k^e w_jk^(c-1) (q1_k, q2_k)        (c >= 1)
// are accumulated in the k loop, where (q1, q2) are the components of Rhat_k in an orthonormal
// basis (e1, e2) of the plane perpendicular to Rhat_j: the Rhat_j component of the vector moment
// sum_k .. Rhat_k equals Z_c[e] and cancels against the -w Rhat_j part of dw/dR_j, so only the
// in-plane part is needed (4 FP64 instructions per (e, c) and k instead of ~7 per basis function).
// The basis functions (a <= b, c) follow from
//     S0 = x^a Z_c[b] + x^b Z_c[a]        S1 = a x^(a-1) Z_c[b] + b x^(b-1) Z_c[a]
//     T  = c (x^a Q_c[b] + x^b Q_c[a])  (in-plane)                   (x = x_j; a = b: one term)
//     g_p = (fc_j' S0 + fc_j x_j' S1) Rhat_j + fc_j/r_j (T1 e1 + T2 e2),     e_p = fc_j S0 / 2.
// The (e, c) plane is cut into units of consecutive c (register budget); a CTA
// runs one unit for 128 consecutive pairs.  Outputs are staged through shared memory in
// groups of 8 functions: g_p goes to the HBM scratch S[pair][scol][3] (scol = unit-major
// column order; contiguous 192-byte runs per pair), and the per-centre-atom sums
// O[site][slot][scol] = sum_{p in out(a)} (g_p, e_p) are reduced in the CTA in row order
// (slot 1 = continuation of a site that started in the previous CTA).  gather3b_kernel then
// needs only the REVERSE entries:  F_b[a] = O[a] - sum_{p in out(a)} g_rev(p),
// Vir_b = + sum_p g_rev(p) (x) R_p  (rev is a bijection and R_rev(p) = -R_p), E_b = sum_a O[a].e
// -- everything in a fixed order, no atomics, 768-byte coalesced reads per (pair, 32 columns).
#include <cstdlib>

#include "basis_3b_shared.cuh"

int ipf_pair_records(ipf_ctx* ctx, const BlockDev& blk, const NeighList& nl, const double** rec, int* rs);
int ipf_assemble_block_3b_cfg(ipf_ctx* ctx, const BlockDev& blk, const ChunkDev& ch, const NeighList& nl,
                              ipf_matrix* Psi, int D, const double* rec, const int16_t* s2c, bool* handled);

namespace {

constexpr int A_THREADS = 128;

struct Fast3Args {
    int64_t p0, np;                // batch pair range [p0, p0 + np)
    const double* rec;             // pair records, stride RS<D> doubles: head (8) + x^0..x^D
    const int64_t* site_off;
    const int64_t* atom_off;
    const int32_t* psite;          // chunk-global site of every pair
    const int64_t* pbeg;           // first pair of that site's list
    const int32_t* pcnt;           // length of that list
    int nb;
    double* S;                     // [np][nb][3], scol order: g_p
    double* O;                     // [nsites_batch][nslot][nb][4]: own sums (gx, gy, gz, e) per centre atom and slot
    int64_t s0;                    // first site of the batch
    int nslot;                     // slot = (batch row / 32) - (site's first batch row / 32): warp-local piece sums
};

constexpr int A_MAXREC = 224;              // pair records staged in shared memory per CTA
constexpr int A_OUTG = 8;                  // functions per staged output group
constexpr int A_OUTLD = 4 * A_OUTG + 2;    // doubles per staged row (+2: conflict-free 16-byte columns)

template <int D>
__host__ __device__ constexpr size_t moments_smem_bytes() {
    constexpr size_t a = sizeof(double) * (size_t)A_MAXREC * RecStride<D>::value;
    constexpr size_t b = sizeof(double) * (size_t)A_THREADS * A_OUTLD;
    return a + b;
}

constexpr int A_EOFF = 3 * A_OUTG;         // staged row: [g(3) x A_OUTG][e x A_OUTG][pad 2]

// sum of column `lane` of the staged tile over rows [r0, r1): four interleaved partial sums (fixed order)
__device__ __forceinline__ double column_sum(const double* __restrict__ stage, int r0, int r1, int lane) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int r = r0;
    for (; r + 3 < r1; r += 4) {
        a0 += stage[r * A_OUTLD + lane];
        a1 += stage[(r + 1) * A_OUTLD + lane];
        a2 += stage[(r + 2) * A_OUTLD + lane];
        a3 += stage[(r + 3) * A_OUTLD + lane];
    }
    for (; r < r1; ++r) a0 += stage[r * A_OUTLD + lane];
    return (a0 + a1) + (a2 + a3);
}

// Warp-synchronous write-out of NG staged functions of the warp's own 32 rows (`wstage`):
//   S[(row)*nb + sc0 .. +NG)[3]  (contiguous 24*NG bytes per row), and
//   O[site][slot][sc0 .. +NG)[4] = column sums over each of the warp's centre-atom pieces, rows in order.
// pmask: bit r set = row r of the warp opens a piece; pout[r]: O row (site, slot) of the piece opening at r.
template <int NG>
__device__ __forceinline__ void flush_group(const Fast3Args& A, const double* __restrict__ wstage, int64_t plw,
                                            int nrw, int sc0, unsigned pmask, int pout) {
    __syncwarp();
    const int lane = threadIdx.x & 31;
    if constexpr (NG % 2 == 0) {
        constexpr int CH = 3 * NG / 2;            // 16-byte chunks per row
        double2* S2 = reinterpret_cast<double2*>(A.S);
        const double2* st2 = reinterpret_cast<const double2*>(wstage);
        for (int q = lane; q < nrw * CH; q += 32) {
            const int row = q / CH, c = q - row * CH;
            S2[((plw + row) * A.nb + sc0) * 3 / 2 + c] = st2[row * (A_OUTLD / 2) + c];
        }
    } else {
        for (int q = lane; q < nrw * 3 * NG; q += 32) {
            const int row = q / (3 * NG), c = q - row * (3 * NG);
            A.S[((plw + row) * A.nb + sc0) * 3 + c] = wstage[row * A_OUTLD + c];
        }
    }
    // column sums: lane = staged column v (0..23: g of function v/3, component v%3; 24..31: e of function v-24)
    const bool vlive = lane < 3 * NG || (lane >= A_EOFF && lane < A_EOFF + NG);
    const int f = lane < A_EOFF ? lane / 3 : lane - A_EOFF;
    const int d = lane < A_EOFF ? lane - 3 * f : 3;
    {
        unsigned m = pmask;
        while (m) {
            const int r0 = __ffs(m) - 1;
            m &= m - 1;
            const int r1 = m ? (__ffs(m) - 1) : nrw;
            const double acc = column_sum(wstage, r0, r1, lane);
            const int orow = __shfl_sync(0xffffffffu, pout, r0);
            if (vlive) A.O[((int64_t)orow * A.nb + sc0 + f) * 4 + d] = acc;
        }
    }
    __syncwarp();
}

// state shared by all units of a CTA (computed once in the kernel prologue)
struct CtaState {
    const double* recp;          // own pair record (global)
    int64_t rk_off;              // offset (doubles) of the centre atom's first neighbour record: in the staged
                                 // window (shared memory) when `staged`, else in A.rec (global)
    bool staged;
    int n, jself;
    double Rj0, Rj1, Rj2, rinvj, dxj, fcj, dfcj;
    double* wstage;              // this warp's staging rows
    int64_t plw;                 // batch row of the warp's first lane
    int nrw;                     // live rows of the warp
    unsigned pmask;
    int pout;
};

// one unit = exponents c in [C0, C1] of w
template <int D, int C0, int C1, int SCB>
__device__ __forceinline__ void unit3b(const Fast3Args& A, const CtaState& T) {
    constexpr int NC = C1 - C0 + 1;
    constexpr int EMAX = D - C0;            // largest x-exponent in this unit
    const int lane = threadIdx.x & 31;
    const double Rj0 = T.Rj0, Rj1 = T.Rj1, Rj2 = T.Rj2, rinvj = T.rinvj, dxj = T.dxj, fcj = T.fcj, dfcj = T.dfcj;
    const double* recp = T.recp;
    double Z[NC][EMAX + 1];
    double U[NC][EMAX + 1][2];
#pragma unroll
    for (int ci = 0; ci < NC; ++ci)
#pragma unroll
        for (int e = 0; e <= EMAX; ++e) { Z[ci][e] = 0.0; U[ci][e][0] = U[ci][e][1] = 0.0; }
    // orthonormal basis (E1, E2) of the plane perpendicular to Rhat_j
    double E1[3], E2[3];
    {
        const bool usex = fabs(Rj0) < 0.8;
        // E1 = normalise(Rhat_j x axis), axis = x or y
        double c0 = usex ? 0.0 : -Rj2, c1 = usex ? Rj2 : 0.0, c2 = usex ? -Rj1 : Rj0;
        const double inv = rsqrt(c0 * c0 + c1 * c1 + c2 * c2);
        E1[0] = c0 * inv; E1[1] = c1 * inv; E1[2] = c2 * inv;
        E2[0] = Rj1 * E1[2] - Rj2 * E1[1];
        E2[1] = Rj2 * E1[0] - Rj0 * E1[2];
        E2[2] = Rj0 * E1[1] - Rj1 * E1[0];
    }

    {
        extern __shared__ __align__(16) double sm_rec[];
        // two instantiations so that the loop body uses LDS (shared) resp. LDG (global), not generic loads
        if (T.staged) moment_loop<D, C0, C1, NC, EMAX>(sm_rec + T.rk_off, T.n, T.jself, Rj0, Rj1, Rj2, E1, E2, Z, U);
        else moment_loop<D, C0, C1, NC, EMAX>(A.rec + T.rk_off, T.n, T.jself, Rj0, Rj1, Rj2, E1, E2, Z, U);
    }

    // ---- combine into basis functions, staged through shared memory in groups of A_OUTG
    double X[EMAX + 1], dX[EMAX + 1];
#pragma unroll
    for (int e = 0; e <= EMAX; ++e) X[e] = __ldg(recp + REC + e);
    dX[0] = 0.0;
#pragma unroll
    for (int e = 1; e <= EMAX; ++e) dX[e] = (double)e * X[e - 1];
    const double tang = fcj * rinvj;
    const double fdx = fcj * dxj;
    double* wstage = T.wstage;
    double* myrow = wstage + lane * A_OUTLD;
    const int64_t plw = T.plw;
    const int nrw = T.nrw;
    const unsigned pmask = T.pmask;
    const int pout = T.pout;
    constexpr int SC0 = SCB;                                       // even: 16-byte aligned scratch rows
    constexpr int NF = scol_base(D, C1 + 1) - scol_base(D, C0);
    int f = 0;
#pragma unroll
    for (int ci = 0; ci < NC; ++ci) {
        const int c = C0 + ci;
        const double cc = (double)c;
        const double tc = tang * cc;
#pragma unroll
        for (int a = 0; a <= (D - c) / 2; ++a) {
#pragma unroll
            for (int b = a; b <= D - c - a; ++b) {
                if (a + b + c == 0) continue;
                double S0, S1, T1, T2;
                if (a == b) {
                    S0 = X[a] * Z[ci][a];
                    S1 = dX[a] * Z[ci][a];
                    T1 = X[a] * U[ci][a][0]; T2 = X[a] * U[ci][a][1];
                } else {
                    S0 = fma(X[a], Z[ci][b], X[b] * Z[ci][a]);
                    S1 = fma(dX[a], Z[ci][b], dX[b] * Z[ci][a]);
                    T1 = fma(X[a], U[ci][b][0], X[b] * U[ci][a][0]);
                    T2 = fma(X[a], U[ci][b][1], X[b] * U[ci][a][1]);
                }
                const double radial = fma(dfcj, S0, fdx * S1);      // coefficient of Rhat_j
                T1 *= tc; T2 *= tc;
                const int slot = f % A_OUTG;
                myrow[3 * slot] = fma(radial, Rj0, fma(T1, E1[0], T2 * E2[0]));
                myrow[3 * slot + 1] = fma(radial, Rj1, fma(T1, E1[1], T2 * E2[1]));
                myrow[3 * slot + 2] = fma(radial, Rj2, fma(T1, E1[2], T2 * E2[2]));
                myrow[A_EOFF + slot] = 0.5 * fcj * S0;
                ++f;
                if (f % A_OUTG == 0) flush_group<A_OUTG>(A, wstage, plw, nrw, SC0 + f - A_OUTG, pmask, pout);
            }
        }
    }
    constexpr int REM = NF % A_OUTG;
    if constexpr (REM > 0) flush_group<REM>(A, wstage, plw, nrw, SC0 + NF - REM, pmask, pout);
}

template <int D> struct UnitTable;

template <int D, int U>
__device__ __forceinline__ void run_unit(const Fast3Args& A, const CtaState& T);
template <int D, int U>
__device__ __forceinline__ void run_all_units(const Fast3Args& A, const CtaState& T);

// CTA = 128 consecutive ordered pairs x one unit (one kernel per unit: own register budget and occupancy;
// running all units in one kernel was measured 1.6x slower)
template <int D, int U>
__global__ void __launch_bounds__(A_THREADS) moments3b_kernel(Fast3Args A) {
    extern __shared__ __align__(16) double sm[];
    __shared__ int64_t s_kbeg;
    __shared__ int s_nrec;
    constexpr int RS = RecStride<D>::value;
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t pl0 = (int64_t)blockIdx.x * A_THREADS;
    const int64_t pl = pl0 + tid;
    const bool live = pl < A.np;
    const int nrow = (int)min((int64_t)A_THREADS, A.np - pl0);
    const int64_t p = A.p0 + (live ? pl : pl0);
    const int64_t k0 = A.pbeg[p];                      // first pair of the centre atom's list
    const int n = live ? A.pcnt[p] : 0;                // its length
    const int site = A.psite[p];
    if (tid == 0) { s_kbeg = k0; }
    if (tid == nrow - 1) { s_nrec = (int)min((int64_t)(A_MAXREC + 1), k0 + A.pcnt[p] - A.pbeg[A.p0 + pl0]); }
    CtaState T;
    // pieces of this warp's rows: a piece opens where the centre atom changes or at the warp's first row
    const int site_prev = __shfl_up_sync(0xffffffffu, site, 1);
    const bool piece_start = live && (lane == 0 || site != site_prev);
    T.pmask = __ballot_sync(0xffffffffu, piece_start);
    // O row of my piece: slot = warp-row-block index relative to the block holding the site's first pair
    T.pout = (int)(site - A.s0) * A.nslot + (int)((pl >> 5) - ((k0 - A.p0) >> 5));
    __syncthreads();
    const int64_t kbeg = s_kbeg;
    T.staged = s_nrec <= A_MAXREC;
    if (T.staged) {
        const double2* src = reinterpret_cast<const double2*>(A.rec + kbeg * RS);
        double2* dst = reinterpret_cast<double2*>(sm);
        const int n2 = s_nrec * (RS / 2);
        for (int i = tid; i < n2; i += A_THREADS) dst[i] = __ldg(src + i);
    }
    T.recp = A.rec + (int64_t)RS * p;
    const double4 o0v = ld4(T.recp);
    const double4 o1v = ld4(T.recp + 4);
    T.Rj0 = o0v.x; T.Rj1 = o0v.y; T.Rj2 = o0v.z; T.rinvj = o0v.w;
    T.dxj = o1v.y; T.fcj = o1v.z; T.dfcj = o1v.w;
    T.n = n;
    T.jself = (int)(p - k0);
    T.rk_off = T.staged ? (k0 - kbeg) * RS : k0 * RS;
    // output staging has its own region behind the records: no CTA barrier after this one
    T.wstage = sm + (size_t)A_MAXREC * RS + (size_t)(tid >> 5) * 32 * A_OUTLD;
    T.plw = pl0 + (tid & ~31);
    T.nrw = max(0, min(32, nrow - (tid & ~31)));
    __syncthreads();
    if constexpr (U >= 0) run_unit<D, U>(A, T);
    else run_all_units<D, 0>(A, T);
}


template <int D, int U>
__device__ __forceinline__ void run_unit(const Fast3Args& A, const CtaState& T) {
    constexpr UnitTable<D> UT{};
    unit3b<D, UT.c0[U], UT.c1[U], UT.base[U]>(A, T);
}

template <int D, int U>
__device__ __forceinline__ void run_all_units(const Fast3Args& A, const CtaState& T) {
    constexpr UnitTable<D> UT{};
    if constexpr (U < UT.n) {
        unit3b<D, UT.c0[U], UT.c1[U], UT.base[U]>(A, T);
        run_all_units<D, U + 1>(A, T);
    }
}

template <int D, int U>
void launch_moments_from(ipf_ctx* ctx, const Fast3Args& A) {
    constexpr UnitTable<D> UT{};
    if constexpr (U < UT.n) {
        const unsigned grid = (unsigned)((A.np + A_THREADS - 1) / A_THREADS);
        auto kern = moments3b_kernel<D, U>;
        constexpr size_t smem = moments_smem_bytes<D>();
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      // per device: set every time
        kern<<<grid, A_THREADS, smem, ctx->stream>>>(A);
        ctx->launches++;
        launch_moments_from<D, U + 1>(ctx, A);
    }
}

template <int D>
void launch_moments(ipf_ctx* ctx, const Fast3Args& A) {
    static const bool all_units = getenv("IPF_3B_ALLUNITS") != nullptr;     // experiment switch
    if (!all_units) { launch_moments_from<D, 0>(ctx, A); return; }
    const unsigned grid = (unsigned)((A.np + A_THREADS - 1) / A_THREADS);
    auto kern = moments3b_kernel<D, -1>;
    constexpr size_t smem = moments_smem_bytes<D>();
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, A_THREADS, smem, ctx->stream>>>(A);
    ctx->launches++;
}

template <int D>
void unit_layout_t(std::vector<int>& ub, std::vector<int>& uc0, std::vector<int>& uc1, int& nbs) {
    constexpr UnitTable<D> T{};
    for (int u = 0; u < T.n; ++u) { ub.push_back(T.base[u]); uc0.push_back(T.c0[u]); uc1.push_back(T.c1[u]); }
    nbs = T.base[T.n];
}

void unit_layout(int D, std::vector<int>& ub, std::vector<int>& uc0, std::vector<int>& uc1, int& nbs) {
    switch (D) {
#define IPF_CASE(DD) case DD: unit_layout_t<DD>(ub, uc0, uc1, nbs); break;
        IPF_CASE(2) IPF_CASE(3) IPF_CASE(4) IPF_CASE(5) IPF_CASE(6) IPF_CASE(7) IPF_CASE(8)
        IPF_CASE(9) IPF_CASE(10) IPF_CASE(11) IPF_CASE(12) IPF_CASE(13) IPF_CASE(14)
#undef IPF_CASE
        default: nbs = 0;
    }
}

}  // namespace

// Returns IPF_OK and sets *handled when the block is the canonical nbpolys(3, desc, D) basis
// with a supported D; otherwise *handled = false and nothing is done.
int ipf_assemble_block_3b(ipf_ctx* ctx, const BlockDev& blk, const ChunkDev& ch, const NeighList& nl,
                          ipf_matrix* Psi, bool* handled) {
    *handled = false;
    if (blk.N != 3 || blk.maxdeg > MAXD) return IPF_OK;
    if (nl.max_nbrs > A_THREADS) return IPF_OK;      // a site must not span more than two producer CTAs
    // ---- canonical check + column map
    int D = 0;
    for (int m = 0; m < blk.nmono; ++m) {
        const uint8_t* e = &blk.h_mono_exp[(size_t)m * IPF_MAX_VARS];
        D = std::max(D, (int)e[0] + e[1] + e[2]);
    }
    if (D < 2 || D > MAXD) return IPF_OK;
    std::vector<int16_t> colmap((MAXD + 1) * (MAXD + 1) * (MAXD + 1), -1);
    int expect = 0;
    for (int deg = 1; deg <= D; ++deg)
        for (int a = 0; a <= deg; ++a)
            for (int b = a; a + b <= deg; ++b) {
                const int c = deg - a - b;
                // orbit sorted: (a,b,c) then (b,a,c) when a < b
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
                colmap[cmidx(c, a, b)] = (int16_t)expect;
                ++expect;
            }
    if (expect != blk.nb) return IPF_OK;
    if (ch.ncfg == 0 || nl.npairs == 0) { *handled = true; return IPF_OK; }

    cudaStream_t st = ctx->stream;
    const double* rec = nullptr;
    int rs = 0;
    IPF_CHECK(ipf_pair_records(ctx, blk, nl, &rec, &rs));
    // scratch column order = the order the unit kernels emit the functions (c ascending, then a, then b),
    // every unit padded to an even number of columns (-1 = padding)
    std::vector<int16_t> h_s2c;
    int nbs = 0;
    {
        std::vector<int> ub, uc0, uc1;
        unit_layout(D, ub, uc0, uc1, nbs);
        h_s2c.assign(nbs, -1);
        int nreal = 0;
        for (size_t u = 0; u < uc0.size(); ++u) {
            int sc = ub[u];
            for (int c = uc0[u]; c <= uc1[u]; ++c)
                for (int a = 0; a <= (D - c) / 2; ++a)
                    for (int b = a; b <= D - c - a; ++b)
                        if (a + b + c > 0) { h_s2c[sc++] = colmap[cmidx(c, a, b)]; ++nreal; }
        }
        if (nreal != blk.nb) return ipf_set_error(ctx, IPF_EINVAL, "internal: 3-body column count mismatch");
    }
    ABuf<int16_t> s2c;
    IPF_CHECK(s2c.alloc(ctx, h_s2c.size()));
    IPF_CUDA(ctx, cudaMemcpyAsync(s2c.p, h_s2c.data(), sizeof(int16_t) * h_s2c.size(), cudaMemcpyHostToDevice, st));
    IPF_CUDA(ctx, cudaStreamSynchronize(st));     // h_s2c is a local

    // configuration-resident kernel (no HBM scratch) when every configuration of the chunk fits in shared memory;
    // IPF_3B_SCRATCH=1 or ipf_set_force_generic(ctx, 2) forces the scratch + gather path below
    static const bool force_scratch = getenv("IPF_3B_SCRATCH") != nullptr;
    if (!force_scratch && !ctx->force_scratch3b) {
        IPF_CHECK(ipf_assemble_block_3b_cfg(ctx, blk, ch, nl, Psi, D, rec, s2c.p, handled));
        if (*handled) return IPF_OK;
    }

    // batches of whole configurations; S (32 B per pair and function) lives in HBM, double buffered:
    // the gather of batch i (HBM bound, stream2) overlaps the moments of batch i+1 (FP64 bound, stream)
    const size_t s_budget = (size_t)3 << 30;
    const int64_t max_pairs_batch = std::max<int64_t>(nl.maxpairs_cfg, (int64_t)(s_budget / (24 * (size_t)nbs)));
    const size_t s_bytes = (size_t)24 * nbs * (size_t)std::min<int64_t>(max_pairs_batch, nl.npairs);
    // own-sum buffer: sites of a batch <= pairs of a batch (every site of the fast path has >= 0 pairs;
    // sized by the largest batch in sites)
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
        double tot = 0.0;
        for (int64_t s = 0; s < nl.nsites; ++s) {
            const double n = (double)(nl.h_site_off[s + 1] - nl.h_site_off[s]);
            tot += 0.5 * n * (n - 1.0);
        }
        ctx->prof["count:clusters_3b"].ms += tot;
        ctx->prof["count:ordered_pairs_3b"].ms += 2.0 * tot;
        ctx->prof["count:pairs_3b"].ms += (double)nl.npairs;
    }
    constexpr size_t g_bytes = sizeof(double) * (32 * G_LDT + G_WARPS * 7 * 32);
    IPF_CUDA(ctx, cudaFuncSetAttribute(gather3b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_bytes));
    static const bool no_overlap = getenv("IPF_NO_OVERLAP") != nullptr;
    cudaStream_t st2 = no_overlap ? ctx->stream : ctx->stream2;
    cudaEvent_t gathered[2] = {nullptr, nullptr};    // gather of the batch that last used buffer b
    // everything queued so far on the main stream (neighbour list, Psi zero fill, ...) precedes the gathers
    int64_t c0 = 0;
    int ib = 0;
    while (c0 < ch.ncfg) {
        int64_t c1 = c0 + 1;
        while (c1 < ch.ncfg && nl.h_cfg_pair_off[c1 + 1] - nl.h_cfg_pair_off[c0] <= max_pairs_batch) ++c1;
        const int buf = ib & 1;
        Fast3Args A;
        A.p0 = nl.h_cfg_pair_off[c0]; A.np = nl.h_cfg_pair_off[c1] - A.p0;
        A.rec = rec; A.site_off = nl.site_off.p; A.atom_off = ch.atom_off.p;
        A.psite = nl.psite.p; A.nb = nbs; A.S = (double*)Sbuf[buf];
        A.O = (double*)Obuf[buf]; A.s0 = ch.h_atom_off[c0]; A.nslot = nslot;
        A.pbeg = nl.pbeg.p; A.pcnt = nl.pcnt.p;
        if (gathered[buf]) {      // buffer reuse: wait for the gather that read it
            IPF_CUDA(ctx, cudaStreamWaitEvent(st, gathered[buf], 0));
            cudaEventDestroy(gathered[buf]);
            gathered[buf] = nullptr;
        }
        if (A.np > 0) {
            ProfScope prof_(ctx, "site_deriv_3b");
            switch (D) {
#define IPF_CASE(DD) case DD: launch_moments<DD>(ctx, A); break;
                IPF_CASE(2) IPF_CASE(3) IPF_CASE(4) IPF_CASE(5) IPF_CASE(6) IPF_CASE(7) IPF_CASE(8)
                IPF_CASE(9) IPF_CASE(10) IPF_CASE(11) IPF_CASE(12) IPF_CASE(13) IPF_CASE(14)
#undef IPF_CASE
                default: return ipf_set_error(ctx, IPF_EINVAL, "3-body degree %d not instantiated", D);
            }
        }
        cudaEvent_t produced;
        IPF_CUDA(ctx, cudaEventCreateWithFlags(&produced, cudaEventDisableTiming));
        IPF_CUDA(ctx, cudaEventRecord(produced, st));
        IPF_CUDA(ctx, cudaStreamWaitEvent(st2, produced, 0));
        cudaEventDestroy(produced);
        Gather3Args G;
        G.p0 = A.p0; G.c0 = c0; G.s0 = A.s0; G.nb = nbs; G.S = (const double*)Sbuf[buf];
        G.pstride = (int64_t)3 * nbs; G.gstride = 0; G.cfg_aligned = 0;
        G.O = (const double*)Obuf[buf]; G.nslot = nslot; G.scol2col = s2c.p;
        G.atom_off = ch.atom_off.p; G.site_off = nl.site_off.p; G.prev = nl.prev.p;
        G.pR = nl.pR.p; G.rowE = ch.rowE.p; G.rowF = ch.rowF.p; G.rowV = ch.rowV.p;
        G.Psi = Psi->d; G.ld = Psi->ld; G.col0 = blk.col0;
        {
            ProfScope prof_(ctx, "gather_3b", st2);
            dim3 grid((unsigned)(c1 - c0), (unsigned)((nbs + 31) / 32));
            gather3b_kernel<<<grid, G_THREADS, g_bytes, st2>>>(G);
            ctx->launches++;
        }
        IPF_CUDA(ctx, cudaEventCreateWithFlags(&gathered[buf], cudaEventDisableTiming));
        IPF_CUDA(ctx, cudaEventRecord(gathered[buf], st2));
        c0 = c1;
        ++ib;
    }
    // later work on the main stream (next block, arena reuse, the solve) must see the gathered rows
    for (int b = 0; b < 2; ++b)
        if (gathered[b]) {
            IPF_CUDA(ctx, cudaStreamWaitEvent(st, gathered[b], 0));
            cudaEventDestroy(gathered[b]);
        }
    IPF_CUDA(ctx, cudaGetLastError());
    *handled = true;
    return IPF_OK;
}
