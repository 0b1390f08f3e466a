// basis_3b_cfg.cu -- K3 "configuration-resident" 3-body kernel: no HBM scratch.
//
// Same own-leg decomposition and moment factorisation as basis_3b.cu, but one CTA owns a whole
// configuration (for one unit of w-exponents) and walks its ordered pairs in passes of C_T rows.
// The force accumulators F[atom][function][3] of the configuration live in shared memory for the
// life of the CTA, so the contributions  F_b[i] += g_p,  F_b[j] -= g_p  of pair p = (i, j) never
// leave the SM: per pass and per group of 8 functions the rows stage (g_p, e_p) in shared memory and
// each warp then updates the atoms it owns (atom mod C_NW == warp) --
//     the rows that TARGET its atoms, in (atom, row) order (lists pre-sorted by tsort_kernel), then
//     the own rows of its atoms (contiguous: pairs are sorted by centre atom), in row order,
// i.e. fixed summation order, no atomics, bit-reproducible.  The virial  -sum_p g_p (x) R_p  and the
// energy  sum_p e_p  are accumulated by the same lanes over the own rows (registers across passes).
// Algorithmic HBM traffic drops from 48 B per (pair, function) [S written + read back] to the pair
// records (read) and the 8 B per entry of Psi (written once).
// Pipelining: the pair records of pass q+1 are fetched with cp.async into the (single) record buffer
// as soon as the moment loop of pass q is done; the small per-pass inputs (bond vectors, sorted
// lists, per-row list offsets) are double buffered and fetched one pass ahead.
//
// Used when every configuration of the chunk fits (24 B * natoms * widest unit + 150 KB of staging
// <= 227 KB, i.e. <= ~75 atoms at D = 11); larger configurations take the scratch path of basis_3b.cu.
#include <climits>
#include <cstdlib>

#include "basis_3b_shared.cuh"

namespace {

constexpr int C_T = 256;                  // rows (ordered pairs) per pass = threads per CTA
constexpr int C_NW = C_T / 32;
constexpr int C_G = 8;                    // functions per staged group
constexpr int C_SLD = 4 * C_G + 2;        // doubles per staged row: [fn][gx, gy, gz, e] + 2 (conflict-free 16-byte columns)
constexpr int C_ZROW = C_T * C_SLD;       // offset of the all-zero stage row the padding entries point at

// small per-pass inputs, one contiguous block per pass (global and shared memory alike), in bytes:
//   [ngmax] groups of 16 bytes {row offset, row offset, row offset, target atom}: the rows of the pass that TARGET
//           an atom, three per group (unused slots point at the all-zero stage row), groups ordered by
//           (owner warp, atom, row); row offset = byte offset of the staged row
//   [16] u16  first group of every owner warp (C_NW + 1 used)
//   [C_T] u16 offset of the row's centre-atom list in the record window
//   [C_T] u16 length of that list
//   [C_T] u16 centre atom of the row (config-local)
constexpr int C_GROWS = 3;
constexpr int C_MAXATOMS = 1024;          // atoms per configuration the kernel accepts (tsort_kernel tables)
__host__ __device__ constexpr int pb_bytes(int ngmax) { return 16 * ngmax + 32 + 6 * C_T; }
__host__ __device__ constexpr int pb_woff(int ngmax) { return 16 * ngmax; }
__host__ __device__ constexpr int pb_k0(int ngmax) { return 16 * ngmax + 32; }
__host__ __device__ constexpr int pb_n(int ngmax) { return 16 * ngmax + 32 + 2 * C_T; }
__host__ __device__ constexpr int pb_atom(int ngmax) { return 16 * ngmax + 32 + 4 * C_T; }

// per-pass header written by tsort_kernel
struct PassHdr {
    int64_t kbeg;      // first pair record of the pass's record window
    int32_t nrec;      // records in the window (neighbour lists of all centre atoms the pass touches)
    int32_t sf, sl;    // first / last centre atom (config-local) with rows in the pass
    int32_t pad;
};

struct Cfg3Args {
    const double* rec;             // pair records (global), stride RecStride<D>
    const int64_t* atom_off;       // [ncfg+1]
    const int64_t* site_off;       // [nsites+1]
    const int64_t* cfg_pair_off;   // [ncfg+1]
    const double* pR;              // [npairs][3] bond vectors
    const PassHdr* hdr;            // [npass]
    const unsigned char* pblock;   // [npass][pb_bytes(ngmax)]
    int ngmax;                     // groups capacity of a pass
    const int32_t* pass_off;       // [ncfg+1] first pass of every configuration
    int ga;                        // atoms capacity (largest configuration of the chunk)
    int maxrec;                    // records staged per pass (window capacity)
    const int16_t* s2c;            // scratch column -> local basis column
    const int64_t* rowE;
    const int64_t* rowF;
    const int64_t* rowV;
    double* Psi;
    int64_t ld, col0;
};

// Per pass: the rows grouped by target atom (three per group, deterministic order: owner warp = atom mod C_NW,
// atom, row), plus the record window, per-row list offsets and the centre-atom range of the pass.
__global__ void __launch_bounds__(C_T) tsort_kernel(const int64_t* __restrict__ cfg_pair_off,
                                                    const int64_t* __restrict__ atom_off,
                                                    const int32_t* __restrict__ pass_off,
                                                    const int32_t* __restrict__ pi, const int32_t* __restrict__ pj,
                                                    const int64_t* __restrict__ pbeg, const int32_t* __restrict__ pcnt,
                                                    int ngmax, PassHdr* __restrict__ hdr,
                                                    unsigned char* __restrict__ pblock) {
    __shared__ int cntA[C_MAXATOMS], gstart[C_MAXATOMS];
    __shared__ unsigned rowmask[C_MAXATOMS][C_NW];      // bit r of word w: row 32 w + r targets the atom
    __shared__ int wo[C_NW + 1];
    const int c = blockIdx.x, q = blockIdx.y, tid = threadIdx.x;
    const int64_t pc0 = cfg_pair_off[c];
    const int64_t npc = cfg_pair_off[c + 1] - pc0;
    if ((int64_t)q * C_T >= npc) return;
    const int nat = (int)(atom_off[c + 1] - atom_off[c]);
    const int nrow = (int)min((int64_t)C_T, npc - (int64_t)q * C_T);
    const int64_t P0 = pc0 + (int64_t)q * C_T;
    const int64_t pass = pass_off[c] + q;
    unsigned char* pb = pblock + pass * pb_bytes(ngmax);
    uint32_t* grp = reinterpret_cast<uint32_t*>(pb);
    uint16_t* woff = reinterpret_cast<uint16_t*>(pb + pb_woff(ngmax));
    uint16_t* k0s = reinterpret_cast<uint16_t*>(pb + pb_k0(ngmax));
    uint16_t* ns = reinterpret_cast<uint16_t*>(pb + pb_n(ngmax));
    uint16_t* ctr = reinterpret_cast<uint16_t*>(pb + pb_atom(ngmax));
    const int tgt = tid < nrow ? pj[P0 + tid] : -1;
    for (int i = tid; i < nat * C_NW; i += C_T) (&rowmask[0][0])[i] = 0u;
    __syncthreads();
    if (tid < nrow) atomicOr(&rowmask[tgt][tid >> 5], 1u << (tid & 31));      // the final masks do not depend on the order
    __syncthreads();
    // k = rows before this one with the same target; cntA = rows of the atom (both from the bit masks)
    int k = 0;
    if (tid < nrow) {
        const int w = tid >> 5;
        for (int u = 0; u < w; ++u) k += __popc(rowmask[tgt][u]);
        k += __popc(rowmask[tgt][w] & ((1u << (tid & 31)) - 1u));
    }
    for (int a = tid; a < nat; a += C_T) {
        int n = 0;
#pragma unroll
        for (int u = 0; u < C_NW; ++u) n += __popc(rowmask[a][u]);
        cntA[a] = n;
    }
    __syncthreads();
    if (tid == 0) {      // first group of every atom, atoms in (owner warp, atom) order
        int run = 0;
        for (int w = 0; w < C_NW; ++w) {
            wo[w] = run;
            for (int a = w; a < nat; a += C_NW) { gstart[a] = run; run += (cntA[a] + C_GROWS - 1) / C_GROWS; }
        }
        wo[C_NW] = run;
    }
    __syncthreads();
    const int64_t kbeg = pbeg[P0];
    if (tid < nrow) {
        const int g = gstart[tgt] + k / C_GROWS, slot = k % C_GROWS;
        grp[4 * g + slot] = (uint32_t)(tid * C_SLD * 8);
        if (slot == 0) {          // group leader: the atom and the unused slots
            grp[4 * g + 3] = (uint32_t)tgt;
            const int left = cntA[tgt] - k;      // rows of the atom from this one on
            for (int s2 = left; s2 < C_GROWS; ++s2) grp[4 * g + s2] = (uint32_t)(C_ZROW * 8);
        }
        k0s[tid] = (uint16_t)(pbeg[P0 + tid] - kbeg);
        ns[tid] = (uint16_t)pcnt[P0 + tid];
        ctr[tid] = (uint16_t)pi[P0 + tid];
    } else {
        k0s[tid] = 0;
        ns[tid] = 0;
        ctr[tid] = 0xffff;
    }
    if (tid < 16) woff[tid] = (uint16_t)(tid <= C_NW ? wo[tid] : 0);
    if (tid == 0) {
        PassHdr h;
        h.kbeg = kbeg;
        h.nrec = (int32_t)(pbeg[P0 + nrow - 1] + pcnt[P0 + nrow - 1] - kbeg);
        h.sf = pi[P0];
        h.sl = pi[P0 + nrow - 1];
        h.pad = 0;
        hdr[pass] = h;
    }
}

// record stride in shared memory: an odd number of 16-byte pieces, so that the records of the two or three
// centre atoms a warp spans fall on different banks
template <int D>
struct SRecStride { static constexpr int value = RecStride<D>::value + ((RecStride<D>::value / 2) % 2 == 0 ? 2 : 0); };

// dynamic shared memory of one CTA
template <int D, int NF>
__host__ __device__ constexpr size_t cfg_smem_bytes(int ga, int maxrec, int ngmax) {
    return sizeof(double) * ((size_t)maxrec * SRecStride<D>::value + (size_t)(C_T + 1) * C_SLD + (((size_t)ga * NF * 3 + 1) & ~(size_t)1) +
                             2 * (size_t)C_T * 4) +
           2 * (size_t)pb_bytes(ngmax) + sizeof(uint16_t) * (2 * (size_t)ga + 8) + 16;
}

struct CfgSmem {
    double* stage;      // [C_T + 1][C_SLD]  (last row: zeros)
    double* F;          // [nat][NF][3]
    uint16_t* own0;     // [ga]
    uint16_t* own1;     // [ga]
};
// per-pass inputs (double buffered)
struct PassBuf {
    double* R;          // [C_T][4]
    unsigned char* pb;  // [pb_bytes(ngmax)]
    int ngmax;
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Scatter of one staged group.  lane = (fn, d): fn = lane / 4 function of the group, d = lane % 4 component
// (gx, gy, gz, e).  Every warp owns the atoms a with a mod C_NW == warp:
//   rows that TARGET its atoms: one flat loop over the warp's slice of the sorted list, four entries per
//     iteration (F[a] -= sum per atom);
//   own rows of its atoms (at most two centre atoms of a pass belong to one warp): F[a] += sum,
//   vE[0..2] += sum_p g_d R_p[beta]  (d < 3),  vE[0] += sum_p e_p  (d = 3).
template <int NG, int NF>
__device__ __forceinline__ void scatter_group(const CfgSmem& M, const PassBuf& B, int sf, int sl, int g0, double (&vE)[3]) {
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fn = lane >> 2, d = lane & 3;
    const bool fl = fn < NG && d < 3;
    const double* __restrict__ st = M.stage + lane;
    double* __restrict__ Fl = M.F + (g0 + fn) * 3 + d;
    {
        const uint4* __restrict__ grp = reinterpret_cast<const uint4*>(B.pb);
        const uint16_t* __restrict__ woff = reinterpret_cast<const uint16_t*>(B.pb + pb_woff(B.ngmax));
        const char* __restrict__ stb = reinterpret_cast<const char*>(st);
        char* __restrict__ Fb = reinterpret_cast<char*>(Fl);
        const int t0 = woff[warp], t1 = woff[warp + 1];
        // one group = up to three staged rows that target the same atom: no per-row branch, one update of F per group.
        // Software pipeline: the group words are fetched two iterations ahead, the staged values one ahead.
        // (Two groups per iteration with both F loads ahead of both stores was measured slower: register moves.)
        if (t0 < t1) {
            uint4 g = grp[t0];
            uint4 gn = grp[min(t0 + 1, t1 - 1)];
            double v0 = *reinterpret_cast<const double*>(stb + g.x);
            double v1 = *reinterpret_cast<const double*>(stb + g.y);
            double v2 = *reinterpret_cast<const double*>(stb + g.z);
            for (int t = t0; t < t1; ++t) {
                const uint4 gnn = grp[min(t + 2, t1 - 1)];
                const double n0 = *reinterpret_cast<const double*>(stb + gn.x);
                const double n1 = *reinterpret_cast<const double*>(stb + gn.y);
                const double n2 = *reinterpret_cast<const double*>(stb + gn.z);
                double* fp = reinterpret_cast<double*>(Fb + g.w * (unsigned)(NF * 24));
                if (fl) *fp -= (v0 + v1) + v2;
                g = gn; gn = gnn; v0 = n0; v1 = n1; v2 = n2;
            }
        }
    }
    const double* __restrict__ Rb = B.R;
    for (int a = sf + ((warp - sf) & (C_NW - 1)); a <= sl; a += C_NW) {
        const int q0 = M.own0[a], q1 = M.own1[a];
        double acc = 0.0, w0 = 0.0, w1 = 0.0, w2 = 0.0;
#pragma unroll 4
        for (int r = q0; r < q1; ++r) {
            const double val = st[r * C_SLD];
            const double2 r01 = *reinterpret_cast<const double2*>(Rb + 4 * r);
            const double r2 = Rb[4 * r + 2];
            acc += val;
            w0 = fma(val, r01.x, w0);
            w1 = fma(val, r01.y, w1);
            w2 = fma(val, r2, w2);
        }
        if (d == 3) {
            vE[0] += acc;
        } else {
            vE[0] += w0; vE[1] += w1; vE[2] += w2;
            if (fl) Fl[a * (NF * 3)] += acc;
        }
    }
    __syncthreads();
}

// CTA = (configuration, unit U of w-exponents): body of one unit
template <int D, int U>
__device__ __forceinline__ void cfg3b_unit(const Cfg3Args& A, const int c) {
    constexpr UnitTable<D> UT{};
    constexpr int C0 = UT.c0[U], C1 = UT.c1[U], SCB = UT.base[U];
    constexpr int NC = C1 - C0 + 1;
    constexpr int EMAX = D - C0;
    constexpr int NF = scol_base(D, C1 + 1) - scol_base(D, C0);
    constexpr int NGRP = (NF + C_G - 1) / C_G;
    constexpr int RS = RecStride<D>::value;
    constexpr int SRS = SRecStride<D>::value;
    constexpr int PPR = RS / 2;               // 16-byte pieces per record
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t a0 = A.atom_off[c];
    const int nat = (int)(A.atom_off[c + 1] - a0);
    const int64_t pc0 = A.cfg_pair_off[c];
    const int64_t npc = A.cfg_pair_off[c + 1] - pc0;
    const int64_t rE = A.rowE[c], rF = A.rowF[c], rV = A.rowV[c];
    if (nat == 0 || (rE == 0 && rF == 0 && rV == 0)) return;

    CfgSmem M;
    PassBuf B0;
    double* srec = sm;                                  // [maxrec][SRS]
    {
        double* q = sm + (size_t)A.maxrec * SRS;
        M.stage = q; q += (size_t)(C_T + 1) * C_SLD;
        M.F = q; q += ((size_t)A.ga * NF * 3 + 1) & ~(size_t)1;      // even: everything behind stays 16-byte aligned
        B0.R = q; q += 2 * (size_t)C_T * 4;
        B0.pb = reinterpret_cast<unsigned char*>(q);
        B0.ngmax = A.ngmax;
        uint16_t* h = reinterpret_cast<uint16_t*>(B0.pb + 2 * (size_t)pb_bytes(A.ngmax));
        M.own0 = h; h += A.ga;
        M.own1 = h;
    }
    auto pass_buf = [&](int b) {
        PassBuf B = B0;
        if (b) { B.R += C_T * 4; B.pb += pb_bytes(A.ngmax); }
        return B;
    };
    const int npass = (int)((npc + C_T - 1) / C_T);
    const int64_t pass0 = A.pass_off[c];

    // asynchronous copies (the caller commits): small per-pass inputs of pass q into buffer b ...
    auto fetch_inputs = [&](int q, int b) {
        const PassBuf B = pass_buf(b);
        const int64_t P0 = pc0 + (int64_t)q * C_T;
        const int nrow = (int)min((int64_t)C_T, npc - (int64_t)q * C_T);
        for (int i = tid; i < 3 * nrow; i += C_T) {
            const int r = i / 3, k = i - 3 * r;
            cp_async8(B.R + 4 * r + k, A.pR + 3 * P0 + i);
        }
        const unsigned char* src = A.pblock + (pass0 + q) * (int64_t)pb_bytes(A.ngmax);
        for (int i = tid; i < pb_bytes(A.ngmax) / 16; i += C_T) cp_async16(B.pb + 16 * i, src + 16 * i);
    };
    // ... and the record window of a pass into the record buffer
    auto fetch_records = [&](const PassHdr& h) {
        if (h.nrec > A.maxrec) return;
        const double* src = A.rec + h.kbeg * RS;
        for (int i = tid; i < h.nrec * PPR; i += C_T) {
            const int r = i / PPR, u = i - r * PPR;
            cp_async16(srec + (size_t)r * SRS + 2 * u, src + 2 * (size_t)i);
        }
    };

    for (int i = tid; i < nat * NF * 3; i += C_T) M.F[i] = 0.0;
    for (int i = tid; i < C_SLD; i += C_T) M.stage[C_ZROW + i] = 0.0;
    for (int i = tid; i < nat; i += C_T) { M.own0[i] = 0; M.own1[i] = 0; }
    double vE[NGRP][3];
#pragma unroll
    for (int g = 0; g < NGRP; ++g) vE[g][0] = vE[g][1] = vE[g][2] = 0.0;

    PassHdr hn;
    hn.kbeg = 0; hn.nrec = 0; hn.sf = 0; hn.sl = -1; hn.pad = 0;
    if (npass > 0) {
        hn = A.hdr[pass0];
        fetch_inputs(0, 0);
        fetch_records(hn);
        cp_async_commit();
    }
    for (int q = 0; q < npass; ++q) {
        const int b = q & 1;
        const PassHdr h = hn;
        if (q + 1 < npass) hn = A.hdr[pass0 + q + 1];
        const int64_t P0 = pc0 + (int64_t)q * C_T;
        const int nrow = (int)min((int64_t)C_T, npc - (int64_t)q * C_T);
        const bool live = tid < nrow;
        const bool staged = h.nrec <= A.maxrec;
        cp_async_wait_all();
        __syncthreads();          // records and inputs of this pass have landed; the previous pass is finished
        if (q + 1 < npass) { fetch_inputs(q + 1, b ^ 1); cp_async_commit(); }
        const PassBuf B = pass_buf(b);
        {   // own-row ranges of the centre atoms of this pass, from the per-row centre atoms (read in the scatters,
            // after >= 1 barrier; atoms without rows are never visited: sf..sl are consecutive atoms WITH rows or
            // atoms without any pair, whose entries keep their initial 0 / 0)
            const uint16_t* ctr = reinterpret_cast<const uint16_t*>(B.pb + pb_atom(A.ngmax));
            if (live) {
                const int at = ctr[tid];
                if (tid == 0 || ctr[tid - 1] != at) M.own0[at] = (uint16_t)tid;
                if (tid == nrow - 1 || ctr[tid + 1] != at) M.own1[at] = (uint16_t)(tid + 1);
            }
        }
        const int k0off = reinterpret_cast<const uint16_t*>(B.pb + pb_k0(A.ngmax))[tid];
        const int n = live ? (int)reinterpret_cast<const uint16_t*>(B.pb + pb_n(A.ngmax))[tid] : 0;
        const int rowoff = (int)(P0 - h.kbeg) + (live ? tid : 0);      // own record in the window
        const int jself = rowoff - k0off;

        double Rj0, Rj1, Rj2, rinvj, dxj, fcj, dfcj;
        {
            double4 o0v, o1v;
            if (staged) {
                const double2* hp = reinterpret_cast<const double2*>(srec + (size_t)rowoff * SRS);
                const double2 h0 = hp[0], h1 = hp[1], h2 = hp[2], h3 = hp[3];
                o0v = make_double4(h0.x, h0.y, h1.x, h1.y);
                o1v = make_double4(h2.x, h2.y, h3.x, h3.y);
            } else {
                const double* recp = A.rec + (h.kbeg + rowoff) * RS;
                o0v = ld4(recp);
                o1v = ld4(recp + 4);
            }
            Rj0 = o0v.x; Rj1 = o0v.y; Rj2 = o0v.z; rinvj = o0v.w;
            dxj = o1v.y; fcj = o1v.z; dfcj = o1v.w;
        }
        double Z[NC][EMAX + 1];
        double Uq[NC][EMAX + 1][2];
#pragma unroll
        for (int ci = 0; ci < NC; ++ci)
#pragma unroll
            for (int e = 0; e <= EMAX; ++e) { Z[ci][e] = 0.0; Uq[ci][e][0] = Uq[ci][e][1] = 0.0; }
        double E1[3], E2[3];
        {
            const bool usex = fabs(Rj0) < 0.8;
            double c0 = usex ? 0.0 : -Rj2, c1 = usex ? Rj2 : 0.0, c2 = usex ? -Rj1 : Rj0;
            const double inv = rsqrt(c0 * c0 + c1 * c1 + c2 * c2);
            E1[0] = c0 * inv; E1[1] = c1 * inv; E1[2] = c2 * inv;
            E2[0] = Rj1 * E1[2] - Rj2 * E1[1];
            E2[1] = Rj2 * E1[0] - Rj0 * E1[2];
            E2[2] = Rj0 * E1[1] - Rj1 * E1[0];
        }
        double X[EMAX + 1], dX[EMAX + 1];
        if (staged) {
            moment_loop<D, C0, C1, NC, EMAX, false, SRS>(srec + (size_t)k0off * SRS, n, jself, Rj0, Rj1, Rj2, E1, E2, Z, Uq);
            const double* xr = srec + (size_t)rowoff * SRS + REC;
#pragma unroll
            for (int e = 0; e <= EMAX; ++e) X[e] = xr[e];
        } else {
            moment_loop<D, C0, C1, NC, EMAX, false>(A.rec + (h.kbeg + k0off) * RS, n, jself, Rj0, Rj1, Rj2, E1, E2, Z, Uq);
            const double* xr = A.rec + (h.kbeg + rowoff) * RS + REC;
#pragma unroll
            for (int e = 0; e <= EMAX; ++e) X[e] = __ldg(xr + e);
        }
        __syncthreads();          // every thread is done with the record buffer: fetch the next window
        if (q + 1 < npass) { fetch_records(hn); cp_async_commit(); }

        // ---- combine into basis functions, staged in groups of C_G, scattered into F
        dX[0] = 0.0;
#pragma unroll
        for (int e = 1; e <= EMAX; ++e) dX[e] = (double)e * X[e - 1];
        const double tang = live ? fcj * rinvj : 0.0;
        const double fdx = fcj * dxj;
        double2* myrow = reinterpret_cast<double2*>(M.stage + (size_t)tid * C_SLD);
        int f = 0;
#pragma unroll
        for (int ci = 0; ci < NC; ++ci) {
            const int cw = C0 + ci;
            const double tc = tang * (double)cw;
#pragma unroll
            for (int a = 0; a <= (D - cw) / 2; ++a) {
#pragma unroll
                for (int bb = a; bb <= D - cw - a; ++bb) {
                    if (a + bb + cw == 0) continue;
                    double S0, S1, T1, T2;
                    if (a == bb) {
                        S0 = X[a] * Z[ci][a];
                        S1 = dX[a] * Z[ci][a];
                        T1 = X[a] * Uq[ci][a][0]; T2 = X[a] * Uq[ci][a][1];
                    } else {
                        S0 = fma(X[a], Z[ci][bb], X[bb] * Z[ci][a]);
                        S1 = fma(dX[a], Z[ci][bb], dX[bb] * Z[ci][a]);
                        T1 = fma(X[a], Uq[ci][bb][0], X[bb] * Uq[ci][a][0]);
                        T2 = fma(X[a], Uq[ci][bb][1], X[bb] * Uq[ci][a][1]);
                    }
                    const double radial = fma(dfcj, S0, fdx * S1);
                    T1 *= tc; T2 *= tc;
                    const int slot = f % C_G;
                    myrow[2 * slot] = make_double2(fma(radial, Rj0, fma(T1, E1[0], T2 * E2[0])),
                                                   fma(radial, Rj1, fma(T1, E1[1], T2 * E2[1])));
                    myrow[2 * slot + 1] = make_double2(fma(radial, Rj2, fma(T1, E1[2], T2 * E2[2])), 0.5 * fcj * S0);
                    ++f;
                    if (f % C_G == 0) scatter_group<C_G, NF>(M, B, h.sf, h.sl, f - C_G, vE[(f - 1) / C_G]);
                }
            }
        }
        constexpr int REM = NF % C_G;
        if constexpr (REM > 0) scatter_group<REM, NF>(M, B, h.sf, h.sl, NF - REM, vE[NGRP - 1]);
    }
    __syncthreads();

    // ---- write-out: F rows (coalesced column strips), then E / V rows (sum over warps in warp order)
    if (rF) {
        const int nr = 3 * nat;
        for (int i = tid; i < NF * nr; i += C_T) {
            const int fn = i / nr, k = i - fn * nr;
            const int a = k / 3, d = k - 3 * a;
            A.Psi[(size_t)(A.col0 + A.s2c[SCB + fn]) * A.ld + (rF - 1) + k] = M.F[(a * NF + fn) * 3 + d];
        }
    }
    double* red = M.stage;        // [C_NW][NGRP][32][3]
#pragma unroll
    for (int g = 0; g < NGRP; ++g) {
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) red[((warp * NGRP + g) * 32 + lane) * 3 + bb] = vE[g][bb];
    }
    __syncthreads();
    for (int i = tid; i < NGRP * 32; i += C_T) {
        const int g = i >> 5, l = i & 31;
        const int fn = g * C_G + (l >> 2), d = l & 3;
        if (fn >= NF) continue;
        double t[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int w = 0; w < C_NW; ++w)
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) t[bb] += red[((w * NGRP + g) * 32 + l) * 3 + bb];
        double* colp = A.Psi + (size_t)(A.col0 + A.s2c[SCB + fn]) * A.ld;
        if (d == 3) { if (rE) colp[rE - 1] = t[0]; }
        else if (rV) {
            // Voigt rows (xx, yy, zz, zy, zx, yx) of  -sum_p g_p (x) R_p ;  this lane holds row alpha = d
            if (d == 0) colp[rV - 1 + 0] = -t[0];
            else if (d == 1) { colp[rV - 1 + 1] = -t[1]; colp[rV - 1 + 5] = -t[0]; }
            else { colp[rV - 1 + 2] = -t[2]; colp[rV - 1 + 3] = -t[1]; colp[rV - 1 + 4] = -t[0]; }
        }
    }
}

template <int D, int U>
__device__ __forceinline__ void cfg3b_dispatch(const Cfg3Args& A, int unit, int c) {
    constexpr UnitTable<D> UT{};
    if constexpr (U < UT.n) {
        if (unit == U) cfg3b_unit<D, U>(A, c);
        else cfg3b_dispatch<D, U + 1>(A, unit, c);
    }
}

// One launch per chunk, all units in one grid.  Block order: groups of `gsz` configurations (one wave of SMs); within
// a group the units follow each other, each over the same gsz configurations.  So an SM keeps running the same unit
// body for a whole wave (instruction cache), the eight waves of a group re-read the group's pair records from L2
// instead of HBM (gsz x ~0.4 MB), and the tail of the grid is one CTA deep instead of one per unit.  One CTA per SM
// either way, so the register count of the heaviest unit costs nothing.  IPF_3B_UNIT_LAUNCHES=1: unit by unit.
template <int D>
__global__ void __launch_bounds__(C_T, 1) cfg3b_kernel(Cfg3Args A, int unit0, int nunits, int ncfg, int gsz) {
    const int per_group = gsz * nunits;
    const int g = (int)blockIdx.x / per_group, r = (int)blockIdx.x - g * per_group;
    const int c0 = g * gsz;
    const int gn = min(gsz, ncfg - c0);             // configurations of this (possibly last, shorter) group
    const int unit = r / gn, c = c0 + r - unit * gn;
    if (unit >= nunits) return;
    cfg3b_dispatch<D, 0>(A, unit0 + unit, c);
}

template <int D>
int launch_cfg(ipf_ctx* ctx, const Cfg3Args& A, int ncfg, size_t smem) {
    constexpr UnitTable<D> UT{};
    auto kern = cfg3b_kernel<D>;
    // the attribute is per device (several ctxs of one process may sit on different GPUs): set it on every launch
    IPF_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static const bool unit_launches = getenv("IPF_3B_UNIT_LAUNCHES") != nullptr;
    if (unit_launches) {
        for (int u = 0; u < UT.n; ++u) { kern<<<(unsigned)ncfg, C_T, smem, ctx->stream>>>(A, u, 1, ncfg, ncfg); ctx->launches++; }
    } else {
        kern<<<(unsigned)ncfg * (unsigned)UT.n, C_T, smem, ctx->stream>>>(A, 0, UT.n, ncfg, std::max(1, ctx->sm_count));
        ctx->launches++;
    }
    return IPF_OK;
}

// largest dynamic shared memory over the units of degree D
template <int D, int U>
size_t cfg_smem_max(int ga, int maxrec, int ngmax) {
    constexpr UnitTable<D> UT{};
    if constexpr (U < UT.n) {
        constexpr int NF = scol_base(D, UT.c1[U] + 1) - scol_base(D, UT.c0[U]);
        return std::max(cfg_smem_bytes<D, NF>(ga, maxrec, ngmax), cfg_smem_max<D, U + 1>(ga, maxrec, ngmax));
    }
    return 0;
}

}  // namespace

// Config-resident 3-body assembly of a canonical nbpolys(3, desc, D) block.  `s2c` = device map scratch column ->
// local basis column (unit-major order, basis_3b.cu).  *handled = false when a configuration does not fit.
int ipf_assemble_block_3b_cfg(ipf_ctx* ctx, const BlockDev& blk, const ChunkDev& ch, const NeighList& nl,
                              ipf_matrix* Psi, int D, const double* rec, const int16_t* s2c, bool* handled) {
    *handled = false;
    int ga = 0;
    for (int64_t c = 0; c < ch.ncfg; ++c) ga = std::max<int64_t>(ga, ch.h_atom_off[c + 1] - ch.h_atom_off[c]);
    if (ga == 0) { *handled = true; return IPF_OK; }
    if (ga > C_MAXATOMS || nl.maxpairs_cfg >= (1 << 30) || nl.max_nbrs > 4096) return IPF_OK;
    // record window of a pass: its rows plus the rest of the first and last centre atom's lists
    const int maxrec = (std::min(C_T + 2 * std::max(0, nl.max_nbrs - 1), 2 * C_T) + 1) & ~1;
    // groups of a pass: sum over target atoms of ceil(rows / 3) <= (C_T + 2 * atoms) / 3
    const int ngmax = std::min(C_T, (C_T + 2 * ga) / C_GROWS + 1);
    size_t need = 0;
    switch (D) {
#define IPF_CASE(DD) case DD: need = cfg_smem_max<DD, 0>(ga, maxrec, ngmax); break;
        IPF_CASE(2) IPF_CASE(3) IPF_CASE(4) IPF_CASE(5) IPF_CASE(6) IPF_CASE(7) IPF_CASE(8)
        IPF_CASE(9) IPF_CASE(10) IPF_CASE(11) IPF_CASE(12) IPF_CASE(13) IPF_CASE(14)
#undef IPF_CASE
        default: return IPF_OK;
    }
    if (need > (size_t)227 * 1024) return IPF_OK;

    cudaStream_t st = ctx->stream;
    // passes of every configuration
    std::vector<int32_t> h_pass(ch.ncfg + 1, 0);
    int maxpass = 0;
    for (int64_t c = 0; c < ch.ncfg; ++c) {
        const int64_t np = nl.h_cfg_pair_off[c + 1] - nl.h_cfg_pair_off[c];
        const int npass = (int)((np + C_T - 1) / C_T);
        h_pass[c + 1] = h_pass[c] + npass;
        maxpass = std::max(maxpass, npass);
    }
    const int64_t npass_tot = h_pass[ch.ncfg];
    ABuf<int32_t> pass_off;
    ABuf<unsigned char> pblock;
    ABuf<PassHdr> hdr;
    IPF_CHECK(pass_off.alloc(ctx, ch.ncfg + 1));
    IPF_CHECK(hdr.alloc(ctx, (size_t)npass_tot));
    IPF_CHECK(pblock.alloc(ctx, (size_t)npass_tot * pb_bytes(ngmax)));
    IPF_CUDA(ctx, cudaMemcpyAsync(pass_off.p, h_pass.data(), sizeof(int32_t) * h_pass.size(), cudaMemcpyHostToDevice, st));
    IPF_CUDA(ctx, cudaStreamSynchronize(st));     // h_pass is a local; nothing long is queued yet
    if (maxpass > 0) {
        ProfScope prof_(ctx, "tsort_3b");
        dim3 grid((unsigned)ch.ncfg, (unsigned)maxpass);
        tsort_kernel<<<grid, C_T, 0, st>>>(nl.cfg_pair_off.p, ch.atom_off.p, pass_off.p, nl.pi.p, nl.pj.p, nl.pbeg.p, nl.pcnt.p,
                                         ngmax, hdr.p, pblock.p);
        ctx->launches++;
    }
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
    Cfg3Args A;
    A.rec = rec; A.atom_off = ch.atom_off.p; A.site_off = nl.site_off.p; A.cfg_pair_off = nl.cfg_pair_off.p;
    A.pR = nl.pR.p; A.hdr = hdr.p; A.pblock = pblock.p; A.ngmax = ngmax;
    A.pass_off = pass_off.p; A.ga = ga; A.maxrec = maxrec; A.s2c = s2c;
    A.rowE = ch.rowE.p; A.rowF = ch.rowF.p; A.rowV = ch.rowV.p;
    A.Psi = Psi->d; A.ld = Psi->ld; A.col0 = blk.col0;
    {
        ProfScope prof_(ctx, "site_deriv_3b");
        switch (D) {
#define IPF_CASE(DD) case DD: IPF_CHECK((launch_cfg<DD>(ctx, A, (int)ch.ncfg, need))); break;
            IPF_CASE(2) IPF_CASE(3) IPF_CASE(4) IPF_CASE(5) IPF_CASE(6) IPF_CASE(7) IPF_CASE(8)
            IPF_CASE(9) IPF_CASE(10) IPF_CASE(11) IPF_CASE(12) IPF_CASE(13) IPF_CASE(14)
#undef IPF_CASE
            default: break;
        }
    }
    IPF_CUDA(ctx, cudaGetLastError());
    *handled = true;
    return IPF_OK;
}
