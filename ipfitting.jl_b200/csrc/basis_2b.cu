// basis_2b.cu -- K2 fast path: canonical pair basis nbpolys(2, desc, D):  V_b(i) = sum_j fc(r_ij) x_ij^b, b = 1..D.
//
// For a pure pair function the reverse-pair derivative is the exact negative of the own one
// (same r, opposite bond vector), so   F_b[a] = sum_{p in out(a)} (g_p - g_rev(p)) = 2 sum_p g_p
// needs no scratch and no reverse lookup:  g_p = (fc' x^b + fc b x^(b-1) x') Rhat_p.
// pair2b_fused_kernel (nb <= 16): one CTA per configuration, one warp per centre atom.  Per chunk of 32 pairs
// the lanes first evaluate the descriptor (x, x', fc, fc', x^0..x^D) of one pair each into a shared-memory
// tile -- no pair records in HBM for this block --, then lane = (function b, pair parity) accumulates
// F (per atom), E and the virial (per configuration, registers) over the chunk; fixed order, no atomics.
// pair2b_kernel (any nb): thread = (basis function, atom) over the pair records in HBM.
// Replaces energy/forces/virial(B2, at) + vec_obs + block placement (src/datatypes.jl:28,40,65;
// src/lsq_db.jl:268-278) for the 2-body block.
#include <cstdlib>

#include "ipf_common.cuh"

int ipf_pair_records(ipf_ctx* ctx, const BlockDev& blk, const NeighList& nl, const double** rec, int* rs);

namespace {

constexpr int REC = 8;
constexpr int P2_THREADS = 256;
constexpr int P2_ACH = 32;       // atoms per pass

struct Pair2Args {
    int nb, rs;
    const double* rec;
    const int64_t* atom_off;
    const int64_t* site_off;
    const double* pR;
    const int64_t* rowE;
    const int64_t* rowF;
    const int64_t* rowV;
    double* Psi;
    int64_t ld, col0;
};

__global__ void __launch_bounds__(P2_THREADS) pair2b_kernel(Pair2Args A) {
    __shared__ double part[P2_THREADS / P2_ACH][P2_ACH][7];
    const int c = blockIdx.x;
    const int64_t a0 = A.atom_off[c];
    const int nat = (int)(A.atom_off[c + 1] - a0);
    const int64_t rE = A.rowE[c], rF = A.rowF[c], rV = A.rowV[c];
    const int tid = threadIdx.x;
    const int al = tid % P2_ACH, bl = tid / P2_ACH;         // lanes run over atoms
    constexpr int BPP = P2_THREADS / P2_ACH;                // basis functions per pass
    for (int b0 = 0; b0 < A.nb; b0 += BPP) {
        const int b = b0 + bl;                              // local function index; exponent = b + 1
        const bool blive = b < A.nb;
        double tot = 0.0;                                   // threads < BPP*7 own one (function, component) total
        for (int ab = 0; ab < nat; ab += P2_ACH) {
            const int at = ab + al;
            double f0 = 0.0, f1 = 0.0, f2 = 0.0, e = 0.0, v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0, v4 = 0.0, v5 = 0.0;
            if (at < nat && blive) {
                const int64_t s = a0 + at;
                const int64_t pe = A.site_off[s + 1];
#pragma unroll 4
                for (int64_t p = A.site_off[s]; p < pe; ++p) {
                    const double* __restrict__ r = A.rec + (int64_t)A.rs * p;
                    const double xb = r[REC + b + 1], xbm1 = r[REC + b];
                    const double radial = r[7] * xb + r[6] * (double)(b + 1) * xbm1 * r[5];
                    const double g0 = radial * r[0], g1 = radial * r[1], g2 = radial * r[2];
                    f0 += g0; f1 += g1; f2 += g2;
                    e += r[6] * xb;
                    const double R0 = A.pR[3 * p], R1 = A.pR[3 * p + 1], R2 = A.pR[3 * p + 2];
                    v0 -= g0 * R0; v1 -= g1 * R1; v2 -= g2 * R2;      // xx, yy, zz
                    v3 -= g2 * R1; v4 -= g2 * R0; v5 -= g1 * R0;      // zy, zx, yx
                }
                if (rF) {
                    double* o = A.Psi + (size_t)(A.col0 + b) * A.ld + (rF - 1) + 3 * at;
                    o[0] = 2.0 * f0; o[1] = 2.0 * f1; o[2] = 2.0 * f2;
                }
            }
            part[bl][al][0] = e; part[bl][al][1] = v0; part[bl][al][2] = v1; part[bl][al][3] = v2;
            part[bl][al][4] = v3; part[bl][al][5] = v4; part[bl][al][6] = v5;
            __syncthreads();
            if (tid < BPP * 7) {
                const int b2 = tid / 7, comp = tid % 7;
                const int lim = min(P2_ACH, nat - ab);
                for (int q = 0; q < lim; ++q) tot += part[b2][q][comp];
            }
            __syncthreads();
        }
        if (tid < BPP * 7) {
            const int b2 = tid / 7, comp = tid % 7;
            if (b0 + b2 < A.nb) {
                double* colp = A.Psi + (size_t)(A.col0 + b0 + b2) * A.ld;
                if (comp == 0) { if (rE) colp[rE - 1] = tot; }
                else if (rV) colp[rV - 1 + comp - 1] = tot;
            }
        }
    }
}

// ---- fused variant ---------------------------------------------------------------------------------
constexpr int F2_THREADS = 128;
constexpr int F2_WARPS = F2_THREADS / 32;
constexpr int F2_MAXNB = 16;
constexpr int F2_HEAD = 10;                            // Rhat(3), R(3), fc, dfc, fc * dx, pad
constexpr int F2_LD = F2_HEAD + F2_MAXNB + 2;          // + x^0..x^16, pad: 28 doubles per staged pair

struct Fused2Args {
    int nb, transform;
    double a, r0, rc1, rc2;
    const int64_t* atom_off;
    const int64_t* site_off;
    const double* pR;
    const int64_t* rowE;
    const int64_t* rowF;
    const int64_t* rowV;
    double* Psi;
    int64_t ld, col0;
};

__global__ void __launch_bounds__(F2_THREADS) pair2b_fused_kernel(Fused2Args A) {
    __shared__ __align__(16) double tile[F2_WARPS][32][F2_LD];
    __shared__ double part[F2_WARPS][F2_MAXNB][7];
    const int c = blockIdx.x;
    const int64_t a0 = A.atom_off[c];
    const int nat = (int)(A.atom_off[c + 1] - a0);
    const int64_t rE = A.rowE[c], rF = A.rowF[c], rV = A.rowV[c];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = lane & 15, h = lane >> 4;            // function (exponent b + 1), pair parity
    const bool blive = b < A.nb;
    const double bp1 = (double)(b + 1);
    double (*T)[F2_LD] = tile[warp];
    double e = 0.0, v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0, v4 = 0.0, v5 = 0.0;
    for (int at = warp; at < nat; at += F2_WARPS) {
        const int64_t q0 = A.site_off[a0 + at], q1 = A.site_off[a0 + at + 1];
        double f0 = 0.0, f1 = 0.0, f2 = 0.0;
        for (int64_t qb = q0; qb < q1; qb += 32) {
            const int cnt = (int)min((int64_t)32, q1 - qb);
            __syncwarp();
            if (lane < cnt) {      // descriptor of pair qb + lane
                const int64_t p = qb + lane;
                const double Rx = A.pR[3 * p], Ry = A.pR[3 * p + 1], Rz = A.pR[3 * p + 2];
                const double r = ipf_bond_length(Rx, Ry, Rz);
                double x, dx, fc, dfc;
                ipf_descriptor(r, A.transform, A.a, A.r0, A.rc1, A.rc2, x, dx, fc, dfc);
                double* o = T[lane];
                o[0] = Rx / r; o[1] = Ry / r; o[2] = Rz / r;
                o[3] = Rx; o[4] = Ry; o[5] = Rz;
                o[6] = fc; o[7] = dfc; o[8] = fc * dx;
                double pw = 1.0;
#pragma unroll
                for (int k = 0; k <= F2_MAXNB; ++k) { o[F2_HEAD + k] = pw; pw *= x; }
            }
            __syncwarp();
            if (blive) {
                for (int k = h; k < cnt; k += 2) {
                    const double* __restrict__ o = T[k];
                    const double xb = o[F2_HEAD + b + 1], xbm1 = o[F2_HEAD + b];
                    const double radial = fma(o[7], xb, o[8] * bp1 * xbm1);
                    const double g0 = radial * o[0], g1 = radial * o[1], g2 = radial * o[2];
                    f0 += g0; f1 += g1; f2 += g2;
                    e = fma(o[6], xb, e);
                    v0 = fma(-g0, o[3], v0); v1 = fma(-g1, o[4], v1); v2 = fma(-g2, o[5], v2);      // xx, yy, zz
                    v3 = fma(-g2, o[4], v3); v4 = fma(-g2, o[3], v4); v5 = fma(-g1, o[3], v5);      // zy, zx, yx
                }
            }
        }
        // the two parities of every function
        f0 += __shfl_xor_sync(0xffffffffu, f0, 16);
        f1 += __shfl_xor_sync(0xffffffffu, f1, 16);
        f2 += __shfl_xor_sync(0xffffffffu, f2, 16);
        if (rF && blive && h == 0) {
            double* o = A.Psi + (size_t)(A.col0 + b) * A.ld + (rF - 1) + 3 * at;
            o[0] = 2.0 * f0; o[1] = 2.0 * f1; o[2] = 2.0 * f2;
        }
    }
    e += __shfl_xor_sync(0xffffffffu, e, 16);
    v0 += __shfl_xor_sync(0xffffffffu, v0, 16); v1 += __shfl_xor_sync(0xffffffffu, v1, 16);
    v2 += __shfl_xor_sync(0xffffffffu, v2, 16); v3 += __shfl_xor_sync(0xffffffffu, v3, 16);
    v4 += __shfl_xor_sync(0xffffffffu, v4, 16); v5 += __shfl_xor_sync(0xffffffffu, v5, 16);
    if (h == 0) {
        part[warp][b][0] = e; part[warp][b][1] = v0; part[warp][b][2] = v1; part[warp][b][3] = v2;
        part[warp][b][4] = v3; part[warp][b][5] = v4; part[warp][b][6] = v5;
    }
    __syncthreads();
    if (threadIdx.x < F2_MAXNB * 7) {
        const int b2 = threadIdx.x / 7, comp = threadIdx.x % 7;
        if (b2 < A.nb) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < F2_WARPS; ++w) tot += part[w][b2][comp];
            double* colp = A.Psi + (size_t)(A.col0 + b2) * A.ld;
            if (comp == 0) { if (rE) colp[rE - 1] = tot; }
            else if (rV) colp[rV - 1 + comp - 1] = tot;
        }
    }
}

}  // namespace

// Canonical nbpolys(2, desc, D): function b (0-based) is the single monomial x^(b+1).
int ipf_assemble_block_2b(ipf_ctx* ctx, const BlockDev& blk, const ChunkDev& ch, const NeighList& nl,
                          ipf_matrix* Psi, bool* handled) {
    *handled = false;
    if (blk.N != 2 || blk.nmono != blk.nb) return IPF_OK;
    for (int b = 0; b < blk.nb; ++b)
        if (blk.h_mono_exp[(size_t)b * IPF_MAX_VARS] != b + 1) return IPF_OK;
    *handled = true;
    if (ch.ncfg == 0) return IPF_OK;
    static const bool no_fused = getenv("IPF_2B_RECORDS") != nullptr;      // experiment switch: kernel over HBM pair records
    if (blk.nb <= F2_MAXNB && !no_fused) {
        Fused2Args F;
        F.nb = blk.nb; F.transform = blk.transform; F.a = blk.a; F.r0 = blk.r0; F.rc1 = blk.rc1; F.rc2 = blk.rc2;
        F.atom_off = ch.atom_off.p; F.site_off = nl.site_off.p; F.pR = nl.pR.p;
        F.rowE = ch.rowE.p; F.rowF = ch.rowF.p; F.rowV = ch.rowV.p; F.Psi = Psi->d; F.ld = Psi->ld; F.col0 = blk.col0;
        {
            ProfScope prof_(ctx, "site_deriv_2b");
            pair2b_fused_kernel<<<(unsigned)ch.ncfg, F2_THREADS, 0, ctx->stream>>>(F);
            ctx->launches++;
        }
        IPF_CUDA(ctx, cudaGetLastError());
        return IPF_OK;
    }
    const double* rec = nullptr;
    int rs = 0;
    IPF_CHECK(ipf_pair_records(ctx, blk, nl, &rec, &rs));
    Pair2Args A;
    A.nb = blk.nb; A.rs = rs; A.rec = rec; A.atom_off = ch.atom_off.p; A.site_off = nl.site_off.p; A.pR = nl.pR.p;
    A.rowE = ch.rowE.p; A.rowF = ch.rowF.p; A.rowV = ch.rowV.p; A.Psi = Psi->d; A.ld = Psi->ld; A.col0 = blk.col0;
    {
        ProfScope prof_(ctx, "site_deriv_2b");
        pair2b_kernel<<<(unsigned)ch.ncfg, P2_THREADS, 0, ctx->stream>>>(A);
        ctx->launches++;
    }
    IPF_CUDA(ctx, cudaGetLastError());
    return IPF_OK;
}
