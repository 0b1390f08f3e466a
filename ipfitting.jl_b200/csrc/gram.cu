// gram.cu -- K5: G = A^T A on the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64).
//
// This is the dense contraction behind the QR of the reference
// (`qr!(Psi)`, src/lsq.jl:468): the CholeskyQR passes in solve.cu factor the
// Gram matrix of the weighted, column-scaled system [A | y].
//
// A is column-major M x ncp (ld multiple of 16, ncp multiple of 128, zero padded), so both
// MMA operands are K-contiguous (K = row index of A):  G[i][j] = sum_k A[k][i] A[k][j]
//   A-operand (.row): rows = columns i of A, K contiguous  -> tile [128 cols][16 rows]
//   B-operand (.col): cols = columns j of A, K contiguous  -> same shape
// Only the upper-triangular 128x128 tiles are computed; the rows are split into
// `nchunk` chunks (split-K) whose partial tiles are summed in a fixed order by
// gram_reduce_kernel (deterministic, no atomics).  CTAs of one chunk run
// concurrently, so every 16x128 operand panel is read from HBM once and from L2
// by the other tiles that need it.
//
// Shared-memory layout: a tile row is 16 doubles = 128 B; the 32-byte group g of
// row r is stored at group g ^ (r & 3), which makes every m8n8k4 fragment load
// (8 rows x 4 consecutive k) hit 16 distinct bank pairs per half-warp.
#include "ipf_common.cuh"

namespace {

constexpr int BT = 128;          // tile edge (columns of A)
constexpr int BK = 16;           // rows of A per stage
constexpr int STAGES = 4;
constexpr int GRAM_THREADS = 256;
constexpr int TILE_DOUBLES = BT * BK;                 // 2048 doubles = 16 KB
constexpr int STAGE_DOUBLES = 2 * TILE_DOUBLES;       // A and B operand
constexpr size_t GRAM_SMEM = sizeof(double) * STAGE_DOUBLES * STAGES;   // 128 KB

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// copy one operand tile (128 columns x 16 rows) into swizzled shared memory
__device__ __forceinline__ void load_tile(double* __restrict__ dst, const double* __restrict__ A, int64_t ld,
                                          int col0, int64_t k0, int tid) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int q = tid + GRAM_THREADS * u;      // 0..1023: 16-byte chunk id
        const int col = q >> 3, c16 = q & 7;
        const int phys = ((((c16 >> 1) ^ (col & 3)) << 1) | (c16 & 1));
        cp_async16(dst + col * BK + phys * 2, A + (int64_t)(col0 + col) * ld + k0 + 2 * c16);
    }
}

// One tile of the Gram matrix for one chunk of rows.  The 8 warps are laid out 2 x 4 over a (16 NA) x (32 NB) sub-tile
// (NA = 8, NB = 4: the full 128 x 128 tile); the last block row / column of G only needs the sub-tile that covers
// n1 - 128 (nt - 1) columns, which saves the MMAs on the zero padding (n1 = 217: 23 % of all MMAs).
template <int NA, int NB>
__device__ __forceinline__ void gram_tile(const double* __restrict__ A, int64_t ld, int64_t Mpad, int nchunk,
                                          int64_t chunk_rows, int ti, int tj, int tile, double* __restrict__ part,
                                          double* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.y;
    const int64_t kbeg = (int64_t)chunk * chunk_rows;
    const int64_t kend = min(Mpad, kbeg + chunk_rows);
    const int nk = kend > kbeg ? (int)((kend - kbeg) / BK) : 0;

    const int wr = warp >> 2, wc = warp & 3;     // warp tile: rows 8 NA wr.., cols 8 NB wc..
    const int fi = lane >> 2, ft = lane & 3;
    double acc[NA][NB][2];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }

    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) {
            double* st = smem + s * STAGE_DOUBLES;
            load_tile(st, A, ld, ti * BT, kbeg + (int64_t)s * BK, tid);
            load_tile(st + TILE_DOUBLES, A, ld, tj * BT, kbeg + (int64_t)s * BK, tid);
        }
        cp_async_commit();
    }
    for (int k = 0; k < nk; ++k) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int kn = k + STAGES - 1;
            if (kn < nk) {
                double* st = smem + (kn % STAGES) * STAGE_DOUBLES;
                load_tile(st, A, ld, ti * BT, kbeg + (int64_t)kn * BK, tid);
                load_tile(st + TILE_DOUBLES, A, ld, tj * BT, kbeg + (int64_t)kn * BK, tid);
            }
            cp_async_commit();
        }
        const double* sA = smem + (k % STAGES) * STAGE_DOUBLES + (wr * 8 * NA) * BK;
        const double* sB = smem + (k % STAGES) * STAGE_DOUBLES + TILE_DOUBLES + (wc * 8 * NB) * BK;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            // logical k = 4s + ft -> 32-byte group s, slot ft; physical group s ^ (row & 3)
            const int off = ((s ^ (fi & 3)) << 2) + ft;
            double af[NA], bf[NB];
#pragma unroll
            for (int a = 0; a < NA; ++a) af[a] = sA[(a * 8 + fi) * BK + off];
#pragma unroll
            for (int b = 0; b < NB; ++b) bf[b] = sB[(b * 8 + fi) * BK + off];
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
    }
    cp_async_wait<0>();
    // partial tile: part[chunk][tile][i][j], i = row in tile (column of A in block ti)
    double* P = part + ((size_t)chunk * gridDim.x + tile) * (BT * BT);
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int i = wr * 8 * NA + a * 8 + fi;
            const int j = wc * 8 * NB + b * 8 + ft * 2;
            *reinterpret_cast<double2*>(P + i * BT + j) = make_double2(acc[a][b][0], acc[a][b][1]);
        }
}

// tiles[t] = (block row ti, block column tj, NA, NB)
__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_kernel(const double* __restrict__ A, int64_t ld, int64_t Mpad, int nt, int nchunk, int64_t chunk_rows,
            const int4* __restrict__ tiles, double* __restrict__ part) {
    extern __shared__ __align__(128) double smem[];
    const int tile = blockIdx.x;
    const int4 T = tiles[tile];
#define IPF_GRAM_CASE(NA_, NB_)                                                                             \
    if (T.z == NA_ && T.w == NB_) { gram_tile<NA_, NB_>(A, ld, Mpad, nchunk, chunk_rows, T.x, T.y, tile, part, smem); return; }
    IPF_GRAM_CASE(8, 4) IPF_GRAM_CASE(8, 3) IPF_GRAM_CASE(8, 2) IPF_GRAM_CASE(8, 1)
    IPF_GRAM_CASE(6, 3) IPF_GRAM_CASE(4, 2) IPF_GRAM_CASE(2, 1)
#undef IPF_GRAM_CASE
}

// G[i][j] (n1 x n1, ld n1, both triangles) = sum over chunks, fixed order
__global__ void gram_reduce_kernel(const double* __restrict__ part, int ntiles, int nchunk,
                                   const int4* __restrict__ tiles, double* __restrict__ G, int n1) {
    const int tile = blockIdx.x;
    const int ti = tiles[tile].x, tj = tiles[tile].y;
    // blockIdx.y selects a band of BT / gridDim.y rows of the tile
    const int band = BT * BT / gridDim.y;
    for (int e = blockIdx.y * band + threadIdx.x; e < (blockIdx.y + 1) * band; e += blockDim.x) {
        const int i = e / BT, j = e % BT;
        const int gi = ti * BT + i, gj = tj * BT + j;
        if (gi >= n1 || gj >= n1) continue;
        if (ti == tj && gj < gi) continue;
        // four interleaved partial sums (fixed order): the loads of four chunks are in flight together
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        const double* pe = part + (size_t)tile * (BT * BT) + e;
        const size_t cs = (size_t)ntiles * (BT * BT);
        int c = 0;
        for (; c + 3 < nchunk; c += 4) {
            s0 += pe[(size_t)c * cs]; s1 += pe[(size_t)(c + 1) * cs]; s2 += pe[(size_t)(c + 2) * cs]; s3 += pe[(size_t)(c + 3) * cs];
        }
        for (; c < nchunk; ++c) s0 += pe[(size_t)c * cs];
        const double s = (s0 + s1) + (s2 + s3);
        G[(size_t)gj * n1 + gi] = s;
        G[(size_t)gi * n1 + gj] = s;
    }
}

}  // namespace

// A: Mpad x ncp column-major (ld), zero padded; G: n1 x n1 (n1 <= ncp), column-major.
int ipf_gram(ipf_ctx* ctx, const double* A, int64_t ld, int64_t Mpad, int ncp, int n1, double* G,
             double* kernel_ms) {
    cudaStream_t st = ctx->stream;
    if (ncp % BT != 0 || ld % BK != 0 || Mpad % BK != 0 || n1 > ncp)
        return ipf_set_error(ctx, IPF_EINVAL, "ipf_gram: bad padding");
    const int nt = (n1 + BT - 1) / BT;
    // last block row / column: only the sub-tile that covers the remaining columns (in steps of 32)
    const int rem = n1 - BT * (nt - 1);
    const int q = (rem + 31) / 32;                      // 1..4
    std::vector<int4> h_tiles;
    for (int i = 0; i < nt; ++i)
        for (int j = i; j < nt; ++j) h_tiles.push_back(make_int4(i, j, i == nt - 1 ? 2 * q : 8, j == nt - 1 ? q : 4));
    const int T = (int)h_tiles.size();
    // split-K: fill whole waves of SMs (one CTA per SM)
    int nchunk = std::max(1, ctx->sm_count / T);
    const int64_t ksteps = Mpad / BK;
    if (nchunk > ksteps) nchunk = (int)std::max<int64_t>(1, ksteps);
    if (ksteps >= 64LL * nchunk) {
        // several waves when there is plenty of work: keeps the tail short
        int waves = (int)std::min<int64_t>(4, ksteps / (64LL * nchunk));
        nchunk = std::max(1, (ctx->sm_count * waves) / T);
    }
    const int64_t chunk_rows = ((ksteps + nchunk - 1) / nchunk) * BK;
    void *d_tiles = nullptr, *part = nullptr;
    IPF_CHECK(ipf_scratch(ctx, 6, sizeof(int4) * T, &d_tiles));
    IPF_CHECK(ipf_scratch(ctx, 7, sizeof(double) * (size_t)nchunk * T * BT * BT, &part));
    // the tile table stays on the device between calls (it depends on n1 only): no upload and no host
    // synchronisation in the steady state
    if (ctx->gram_tiles_n1 != n1 || ctx->gram_tiles_ptr != d_tiles) {
        IPF_CUDA(ctx, cudaMemcpyAsync(d_tiles, h_tiles.data(), sizeof(int4) * T, cudaMemcpyHostToDevice, st));
        IPF_CUDA(ctx, cudaStreamSynchronize(st));     // h_tiles must outlive the copy
        ctx->gram_tiles_n1 = n1; ctx->gram_tiles_ptr = d_tiles;
    }
    IPF_CUDA(ctx, cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRAM_SMEM));
    (void)kernel_ms;
    {
        ProfScope prof_(ctx, "gram");
        gram_kernel<<<dim3(T, nchunk), GRAM_THREADS, GRAM_SMEM, st>>>(A, ld, Mpad, nt, nchunk, chunk_rows,
                                                                      (const int4*)d_tiles, (double*)part);
    }
    gram_reduce_kernel<<<dim3(T, 64), 256, 0, st>>>((const double*)part, T, nchunk, (const int4*)d_tiles, G, n1);
    ctx->launches += 2;
    IPF_CUDA(ctx, cudaGetLastError());
    return IPF_OK;
}
