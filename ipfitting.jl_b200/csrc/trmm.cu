// trmm.cu -- K6: the re-orthogonalisation update  A <- A L^-T  of a CholeskyQR pass on the FP64 tensor cores.
//
// The triangular solve with M (~4e5) right-hand-side rows is done as a product with the explicit inverse:
//   W = [ L^-T  0 ]      (n1 x n1, upper triangular; the weighted-y column n passes through)
//       [  0    1 ]
//   trinv_kernel : one warp per column of L^-1 (forward substitution; blocks of 32 rows, the in-block
//                  recurrence through shuffles, reciprocal diagonal precomputed)
//   applyw_kernel: A[m0:m0+128, :] <- A[m0:m0+128, :] W  with mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), one CTA per
//                  128 rows, IN PLACE: the 128-column output tiles are produced right to left, tile tj needs the
//                  input columns k < 128 (tj + 1) only (W is upper triangular), which are still unmodified.
// Replaces the thread-per-row substitution (trsm_rows_kernel, FP64 FMA pipe at 25 %): same flops, but on the tensor
// pipe and with every row of A read twice instead of four times.
// Part of the `qr!` replacement (src/lsq.jl:468), see solve.cu.
#include "ipf_common.cuh"

namespace {

constexpr int BT = 128;          // tile edge
constexpr int BK = 16;           // k per stage
constexpr int STAGES = 4;
constexpr int AW_THREADS = 256;
constexpr int LDA_S = BT + 4;    // shared-memory row stride of the A tile ([k][m], m contiguous): conflict-free fragments
constexpr int ATILE_DOUBLES = BK * LDA_S;
constexpr int WTILE_DOUBLES = BT * BK;
constexpr int STAGE_DOUBLES = ATILE_DOUBLES + WTILE_DOUBLES;
constexpr size_t AW_SMEM = sizeof(double) * STAGE_DOUBLES * STAGES;

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

// W tile: 128 columns j x 16 k, k contiguous in global (column-major W), 32-byte groups XOR-swizzled (as gram.cu)
__device__ __forceinline__ void load_w_tile(double* __restrict__ dst, const double* __restrict__ W, int64_t ldw,
                                            int col0, int k0, int tid) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int q = tid + AW_THREADS * u;
        const int col = q >> 3, c16 = q & 7;
        const int phys = ((((c16 >> 1) ^ (col & 3)) << 1) | (c16 & 1));
        cp_async16(dst + col * BK + phys * 2, W + (int64_t)(col0 + col) * ldw + k0 + 2 * c16);
    }
}
// A tile: 16 columns k x 128 rows m, m contiguous in global
__device__ __forceinline__ void load_a_tile(double* __restrict__ dst, const double* __restrict__ A, int64_t ld,
                                            int64_t m0, int64_t Mpad, int k0, int tid) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int q = tid + AW_THREADS * u;
        const int kk = q >> 6, c16 = q & 63;
        double* d = dst + kk * LDA_S + 2 * c16;
        if (m0 + 2 * c16 < Mpad) cp_async16(d, A + (int64_t)(k0 + kk) * ld + m0 + 2 * c16);
        else *reinterpret_cast<double2*>(d) = make_double2(0.0, 0.0);
    }
}

__global__ void __launch_bounds__(AW_THREADS, 1)
applyw_kernel(double* __restrict__ A, int64_t ld, int64_t Mpad, int n1, int kpad, const double* __restrict__ W) {
    extern __shared__ __align__(128) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t m0 = (int64_t)blockIdx.x * BT;
    const int wr = warp >> 2, wc = warp & 3;     // warp tile: rows 64*wr.., cols 32*wc..
    const int fi = lane >> 2, ft = lane & 3;
    const int nt = (n1 + BT - 1) / BT;
    for (int tj = nt - 1; tj >= 0; --tj) {
        const int nk = min(kpad, BT * (tj + 1)) / BK;      // W[k][j] = 0 for k > j
        double acc[8][4][2];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < nk) {
                double* st = smem + s * STAGE_DOUBLES;
                load_a_tile(st, A, ld, m0, Mpad, s * BK, tid);
                load_w_tile(st + ATILE_DOUBLES, W, kpad, tj * BT, s * BK, tid);
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
                    load_a_tile(st, A, ld, m0, Mpad, kn * BK, tid);
                    load_w_tile(st + ATILE_DOUBLES, W, kpad, tj * BT, kn * BK, tid);
                }
                cp_async_commit();
            }
            const double* sA = smem + (k % STAGES) * STAGE_DOUBLES + wr * 64;
            const double* sW = smem + (k % STAGES) * STAGE_DOUBLES + ATILE_DOUBLES + (wc * 32) * BK;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int off = ((s ^ (fi & 3)) << 2) + ft;
                double af[8], bf[4];
#pragma unroll
                for (int a = 0; a < 8; ++a) af[a] = sA[(4 * s + ft) * LDA_S + a * 8 + fi];
#pragma unroll
                for (int b = 0; b < 4; ++b) bf[b] = sW[(b * 8 + fi) * BK + off];
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
            }
        }
        cp_async_wait<0>();
        __syncthreads();          // every warp has read its last stage: the tile's columns may be overwritten
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int64_t m = m0 + wr * 64 + a * 8 + fi;
            if (m >= Mpad) continue;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int j = tj * BT + wc * 32 + b * 8 + ft * 2;
                if (j < n1) A[(int64_t)j * ld + m] = acc[a][b][0];
                if (j + 1 < n1) A[(int64_t)(j + 1) * ld + m] = acc[a][b][1];
            }
        }
    }
}

// W (kpad x ncp, column-major, zero filled by the caller): W[k][j] = (L^-1)[j][k] for k <= j < n, W[n][n] = 1.
// One warp per column c of X = L^-1 (X[i][c], i >= c); L is column-major n x n (lower).
constexpr int TI_WARPS = 8;
__global__ void __launch_bounds__(32 * TI_WARPS) trinv_kernel(const double* __restrict__ L, int n, double* __restrict__ W,
                                                              int kpad) {
    extern __shared__ double ti_smem[];
    double* rdiag = ti_smem;                               // [n] reciprocal diagonal
    double* xs = ti_smem + n + (threadIdx.x >> 5) * n;     // [n] this warp's column
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < n; i += blockDim.x) rdiag[i] = 1.0 / L[(size_t)i * n + i];
    __syncthreads();
    const int c = blockIdx.x * TI_WARPS + warp;
    if (c > n) return;
    if (c == n) { if (lane == 0) W[(size_t)n * kpad + n] = 1.0; return; }
    for (int b0 = (c / 32) * 32; b0 < n; b0 += 32) {
        const int i = b0 + lane;
        const bool live = i < n && i >= c;
        double s = (i == c) ? 1.0 : 0.0;
        if (live)
            for (int k = c; k < b0; ++k) s = fma(-L[(size_t)k * n + i], xs[k], s);
        // in-block recurrence: x_t = s_t / L[t][t];  s_i -= L[i][t] x_t  (i > t).  Lane i preloads row i of the
        // diagonal block (32 coalesced loads in flight at once); the recurrence runs out of registers.
        double d[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) d[t] = (i < n && b0 + t < n) ? L[(size_t)(b0 + t) * n + i] : 0.0;
        const double rdi = rdiag[min(i, n - 1)];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
            const int it = b0 + t;
            const double xt = __shfl_sync(0xffffffffu, s * rdi, t);
            if (it >= n || it < c) continue;      // uniform
            if (lane == t) s = xt;
            else if (live && i > it) s = fma(-d[t], xt, s);
        }
        if (live) {
            xs[i] = s;
            W[(size_t)i * kpad + c] = s;        // W[k = c][j = i] = X[i][c]
        }
        __syncwarp();
    }
}

}  // namespace

// A (Mpad x ncp column-major, ld) <- A W with W = blockdiag(L^-T, 1);  L: n x n column-major lower.
// Wsave (n x n column-major, may be NULL) receives the L^-T block that was actually applied.
int ipf_apply_inverse_transpose(ipf_ctx* ctx, double* A, int64_t ld, int64_t Mpad, int ncp, int n, const double* L,
                                double* Wsave) {
    cudaStream_t st = ctx->stream;
    const int n1 = n + 1;
    const int kpad = (n1 + BK - 1) / BK * BK;
    if (ncp % BT != 0 || n1 > ncp || kpad > ncp) return ipf_set_error(ctx, IPF_EINVAL, "ipf_apply_inverse_transpose: bad padding");
    void* W = nullptr;
    IPF_CHECK(ipf_scratch(ctx, 2, sizeof(double) * (size_t)kpad * ncp, &W));
    IPF_CUDA(ctx, cudaMemsetAsync(W, 0, sizeof(double) * (size_t)kpad * ncp, st));
    const size_t ti_bytes = sizeof(double) * (size_t)n * (TI_WARPS + 1);
    if (ti_bytes > (size_t)227 * 1024)
        return ipf_set_error(ctx, IPF_EINVAL, "ipf_apply_inverse_transpose: n = %d exceeds the shared-memory limit of trinv_kernel", n);
    if (ti_bytes > 48 * 1024)
        IPF_CUDA(ctx, cudaFuncSetAttribute(trinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ti_bytes));
    trinv_kernel<<<(n1 + TI_WARPS - 1) / TI_WARPS, 32 * TI_WARPS, ti_bytes, st>>>(L, n, (double*)W, kpad);
    IPF_CUDA(ctx, cudaGetLastError());
    if (Wsave)
        IPF_CUDA(ctx, cudaMemcpy2DAsync(Wsave, sizeof(double) * n, W, sizeof(double) * kpad, sizeof(double) * n, n,
                                        cudaMemcpyDeviceToDevice, st));
    IPF_CUDA(ctx, cudaFuncSetAttribute(applyw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AW_SMEM));
    applyw_kernel<<<(unsigned)((Mpad + BT - 1) / BT), AW_THREADS, AW_SMEM, st>>>(A, ld, Mpad, n1, kpad, (const double*)W);
    ctx->launches += 2;
    IPF_CUDA(ctx, cudaGetLastError());
    return IPF_OK;
}
