// peaks.cu -- measures the FP64 roofline denominators MEASURED_PEAKS.json lacks:
// FP64 vector FMA peak (DFMA chains) and FP64 tensor peak (DMMA m8n8k4 chains).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peaks tools/peaks.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void dfma_chain(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void dmma_chain(double* out, int iters, double a, double b) {
    double c[8][2];
    for (int k = 0; k < 8; ++k) { c[k][0] = threadIdx.x + k; c[k][1] = k; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out;
    cudaMalloc(&out, sizeof(double) * 148 * 16 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {256, 512, 1024}) {
        for (int bps : {1, 2}) {
            if (threads * bps > 2048) continue;
            int grid = sms * bps;
            const int iters = 1 << 16;
            float best = 1e30f;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(e0);
                dfma_chain<<<grid, threads>>>(out, iters, 1.0000001, 1e-9);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            double fl = 2.0 * 8 * iters * (double)grid * threads;
            printf("{\"kind\":\"dfma\",\"threads\":%d,\"blocks_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n", threads, bps, best, fl / best * 1e-9);
            best = 1e30f;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(e0);
                dmma_chain<<<grid, threads>>>(out, iters / 4, 1.0000001, 1e-9);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            fl = 2.0 * 8 * 8 * 4 * 8 * (iters / 4) * (double)grid * (threads / 32);
            printf("{\"kind\":\"dmma_m8n8k4\",\"threads\":%d,\"blocks_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n", threads, bps, best, fl / best * 1e-9);
        }
    }
    cudaError_t e = cudaGetLastError();
    printf("{\"sms\":%d,\"status\":\"%s\"}\n", sms, cudaGetErrorString(e));
    return 0;
}
