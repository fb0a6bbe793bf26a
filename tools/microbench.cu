// Throughput microbenchmarks that decide the LP kernel's design:
// DFMA vs DMMA (mma.sync m8n8k4 f64) vs 64-bit SHFL vs LDS.64 per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__global__ void k_dfma(double* out, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0+4, x5=x0+5, x6=x0+6, x7=x0+7;
    for (int i = 0; i < ITERS; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void k_dmma(double* out, double a, double b) {
    double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
    double av = a + threadIdx.x, bv = b;
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[j]), "+d"(c1[j]) : "d"(av), "d"(bv));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c1[0] + c0[1] + c1[1] + c0[2] + c1[2] + c0[3] + c1[3];
}
__global__ void k_shfl(double* out, double a) {
    double x0 = a + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    for (int i = 0; i < ITERS; ++i) {
        x0 += __shfl_xor_sync(0xffffffffu, x0, 1); x1 += __shfl_xor_sync(0xffffffffu, x1, 2);
        x2 += __shfl_xor_sync(0xffffffffu, x2, 4); x3 += __shfl_xor_sync(0xffffffffu, x3, 8);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}
__global__ void k_lds(double* out, int stride) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double x0 = 0, x1 = 0, x2 = 0, x3 = 0;
    int base = (threadIdx.x & 31) * stride;
    for (int i = 0; i < ITERS; ++i) {
        x0 += sm[(base + i) & 4095]; x1 += sm[(base + i + 64) & 4095];
        x2 += sm[(base + i + 128) & 4095]; x3 += sm[(base + i + 192) & 4095];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}
__global__ void k_rsqrt(double* out, double a) {
    double x0 = a + threadIdx.x + 1, x1 = x0 + 1;
    for (int i = 0; i < ITERS / 8; ++i) { x0 = rsqrt(x0) + 2.0; x1 = 1.0 / x1 + 2.0; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; printf("%s SMs=%d clock=%d kHz smem/SM=%zu\n", p.name, sms, p.clockRate, p.sharedMemPerMultiprocessor);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    for (int wps : {4, 8, 16, 32}) {
        int threads = 32 * wps; dim3 g(sms * 2), b(threads / 2 < 32 ? 32 : threads / 2);
        double nthr = (double)g.x * b.x;
        float ms = timeit([&] { k_dfma<<<g, b>>>(out, 1.0000001, 1e-9); });
        printf("warps/SM=%2d DFMA  %8.2f TFLOP/s  (%.2f warp-instr/clk/SM @1.9GHz)\n", wps, nthr * ITERS * 8 * 2 / ms / 1e9, nthr / 32 * ITERS * 8 / (ms * 1e-3) / sms / 1.9e9);
        ms = timeit([&] { k_dmma<<<g, b>>>(out, 1.0000001, 1e-9); });
        printf("warps/SM=%2d DMMA  %8.2f TFLOP/s  (%.3f mma/clk/SM)\n", wps, nthr / 32 * ITERS * 4 * 512 / ms / 1e9, nthr / 32 * ITERS * 4 / (ms * 1e-3) / sms / 1.9e9);
        ms = timeit([&] { k_shfl<<<g, b>>>(out, 1.0); });
        printf("warps/SM=%2d SHFL64 %7.3f shfl64/clk/SM\n", wps, nthr / 32 * ITERS * 4 / (ms * 1e-3) / sms / 1.9e9);
        for (int stride : {1, 9, 16}) {
            ms = timeit([&] { k_lds<<<g, b, 4096 * 8>>>(out, stride); });
            printf("warps/SM=%2d LDS64 stride %2d %7.3f lds/clk/SM\n", wps, stride, nthr / 32 * ITERS * 4 / (ms * 1e-3) / sms / 1.9e9);
        }
        ms = timeit([&] { k_rsqrt<<<g, b>>>(out, 1.0); });
        printf("warps/SM=%2d rsqrt+div pair %7.4f pairs/clk/SM\n", wps, nthr / 32 * (ITERS / 8) / (ms * 1e-3) / sms / 1.9e9);
    }
    return 0;
}
