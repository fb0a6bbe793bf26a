// Measured fp64 peak of the device the library runs on: dependent-free DFMA chains, enough
// warps to fill every scheduler (tools/microbench.cu is the stand-alone version with the
// DMMA / SHFL / LDS numbers).  bench.py divides the kernels' algorithmic flops by this figure
// instead of a datasheet number (MEASURED_PEAKS.json has no fp64 entry).
#include "common.cuh"

namespace pb200 {

constexpr int PEAK_ITERS = 4096;

__global__ void dfma_peak_kernel(double* out, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < PEAK_ITERS; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_measure_dfma_tflops(double* scratch, size_t scratch_doubles, double* tflops, void* stream) {
    if (!scratch || !tflops) return fail(PB200_EINVAL, "pb200_measure_dfma_tflops: null pointer");
    const int sms = sm_count();
    if (!sms) return PB200_ECUDA;
    const int threads = 512, blocks = sms * 2;              // 32 warps per SM
    if (scratch_doubles < (size_t)threads * blocks) return fail(PB200_EWORKSPACE, "pb200_measure_dfma_tflops: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    PB_CHECK_CUDA(cudaEventCreate(&e0));
    PB_CHECK_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {                     // first repetition warms up
        PB_CHECK_CUDA(cudaEventRecord(e0, st));
        dfma_peak_kernel<<<blocks, threads, 0, st>>>(scratch, 1.0000001, 1e-9);
        PB_CHECK_CUDA(cudaEventRecord(e1, st));
        PB_CHECK_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        PB_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    count_launch(4);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = (double)threads * blocks * PEAK_ITERS * 8 * 2 / (best * 1e-3) / 1e12;
    return PB200_OK;
}
