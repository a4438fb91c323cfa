// Measurement aid: the FP32 FMA-pipe peak of the device this library runs on, measured live (bench.py quotes the scan's
// roofline fraction against it as well as against the nominal 148 SM x 128 lanes x 2 x clock).  Not part of the hot path.
#include "xs_common.cuh"

namespace xs {

constexpr int kPeakChains = 16;

// FFMA2 (fma.rn.f32x2, the instruction the scan issues): kPeakChains independent chains per thread, operands in registers
__global__ void __launch_bounds__(256) k_peak_ffma2(float *out, int iters, float a, float b) {
    unsigned long long p[kPeakChains];
    unsigned long long pa, pb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
#pragma unroll
    for (int i = 0; i < kPeakChains; ++i) {
        const float x = a * (float)(threadIdx.x + i);
        asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(x), "f"(x + b));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kPeakChains; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kPeakChains; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
        s += lo + hi;
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace xs

extern "C" int xs_bench_fp32_peak(double *tflops, void *stream) {
    using namespace xs;
    if (!tflops) {
        set_error("xs_bench_fp32_peak: null argument");
        return XS_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 8, threads = 256, iters = 40000;
    float *out = nullptr;
    XS_CUDA(cudaMalloc(&out, sizeof(float) * (size_t)grid * threads));
    cudaEvent_t e0, e1;
    XS_CUDA(cudaEventCreate(&e0));
    XS_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {  // the first repetition warms the clocks up
        XS_CUDA(cudaEventRecord(e0, st));
        XS_LAUNCH(k_peak_ffma2, grid, threads, 0, st, out, iters, 1.0001f, 0.5f);
        XS_CUDA(cudaEventRecord(e1, st));
        XS_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        XS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 2.0 * kPeakChains * (double)iters * threads * grid;  // 2 lanes x (mul + add)
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return XS_OK;
}
