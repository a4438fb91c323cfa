// FP32 CUDA-core pipe micro-benchmark for B200 (sm_100a): measures the issue/pipe ceilings the co-pol scan
// kernel is designed against (FFMA, FFMA2, FADD2, FMNMX, FMNMX3 and the scan's instruction mix).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Output: one JSON line per test: lane-ops per clock per SM (from clock64) and Gop/s (from CUDA events).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
    u64 d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float ffma(float a, float b, float c) {
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float fmin2(float a, float b) {
    float d;
    asm volatile("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ u64 pack(float x, float y) {
    return ((u64)__float_as_uint(y) << 32) | (u64)__float_as_uint(x);
}
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }

constexpr int NCH = 16;  // independent chains per thread

template <int TEST>
__global__ void bench(float *out, u64 *cycles, int iters, float a, float b) {
    float x[NCH];
    u64 p[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        x[i] = a * (threadIdx.x + i);
        p[i] = pack(x[i], x[i] + b);
    }
    const u64 pa = pack(a, a), pb = pack(b, b);
    float m0 = 1e30f, m1 = 1e30f, m2 = 1e30f, m3 = 1e30f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (TEST == 0) {  // FFMA, 3 distinct register operands
#pragma unroll
            for (int i = 0; i < NCH; ++i) x[i] = ffma(x[i], a, b);
        } else if (TEST == 1) {  // FFMA2
#pragma unroll
            for (int i = 0; i < NCH; ++i) p[i] = ffma2(p[i], pa, pb);
        } else if (TEST == 2) {  // FADD2
#pragma unroll
            for (int i = 0; i < NCH; ++i) p[i] = fadd2(p[i], pa);
        } else if (TEST == 3) {  // FMNMX (2-input)
#pragma unroll
            for (int i = 0; i < NCH; ++i) x[i] = fmin2(x[i], a);
        } else if (TEST == 4) {  // FMNMX3
#pragma unroll
            for (int i = 0; i < NCH; ++i) x[i] = fmin3(x[i], a, b);
        } else if (TEST == 5) {  // scan mix, register operands: per pair FADD2, FFMA2, FFMA2, FMNMX3
#pragma unroll
            for (int i = 0; i < NCH; i += 4) {
                u64 d0 = fadd2(p[i], pa), d1 = fadd2(p[i + 1], pa), d2 = fadd2(p[i + 2], pa), d3 = fadd2(p[i + 3], pa);
                u64 t0_ = ffma2(pb, p[i], pa), t1 = ffma2(pb, p[i + 1], pa), t2 = ffma2(pb, p[i + 2], pa),
                    t3 = ffma2(pb, p[i + 3], pa);
                u64 j0 = ffma2(d0, d0, t0_), j1 = ffma2(d1, d1, t1), j2 = ffma2(d2, d2, t2), j3 = ffma2(d3, d3, t3);
                m0 = fmin3(m0, lo(j0), hi(j0));
                m1 = fmin3(m1, lo(j1), hi(j1));
                m2 = fmin3(m2, lo(j2), hi(j2));
                m3 = fmin3(m3, lo(j3), hi(j3));
            }
        } else if (TEST == 6) {  // scan mix, scalar: per candidate FADD, FFMA, FFMA, FMNMX
#pragma unroll
            for (int i = 0; i < NCH; i += 4) {
                float d0 = x[i] + a, d1 = x[i + 1] + a, d2 = x[i + 2] + a, d3 = x[i + 3] + a;
                float t0_ = ffma(b, x[i], a), t1 = ffma(b, x[i + 1], a), t2 = ffma(b, x[i + 2], a),
                      t3 = ffma(b, x[i + 3], a);
                m0 = fmin2(m0, ffma(d0, d0, t0_));
                m1 = fmin2(m1, ffma(d1, d1, t1));
                m2 = fmin2(m2, ffma(d2, d2, t2));
                m3 = fmin2(m3, ffma(d3, d3, t3));
            }
        } else if (TEST == 7) {  // FFMA2 + FMNMX3 interleaved 3:1 (are the fma and alu pipes concurrent?)
#pragma unroll
            for (int i = 0; i < NCH; i += 4) {
                p[i] = ffma2(p[i], pa, pb);
                p[i + 1] = ffma2(p[i + 1], pa, pb);
                p[i + 2] = ffma2(p[i + 2], pa, pb);
                x[i] = fmin3(x[i], a, b);
            }
        } else if (TEST == 8) {  // FFMA with one operand reused (2 distinct sources): x = x*x + a
#pragma unroll
            for (int i = 0; i < NCH; ++i) x[i] = ffma(x[i], x[i], a);
        }
    }
    const long long t1 = clock64();
    float s = m0 + m1 + m2 + m3;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += x[i] + lo(p[i]) + hi(p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (u64)(t1 - t0);
}

struct Test {
    const char *name;
    double lane_ops_per_thread_iter;  // FP32 lane operations (an f32x2 op = 2, FMNMX3 = 1)
    double instr_per_thread_iter;
};

template <int TEST>
static void run(const Test &t, int ctas_per_sm, int threads, int iters) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * ctas_per_sm;
    float *out;
    u64 *cyc;
    cudaMalloc(&out, sizeof(float) * grid * threads);
    cudaMalloc(&cyc, sizeof(u64) * grid);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    bench<TEST><<<grid, threads>>>(out, cyc, iters / 10, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench<TEST><<<grid, threads>>>(out, cyc, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    u64 *h = (u64 *)malloc(sizeof(u64) * grid);
    cudaMemcpy(h, cyc, sizeof(u64) * grid, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < grid; ++i) mean += (double)h[i];
    mean /= grid;
    const double lane_ops_sm = t.lane_ops_per_thread_iter * iters * (double)threads * ctas_per_sm;
    const double instr_sm = t.instr_per_thread_iter * iters * (double)threads / 32.0 * ctas_per_sm;
    printf("{\"test\": \"%s\", \"ctas_per_sm\": %d, \"threads\": %d, \"lane_ops_per_clk_per_sm\": %.1f, "
           "\"warp_instr_per_clk_per_sm\": %.2f, \"Glaneops_per_s\": %.0f, \"ms\": %.3f, \"mhz_eff\": %.0f, \"err\": \"%s\"}\n",
           t.name, ctas_per_sm, threads, lane_ops_sm / mean, instr_sm / mean, lane_ops_sm * sms / (ms * 1e-3) / 1e9, ms,
           mean / (ms * 1e-3) / 1e6, cudaGetErrorString(cudaGetLastError()));
    free(h);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    const int iters = 20000;
    for (int cfg = 0; cfg < 2; ++cfg) {
        const int ctas = cfg == 0 ? 2 : 4, thr = 256;
        run<0>({"ffma_3reg", NCH, NCH}, ctas, thr, iters);
        run<8>({"ffma_2src", NCH, NCH}, ctas, thr, iters);
        run<1>({"ffma2", 2.0 * NCH, NCH}, ctas, thr, iters);
        run<2>({"fadd2", 2.0 * NCH, NCH}, ctas, thr, iters);
        run<3>({"fmnmx", NCH, NCH}, ctas, thr, iters);
        run<4>({"fmnmx3", NCH, NCH}, ctas, thr, iters);
        run<5>({"scan_mix_f32x2(lane_ops=fma-pipe only)", 6.0 * NCH, 4.0 * NCH}, ctas, thr, iters);
        run<6>({"scan_mix_scalar(lane_ops=fma-pipe only)", 3.0 * NCH, 4.0 * NCH}, ctas, thr, iters);
        run<7>({"ffma2x3+fmnmx3", 6.0 * NCH / 4 + NCH / 4, NCH}, ctas, thr, iters);
    }
    return 0;
}
