// Inner-loop experiments for the shared-sigma0 mode of k_scan_co (dev tool): per pixel and candidate pair one FFMA2 and one
// FMNMX3 -- which resource bounds the loop, and does any re-formulation of the min go faster?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/shared_microbench tools/shared_microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float x, float y) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y)); return d; }
__device__ __forceinline__ void unpack2(u64 v, float &x, float &y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int KP = 3, ROWS = 32;
// MODE 0: as the kernel.  1: two accumulators per pixel, 2-input FMNMX.  2: scalar FFMA x2 + FMNMX3.  3: FMA pipe only (packed
// add instead of the min).  4: ALU only (min of the shared M, no per-pixel FFMA2).  5: min3 of (m, J.lo, J.hi) alternating
// between two accumulators.  6: exact-k mode of the kernel (2 FFMA2 + FMNMX3) for comparison.  7: per-pixel FFMA2, min by
// integer VIMNMX3 on the bits.  8: packed min through two predicated moves (FSETP + SEL x2).
template <int MODE, int P, int NT, int MB>
__global__ void __launch_bounds__(NT, MB) k(const float *__restrict__ src, const float2 *__restrict__ rowtab, float *out, int reps) {
    __shared__ __align__(16) float ring[ROWS * 64 * KP];
    __shared__ float2 rt_s[ROWS];
    for (int i = threadIdx.x; i < ROWS * 64 * KP; i += blockDim.x) ring[i] = src[i];
    for (int i = threadIdx.x; i < ROWS; i += blockDim.x) rt_s[i] = rowtab[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    u64 g[P][KP], acc[P];
    float nqs[P], m[P], m2[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        nqs[p] = src[(p * 37 + lane) % 977];
        m[p] = m2[p] = 1e30f;
        acc[p] = 0;
#pragma unroll
        for (int j = 0; j < KP; ++j) g[p][j] = pack2(src[(p * 64 + j * 8 + lane) % 977], src[(p * 32 + j * 16 + lane + 3) % 977]);
    }
    const u64 ncs2 = pack2(src[5], src[5]);
    const u64 *rows = reinterpret_cast<const u64 *>(ring);
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
        for (int r = 0; r < ROWS; ++r) {
            const float2 rt = rt_s[r];
            const u64 nwh = pack2(rt.x, rt.x), w2q = pack2(rt.y, rt.y);
            u64 L[KP], M[KP];
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                L[j] = fadd2(rows[r * (32 * KP) + lane + 32 * j], ncs2);
                M[j] = ffma2(L[j], L[j], w2q);
            }
#pragma unroll
            for (int p = 0; p < P; ++p) {
#pragma unroll
                for (int j = 0; j < KP; ++j) {
                    float j0, j1;
                    if (MODE == 2) {
                        float g0, g1, M0, M1;
                        unpack2(g[p][j], g0, g1);
                        unpack2(M[j], M0, M1);
                        j0 = fmaf(rt.x, g0, M0);
                        j1 = fmaf(rt.x, g1, M1);
                        m[p] = fmin3(m[p], j0, j1);
                        continue;
                    }
                    if (MODE == 4) {
                        unpack2(M[j], j0, j1);
                        m[p] = fmin3(m[p], j0 + nqs[p] * 0.f, j1);  // (the add folds away: ALU only)
                        continue;
                    }
                    const u64 J = MODE == 6 ? ffma2(nwh, g[p][j], ffma2(pack2(nqs[p], nqs[p]), L[j], M[j])) : ffma2(nwh, g[p][j], M[j]);
                    if (MODE == 3) {
                        acc[p] = fadd2(acc[p], J);
                        continue;
                    }
                    unpack2(J, j0, j1);
                    if (MODE == 1) {
                        m[p] = fminf(m[p], j0);
                        m2[p] = fminf(m2[p], j1);
                    } else if (MODE == 5) {
                        if (j & 1) m2[p] = fmin3(m2[p], j0, j1);
                        else m[p] = fmin3(m[p], j0, j1);
                    } else if (MODE == 7) {
                        m[p] = __int_as_float(__vimin3_s32(__float_as_int(m[p]), __float_as_int(j0), __float_as_int(j1)));
                    } else if (MODE == 8) {
                        const float t = j0 < j1 ? j0 : j1;
                        m[p] = t < m[p] ? t : m[p];
                    } else
                        m[p] = fmin3(m[p], j0, j1);
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        float a0, a1;
        unpack2(acc[p], a0, a1);
        s += m[p] + m2[p] + a0 + a1;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int P, int NT, int MB>
void run(const char *name, const float *src, const float2 *rt, float *out) {
    const int reps = 800, grid = 148 * MB;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE, P, NT, MB><<<grid, NT>>>(src, rt, out, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE, P, NT, MB><<<grid, NT>>>(src, rt, out, reps);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cand = (double)grid * (NT / 32) * P * reps * ROWS * 64.0 * KP;  // candidate evaluations (incl. padding slots)
    printf("{\"mode\": %d, \"P\": %d, \"threads\": %d, \"ctas_per_sm\": %d, \"name\": \"%s\", \"ms\": %.3f, \"Gcand_per_s\": %.1f, "
           "\"cycles_per_pair_per_smsp@1.965GHz\": %.2f, \"err\": \"%s\"}\n",
           MODE, P, NT, MB, name, ms, cand / ms / 1e6, (ms * 1e-3 * 1.965e9) / (cand / 2 / 32 / (148 * 4)), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *src, *out;
    float2 *rt;
    cudaMalloc(&src, sizeof(float) * ROWS * 64 * KP);
    cudaMalloc(&rt, sizeof(float2) * ROWS);
    cudaMalloc(&out, sizeof(float) * 148 * 8 * 256);
    static float h[ROWS * 64 * KP];
    float2 hr[ROWS];
    for (int i = 0; i < ROWS * 64 * KP; ++i) h[i] = -300.f + 0.01f * (i % 977);
    for (int i = 0; i < ROWS; ++i) hr[i] = make_float2(-0.05f * i, 0.0025f * i * i);
    cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMemcpy(rt, hr, sizeof(hr), cudaMemcpyHostToDevice);
    for (int round = 0; round < 2; ++round) {
        run<0, 8, 128, 4>("shared mode as k_scan_co: FFMA2 + FMNMX3", src, rt, out);
        run<1, 8, 128, 4>("two accumulators, 2 x FMNMX", src, rt, out);
        run<2, 8, 128, 4>("2 x scalar FFMA + FMNMX3", src, rt, out);
        run<3, 8, 128, 4>("FMA pipe only (FFMA2 + FADD2)", src, rt, out);
        run<4, 8, 128, 4>("ALU only (FMNMX3 of the shared M)", src, rt, out);
        run<5, 8, 128, 4>("FMNMX3 alternating between two accumulators", src, rt, out);
        run<6, 8, 128, 4>("exact-k mode: 2 FFMA2 + FMNMX3", src, rt, out);
        run<7, 8, 128, 4>("VIMNMX3.S32 on the bits", src, rt, out);
        run<8, 8, 128, 4>("min by compare + select", src, rt, out);
        run<0, 8, 128, 3>("shared mode, 3 CTAs/SM", src, rt, out);
        run<0, 8, 256, 2>("shared mode, 2 CTAs x 8 warps", src, rt, out);
        run<0, 12, 128, 3>("shared mode, P = 12, 3 CTAs/SM", src, rt, out);
        run<0, 16, 128, 2>("shared mode, P = 16, 2 CTAs/SM", src, rt, out);
        run<0, 4, 128, 4>("shared mode, P = 4", src, rt, out);
    }
    return 0;
}
