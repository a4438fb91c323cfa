// Inner-loop experiments for k_scan_co (dev tool): which resource bounds the FADD2/FFMA2/FFMA2/FMNMX3 mix?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/scan_microbench tools/scan_microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float x, float y) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y)); return d; }
__device__ __forceinline__ void unpack2(u64 v, float &x, float &y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int KP = 3, ROWS = 32;
// MODE 0: as k_scan_co. 1: FMNMX3 -> 2x FMNMX. 2: t hoisted out of the row loop (2 FMA-pipe ops / pair).
// 3: scalar FADD/FFMA/FFMA/FMNMX (no f32x2). 4: as 0 but nwh/w2q as packed pairs (no .F32 broadcast operand)
// 8: hybrid A: t by two scalar FFMA (row constants hit the operand-reuse cache), d and J packed.
// 9: hybrid B: t and d scalar, J packed.
// 10: centred expansion: per pair Lc = L - c and M = Lc^2 + w2q are shared by the P pixels; per pixel two FFMA2
//     (a = k_p*Lc + M, J = nwh*g + a) and FMNMX3.
// 11: w^2/4 folded out of the per-candidate work: t' = nwh*g (FMUL2), J = d*d + t', row minimum r, then m = min(m, r + w2q).
// 12-15: the running minimum taken by integer min/max instructions on the float bit patterns (timing only).
// 5: as 0 without the min (sum into m with FADD: all FMA pipe). 6: only the loads + FMNMX3 (no FMA-pipe work)
template <int MODE, int P>
__global__ void __launch_bounds__(256, 2) k(const float *__restrict__ src, const float2 *__restrict__ rowtab, float *out, int reps) {
    __shared__ __align__(16) float ring[ROWS * 64 * KP];
    __shared__ float2 rt_s[ROWS];
    for (int i = threadIdx.x; i < ROWS * 64 * KP; i += blockDim.x) ring[i] = src[i];
    for (int i = threadIdx.x; i < ROWS; i += blockDim.x) rt_s[i] = rowtab[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    u64 g[P][KP];
    float nqs[P], m[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        nqs[p] = src[(p * 37 + lane) % 977];
        m[p] = 1e30f;
#pragma unroll
        for (int j = 0; j < KP; ++j) g[p][j] = pack2(src[(p * 64 + j * 8 + lane) % 977], src[(p * 32 + j * 16 + lane + 3) % 977]);
    }
    const u64 *rows = reinterpret_cast<const u64 *>(ring);
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
        for (int r = 0; r < ROWS; ++r) {
            const float2 rt = rt_s[r];
            const u64 nwh = pack2(rt.x, MODE == 4 ? rt.y : rt.x), w2q = pack2(rt.y, MODE == 4 ? rt.x : rt.y);
            u64 L[KP];
#pragma unroll
            for (int j = 0; j < KP; ++j) L[j] = rows[r * (32 * KP) + lane + 32 * j];
            if (MODE == 11) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const u64 q2 = pack2(nqs[p], nqs[p]);
                    float rmin;
#pragma unroll
                    for (int j = 0; j < KP; ++j) {
                        const u64 d = fadd2(L[j], q2);
                        u64 tp;
                        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(tp) : "l"(nwh), "l"(g[p][j]));
                        const u64 J = ffma2(d, d, tp);
                        float j0, j1;
                        unpack2(J, j0, j1);
                        rmin = j == 0 ? fminf(j0, j1) : fmin3(rmin, j0, j1);
                    }
                    m[p] = fminf(m[p], rmin + rt.y);
                }
                continue;
            }
            u64 Lc[KP], M[KP];
            if (MODE == 10) {
#pragma unroll
                for (int j = 0; j < KP; ++j) {
                    Lc[j] = fadd2(L[j], pack2(rt.y, rt.y));
                    M[j] = ffma2(Lc[j], Lc[j], w2q);
                }
            }
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const u64 q2 = pack2(nqs[p], nqs[p]);
#pragma unroll
                for (int j = 0; j < KP; ++j) {
                    if (MODE == 10) {
                        const u64 a = ffma2(q2, Lc[j], M[j]);
                        const u64 J = ffma2(nwh, g[p][j], a);
                        float j0, j1;
                        unpack2(J, j0, j1);
                        m[p] = fmin3(m[p], j0, j1);
                    } else if (MODE == 3) {
                        float l0, l1, g0, g1;
                        unpack2(L[j], l0, l1);
                        unpack2(g[p][j], g0, g1);
                        const float d0 = l0 + nqs[p], d1 = l1 + nqs[p];
                        const float t0 = fmaf(rt.x, g0, rt.y), t1 = fmaf(rt.x, g1, rt.y);
                        m[p] = fminf(fminf(m[p], fmaf(d0, d0, t0)), fmaf(d1, d1, t1));
                    } else if (MODE == 8 || MODE == 9) {
                        float g0, g1;
                        unpack2(g[p][j], g0, g1);
                        const u64 t = pack2(fmaf(rt.x, g0, rt.y), fmaf(rt.x, g1, rt.y));
                        u64 d;
                        if (MODE == 9) {
                            float l0, l1;
                            unpack2(L[j], l0, l1);
                            d = pack2(l0 + nqs[p], l1 + nqs[p]);
                        } else
                            d = fadd2(L[j], q2);
                        const u64 J = ffma2(d, d, t);
                        float j0, j1;
                        unpack2(J, j0, j1);
                        m[p] = fmin3(m[p], j0, j1);
                    } else if (MODE == 6) {
                        float l0, l1;
                        unpack2(L[j], l0, l1);
                        m[p] = fmin3(m[p], l0, l1);
                    } else {
                        const u64 d = fadd2(L[j], q2);
                        const u64 t = MODE == 2 ? g[p][j] : ffma2(nwh, g[p][j], w2q);
                        const u64 J = ffma2(d, d, t);
                        float j0, j1;
                        unpack2(J, j0, j1);
                        if (MODE == 12) m[p] = __int_as_float(__vimin3_s32(__float_as_int(m[p]), __float_as_int(j0), __float_as_int(j1)));
                        else if (MODE == 13) m[p] = __int_as_float(min(min(__float_as_int(m[p]), __float_as_int(j0)), __float_as_int(j1)));
                        else if (MODE == 14) m[p] = __uint_as_float(__vimin3_u32(__float_as_uint(m[p]), __float_as_uint(j0), __float_as_uint(j1)));
                        else if (MODE == 15) m[p] = __int_as_float(__vimax3_s32(__float_as_int(m[p]), __float_as_int(j0), __float_as_int(j1)));
                        else if (MODE == 16) { if (j & 1) m[p] = fmin3(m[p], j0, j1); else m[p] = fminf(m[p], j0 + j1); }
                        else if (MODE == 1) m[p] = fminf(fminf(m[p], j0), j1);
                        else if (MODE == 5) m[p] = m[p] + j0 + j1;
                        else m[p] = fmin3(m[p], j0, j1);
                    }
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) s += m[p];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// MODE 7: candidates paired along wspd (two rows per f32x2), so g is a broadcast scalar and the row constants are
// packed pairs that stay the same for all 6*P FFMA2 of a row pair (operand-reuse friendly).
template <int P>
__global__ void __launch_bounds__(256, 2) k7(const float *__restrict__ src, const float2 *__restrict__ rowtab, float *out, int reps) {
    __shared__ __align__(16) float ring[ROWS * 64 * KP];
    __shared__ float4 rt_s[ROWS / 2];
    for (int i = threadIdx.x; i < ROWS * 64 * KP; i += blockDim.x) ring[i] = src[i];
    for (int i = threadIdx.x; i < ROWS / 2; i += blockDim.x) rt_s[i] = make_float4(rowtab[2 * i].x, rowtab[2 * i + 1].x, rowtab[2 * i].y, rowtab[2 * i + 1].y);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float g[P][2 * KP];
    float nqs[P], m[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        nqs[p] = src[(p * 37 + lane) % 977];
        m[p] = 1e30f;
#pragma unroll
        for (int j = 0; j < 2 * KP; ++j) g[p][j] = src[(p * 64 + j * 8 + lane) % 977];
    }
    const u64 *rows = reinterpret_cast<const u64 *>(ring);
    for (int rep = 0; rep < reps; ++rep) {
        for (int r = 0; r < ROWS / 2; ++r) {
            const float4 rt = rt_s[r];
            const u64 nwh = pack2(rt.x, rt.y), w2q = pack2(rt.z, rt.w);
#pragma unroll
            for (int j = 0; j < 2 * KP; ++j) {
                const u64 L = rows[r * (64 * KP) + lane + 32 * j];
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const u64 d = fadd2(L, pack2(nqs[p], nqs[p]));
                    const u64 t = ffma2(nwh, pack2(g[p][j], g[p][j]), w2q);
                    const u64 J = ffma2(d, d, t);
                    float j0, j1;
                    unpack2(J, j0, j1);
                    m[p] = fmin3(m[p], j0, j1);
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) s += m[p];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int P>
void run7(const char *name, const float *src, const float2 *rt, float *out) {
    const int reps = 800;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k7<P><<<296, 256>>>(src, rt, out, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k7<P><<<296, 256>>>(src, rt, out, reps);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cand = 296.0 * 8 * P * reps * ROWS * 64.0 * KP;
    printf("{\"mode\": 7, \"P\": %d, \"name\": \"%s\", \"ms\": %.3f, \"Gcand_per_s\": %.1f, \"cycles_per_pair_per_smsp@1.965GHz\": %.2f, \"err\": \"%s\"}\n",
           P, name, ms, cand / ms / 1e6, (ms * 1e-3 * 1.965e9) / (cand / 2 / 32 / (148 * 4)), cudaGetErrorString(cudaGetLastError()));
}

template <int MODE, int P>
void run(const char *name, const float *src, const float2 *rt, float *out) {
    const int reps = 800;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE, P><<<296, 256>>>(src, rt, out, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE, P><<<296, 256>>>(src, rt, out, reps);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cand = 296.0 * 8 * P * reps * ROWS * 64.0 * KP;  // candidate evaluations (incl. padding slots)
    printf("{\"mode\": %d, \"P\": %d, \"name\": \"%s\", \"ms\": %.3f, \"Gcand_per_s\": %.1f, \"cycles_per_pair_per_smsp@1.965GHz\": %.2f, \"err\": \"%s\"}\n",
           MODE, P, name, ms, cand / ms / 1e6, (ms * 1e-3 * 1.965e9) / (cand / 2 / 32 / (148 * 4)), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *src, *out;
    float2 *rt;
    cudaMalloc(&src, sizeof(float) * ROWS * 64 * KP);
    cudaMalloc(&rt, sizeof(float2) * ROWS);
    cudaMalloc(&out, sizeof(float) * 296 * 256);
    float h[ROWS * 64 * KP];
    float2 hr[ROWS];
    for (int i = 0; i < ROWS * 64 * KP; ++i) h[i] = -300.f + 0.01f * (i % 977);
    for (int i = 0; i < ROWS; ++i) hr[i] = make_float2(-0.05f * i, 0.0025f * i * i);
    cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMemcpy(rt, hr, sizeof(hr), cudaMemcpyHostToDevice);
    for (int round = 0; round < 2; ++round) {
        run<0, 8>("as k_scan_co", src, rt, out);
        run<1, 8>("FMNMX3 -> 2 FMNMX", src, rt, out);
        run<2, 8>("t hoisted (2 fma-pipe ops/pair)", src, rt, out);
        run<3, 8>("scalar FADD/FFMA/FFMA/FMNMX", src, rt, out);
        run<4, 8>("no .F32 broadcast operands", src, rt, out);
        run<5, 8>("sum instead of min (all fma pipe)", src, rt, out);
        run<6, 8>("LDS + FMNMX3 only", src, rt, out);
        run7<8>("pairs along wspd, g scalar", src, rt, out);
        run<8, 8>("hybrid A: scalar t, packed d and J", src, rt, out);
        run<9, 8>("hybrid B: scalar t and d, packed J", src, rt, out);
        run<10, 8>("centred expansion: 2 FFMA2 + FMNMX3 per pixel pair", src, rt, out);
        run<11, 8>("w2q folded out: FADD2 + FMUL2 + FFMA2 + row min", src, rt, out);
        run<12, 8>("min as VIMNMX3.S32 on the float bits", src, rt, out);
        run<13, 8>("min as 2 x IMNMX.S32", src, rt, out);
        run<14, 8>("min as VIMNMX3.U32", src, rt, out);
        run<15, 8>("max as VIMNMX3.S32", src, rt, out);
        run<0, 4>("as k_scan_co, P=4", src, rt, out);
        run<3, 4>("scalar, P=4", src, rt, out);
    }
    return 0;
}
