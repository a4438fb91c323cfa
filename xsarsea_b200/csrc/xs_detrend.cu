// K7: sigma0 detrending (reference detrend.py:55-64): out[l][s] = sigma0[l][s] / (gmf[s] / nanmean(gmf)).
// HBM-bound: one streaming read and one streaming write of the raster; the [W] ratio vector stays in L2.
#include <math_constants.h>
#include <stdlib.h>

#include "xs_common.cuh"

namespace xs {

// ratio[s] = gmf[s] / nanmean(gmf); single CTA, FP64 tree reduction
__global__ void __launch_bounds__(1024) k_detrend_ratio(const double *__restrict__ gmf, int64_t w, double *__restrict__ ratio,
                                                        double *__restrict__ rinv) {
    __shared__ double sh_sum[32];
    __shared__ long long sh_cnt[32];
    __shared__ double mean_s;
    double sum = 0.0;
    long long cnt = 0;
    for (int64_t i = threadIdx.x; i < w; i += blockDim.x) {
        const double v = gmf[i];
        if (!isnan(v)) {
            sum += v;
            ++cnt;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sh_sum[threadIdx.x >> 5] = sum;
        sh_cnt[threadIdx.x >> 5] = cnt;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        sum = threadIdx.x < (blockDim.x >> 5) ? sh_sum[threadIdx.x] : 0.0;
        cnt = threadIdx.x < (blockDim.x >> 5) ? sh_cnt[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if (threadIdx.x == 0) mean_s = cnt ? sum / (double)cnt : CUDART_NAN;
    }
    __syncthreads();
    const double mean = mean_s;
    for (int64_t i = threadIdx.x; i < w; i += blockDim.x) {
        const double r = gmf[i] / mean;
        ratio[i] = r;
        rinv[i] = 1.0 / r;
    }
}

// Flat grid-stride pass over 16-byte words (VEC samples each), UNROLL independent loads in flight per thread.  The
// sample index of a word is tracked incrementally (no integer division in the loop); the [W] ratio vector and its
// reciprocal stay in L1/L2.  x / r is evaluated as q0 = x*(1/r) plus one Newton/Markstein correction
// q = q0 + (x - q0*r)*(1/r) with the residual exact by FMA: the correctly rounded quotient but for rare double-rounding
// cases (<= 1 ulp), at 3 FP64 operations instead of a full division per pixel.
template <typename T, typename V, int VEC, int UNROLL>
__global__ void __launch_bounds__(256) k_detrend_vec(const V *__restrict__ s0, const double *__restrict__ ratio,
                                                     const double *__restrict__ rinv, int64_t n_words, int64_t wv,
                                                     V *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t step_col = stride % wv;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t col = i % wv;
    for (; i < n_words; i += stride * UNROLL) {
        V v[UNROLL];
        int64_t c[UNROLL];
#pragma unroll
        for (int q = 0; q < UNROLL; ++q) {
            c[q] = col;
            col += step_col;
            if (col >= wv) col -= wv;
            if (i + q * stride < n_words) v[q] = __ldcs(&s0[i + q * stride]);  // streaming: read once
        }
#pragma unroll
        for (int q = 0; q < UNROLL; ++q)
            if (i + q * stride < n_words) {
                T *e = reinterpret_cast<T *>(&v[q]);
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const double r = ratio[c[q] * VEC + k], ri = rinv[c[q] * VEC + k];
                    const double x = (double)e[k];
                    const double q0 = x * ri;
                    const double res = fma(-q0, r, x);
                    // non-finite operands (NaN/inf sigma0, zero or NaN ratio) take the plain division
                    e[k] = (T)((isfinite(q0) && isfinite(ri)) ? fma(res, ri, q0) : x / r);
                }
                __stcs(&out[i + q * stride], v[q]);
            }
    }
}

// Column-stationary variant: a thread owns VEC adjacent samples (one 16-byte word of a line), keeps their ratio and
// reciprocal in registers and walks down the lines of its CTA's row block with UNROLL independent loads in flight.  The
// [W] vectors are read once per thread instead of once per pixel (16 B of L1 traffic per pixel before, which is what
// bounded the float32 case at half the HBM rate).
template <typename T, typename V, int VEC, int UNROLL>
__global__ void __launch_bounds__(256) k_detrend_cols(const V *__restrict__ s0, const double *__restrict__ ratio,
                                                      const double *__restrict__ rinv, int64_t n_lines, int64_t wv,
                                                      int64_t lines_per_cta, V *__restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= wv) return;
    double r[VEC], ri[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        r[k] = ratio[c * VEC + k];
        ri[k] = rinv[c * VEC + k];
    }
    const int64_t l0 = (int64_t)blockIdx.y * lines_per_cta, l1 = min(l0 + lines_per_cta, n_lines);
    for (int64_t l = l0; l < l1; l += UNROLL) {
        V v[UNROLL];
#pragma unroll
        for (int q = 0; q < UNROLL; ++q)
            if (l + q < l1) v[q] = __ldcs(&s0[(l + q) * wv + c]);  // streaming: read once
#pragma unroll
        for (int q = 0; q < UNROLL; ++q)
            if (l + q < l1) {
                T *e = reinterpret_cast<T *>(&v[q]);
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const double x = (double)e[k];
                    const double q0 = x * ri[k];
                    const double res = fma(-q0, r[k], x);
                    // non-finite operands (NaN/inf sigma0, zero or NaN ratio) take the plain division
                    e[k] = (T)((isfinite(q0) && isfinite(ri[k])) ? fma(res, ri[k], q0) : x / r[k]);
                }
                __stcs(&out[(l + q) * wv + c], v[q]);
            }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_detrend_scalar(const T *__restrict__ s0, const double *__restrict__ ratio, int64_t h,
                                                        int64_t w, T *__restrict__ out) {
    const int64_t n = h * w;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (T)((double)s0[i] / ratio[i % w]);
}

}  // namespace xs

extern "C" int xs_detrend(const void *sigma0, const double *gmf_line, int64_t n_lines, int64_t n_samples, int dtype,
                          void *out, void *stream) {
    using namespace xs;
    if (!sigma0 || !gmf_line || !out || n_lines < 0 || n_samples <= 0 || (dtype != XS_F64 && dtype != XS_F32)) {
        set_error("xs_detrend: invalid argument");
        return XS_E_INVALID;
    }
    if (n_lines == 0) return XS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    double *ratio = nullptr;
    keep_async_pool();
    XS_CUDA(cudaMallocAsync(&ratio, 2 * sizeof(double) * (size_t)n_samples, st));
    double *rinv = ratio + n_samples;
    XS_LAUNCH(k_detrend_ratio, 1, 1024, 0, stream, gmf_line, n_samples, ratio, rinv);
    const int64_t n = n_lines * n_samples;
    const bool aligned = (((uintptr_t)sigma0 | (uintptr_t)out) & 15) == 0;
    auto grid_for = [](int64_t items) {
        int64_t g = ceil_div(items, 256);
        const int64_t cap = (int64_t)kNumSMs * 32;
        return (int)(g < 1 ? 1 : (g > cap ? cap : g));
    };
    // column-stationary launch shape: enough CTAs for ~16 per SM, at least 64 lines per CTA
    auto rows_shape = [&](int64_t wv, dim3 *grid, int64_t *lines_per_cta) {
        const int64_t gx = ceil_div(wv, 256);
        int64_t gy = ceil_div((int64_t)kNumSMs * 16, gx);
        if (gy > ceil_div(n_lines, 64)) gy = ceil_div(n_lines, 64);
        if (gy < 1) gy = 1;
        if (gy > 65535) gy = 65535;
        *lines_per_cta = ceil_div(n_lines, gy);
        *grid = dim3((unsigned)gx, (unsigned)ceil_div(n_lines, *lines_per_cta));
    };
    dim3 grid2;
    int64_t lpc = 0;
    if (dtype == XS_F64) {
        if (aligned && n_samples % 2 == 0 && n_lines >= 256) {
            rows_shape(n_samples / 2, &grid2, &lpc);
            XS_LAUNCH((k_detrend_cols<double, double2, 2, 4>), grid2, 256, 0, stream, (const double2 *)sigma0, ratio, rinv, n_lines,
                      n_samples / 2, lpc, (double2 *)out);
        } else if (aligned && n_samples % 2 == 0)
            XS_LAUNCH((k_detrend_vec<double, double2, 2, 4>), grid_for(n / 2 / 4), 256, 0, stream, (const double2 *)sigma0,
                      ratio, rinv, n / 2, n_samples / 2, (double2 *)out);
        else
            XS_LAUNCH(k_detrend_scalar<double>, grid_for(n), 256, 0, stream, (const double *)sigma0, ratio, n_lines, n_samples,
                      (double *)out);
    } else {
        if (aligned && n_samples % 4 == 0 && n_lines >= 256) {
            rows_shape(n_samples / 4, &grid2, &lpc);
            XS_LAUNCH((k_detrend_cols<float, float4, 4, 4>), grid2, 256, 0, stream, (const float4 *)sigma0, ratio, rinv, n_lines,
                      n_samples / 4, lpc, (float4 *)out);
        } else if (aligned && n_samples % 4 == 0)
            XS_LAUNCH((k_detrend_vec<float, float4, 4, 4>), grid_for(n / 4 / 4), 256, 0, stream, (const float4 *)sigma0,
                      ratio, rinv, n / 4, n_samples / 4, (float4 *)out);
        else
            XS_LAUNCH(k_detrend_scalar<float>, grid_for(n), 256, 0, stream, (const float *)sigma0, ratio, n_lines, n_samples,
                      (float *)out);
    }
    XS_CUDA(cudaFreeAsync(ratio, st));
    return XS_OK;
}
