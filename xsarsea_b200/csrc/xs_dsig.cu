// Row F1 of SURVEY.md section 8: the pre-processors that produce the `dsig_cr` raster consumed by xs_invert
// (reference windspeed/utils.py): get_dsig (:47-91), get_dsig_wspd (:18-44), nesz_flattening (:94-163).
// All are HBM-bound FP64 CUDA-core work: element-wise pow/exp for the first two, and for the flattening a
// column-mean pass plus one read and one write of the noise raster (the per-line order-1 least-squares fit is done
// in a single pass with shifted sums, one CTA per line).
#include <math_constants.h>

#include "xs_common.cuh"

namespace xs {

template <typename T>
__device__ __forceinline__ double ld(const void *p, int64_t i) {
    return (double)__ldcs(reinterpret_cast<const T *>(p) + i);
}

// ---- get_dsig ------------------------------------------------------------------------------------------------------
// utils.py:66-86.  ID 0: 1/sqrt((s/n)**c), c = d0 + d1/(1+exp(-c0*(inc-c1)));  1: 1/sqrt((s/n)**8);  2: (1.25/(s/n))**4
// The kernels are bound by FP64 transcendental throughput, not HBM, so the powers are evaluated in their cheapest exact
// form: integer powers by repeated squaring (<= 3.5 ulp), 1/sqrt(x**c) as exp(-c/2 * log x) (error ~ c/2*|log x| ulp,
// < 20 ulp over the physical range); NaN / zero / negative / infinite ratios give what numpy's pow/sqrt give.
template <typename T, int ID>
__global__ void __launch_bounds__(256) k_dsig(const void *__restrict__ inc, const void *__restrict__ s0, const void *__restrict__ nesz,
                                              double *__restrict__ out, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double snr = __ddiv_rn(ld<T>(s0, i), ld<T>(nesz, i));
        double r;
        if (ID == 0) {
            const double c0 = 1.57952257, c1 = 25.61843791, d0 = 1.46852088, d1 = 1.4058646;
            const double x = ld<T>(inc, i);
            const double c = __dadd_rn(d0, __ddiv_rn(d1, __dadd_rn(1.0, exp(__dmul_rn(-c0, __dsub_rn(x, c1))))));
            const double t = c * log(snr);  // log of (s/n)**c
            // where numpy's power over/underflows the reference returns 1/sqrt(inf) = 0 resp. 1/sqrt(0) = inf
            r = t > 709.782712893384 ? 0.0 : (t < -745.1332191019412 ? CUDART_INF : exp(-0.5 * t));
        } else if (ID == 1) {
            const double x2 = snr * snr, x4 = x2 * x2, x8 = x4 * x4;
            // 1/sqrt(x**8) = 1/x**4 unless x**8 over/underflows (then the reference's 0 / inf / subnormal path is kept)
            r = (x8 < 2.2250738585072014e-308 || isinf(x8)) ? __ddiv_rn(1.0, sqrt(x8)) : __ddiv_rn(1.0, x4);
        } else {
            const double q = __ddiv_rn(1.25, snr), q2 = q * q;
            r = q2 * q2;
        }
        __stcs(out + i, r);
    }
}

// ---- get_dsig_wspd -------------------------------------------------------------------------------------------------
// utils.py:19-24: clip( 1/(1+exp(-b*(U-(c0-gamma*SNR)))) * 1/(1+exp((U-30)*k)), 0, 1 ); NaN propagates (np.clip).
struct WspdCoef {
    double b, c0, gamma, k;
};
template <typename T>
__global__ void __launch_bounds__(256) k_dsig_wspd(WspdCoef q, const void *__restrict__ u, const void *__restrict__ snr,
                                                   double *__restrict__ out, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double U = ld<T>(u, i), S = ld<T>(snr, i);
        const double c0 = __dsub_rn(q.c0, __dmul_rn(q.gamma, S));
        const double x = __dsub_rn(U, c0);
        const double core = __ddiv_rn(1.0, __dadd_rn(1.0, exp(__dmul_rn(-q.b, x))));
        const double drop = __ddiv_rn(1.0, __dadd_rn(1.0, exp(__dmul_rn(__dsub_rn(U, 30.0), q.k))));
        double v = __dmul_rn(core, drop);
        if (!isnan(v)) v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
        __stcs(out + i, v);
    }
}

// ---- nesz_flattening -----------------------------------------------------------------------------------------------
// Pass A: column partial sums (NaN skipped) of noise and incidence over a segment of lines; grid (col blocks, segments).
// part layout: [seg][4][w] = {sum_noise, cnt_noise, sum_inc, cnt_inc}
template <typename T>
__global__ void __launch_bounds__(256) k_colsum(const void *__restrict__ noise, const void *__restrict__ inc, int64_t h, int64_t w,
                                                int64_t lines_per_seg, double *__restrict__ part) {
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= w) return;
    const int64_t l0 = (int64_t)blockIdx.y * lines_per_seg;
    const int64_t l1 = min(h, l0 + lines_per_seg);
    double sn = 0.0, cn = 0.0, si = 0.0, ci = 0.0;
    constexpr int U = 4;
    int64_t l = l0;
    for (; l + U <= l1; l += U) {
        double a[U], b[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
            a[q] = ld<T>(noise, (l + q) * w + col);
            b[q] = ld<T>(inc, (l + q) * w + col);
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            if (!isnan(a[q])) {
                sn += a[q];
                cn += 1.0;
            }
            if (!isnan(b[q])) {
                si += b[q];
                ci += 1.0;
            }
        }
    }
    for (; l < l1; ++l) {
        const double a = ld<T>(noise, l * w + col), b = ld<T>(inc, l * w + col);
        if (!isnan(a)) {
            sn += a;
            cn += 1.0;
        }
        if (!isnan(b)) {
            si += b;
            ci += 1.0;
        }
    }
    double *p = part + (int64_t)blockIdx.y * 4 * w;
    p[col] = sn;
    p[w + col] = cn;
    p[2 * w + col] = si;
    p[3 * w + col] = ci;
}

// Pass B: means[0][w] = nanmean(noise, axis 0), means[1][w] = nanmean(inc, axis 0) (0/0 -> NaN like numpy)
__global__ void __launch_bounds__(256) k_colmean(const double *__restrict__ part, int n_seg, int64_t w, double *__restrict__ means) {
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= w) return;
    double sn = 0.0, cn = 0.0, si = 0.0, ci = 0.0;
    for (int s = 0; s < n_seg; ++s) {
        const double *p = part + (int64_t)s * 4 * w;
        sn += p[col];
        cn += p[w + col];
        si += p[2 * w + col];
        ci += p[3 * w + col];
    }
    means[col] = sn / cn;
    means[w + col] = si / ci;
}

__device__ __forceinline__ double block_sum(double v, double *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();  // sh may still be read from the previous call
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];  // same order in every thread: identical result
    return t;
}

// Pass C: one CTA per line.  y = 10*log10(noise, NaN replaced by the column mean); least-squares line through the finite
// (x = mean incidence, y) points from sums shifted by (x0, y0) = the line's first finite point, so that the centred
// moments do not cancel; then out = 10**((x*a + b - 1)/10) for every column (utils.py:140-160).
template <typename T>
__global__ void __launch_bounds__(256) k_flatten_line(const void *__restrict__ noise, const double *__restrict__ means, int64_t w,
                                                      double *__restrict__ out) {
    __shared__ double sh[8];
    __shared__ int first_s;
    const int64_t line = blockIdx.x;
    const double *cm = means, *xr = means + w;
    if (threadIdx.x == 0) first_s = 0x7fffffff;
    __syncthreads();
    // shift point: first column with a finite y (any point inside the data range would do; this one is deterministic)
    int mine = 0x7fffffff;
    for (int64_t s = threadIdx.x; s < w && mine == 0x7fffffff; s += blockDim.x) {
        double v = (double)reinterpret_cast<const T *>(noise)[line * w + s];
        if (isnan(v)) v = cm[s];
        const double y = 10.0 * log10(v);
        if (isfinite(y)) mine = (int)s;
    }
    if (mine != 0x7fffffff) atomicMin(&first_s, mine);
    __syncthreads();
    const int f = first_s;
    double *orow = out + line * w;
    if (f == 0x7fffffff) {  // no finite point: np.polyfit raises TypeError -> row of NaN (utils.py:153-155)
        for (int64_t s = threadIdx.x; s < w; s += blockDim.x) orow[s] = CUDART_NAN;
        return;
    }
    double v0 = (double)reinterpret_cast<const T *>(noise)[line * w + f];
    if (isnan(v0)) v0 = cm[f];
    const double y0 = 10.0 * log10(v0), x0 = xr[f];
    double n = 0.0, sx = 0.0, sy = 0.0, sxx = 0.0, sxy = 0.0;
    for (int64_t s = threadIdx.x; s < w; s += blockDim.x) {
        double v = ld<T>(noise, line * w + s);
        if (isnan(v)) v = cm[s];
        const double y = 10.0 * log10(v);
        if (isfinite(y)) {
            const double dx = xr[s] - x0, dy = y - y0;
            n += 1.0;
            sx += dx;
            sy += dy;
            sxx = fma(dx, dx, sxx);
            sxy = fma(dx, dy, sxy);
        }
    }
    n = block_sum(n, sh);
    sx = block_sum(sx, sh);
    sy = block_sum(sy, sh);
    sxx = block_sum(sxx, sh);
    sxy = block_sum(sxy, sh);
    const double mx = sx / n, my = sy / n;
    const double cxx = sxx - sx * mx, cxy = sxy - sx * my;
    double a, b;
    if (cxx > 0.0 && isfinite(cxx)) {  // isfinite also screens an inf abscissa
        a = cxy / cxx;
        b = (y0 + my) - a * (x0 + mx);
    } else {
        // all abscissae equal: np.polyfit scales the Vandermonde columns to unit norm, both become 1/sqrt(n), and lstsq
        // returns the minimum-norm solution of the rank-1 system, i.e. a = ybar/(2x), b = ybar/2 once the scaling is
        // undone.  A NaN abscissa among the fitted points (numpy: SVD failure) gives a NaN line here.
        const double xb = x0 + mx, yb = y0 + my;
        a = isnan(cxx) ? CUDART_NAN : yb / (2.0 * xb);
        b = isnan(cxx) ? CUDART_NAN : yb / 2.0;
    }
    for (int64_t s = threadIdx.x; s < w; s += blockDim.x)
        __stcs(orow + s, exp10((xr[s] * a + b - 1.0) / 10.0));
}

static int grid_for(int64_t n) {
    int64_t g = ceil_div(n, 256 * 4);
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace xs

extern "C" int xs_dsig(int dsig_id, int dtype, const void *inc, const void *sigma0_cr, const void *nesz_cr, double *out,
                       int64_t n, void *stream) {
    using namespace xs;
    if (dsig_id < 0 || dsig_id >= XS_DSIG_COUNT || (dtype != XS_F64 && dtype != XS_F32) || !sigma0_cr || !nesz_cr || !out ||
        n < 0 || (dsig_id == XS_DSIG_GMF_S1_V2 && !inc)) {
        set_error("xs_dsig: invalid argument");
        return XS_E_INVALID;
    }
    if (n == 0) return XS_OK;
    const int g = grid_for(n);
#define XS_DSIG_CASE(T)                                                                                 \
    switch (dsig_id) {                                                                                  \
        case 0: XS_LAUNCH((k_dsig<T, 0>), g, 256, 0, stream, inc, sigma0_cr, nesz_cr, out, n); break;   \
        case 1: XS_LAUNCH((k_dsig<T, 1>), g, 256, 0, stream, inc, sigma0_cr, nesz_cr, out, n); break;   \
        default: XS_LAUNCH((k_dsig<T, 2>), g, 256, 0, stream, inc, sigma0_cr, nesz_cr, out, n); break;  \
    }
    if (dtype == XS_F64) {
        XS_DSIG_CASE(double)
    } else {
        XS_DSIG_CASE(float)
    }
#undef XS_DSIG_CASE
    return XS_OK;
}

extern "C" int xs_dsig_wspd(int dsig_wspd_id, int dtype, const void *u_crosspol, const void *snr_cr, double *out, int64_t n,
                            void *stream) {
    using namespace xs;
    // (b, c0, gamma, k) per name, utils.py:26-42
    static const WspdCoef coef[XS_DSIG_WSPD_COUNT] = {
        {-0.4908643753212401, 16.763199934792965, 1.3891445172991084, 20.616914824394343},
        {-0.5858970325653666, 16.50039320910609, 1.1032031322520397, 7.434663633997121},
        {-0.7920301376936547, 15.8288289109038, 0.24040294696606557, 0.2538177092195224},
    };
    if (dsig_wspd_id < 0 || dsig_wspd_id >= XS_DSIG_WSPD_COUNT || (dtype != XS_F64 && dtype != XS_F32) || !u_crosspol ||
        !snr_cr || !out || n < 0) {
        set_error("xs_dsig_wspd: invalid argument");
        return XS_E_INVALID;
    }
    if (n == 0) return XS_OK;
    if (dtype == XS_F64)
        XS_LAUNCH(k_dsig_wspd<double>, grid_for(n), 256, 0, stream, coef[dsig_wspd_id], u_crosspol, snr_cr, out, n);
    else
        XS_LAUNCH(k_dsig_wspd<float>, grid_for(n), 256, 0, stream, coef[dsig_wspd_id], u_crosspol, snr_cr, out, n);
    return XS_OK;
}

static int nesz_segments(int64_t n_lines, int64_t n_samples) {
    // enough (column block, segment) CTAs to fill the chip a few times over, at least 64 lines per segment
    const int64_t col_blocks = xs::ceil_div(n_samples, 256);
    int64_t seg = xs::ceil_div((int64_t)xs::kNumSMs * 8, col_blocks);
    const int64_t max_seg = xs::ceil_div(n_lines, 64);
    if (seg > max_seg) seg = max_seg;
    if (seg < 1) seg = 1;
    return (int)seg;
}

extern "C" size_t xs_nesz_flatten_workspace_bytes(int64_t n_lines, int64_t n_samples) {
    if (n_lines <= 0 || n_samples <= 0) return 0;
    return (size_t)(nesz_segments(n_lines, n_samples) * 4 + 2) * (size_t)n_samples * sizeof(double);
}

extern "C" int xs_nesz_flatten(const void *noise, const void *inc, int64_t n_lines, int64_t n_samples, int dtype, double *out,
                               void *workspace, size_t workspace_bytes, void *stream) {
    using namespace xs;
    if (!noise || !inc || !out || n_lines < 0 || n_samples <= 0 || (dtype != XS_F64 && dtype != XS_F32)) {
        set_error("xs_nesz_flatten: invalid argument");
        return XS_E_INVALID;
    }
    if (n_lines == 0) return XS_OK;
    if (!workspace || workspace_bytes < xs_nesz_flatten_workspace_bytes(n_lines, n_samples)) {
        set_error("xs_nesz_flatten: workspace too small");
        return XS_E_WORKSPACE;
    }
    if (n_lines > 0x7fffffff) {
        set_error("xs_nesz_flatten: more than 2^31-1 lines");
        return XS_E_UNSUPPORTED;
    }
    const int n_seg = nesz_segments(n_lines, n_samples);
    const int64_t lines_per_seg = ceil_div(n_lines, n_seg);
    double *means = reinterpret_cast<double *>(workspace);
    double *part = means + 2 * n_samples;
    const dim3 gA((unsigned)ceil_div(n_samples, 256), (unsigned)n_seg);
    const int gB = (int)ceil_div(n_samples, 256);
    if (dtype == XS_F64) {
        XS_LAUNCH(k_colsum<double>, gA, 256, 0, stream, noise, inc, n_lines, n_samples, lines_per_seg, part);
        XS_LAUNCH(k_colmean, gB, 256, 0, stream, part, n_seg, n_samples, means);
        XS_LAUNCH(k_flatten_line<double>, (unsigned)n_lines, 256, 0, stream, noise, means, n_samples, out);
    } else {
        XS_LAUNCH(k_colsum<float>, gA, 256, 0, stream, noise, inc, n_lines, n_samples, lines_per_seg, part);
        XS_LAUNCH(k_colmean, gB, 256, 0, stream, part, n_seg, n_samples, means);
        XS_LAUNCH(k_flatten_line<float>, (unsigned)n_lines, 256, 0, stream, noise, means, n_samples, out);
    }
    return XS_OK;
}
