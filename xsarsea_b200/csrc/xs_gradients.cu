// Row F4 of SURVEY.md section 8: the local-gradient stage that consumes sigma0_detrend's output
// (reference gradients.py: local_gradients :588-634, R2 :676-722).
//
//   grad   = Scharr_x(image) + i Scharr_y(image)            (cv2.Scharr, 3x3, BORDER_REFLECT_101)
//   grad12 = grad**2                                          (squared: a gradient and its negative coincide)
//   grad2  = R2(grad12), grad3 = R2(|grad12|)                 R2 = 5x5 binomial pre-smoothing ('symm' boundary),
//                                                             2x2 NaN-skipping mean, 3x3 binomial post-smoothing
//   c      = |grad2| / (grad3 + 1e-5), values > 1 or NaN -> 0
//   out    = sqrt(grad2), grad3, c                            (half size)
//
// Two kernels.  k_grad_reduce fuses Scharr, the complex square, the separable 5x5 pre-smoothing and the 2x2 mean:
// one CTA per 8 x 32 tile of the half-size image, the image tile, the three grad12 planes and the row-smoothed planes
// live in shared memory, so the full-size raster is read once from HBM (8 B/px) and only the three half-size planes
// are written (6 B/px).  k_grad_finish does the 3x3 post-smoothing, c and the complex square root on the half-size
// planes (6 B/px read, 8 B/px written).  FP64 CUDA-core math throughout (the reference is FP64).
#include <math_constants.h>
#include <stdlib.h>

#include "xs_common.cuh"

namespace xs {

constexpr int kGT_H = 8, kGT_W = 32;             // tile of the half-size image
constexpr int kGP_H = 2 * kGT_H, kGP_W = 2 * kGT_W;  // pre-smoothed full-size pixels of the tile
constexpr int kGG_H = kGP_H + 4, kGG_W = kGP_W + 4;  // grad12 with the 5x5 halo
constexpr int kGI_H = kGG_H + 2, kGI_W = kGG_W + 2;  // image with the Scharr halo

static_assert(kGG_H % 2 == 0 && kGG_W % 2 == 0 && kGI_W % 2 == 0, "16-byte shared-memory accesses need even extents");
struct GradSmem {
    double g[3][kGG_H][kGG_W];                 // re, im, |.| of grad12
    union {
        double img[kGI_H][kGI_W];              // dead once g is built
        double hs[3][kGG_H][kGP_W];            // row-smoothed planes
    };
};

__device__ __forceinline__ int reflect101(int q, int n) {  // cv2 BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba
    if (q < 0) q = -q;
    if (q >= n) q = 2 * n - 2 - q;
    return min(max(q, 0), n - 1);
}
__device__ __forceinline__ int reflect_symm(int q, int n) {  // scipy boundary='symm': dcba|abcd|dcba
    if (q < 0) q = -1 - q;
    if (q >= n) q = 2 * n - 1 - q;
    return min(max(q, 0), n - 1);
}

template <typename T>
__global__ void __launch_bounds__(256) k_grad_reduce(const T *__restrict__ image, int h, int w, int h2, int w2,
                                                     double *__restrict__ c_re, double *__restrict__ c_im,
                                                     double *__restrict__ c_abs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GradSmem &sm = *reinterpret_cast<GradSmem *>(smem_raw);
    const int cy0 = blockIdx.y * kGT_H, cx0 = blockIdx.x * kGT_W;
    const int gy0 = 2 * cy0 - 2, gx0 = 2 * cx0 - 2;  // raw coordinates of g[.][0][0]
    // ---- image tile (raw origin gy0-1, gx0-1), Scharr's border rule applied on load ----
    for (int e = threadIdx.x; e < kGI_H * kGI_W; e += blockDim.x) {
        const int i = e / kGI_W, j = e - i * kGI_W;
        const int y = reflect101(gy0 - 1 + i, h), x = reflect101(gx0 - 1 + j, w);
        sm.img[i][j] = (double)__ldg(image + (int64_t)y * w + x);
    }
    __syncthreads();
    // ---- grad12 at the in-range positions of the tile: a 2x2 block per thread from a 4x4 image patch (16-byte loads,
    //      consecutive lanes 16 bytes apart: no bank conflicts, 32 B of shared-memory reads per value) ----
    for (int e = threadIdx.x; e < (kGG_H / 2) * (kGG_W / 2); e += blockDim.x) {
        const int bi = e / (kGG_W / 2), bj = e - bi * (kGG_W / 2);
        const int i = 2 * bi, j = 2 * bj;
        double I[4][4];  // image tile rows i..i+3, columns j..j+3 (tile position = g position + 1 on both axes)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const double2 lo = *reinterpret_cast<const double2 *>(&sm.img[i + r][j]);
            const double2 hi = *reinterpret_cast<const double2 *>(&sm.img[i + r][j + 2]);
            I[r][0] = lo.x, I[r][1] = lo.y, I[r][2] = hi.x, I[r][3] = hi.y;
        }
#pragma unroll
        for (int di = 0; di < 2; ++di)
#pragma unroll
            for (int dj = 0; dj < 2; ++dj) {
                const int y = gy0 + i + di, x = gx0 + j + dj;
                if (y < 0 || y >= h || x < 0 || x >= w) continue;
                // separable Scharr as cv2 evaluates it: difference along one axis, then 10*centre + 3*(sum of neighbours)
                const int a = di + 1, b = dj + 1;
                const double dx0 = I[a - 1][b + 1] - I[a - 1][b - 1], dx1 = I[a][b + 1] - I[a][b - 1], dx2 = I[a + 1][b + 1] - I[a + 1][b - 1];
                const double dy0 = I[a + 1][b - 1] - I[a - 1][b - 1], dy1 = I[a + 1][b] - I[a - 1][b], dy2 = I[a + 1][b + 1] - I[a - 1][b + 1];
                const double gr = __dadd_rn(__dmul_rn(10.0, dx1), __dmul_rn(3.0, __dadd_rn(dx0, dx2)));
                const double gi = __dadd_rn(__dmul_rn(10.0, dy1), __dmul_rn(3.0, __dadd_rn(dy0, dy2)));
                // numpy complex product (a+ib)(a+ib): re = a*a - b*b, im = a*b + b*a (keeps the sign of a zero imaginary part)
                sm.g[0][i + di][j + dj] = __dsub_rn(__dmul_rn(gr, gr), __dmul_rn(gi, gi));
                sm.g[1][i + di][j + dj] = __dadd_rn(__dmul_rn(gr, gi), __dmul_rn(gi, gr));
                sm.g[2][i + di][j + dj] = fma(gr, gr, gi * gi);  // |grad**2| = |grad|**2 (np.abs(grad12) to rounding, no hypot)
            }
    }
    __syncthreads();
    // ---- out-of-range halo = 'symm' reflection of the in-range values (the reflected positions are in the tile) ----
    const bool interior = gy0 >= 0 && gx0 >= 0 && gy0 + kGG_H <= h && gx0 + kGG_W <= w;  // CTA-uniform: no halo to mirror
    for (int e = threadIdx.x; !interior && e < kGG_H * kGG_W; e += blockDim.x) {
        const int i = e / kGG_W, j = e - i * kGG_W;
        const int y = gy0 + i, x = gx0 + j;
        if (y >= 0 && y < h && x >= 0 && x < w) continue;
        const int si = reflect_symm(y, h) - gy0, sj = reflect_symm(x, w) - gx0;
        const bool ok = si >= 0 && si < kGG_H && sj >= 0 && sj < kGG_W;  // false only for halo cells no output uses
#pragma unroll
        for (int k = 0; k < 3; ++k) sm.g[k][i][j] = ok ? sm.g[k][si][sj] : 0.0;
    }
    if (!interior) __syncthreads();
    // ---- 5-tap binomial along the sample axis (img is dead: hs aliases it) ----
    // two outputs per thread from three 16-byte loads, consecutive lanes 16 bytes apart (no bank conflicts)
    for (int e = threadIdx.x; e < 3 * kGG_H * (kGP_W / 2); e += blockDim.x) {
        const int k = e / (kGG_H * (kGP_W / 2)), r = e - k * (kGG_H * (kGP_W / 2));
        const int i = r / (kGP_W / 2), j = 2 * (r - i * (kGP_W / 2));
        const double2 *row = reinterpret_cast<const double2 *>(&sm.g[k][i][j]);  // 16-byte aligned: j and kGG_W are even
        const double2 v0 = row[0], v1 = row[1], v2 = row[2];
        const double o0 = (v0.x + v2.x) * 0.0625 + (v0.y + v1.y) * 0.25 + v1.x * 0.375;
        const double o1 = (v0.y + v2.y) * 0.0625 + (v1.x + v2.x) * 0.25 + v1.y * 0.375;
        *reinterpret_cast<double2 *>(&sm.hs[k][i][j]) = make_double2(o0, o1);
    }
    __syncthreads();
    // ---- 5-tap binomial along the line axis + NaN-skipping 2x2 mean ----
    for (int e = threadIdx.x; e < 3 * kGT_H * kGT_W; e += blockDim.x) {
        const int k = e / (kGT_H * kGT_W), r = e - k * (kGT_H * kGT_W);
        const int ty = r / kGT_W, tx = r - ty * kGT_W;
        const int cy = cy0 + ty, cx = cx0 + tx;
        if (cy >= h2 || cx >= w2) continue;
        double sum = 0.0;
        int cnt = 0;
        double2 col[6];  // hs rows 2ty .. 2ty+5 (row i is the top neighbour, -2, of pre-smoothed row i), columns 2tx, 2tx+1
#pragma unroll
        for (int q = 0; q < 6; ++q) col[q] = *reinterpret_cast<const double2 *>(&sm.hs[k][2 * ty + q][2 * tx]);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const double px = (col[dy].x + col[dy + 4].x) * 0.0625 + (col[dy + 1].x + col[dy + 3].x) * 0.25 + col[dy + 2].x * 0.375;
            const double py = (col[dy].y + col[dy + 4].y) * 0.0625 + (col[dy + 1].y + col[dy + 3].y) * 0.25 + col[dy + 2].y * 0.375;
            if (!isnan(px)) {
                sum += px;
                ++cnt;
            }
            if (!isnan(py)) {
                sum += py;
                ++cnt;
            }
        }
        const double v = cnt ? sum / (double)cnt : CUDART_NAN;
        double *dst = k == 0 ? c_re : (k == 1 ? c_im : c_abs);
        dst[(int64_t)cy * w2 + cx] = v;
    }
}

// ---- streaming variant (round 2) -----------------------------------------------------------------------------------
// The same arithmetic as k_grad_reduce, organised as a register sliding window: a CTA owns a strip of 124 image columns
// (62 threads x 2 adjacent columns + one halo thread on either side) and walks down a segment of lines.  Per line a thread
//   * loads the 3 x 4 image patch around its two columns (L1-resident after the first of the three visits) and evaluates the
//     Scharr gradient and the three planes (re, im, |.|) of its square for its two columns -- at the 'symm'-mirrored position
//     when the line or the column lies outside the image, so the borders need no special pass;
//   * pushes them into a 5-line window held in registers (5 phases unrolled: the window slots are compile-time indices) and
//     forms the 5-tap binomial along the line axis;
//   * exchanges those 6 values with its neighbours through shared memory (double-buffered: one barrier per line), forms
//     the 5-tap binomial along the sample axis, and every second line emits the NaN-skipping 2 x 2 mean.
// Shared-memory traffic 72 B per pixel (the tiled kernel: > 300 B, its bound); what binds this one is the FP64 pipe
// (~55 FP64 instructions per pixel).
constexpr int kGS_Threads = 64;   // small CTAs: the kernel is a chain of dependent steps per line, many CTAs per SM overlap them
constexpr int kGS_Core = kGS_Threads - 2;     // threads that produce output columns
constexpr int kGS_SegRows = 64;               // half-size lines per CTA

template <typename T>
__global__ void __launch_bounds__(kGS_Threads) k_grad_stream(const T *__restrict__ image, int h, int w, int h2, int w2,
                                                              double *__restrict__ c_re, double *__restrict__ c_im,
                                                              double *__restrict__ c_abs) {
    __shared__ double sv[2][3][2 * kGS_Threads];
    const int t = threadIdx.x;
    const int cx = blockIdx.x * kGS_Core + t - 1;       // half-size column of this thread (halo threads: -1 / beyond)
    const int x0 = 2 * cx;                               // its image columns x0, x0 + 1
    const int cy0 = blockIdx.y * kGS_SegRows, cy1 = min(cy0 + kGS_SegRows, h2);
    // effective (mirrored) columns of the two own columns and the reflect-101 neighbours the Scharr kernel reads around them
    int col[2][3];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int xe = reflect_symm(x0 + k, w);
        col[k][0] = reflect101(xe - 1, w);
        col[k][1] = xe;
        col[k][2] = reflect101(xe + 1, w);
    }
    const int v0 = 2 * cy0 - 2, v1 = 2 * cy1 + 2;  // virtual lines of the squared-gradient planes this segment needs
    // neither a line of the segment nor one of the warp's columns (nor a Scharr neighbour) is mirrored
    const bool fast = v0 - 1 >= 0 && v1 <= h - 1 && __all_sync(0xffffffffu, x0 - 1 >= 0 && x0 + 2 < w);
    const T *pcur = image + (int64_t)max(v0, 0) * w + (x0 - 1);  // fast path: line v, column x0 - 1
    double win[5][6];   // g planes of the last five virtual lines: [slot][2 * plane + column]
    double prevP[6];    // pre-smoothed values of the even line of the current pair
#pragma unroll
    for (int i = 0; i < 6; ++i) prevP[i] = 0.0;
    for (int vb = v0; vb < v1; vb += 5) {
#pragma unroll
        for (int K = 0; K < 5; ++K) {
            const int v = vb + K;
            if (v >= v1) break;  // CTA-uniform
            // ---- squared gradient at virtual line v (effective line: 'symm' mirror), own two columns ----
            double Ia[3][3], Ib[3][3];  // image patches around the first / second own column
            if (fast) {
                // interior lines (CTA-uniform) and interior columns (warp-uniform): three consecutive lines, four consecutive
                // columns, immediate offsets from one pointer that moves down a line per step
                const T *pm = pcur - w, *pp = pcur + w;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(v + 4 < h ? pcur + 4 * (int64_t)w + 1 : pcur));
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const T *line = r == 0 ? pm : (r == 1 ? pcur : pp);
                    Ia[r][0] = (double)__ldg(line);
                    Ia[r][1] = Ib[r][0] = (double)__ldg(line + 1);
                    Ia[r][2] = Ib[r][1] = (double)__ldg(line + 2);
                    Ib[r][2] = (double)__ldg(line + 3);
                }
                pcur += w;
            } else {
                const int ye = reflect_symm(v, h);
                const int rows[3] = {reflect101(ye - 1, h), ye, reflect101(ye + 1, h)};
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const T *line = image + (int64_t)rows[r] * w;
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        Ia[r][q] = (double)__ldg(line + col[0][q]);
                        Ib[r][q] = (double)__ldg(line + col[1][q]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const double (&I)[3][3] = k ? Ib : Ia;
                // separable Scharr as cv2 evaluates it: difference along one axis, then 10*centre + 3*(sum of neighbours)
                const double dx0 = I[0][2] - I[0][0], dx1 = I[1][2] - I[1][0], dx2 = I[2][2] - I[2][0];
                const double dy0 = I[2][0] - I[0][0], dy1 = I[2][1] - I[0][1], dy2 = I[2][2] - I[0][2];
                const double gr = __dadd_rn(__dmul_rn(10.0, dx1), __dmul_rn(3.0, __dadd_rn(dx0, dx2)));
                const double gi = __dadd_rn(__dmul_rn(10.0, dy1), __dmul_rn(3.0, __dadd_rn(dy0, dy2)));
                win[K][0 + k] = __dsub_rn(__dmul_rn(gr, gr), __dmul_rn(gi, gi));
                win[K][2 + k] = __dadd_rn(__dmul_rn(gr, gi), __dmul_rn(gi, gr));
                win[K][4 + k] = fma(gr, gr, gi * gi);
            }
            if (v < v0 + 4) continue;  // the window is not full yet (CTA-uniform)
            // ---- 5-tap binomial along the line axis: line Y = v - 2 (slots K-4 .. K modulo 5) ----
            const int Y = v - 2;
            const int buf = Y & 1;
            double V[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                V[i] = (win[(K + 1) % 5][i] + win[K][i]) * 0.0625 + (win[(K + 2) % 5][i] + win[(K + 4) % 5][i]) * 0.25 + win[(K + 3) % 5][i] * 0.375;
                sv[buf][i >> 1][2 * t + (i & 1)] = V[i];
            }
            __syncthreads();
            if (t >= 1 && t <= kGS_Core) {
                // ---- 5-tap binomial along the sample axis + NaN-skipping 2 x 2 mean ----
                double P[6];
#pragma unroll
                for (int pl = 0; pl < 3; ++pl) {
                    const double *row = &sv[buf][pl][2 * t];
                    const double m2 = row[-2], m1 = row[-1], p2 = row[2], p3 = row[3];
                    P[2 * pl] = (m2 + p2) * 0.0625 + (m1 + V[2 * pl + 1]) * 0.25 + V[2 * pl] * 0.375;
                    P[2 * pl + 1] = (m1 + p3) * 0.0625 + (V[2 * pl] + p2) * 0.25 + V[2 * pl + 1] * 0.375;
                }
                if (Y & 1) {
                    const int cy = Y >> 1;
                    if (cx < w2 && cy < h2) {
#pragma unroll
                        for (int pl = 0; pl < 3; ++pl) {
                            double sum = 0.0;
                            int cnt = 0;
                            const double q[4] = {prevP[2 * pl], prevP[2 * pl + 1], P[2 * pl], P[2 * pl + 1]};
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (!isnan(q[i])) {
                                    sum += q[i];
                                    ++cnt;
                                }
                            double *dst = pl == 0 ? c_re : (pl == 1 ? c_im : c_abs);
                            // sum / cnt: exact scalings for 4, 2 and 1 summands
                            dst[(int64_t)cy * w2 + cx] = cnt == 4 ? sum * 0.25 : (cnt == 2 ? sum * 0.5 : (cnt == 1 ? sum : (cnt == 3 ? sum / 3.0 : CUDART_NAN)));
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 6; ++i) prevP[i] = P[i];
                }
            }
        }
    }
}

// numpy's complex square root (npy_csqrt, the FreeBSD msun algorithm) without the overflow rescaling, which the
// magnitudes of a squared gradient never need.
__device__ __forceinline__ double2 csqrt_np(double a, double b, double mod) {  // mod = hypot(a, b)
    if (a == 0.0 && b == 0.0) return make_double2(0.0, b);
    if (isinf(b)) return make_double2(CUDART_INF, b);
    if (isnan(a)) return make_double2(a, CUDART_NAN);
    if (isinf(a)) {
        if (a < 0.0) return make_double2(fabs(b - b), copysign(a, b));
        return make_double2(a, copysign(b - b, b));
    }
    if (isnan(b)) return make_double2(CUDART_NAN, CUDART_NAN);
    if (a >= 0.0) {
        const double t = sqrt((a + mod) * 0.5);
        return make_double2(t, b / (2.0 * t));
    }
    const double t = sqrt((-a + mod) * 0.5);
    return make_double2(fabs(b) / (2.0 * t), copysign(t, b));
}

__global__ void __launch_bounds__(256) k_grad_finish(const double *__restrict__ c_re, const double *__restrict__ c_im,
                                                     const double *__restrict__ c_abs, int h2, int w2, double2 *__restrict__ g2,
                                                     double *__restrict__ g3, double *__restrict__ cq) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w2 || y >= h2) return;
    int ys[3], xs_[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        ys[d] = reflect_symm(y + d - 1, h2);
        xs_[d] = reflect_symm(x + d - 1, w2);
    }
    double acc[3];
    const double *src[3] = {c_re, c_im, c_abs};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double rows[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double *r = src[k] + (int64_t)ys[d] * w2;
            rows[d] = (r[xs_[0]] + r[xs_[2]]) * 0.0625 + r[xs_[1]] * 0.125;
        }
        acc[k] = (rows[0] + rows[2]) + rows[1] * 2.0;
    }
    const double re = acc[0], im = acc[1], a3 = acc[2];
    const int64_t o = (int64_t)y * w2 + x;
    const double mod = hypot(re, im);
    g2[o] = csqrt_np(re, im, mod);
    g3[o] = a3;
    const double c = mod / (a3 + 0.00001);
    cq[o] = (c <= 1.0) ? c : 0.0;  // where(c <= 1).fillna(0): NaN and > 1 become 0
}

}  // namespace xs

extern "C" size_t xs_local_gradients_workspace_bytes(int64_t n_lines, int64_t n_samples) {
    if (n_lines < 2 || n_samples < 2) return 0;
    return (size_t)3 * (size_t)(n_lines / 2) * (size_t)(n_samples / 2) * sizeof(double);
}

extern "C" int xs_local_gradients(const void *image, int64_t n_lines, int64_t n_samples, int dtype, void *g2, double *g3,
                                  double *c, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace xs;
    if (!image || n_lines < 0 || n_samples < 0 || (dtype != XS_F64 && dtype != XS_F32)) {
        set_error("xs_local_gradients: invalid argument");
        return XS_E_INVALID;
    }
    if (n_lines > 0x3fffffff || n_samples > 0x3fffffff) {
        set_error("xs_local_gradients: raster side above 2^30");
        return XS_E_UNSUPPORTED;
    }
    const int h = (int)n_lines, w = (int)n_samples, h2 = h / 2, w2 = w / 2;
    if (h2 == 0 || w2 == 0) return XS_OK;  // empty result
    if (!g2 || !g3 || !c) {
        set_error("xs_local_gradients: null output");
        return XS_E_INVALID;
    }
    if (!workspace || workspace_bytes < xs_local_gradients_workspace_bytes(n_lines, n_samples)) {
        set_error("xs_local_gradients: workspace too small");
        return XS_E_WORKSPACE;
    }
    double *c_re = reinterpret_cast<double *>(workspace);
    double *c_im = c_re + (size_t)h2 * w2, *c_abs = c_im + (size_t)h2 * w2;
    static int tiled = -1;  // XS_GRAD_TILED=1: the round-1 tiled kernel (development aid / cross-check)
    if (tiled < 0) {
        const char *e = getenv("XS_GRAD_TILED");
        tiled = e ? atoi(e) : 0;
    }
    if (!tiled) {
        const dim3 grid((unsigned)ceil_div(w2, kGS_Core), (unsigned)ceil_div(h2, kGS_SegRows));
        if (grid.y > 65535u) {
            set_error("xs_local_gradients: more than 65535 x 128 lines");
            return XS_E_UNSUPPORTED;
        }
        if (dtype == XS_F64)
            XS_LAUNCH(k_grad_stream<double>, grid, kGS_Threads, 0, stream, (const double *)image, h, w, h2, w2, c_re, c_im, c_abs);
        else
            XS_LAUNCH(k_grad_stream<float>, grid, kGS_Threads, 0, stream, (const float *)image, h, w, h2, w2, c_re, c_im, c_abs);
    } else {
    const dim3 grid((unsigned)ceil_div(w2, kGT_W), (unsigned)ceil_div(h2, kGT_H));
    if (grid.y > 65535u) {
        set_error("xs_local_gradients: more than 65535 x 16 lines");
        return XS_E_UNSUPPORTED;
    }
    const size_t smem = sizeof(GradSmem);
    // per launch, not once per process: the attribute belongs to the current device's context
    if (dtype == XS_F64)
        XS_CUDA(cudaFuncSetAttribute(k_grad_reduce<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
        XS_CUDA(cudaFuncSetAttribute(k_grad_reduce<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dtype == XS_F64)
        XS_LAUNCH(k_grad_reduce<double>, grid, 256, smem, stream, (const double *)image, h, w, h2, w2, c_re, c_im, c_abs);
    else
        XS_LAUNCH(k_grad_reduce<float>, grid, 256, smem, stream, (const float *)image, h, w, h2, w2, c_re, c_im, c_abs);
    }
    const dim3 grid2((unsigned)ceil_div(w2, 256), (unsigned)h2);
    XS_LAUNCH(k_grad_finish, grid2, 256, 0, stream, c_re, c_im, c_abs, h2, w2, (double2 *)g2, g3, c);
    return XS_OK;
}
