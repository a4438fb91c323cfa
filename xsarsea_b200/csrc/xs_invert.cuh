// Internal declarations shared by the inversion translation units of libxsarsea_b200 (sm_100a).
#pragma once
#include <math_constants.h>

#include "xs_common.cuh"

struct xs_plan {
    // ---- co-pol model (n_inc == 0 when absent) ----
    int n_inc, n_wspd, n_phi;
    int phi_180;      // windspeed.py:152-156
    double dsig_co;   // windspeed.py:24
    double w_absmax;  // max |wspd grid| (error-bound input)
    const double *co_lut;  // caller-owned [n_inc][n_wspd][n_phi] dB
    double *inc_grid, *wspd_grid, *phi_grid, *cos_phi, *sin_phi;  // device copies of the host grids
    // scan image: [n_inc][n_wspd_pad][nph_pad] float, value = L_dB / dsig_co, +inf in the padding
    float *scan;
    float2 *rowtab;     // [n_wspd_pad] {-w/2, w*w/4} (0,0 in the padding)
    int *first_nan;     // [n_inc] flat index (w*n_phi+p) of the first NaN of the slab, or -1
    float *slab_absmax; // [n_inc] max finite |scan value| of the slab
    int kp;             // float2 pairs per lane: nph_pad = 64*kp
    int nph_pad, n_wspd_pad;
    int fast_ok;        // the FP32 scan can be used for this plan
    int inc_sorted;     // inc_grid strictly ascending (binary search allowed)
    // ---- cross-pol model (n_inc_cr == 0 when absent) ----
    int n_inc_cr, n_wspd_cr;
    const double *cr_lut;  // caller-owned [n_inc_cr][n_wspd_cr] dB
    double *inc_cr_grid, *wspd_cr_grid;
    float *wspd_cr_half;   // [n_wspd_cr] (float)(w/2), for the FP32 filter pass of the cross-pol scan
    float *cr_scan;        // [n_inc_cr][n_wspd_cr] (float) LUT dB
    float *cr_absmax;      // [n_inc_cr] max |LUT dB| of the incidence row
    double w_cr_absmax;    // max |wspd_cr grid|
    int *cr_finite;        // [n_inc_cr] bit 0: every LUT value of the incidence row is finite; bit 1: and the row is
                           // non-decreasing in wspd (the exact interval search of k_cross applies)
    int inc_cr_sorted;
    int wspd_cr_sorted;    // wspd_cr_grid strictly ascending
    // ---- counters of the last xs_invert (device) ----
    unsigned long long *stats;  // [8]
    int device;
    cudaEvent_t ev_scan0, ev_scan1;  // around the last k_scan_co launch
    int scan_timed;
};

namespace xs {

constexpr int kChunkRows = 16;     // wspd rows per staged chunk = granularity of the argmin bookkeeping
constexpr int kRowPad = 8;         // the scan image pads the wspd axis to a multiple of this (+inf rows)
constexpr int kStages = 4;         // shared-memory ring depth
constexpr int kMaxIncBins = 8192;  // bins that fit the shared-memory histograms

// raster element access: XS_F64 / XS_F32, promoted to double on load
__device__ __forceinline__ double load_real(const void *p, int64_t i, int dtype) {
    return dtype == XS_F64 ? reinterpret_cast<const double *>(p)[i] : (double)reinterpret_cast<const float *>(p)[i];
}
__device__ __forceinline__ double2 load_cplx(const void *p, int64_t i, int dtype) {
    if (dtype == XS_F64) return reinterpret_cast<const double2 *>(p)[i];
    const float2 v = reinterpret_cast<const float2 *>(p)[i];
    return make_double2((double)v.x, (double)v.y);
}

// np.argmin(np.abs(grid - v)) (windspeed.py:212, :254): first minimum, NaN-free grid.
__device__ __forceinline__ int nearest_bin(const double *__restrict__ grid, int n, double v, int sorted) {
    if (isinf(v)) return 0;  // every |grid - v| is inf: first index wins
    if (!sorted) {
        int best = 0;
        double bd = fabs(grid[0] - v);
        for (int j = 1; j < n; ++j) {
            const double d = fabs(grid[j] - v);
            if (d < bd) {
                bd = d;
                best = j;
            }
        }
        return best;
    }
    int lo = 0, hi = n;  // first index with grid[idx] >= v
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (grid[mid] < v)
            lo = mid + 1;
        else
            hi = mid;
    }
    if (lo == 0) return 0;
    if (lo == n) return n - 1;
    const double dl = fabs(grid[lo - 1] - v), dr = fabs(grid[lo] - v);
    return (dl <= dr) ? lo - 1 : lo;
}

// isnan(np.abs(z)) for complex z: np.abs is hypot, which is inf (not NaN) when either part is infinite.
__device__ __forceinline__ bool cplx_abs_is_nan(double2 z) {
    return (isnan(z.x) || isnan(z.y)) && !(isinf(z.x) || isinf(z.y));
}

// Per-pixel inputs after the dB prologue, and what has to happen to the pixel.
struct Pixel {
    double inc, s_co, s_cr, dsig_cr;  // s_* in dB (NaN = absent)
    double2 anc;
    int cls;  // 0: both outputs NaN (+0j); 1: normal
    int co;   // co-pol inversion happens (windspeed.py:209)
};

struct RasterArgs {
    const void *inc, *s_co, *s_cr, *dsig_cr, *anc;
    double dsig_cr_scalar;
    int dtype;
    unsigned flags;
};

// windspeed.py:126-128
__device__ __forceinline__ double to_db(double s) { return 10.0 * log10(s + 1e-15); }

// Classification only (no log10): which pixels take the co-pol inversion and in which incidence bin.
__device__ __forceinline__ bool pixel_co_bin(const xs_plan &pl, const RasterArgs &a, int64_t i, int *bin) {
    if (pl.n_inc == 0 || !a.s_co) return false;
    const double inc = load_real(a.inc, i, a.dtype);
    if (isnan(inc)) return false;  // :198
    const double s = load_real(a.s_co, i, a.dtype);
    const bool s_nan = (a.flags & XS_FLAG_SIGMA0_DB) ? isnan(s) : (isnan(s) || (s + 1e-15 < 0.0));
    if (s_nan) return false;  // :209
    if (!a.anc) return false;  // all-NaN ancillary: :204
    if (cplx_abs_is_nan(load_cplx(a.anc, i, a.dtype))) return false;  // :204
    *bin = nearest_bin(pl.inc_grid, pl.n_inc, inc, pl.inc_sorted);
    return true;
}

__device__ __forceinline__ Pixel load_pixel(const xs_plan &pl, const RasterArgs &a, int64_t i) {
    Pixel p;
    p.inc = load_real(a.inc, i, a.dtype);
    const bool db = a.flags & XS_FLAG_SIGMA0_DB;
    p.s_co = CUDART_NAN;
    p.s_cr = CUDART_NAN;
    if (a.s_co) {
        const double s = load_real(a.s_co, i, a.dtype);
        p.s_co = db ? s : to_db(s);
    }
    double s_cr_raw = CUDART_NAN;
    if (a.s_cr) {
        s_cr_raw = load_real(a.s_cr, i, a.dtype);
        p.s_cr = db ? s_cr_raw : to_db(s_cr_raw);
    }
    // a scalar dsig_cr is `sigma0_cr * 0 + dsig_cr` in the reference (windspeed.py:122-123): NaN where sigma0_cr is
    // not finite
    p.dsig_cr = a.dsig_cr ? load_real(a.dsig_cr, i, a.dtype) : (isfinite(s_cr_raw) ? a.dsig_cr_scalar : CUDART_NAN);
    p.anc = a.anc ? load_cplx(a.anc, i, a.dtype) : make_double2(CUDART_NAN, CUDART_NAN);
    const bool anc_nan = cplx_abs_is_nan(p.anc);
    p.cls = 1;
    if (isnan(p.inc)) p.cls = 0;                     // :198-201
    else if (!isnan(p.s_co) && anc_nan) p.cls = 0;   // :204-207
    p.co = p.cls == 1 && !isnan(p.s_co) && pl.n_inc > 0;
    return p;
}

// The reference's FP64 cost of one co-pol candidate, operation for operation (windspeed.py:220-225):
//   ((w cos(phi) - m_antenna)/2)**2 + ((w sin(phi) - m_azi)/2)**2 + ((L - s)/dsig_co)**2
// The explicit _rn intrinsics forbid FMA contraction (numba compiles the reference without fast-math).
__device__ __forceinline__ double exact_cost_co(double w, double cphi, double sphi, double L, double qa, double qb,
                                                double s, double dsig_co) {
    const double ta = __dmul_rn(__dsub_rn(__dmul_rn(w, cphi), qa), 0.5);
    const double tz = __dmul_rn(__dsub_rn(__dmul_rn(w, sphi), qb), 0.5);
    const double jw = __dadd_rn(__dmul_rn(ta, ta), __dmul_rn(tz, tz));
    const double ts = __ddiv_rn(__dsub_rn(L, s), dsig_co);
    return __dadd_rn(jw, __dmul_rn(ts, ts));
}

// numba's np.argmin: the first NaN wins, otherwise the first minimum (numba/np/arraymath.py:645-663).
struct ArgMin {
    double j;
    int idx;      // INT_MAX = nothing seen
    int nan_idx;  // INT_MAX = no NaN seen
    __device__ __forceinline__ void init() {
        j = CUDART_INF;
        idx = 0x7fffffff;
        nan_idx = 0x7fffffff;
    }
    __device__ __forceinline__ void feed(double v, int c) {
        if (isnan(v)) {
            nan_idx = min(nan_idx, c);
        } else if (v < j || (v == j && c < idx)) {
            j = v;
            idx = c;
        }
    }
    __device__ __forceinline__ void warp_reduce() {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double oj = __shfl_xor_sync(0xffffffffu, j, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            const int on = __shfl_xor_sync(0xffffffffu, nan_idx, o);
            nan_idx = min(nan_idx, on);
            if (oj < j || (oj == j && oi < idx)) {
                j = oj;
                idx = oi;
            }
        }
    }
    __device__ __forceinline__ int result() const { return nan_idx != 0x7fffffff ? nan_idx : idx; }
};

}  // namespace xs
