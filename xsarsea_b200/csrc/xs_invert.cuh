// Internal declarations shared by the inversion translation units of libxsarsea_b200 (sm_100a).
#pragma once
#include <math_constants.h>
#include <string.h>

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
    // the same values regrouped for the refinement (k_refine_easy): the 16 rows x 2 kp slots a scan lane L sees in chunk c are
    // contiguous, [n_inc][n_chunks][32 lanes][kp quads][8 row-lanes][4 floats] -- quad m of row-lane `sub` holds elements
    // 4m .. 4m+3 of its 4 kp candidates ordered (row sub | row sub + 8) x slot -- so a cell is kp full 128-byte lines
    float *cell;
    float2 *rowtab;     // [n_wspd_pad] {-w/2, w*w/4} (0,0 in the padding)
    int *first_nan;     // [n_inc] flat index (w*n_phi+p) of the first NaN of the slab, or -1
    float *slab_absmax; // [n_inc] max finite |scan value| of the slab
    int *slab_range;    // [2 n_inc] {smallest, largest} finite scan value of the slab as order-preserving int keys
    int kp;             // float2 pairs per lane: nph_pad = 64*kp
    int nph_pad, n_wspd_pad;
    int fast_ok;        // the FP32 scan can be used for this plan
    // exact pruning (k_tile_plan): value range of every cell = (16-row chunk, phi group) of every slab (FP64 dB; -inf / +inf
    // when the cell holds a non-finite value) and the range of |wspd| over the chunk's rows.  A phi group is what a scan lane
    // walks with one (or, kp > 3, two) of its kp float2 slots: slots j with j * n_groups / kp == g, i.e. the phi nodes
    // [64 j, 64 j + 63]
    int n_chunks, mask_sh;  // chunks per slab; a bit of the 32-bit chunk masks covers 2^mask_sh chunks
    int n_groups;           // min(kp, 3)
    double *chunk_lo, *chunk_hi;    // [n_inc][n_chunks][n_groups]
    double *chunk_wlo, *chunk_whi;  // [n_chunks]
    unsigned short *seed_rmax;      // [n_inc][64] row of the largest LUT value on the j-th seed phi node of k_tile_plan (node j * seed_stride)
    int seed_stride;                // ceil(n_phi / 64)
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
    // inverse index of the monotone cross-pol rows: cr_inv[row][b] = first wspd index whose LUT value is >= cr_vlo[row] +
    // b / cr_vscale[row], b = 0 .. kCrInvBuckets (the last entry is n_wspd_cr): one table look-up replaces most of a bisection
    unsigned short *cr_inv;
    double *cr_vlo, *cr_vscale;
    // cross-pol-only step function (k_cross_only): on a strictly increasing row the argmin of ((L - s)/dsig)^2 is the number
    // of midpoints (L[w-1] + L[w])/2, w = 1 .. n_wspd_cr - 1, below s.  cr_step_db[row][w] holds the midpoints in dB,
    // cr_step_lin[row][w] = 10^(midpoint/10), compared with sigma0 + 1e-15 so that no log10 is taken per pixel (entry 0
    // of a row is unused); cr_finite bit 2 marks the rows this applies to.
    double *cr_step_db, *cr_step_lin;
    int cr_step_rows;  // number of such rows (0: k_cross_only is not launched)
    // uniform grids (np.linspace): a direct index guess replaces the bisection of nearest_bin / of the wspd sign change
    double inc_cr_g0, inc_cr_inv_step, wspd_cr_g0, wspd_cr_inv_step;
    int inc_cr_uniform, wspd_cr_uniform;
    int device;
    // nothing mutable lives here: counters, timers and workspaces belong to the xs_invert call (ABI 2), so one plan
    // serves concurrent calls from several host threads / streams
};

struct xs_timer {
    cudaEvent_t ev[3];  // before k_scan_co, after k_scan_co, after k_refine_easy
    int recorded;
};

namespace xs {

constexpr int kChunkRows = 16;     // wspd rows per staged chunk = granularity of the argmin bookkeeping
constexpr int kRowPad = 8;         // the scan image pads the wspd axis to a multiple of this (+inf rows)
constexpr int kStages = 4;         // shared-memory ring depth of the scan (4 x 12 KB per CTA at 192 phi slots; 4 CTAs per SM)
constexpr int kCrInvBuckets = 1024;
constexpr float kBandMargin = 0.5f;  // every accepted error band is narrower than this (2 E < kBandMargin)
constexpr int kTilePad = 64;       // upper bound of the pixels per scan tile (the bin segments of the pixel list are padded to tiles)
constexpr int kPlanGroups = 3;     // phi groups of the pruning (xs_plan::n_groups <= this)
constexpr int kPlanWords = 16;     // 64-byte scan plan of a tile: [0] chunks the CTA streams, [1 + 3 w + g] chunks on which warp w computes phi group g
constexpr int kMinTilePx = 16;     // smallest scan tile (sizes the tile-plan array)
constexpr int kMaxIncBins = 6144;  // bins whose two shared-memory histograms (k_bin_scatter) fit the default 48 KB

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

// the same with a direct first guess on a (nearly) uniform ascending grid: the guess is moved to the local minimum of
// |grid - v|, which is the global one on a sorted grid; ties to the lower index like np.argmin
__device__ __forceinline__ int nearest_bin_uniform(const double *__restrict__ grid, int n, double v, double g0, double inv_step) {
    if (isinf(v)) return 0;
    const double t = (v - g0) * inv_step;
    int i = t <= 0.0 ? 0 : (t >= (double)(n - 1) ? n - 1 : (int)(t + 0.5));
    double d = fabs(grid[i] - v);
    while (i > 0 && fabs(grid[i - 1] - v) <= d) {
        --i;
        d = fabs(grid[i] - v);
    }
    while (i + 1 < n && fabs(grid[i + 1] - v) < d) {
        ++i;
        d = fabs(grid[i] - v);
    }
    return i;
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

// co_db = false: s_co only tells NaN (no co-pol inversion) from not-NaN, without the log10 (the cross-pol pass)
__device__ __forceinline__ Pixel load_pixel(const xs_plan &pl, const RasterArgs &a, int64_t i, bool co_db = true) {
    Pixel p;
    p.inc = load_real(a.inc, i, a.dtype);
    const bool db = a.flags & XS_FLAG_SIGMA0_DB;
    p.s_co = CUDART_NAN;
    p.s_cr = CUDART_NAN;
    if (a.s_co) {
        const double s = load_real(a.s_co, i, a.dtype);
        if (co_db)
            p.s_co = db ? s : to_db(s);
        else  // log10(s + 1e-15) is NaN exactly for a NaN or negative argument
            p.s_co = (isnan(s) || (!db && s + 1e-15 < 0.0)) ? CUDART_NAN : 0.0;
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


typedef unsigned long long u64;

// ---- packed FP32 helpers (sm_100 FADD2 / FFMA2 / FMNMX3) -------------------------------------------------
__device__ __forceinline__ u64 pack2(float x, float y) {
    u64 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y));
    return d;
}
__device__ __forceinline__ void unpack2(u64 v, float &x, float &y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// g(phi) = a cos(phi) + b sin(phi) rounded to FP32; one definition so that the scan and the refinement that re-creates
// the scan's FP32 costs use bit-identical values
__device__ __forceinline__ float g32(double qa, double qb, double c, double s) {
    return (float)__fma_rn(qa, c, __dmul_rn(qb, s));
}

// ---- mbarrier / bulk-async copy (TMA) ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared-memory arrival counter (relaxed: the one thread that acts on the count fences before it does)
__device__ __forceinline__ unsigned atom_add_shared(unsigned *p, unsigned v) {
    unsigned old;
    asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// order-preserving float <-> int map (for atomicMin / atomicMax on floats)
__host__ __device__ __forceinline__ int float_order_key(float v) {
#ifdef __CUDA_ARCH__
    const int k = __float_as_int(v);
#else
    int k;
    memcpy(&k, &v, 4);
#endif
    return k ^ ((k >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float float_from_order_key(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

// ---- per-call workspace ------------------------------------------------------------------------------------------
// The pixels that take the co-pol inversion are sorted by (incidence bin, sigma0); every bin is padded to whole scan tiles,
// so tile t is the record positions [t*tile_px, (t+1)*tile_px), all of a tile's pixels share one LUT slab and the pixels
// of a warp have nearly the same sigma0.  k_list_prepare materialises one PixRec per record position (what the scan
// streams in with one bulk copy per tile); k_scan_co leaves one RefRec per scanned pixel (the error band and the
// contending (lane, chunk) cells) for the refinement kernels.
struct __align__(16) PixRec {  // 32 B
    double qa, qb, s;          // m_antenna, m_azi (|.| if phi_180), sigma0 in dB
    unsigned px;               // raster index
    unsigned short bin;
    unsigned char state;       // 0: padding, 1: scan, 2: the slab holds NaN (answer = first NaN), 3: exhaustive FP64 needed
    unsigned char neg;         // Im(ancillary) < 0 (direction sign, windspeed.py:234-242)
};
struct __align__(16) RefRec {  // 32 B
    float thr, cs, nq;         // band threshold m32 + 2E, warp centre, k = -2 (s/dsig - cs) (0: shared-sigma0 mode): what
                               // re-creates the scanned FP32 costs
    float efp;                 // bound of the full centred form's FP32 error (second filter of shared-sigma0 records)
    unsigned cont;             // lanes holding band members (0: the scan could not bound its error -> exhaustive FP64)
    unsigned mask[3];          // chunk masks (bit = chunk >> mask_sh) of the first two cont lanes and the union of the others'
};
static_assert(sizeof(PixRec) == 32 && sizeof(RefRec) == 32, "record layout");

// counters (u64): see XS_N_COUNTERS in the public header
struct Workspace {
    u64 *counters;         // [16]
    unsigned *hist;        // [n_inc] listed pixels per bin
    unsigned *bin_start;   // [n_inc + 1] record position of the bin's first pixel (bins padded to whole tiles)
    unsigned *ubase;       // [n_inc] position of the bin's first pixel in the sorted (key, pixel) arrays
    unsigned *tile_start;  // [n_inc + 1]
    unsigned *key[2];      // [n_px] sort keys (bin | quantised sigma0) and their alternate buffer
    unsigned *val[2];      // [n_px] pixel indices carried by the sort
    void *sort_temp;       // CUB's scratch
    size_t sort_temp_bytes;
    unsigned *fallback;    // [n_px] pixels for the exhaustive kernel; in a cross-pol-only call: the pixels k_cross_only left to k_cross
    PixRec *pix;           // [n_list]
    RefRec *rec;           // [n_list]
    unsigned *tile_plan;   // [tiles][kPlanWords] chunk masks of every tile (k_tile_plan)
    int *idx_tmp;          // [n_px] co-pol argmin when the caller gave no idx_co and the outputs are speed/direction planes
    int64_t n_list;        // n_px + kTilePad * n_inc rounded up to a sort run
};
size_t ws_layout(int n_inc, int64_t n_px, unsigned flags, size_t sort_bytes, bool cr_list, char *base, Workspace *w);

// ---- outputs ---------------------------------------------------------------------------------------------------------
// complex128 per pixel (the reference's return type), or -- XS_FLAG_OUT_SPEED_DIR -- two planes [speed | direction] of
// float64 / float32 (row F2: the abs / angle / dir_sample_to_meteo post-processing every caller applies, fused).
struct OutSpec {
    void *co, *cr;
    int *idx_co, *idx_cr;
    const void *gh;      // ground heading raster or NULL
    double gh_scalar;
    int64_t n_px;
    unsigned flags;
    int dtype;           // raster dtype (of gh)
};

// windspeed.py:231-247 from the flat argmin: w * (cos phi, +- sin phi); the reference picks +phi or -phi by comparing
// |angle(anc/sol)| and |angle(anc/sol2)| (ties keep +phi); for phi in [0,180] that is Im(anc) >= 0.
__device__ __forceinline__ double2 co_from_idx(const xs_plan &pl, int idx, bool anc_im_neg) {
    const int iw = idx / pl.n_phi, ip = idx - iw * pl.n_phi;
    const double w = pl.wspd_grid[iw];
    double re = w * pl.cos_phi[ip], im = w * pl.sin_phi[ip];
    if (pl.phi_180 && anc_im_neg) im = -im;
    return make_double2(re, im);
}

// np.abs / np.angle(deg=True) / (90 - angle + ground_heading) % 360 (docs/examples/windspeed_retrieval_L1.ipynb cell 33,
// detrend.py:114-130) of one result
__device__ __forceinline__ void store_wind(const OutSpec &o, void *base, int64_t px, double2 z) {
    if (!(o.flags & XS_FLAG_OUT_SPEED_DIR)) {
        reinterpret_cast<double2 *>(base)[px] = z;
        return;
    }
    const double spd = hypot(z.x, z.y);
    double dir = atan2(z.y, z.x) * (180.0 / 3.14159265358979323846);  // np.angle(z, deg=True)
    if (o.flags & XS_FLAG_DIR_METEO) {
        const double gh = o.gh ? load_real(o.gh, px, o.dtype) : o.gh_scalar;
        dir = __dadd_rn(__dsub_rn(90.0, dir), gh);  // dir_sample_to_meteo
        double r = fmod(dir, 360.0);                // np.mod: the result takes the sign of the divisor
        if (r != 0.0 && r < 0.0) r += 360.0;
        dir = r;
    }
    if (o.flags & XS_FLAG_OUT_F32) {
        float *f = reinterpret_cast<float *>(base);
        f[px] = (float)spd;
        f[o.n_px + px] = (float)dir;
    } else {
        double *d = reinterpret_cast<double *>(base);
        d[px] = spd;
        d[o.n_px + px] = dir;
    }
}

// co-pol result of one pixel: complex output (the cross-pol pass reads it back) or, with plane outputs, only the index
__device__ __forceinline__ void write_co(const xs_plan &pl, const OutSpec &o, int idx, bool anc_im_neg, int64_t px) {
    if (o.idx_co) o.idx_co[px] = idx;
    if (!(o.flags & XS_FLAG_OUT_SPEED_DIR)) reinterpret_cast<double2 *>(o.co)[px] = co_from_idx(pl, idx, anc_im_neg);
}

size_t sort_temp_bytes(int64_t n);  // CUB radix sort scratch for n (key, value) pairs with double buffers (host-side query)
int sort_pairs_u32(unsigned *keys[2], unsigned *vals[2], int64_t n, void *temp, size_t temp_bytes, cudaStream_t st, int *which);
int launch_scan_pipeline(const xs_plan *pl, const RasterArgs &ra, const Workspace &ws, const unsigned *sorted_px,
                         const OutSpec &out, int64_t n_px, xs_timer *timer, cudaStream_t st);
int scan_tile_px(int kp);

}  // namespace xs
