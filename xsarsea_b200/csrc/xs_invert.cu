// K1: wind inversion (reference windspeed/windspeed.py:132-331) for B200 / sm_100a.
//
// Pipeline of one xs_invert call (all on the caller's stream, no host synchronisation; every piece of mutable state --
// workspace, counters, timer -- belongs to the call, so one plan serves concurrent calls):
//   k_bin_keys / k_bin_offsets / radix sort (xs_sort.cu)   the co-pol pixels ordered by (incidence bin, sigma0); every bin's
//                  segment of the list is padded to whole scan tiles
//   k_list_prepare / k_tile_plan / k_scan_co / k_refine_easy    the co-pol argmin: exact pruning of the slab, FP32 FFMA2 scan
//                  of what is kept, exact refinement (xs_scan.cu)
//   k_exact        exhaustive FP64 scan (warp per pixel) of the pixels the fast path cannot handle (non-finite
//                  inputs, magnitudes outside the error bound's range) and of XS_MODE_FP64
//   k_cross        cross-pol / dual-pol pass (windspeed.py:252-279), merge (:426-428), NaN classes, and the output
//                  epilogue (complex128, or speed / direction planes: row F2); cross-pol-only calls on strictly increasing LUT
//                  rows go through k_cross_only first (the argmin as a step function of sigma0)
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <cmath>

#include "xs_invert.cuh"

namespace xs {

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
size_t ws_layout(int n_inc, int64_t n_px, unsigned flags, size_t sort_bytes, bool cr_list, char *base, Workspace *w) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char *p = base ? base + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    // record positions: every bin is padded to whole tiles
    const int64_t n_list = n_inc > 0 ? n_px + (int64_t)kTilePad * n_inc : 0;
    const int64_t n_sort = n_inc > 0 ? n_px : 0;
    char *c = take(XS_N_COUNTERS * sizeof(u64));
    char *h = take(sizeof(unsigned) * (size_t)(n_inc + 1));
    char *bs = take(sizeof(unsigned) * (size_t)(n_inc + 1));
    char *ub = take(sizeof(unsigned) * (size_t)(n_inc + 1));
    char *ts = take(sizeof(unsigned) * (size_t)(n_inc + 1));
    char *k0 = take(sizeof(unsigned) * (size_t)n_sort);
    char *k1 = take(sizeof(unsigned) * (size_t)n_sort);
    char *v0 = take(sizeof(unsigned) * (size_t)n_sort);
    char *v1 = take(sizeof(unsigned) * (size_t)n_sort);
    if (n_inc <= 0) sort_bytes = 0;
    char *st = take(sort_bytes);
    char *fb = take(sizeof(unsigned) * (size_t)((n_inc > 0 || cr_list) ? n_px : 0));  // also k_cross_only's slow list
    char *pr = take(sizeof(PixRec) * (size_t)n_list);
    char *rr = take(sizeof(RefRec) * (size_t)n_list);
    char *tp = take(sizeof(unsigned) * kPlanWords * (size_t)(n_list / kMinTilePx + 1));
    char *it = take((flags & XS_FLAG_OUT_SPEED_DIR) ? sizeof(int) * (size_t)n_px : 0);
    if (w) {
        w->counters = (u64 *)c;
        w->hist = (unsigned *)h;
        w->bin_start = (unsigned *)bs;
        w->ubase = (unsigned *)ub;
        w->tile_start = (unsigned *)ts;
        w->key[0] = (unsigned *)k0;
        w->key[1] = (unsigned *)k1;
        w->val[0] = (unsigned *)v0;
        w->val[1] = (unsigned *)v1;
        w->sort_temp = st;
        w->sort_temp_bytes = sort_bytes;
        w->fallback = (unsigned *)fb;
        w->pix = (PixRec *)pr;
        w->rec = (RefRec *)rr;
        w->tile_plan = (unsigned *)tp;
        w->idx_tmp = (int *)it;
        w->n_list = n_list;
    }
    return off;
}

// ---- plan construction kernels -----------------------------------------------------------------------------
// scan[bin][row][slot] = (float)(L/dsig_co), +inf in the padding; per-slab first NaN and max finite magnitude
__global__ void k_build_scan(xs_plan pl) {
    const int64_t per_slab = (int64_t)pl.n_wspd_pad * pl.nph_pad;
    const int64_t n = per_slab * pl.n_inc;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int bin = (int)(i / per_slab);
        const int64_t r = i % per_slab;
        const int row = (int)(r / pl.nph_pad), slot = (int)(r % pl.nph_pad);
        float v = CUDART_INF_F;
        if (row < pl.n_wspd && slot < pl.n_phi) {
            const double L = pl.co_lut[((int64_t)bin * pl.n_wspd + row) * pl.n_phi + slot];
            v = (float)(L / pl.dsig_co);
            if (isnan(L))
                atomicMin(&pl.first_nan[bin], row * pl.n_phi + slot);
            else if (!isinf(v)) {
                atomicMax(reinterpret_cast<unsigned *>(&pl.slab_absmax[bin]), __float_as_uint(fabsf(v)));
                atomicMin(&pl.slab_range[2 * bin], float_order_key(v));
                atomicMax(&pl.slab_range[2 * bin + 1], float_order_key(v));
            }
        }
        pl.scan[i] = v;
    }
}
// cell image of the refinement (layout: xs_plan::cell), from the scan image
__global__ void k_build_cell(xs_plan pl) {
    const int kp = pl.kp;
    const int64_t per_slab = (int64_t)pl.n_chunks * kChunkRows * pl.nph_pad;
    const int64_t n = per_slab * pl.n_inc;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const int bin = (int)(e / per_slab);
        int64_t r = e % per_slab;
        const int i = (int)(r & 3);
        r >>= 2;
        const int sub = (int)(r & 7);
        r >>= 3;
        const int m = (int)(r % kp);
        r /= kp;
        const int L = (int)(r & 31);
        const int c = (int)(r >> 5);
        const int idx = 4 * m + i, h = idx / (2 * kp), k = idx - h * 2 * kp;
        const int row = c * kChunkRows + sub + 8 * h, slot = 2 * (L + 32 * (k >> 1)) + (k & 1);
        pl.cell[e] = row < pl.n_wspd_pad ? pl.scan[((int64_t)bin * pl.n_wspd_pad + row) * pl.nph_pad + slot] : CUDART_INF_F;
    }
}
__global__ void k_init_slab_range(xs_plan pl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pl.n_inc) {
        pl.slab_range[2 * i] = 0x7fffffff;
        pl.slab_range[2 * i + 1] = (int)0x80000000;
    }
}
__global__ void k_build_rowtab(xs_plan pl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pl.n_wspd_pad) return;
    float2 v = make_float2(0.f, 0.f);
    if (i < pl.n_wspd) {
        const double w = pl.wspd_grid[i];
        v = make_float2((float)(-0.5 * w), (float)(0.25 * w * w));
    }
    pl.rowtab[i] = v;
}
// value range of every cell (chunk x phi group) of every slab over the chunk's valid rows, and the |wspd| range of every chunk:
// what k_tile_plan's lower bounds are made of.  One warp per (bin, chunk, group).
__global__ void k_build_chunk_ranges(xs_plan pl) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= (int64_t)pl.n_inc * pl.n_chunks * pl.n_groups) return;
    const int g = (int)(warp % pl.n_groups);
    const int c = (int)((warp / pl.n_groups) % pl.n_chunks), bin = (int)(warp / pl.n_groups / pl.n_chunks);
    const int r0 = c * kChunkRows, r1 = min(r0 + kChunkRows, pl.n_wspd);
    // phi nodes of the group: slots j with j * n_groups / kp == g
    int j0 = 0, j1 = 0;
    for (int j = 0; j < pl.kp; ++j)
        if (j * pl.n_groups / pl.kp == g) {
            if (j1 == 0) j0 = j;
            j1 = j + 1;
        }
    const int p0 = min(64 * j0, pl.n_phi), p1 = min(64 * j1, pl.n_phi);
    double lo = CUDART_INF, hi = -CUDART_INF, wlo = CUDART_INF, whi = -CUDART_INF;
    bool finite = true;
    const double *slab = pl.co_lut + (int64_t)bin * pl.n_wspd * pl.n_phi;
    for (int r = r0; r < r1; ++r)
        for (int ip = p0 + lane; ip < p1; ip += 32) {
            const double v = slab[(int64_t)r * pl.n_phi + ip];
            finite &= isfinite(v);
            lo = fmin(lo, v);
            hi = fmax(hi, v);
        }
    for (int r = r0 + lane; r < r1; r += 32) {
        const double w = fabs(pl.wspd_grid[r]);
        wlo = fmin(wlo, w);
        whi = fmax(whi, w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        wlo = fmin(wlo, __shfl_xor_sync(0xffffffffu, wlo, o));
        whi = fmax(whi, __shfl_xor_sync(0xffffffffu, whi, o));
    }
    finite = __all_sync(0xffffffffu, finite);
    if (lane == 0) {
        // a cell with a non-finite value gives no sigma0 bound; a cell without candidates can always be skipped
        const bool empty = r1 <= r0 || p1 <= p0;
        pl.chunk_lo[warp] = empty ? CUDART_INF : (finite ? lo : -CUDART_INF);
        pl.chunk_hi[warp] = empty ? CUDART_INF : (finite ? hi : CUDART_INF);
        if (bin == 0 && g == 0) {
            pl.chunk_wlo[c] = r1 > r0 ? wlo : 0.0;
            pl.chunk_whi[c] = r1 > r0 ? whi : CUDART_INF;
        }
    }
}
// row of the largest LUT value on every seed phi node of every slab (GMFs saturate: the largest sigma0 is not at the largest
// wind speed for every direction); the seed of a pixel whose sigma0 lies above the column (k_tile_plan)
__global__ void k_build_seed_rmax(xs_plan pl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pl.n_inc * 64) return;
    const int bin = i >> 6, ip = (i & 63) * pl.seed_stride;
    int best = 0;
    if (ip < pl.n_phi) {
        const double *slab = pl.co_lut + (int64_t)bin * pl.n_wspd * pl.n_phi;
        double bv = -CUDART_INF;
        for (int r = 0; r < pl.n_wspd; ++r) {
            const double v = slab[(int64_t)r * pl.n_phi + ip];
            if (v > bv) {
                bv = v;
                best = r;
            }
        }
    }
    pl.seed_rmax[i] = (unsigned short)best;
}
// plans without a scan image still need the first-NaN table
__global__ void k_find_first_nan(xs_plan pl) {
    const int64_t per_slab = (int64_t)pl.n_wspd * pl.n_phi;
    const int64_t n = per_slab * pl.n_inc;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (isnan(pl.co_lut[i])) atomicMin(&pl.first_nan[i / per_slab], (int)(i % per_slab));
}
__global__ void k_build_cr_tables(xs_plan pl, int *n_step_rows) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pl.n_wspd_cr) pl.wspd_cr_half[i] = (float)(0.5 * pl.wspd_cr_grid[i]);
    if (i < pl.n_inc_cr) {
        int ok = 1, mono = 1;
        float amax = 0.f;
        double prev = -CUDART_INF;
        for (int w = 0; w < pl.n_wspd_cr; ++w) {
            const double v = pl.cr_lut[(int64_t)i * pl.n_wspd_cr + w];
            ok &= isfinite(v) ? 1 : 0;
            mono &= (v >= prev) ? 1 : 0;
            prev = v;
            pl.cr_scan[(int64_t)i * pl.n_wspd_cr + w] = (float)v;
            if (isfinite(v)) amax = fmaxf(amax, (float)fabs(v) * 1.0000002f);
        }
        pl.cr_absmax[i] = amax;
        // step function of the cross-pol-only argmin (xs_plan::cr_step_db): rows whose neighbouring values are clearly apart
        int step = ok & mono;
        {
            const double *row = pl.cr_lut + (int64_t)i * pl.n_wspd_cr;
            double *sdb = pl.cr_step_db + (int64_t)i * pl.n_wspd_cr, *slin = pl.cr_step_lin + (int64_t)i * pl.n_wspd_cr;
            sdb[0] = slin[0] = 0.0;
            for (int w = 1; w < pl.n_wspd_cr; ++w) {
                const double a = row[w - 1], b = row[w];
                step &= (b - a > 1e-9 * (fabs(a) + fabs(b) + 1.0)) ? 1 : 0;
                const double mid = 0.5 * a + 0.5 * b;
                const double lin = exp10(mid * 0.1);
                step &= (fabs(mid) < 3000.0 && lin > 1e-290 && lin < 1e290) ? 1 : 0;
                sdb[w] = mid;
                slin[w] = lin;
            }
        }
        if (step) atomicAdd(n_step_rows, 1);
        pl.cr_finite[i] = ok | ((ok & mono) << 1) | (step << 2);
        // inverse index of a monotone row (see xs_plan::cr_inv)
        const double *col = pl.cr_lut + (int64_t)i * pl.n_wspd_cr;
        unsigned short *inv = pl.cr_inv + (int64_t)i * (kCrInvBuckets + 1);
        const double vlo = col[0], vhi = col[pl.n_wspd_cr - 1];
        const bool usable = ok && mono && vhi > vlo && isfinite(vhi - vlo);
        const double scale = usable ? (double)kCrInvBuckets / (vhi - vlo) : 0.0;
        pl.cr_vlo[i] = usable ? vlo : 0.0;
        pl.cr_vscale[i] = scale;
        int w = 0;
        for (int b = 0; b < kCrInvBuckets; ++b) {
            const double edge = vlo + (double)b / scale;
            while (usable && w < pl.n_wspd_cr && col[w] < edge) ++w;
            inv[b] = (unsigned short)(usable ? w : 0);
        }
        inv[kCrInvBuckets] = (unsigned short)pl.n_wspd_cr;
    }
}
__global__ void k_fix_first_nan(xs_plan pl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pl.n_inc && pl.first_nan[i] == 0x7f7f7f7f) pl.first_nan[i] = -1;
}

// ---- ordering of the co-pol pixels by (incidence bin, sigma0) ----------------------------------------------------
// k_bin_keys: one 32-bit sort key per pixel -- the bin in the top bits, sigma0 quantised monotonically below (23 bits for a
// 501-bin LUT: steps of ~3e-5 dB) -- plus the per-bin histogram; pixels that take no co-pol inversion get the largest key
// and sort to the end.  xs_sort.cu sorts (key, pixel index) pairs; k_bin_offsets turns the histogram into padded tile
// ranges; k_list_prepare (xs_scan.cu) gathers the sorted pixels into the padded per-tile records.
constexpr int kBinThreads = 256;
constexpr int kBinPxPerCta = 256 * 32;

__device__ __forceinline__ unsigned sigma0_order_key(double s, bool is_db, int bits) {
    unsigned q;  // 23-bit monotone key
    if (is_db) {  // dB in: linear quantisation of [-200, 60] dB
        const double t = fmin(fmax((s + 200.0) * (1.0 / 260.0), 0.0), 1.0);
        q = (unsigned)(t * 8388607.0);
    } else {  // linear in: 6 exponent bits (2^-57 .. 2^6) + 17 mantissa bits of the float = steps of <= 3.3e-5 dB
        const unsigned b = __float_as_uint(fmaxf((float)s, 0.f));
        const int e = min(max((int)(b >> 23) - 70, 0), 63);
        q = ((unsigned)e << 17) | ((b >> 6) & 0x1ffffu);
        if ((int)(b >> 23) - 70 < 0) q = 0;
        if ((int)(b >> 23) - 70 > 63) q = 0x7fffffu;
    }
    return bits >= 23 ? q << (bits - 23) : q >> (23 - bits);
}

__global__ void __launch_bounds__(kBinThreads) k_bin_keys(xs_plan pl, RasterArgs a, int64_t n_px, Workspace ws, int bin_bits) {
    extern __shared__ unsigned sh_hist[];
    for (int b = threadIdx.x; b < pl.n_inc; b += blockDim.x) sh_hist[b] = 0;
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * kBinPxPerCta;
    const int64_t hi = min(lo + (int64_t)kBinPxPerCta, n_px);
    const int s_bits = 32 - bin_bits;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        int bin;
        unsigned key = 0xffffffffu;
        if (pixel_co_bin(pl, a, i, &bin)) {
            atomicAdd(&sh_hist[bin], 1u);
            key = ((unsigned)bin << s_bits) | sigma0_order_key(load_real(a.s_co, i, a.dtype), a.flags & XS_FLAG_SIGMA0_DB, s_bits);
            if (key == 0xffffffffu) key = 0xfffffffeu;
        }
        ws.key[0][i] = key;
        ws.val[0][i] = (unsigned)i;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < pl.n_inc; b += blockDim.x)
        if (sh_hist[b]) atomicAdd(&ws.hist[b], sh_hist[b]);
}

// single CTA: exclusive scans of the bins' pixel and tile counts; a bin's records start at tile_start * tile_px (every
// bin is padded to whole tiles, so tile t covers the record positions [t * tile_px, (t + 1) * tile_px) of one bin)
__global__ void k_bin_offsets(int n_inc, int tile_px, Workspace ws) {
    __shared__ unsigned carry_px, carry_tiles;
    __shared__ unsigned sh_px[1024], sh_tl[1024];
    if (threadIdx.x == 0) carry_px = carry_tiles = 0;
    __syncthreads();
    for (int base = 0; base < n_inc; base += blockDim.x) {
        const int b = base + threadIdx.x;
        const unsigned cnt = b < n_inc ? ws.hist[b] : 0u;
        const unsigned tiles = (cnt + tile_px - 1) / tile_px;
        sh_px[threadIdx.x] = cnt;
        sh_tl[threadIdx.x] = tiles;
        __syncthreads();
        for (int o = 1; o < blockDim.x; o <<= 1) {  // Hillis-Steele inclusive scan
            unsigned vp = 0, vt = 0;
            if ((int)threadIdx.x >= o) {
                vp = sh_px[threadIdx.x - o];
                vt = sh_tl[threadIdx.x - o];
            }
            __syncthreads();
            sh_px[threadIdx.x] += vp;
            sh_tl[threadIdx.x] += vt;
            __syncthreads();
        }
        if (b < n_inc) {
            const unsigned et = carry_tiles + sh_tl[threadIdx.x] - tiles;
            ws.bin_start[b] = et * tile_px;
            ws.ubase[b] = carry_px + sh_px[threadIdx.x] - cnt;  // position of the bin's first pixel in the sorted array
            ws.tile_start[b] = et;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) {
            carry_px += sh_px[threadIdx.x];
            carry_tiles += sh_tl[threadIdx.x];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ws.bin_start[n_inc] = carry_tiles * tile_px;
        ws.tile_start[n_inc] = carry_tiles;
        ws.counters[0] = carry_tiles;
    }
}

// Exhaustive FP64 argmin of one pixel by one warp (reference semantics incl. NaN).  lane <-> phi index.
__device__ int exact_scan_co(const xs_plan &pl, int bin, double qa, double qb, double s, int lane) {
    const bool finite_q = isfinite(qa) && isfinite(qb) && isfinite(s);
    if (finite_q && pl.first_nan[bin] >= 0) return pl.first_nan[bin];  // J is NaN exactly where L is NaN
    ArgMin am;
    am.init();
    const double *slab = pl.co_lut + (int64_t)bin * pl.n_wspd * pl.n_phi;
    for (int ip = lane; ip < pl.n_phi; ip += 32) {
        const double c = pl.cos_phi[ip], sn = pl.sin_phi[ip];
        for (int iw = 0; iw < pl.n_wspd; ++iw) {
            const double J = exact_cost_co(pl.wspd_grid[iw], c, sn, slab[(int64_t)iw * pl.n_phi + ip], qa, qb, s, pl.dsig_co);
            am.feed(J, iw * pl.n_phi + ip);
        }
    }
    am.warp_reduce();
    return am.result();
}

// MODE_FP64 (list == nullptr: every pixel) and the fallback list of the fast path.
__global__ void __launch_bounds__(256) k_exact(xs_plan pl, RasterArgs a, int64_t n_px, const unsigned *list,
                                               const u64 *list_count, OutSpec out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n = list ? (int64_t)*list_count : n_px;
    for (int64_t e = warp; e < n; e += n_warps) {
        const int64_t px = list ? (int64_t)list[e] : e;
        const Pixel p = load_pixel(pl, a, px);
        if (!p.co) continue;
        const int bin = nearest_bin(pl.inc_grid, pl.n_inc, p.inc, pl.inc_sorted);
        const double qb = pl.phi_180 ? fabs(p.anc.y) : p.anc.y;
        const int idx = exact_scan_co(pl, bin, p.anc.x, qb, p.s_co, lane);
        if (lane == 0) write_co(pl, out, idx, p.anc.y < 0.0, px);
    }
}

// The reference's FP64 cross-pol cost of one candidate, operation for operation (windspeed.py:257-264).
__device__ __forceinline__ double exact_cost_cr(double L, double s, double dsig, double w, double mag, bool has_co) {
    const double ts = __ddiv_rn(__dsub_rn(L, s), dsig);
    double J = __dmul_rn(ts, ts);
    if (has_co) {
        const double tw = __dmul_rn(__dsub_rn(w, mag), 0.5);
        J = __dadd_rn(J, __dmul_rn(tw, tw));
    }
    return J;
}

// Exact cross-pol argmin of one pixel by interval search, for an incidence row that is finite and non-decreasing in
// wspd (true of every built-in cross-pol model) with dsig > 0 and a strictly ascending wspd grid.  The reference cost
// is J(w) = fl(a(w) + b(w)), a = fl(ts*ts), ts = fl(fl(L[w]-s)/dsig), b = fl(tw*tw), tw = fl(fl(w-mag)*0.5)
// (windspeed.py:257-264; b absent without a co-pol solution).  IEEE rounding is monotone, so ts and tw are
// non-decreasing in w, a and b are "valley" shaped, and J >= max(a, b).  With m0 = J of any candidate, every candidate
// that can be the argmin (J <= m0, ties included) has a <= m0 and b <= m0, and each of these sets is an index interval
// whose ends are found by bisection.  The interval is then scanned in index order with the reference's operations
// (first minimum wins, like np.argmin).  Returns -1 when no finite bound exists (caller falls back to the full scan).
// Cross-pol only (no co-pol solution), monotone finite row, dsig > 0: the argmin from the nodes around the sign change of
// L[w] - s, found through the row's inverse index; -2 when a comparison falls into a rounding sliver or the magnitudes
// leave the normal range (the caller then runs the general search).  Derivation: see the `!hc` block of
// cross_interval_search, which repeats these decisions.
__device__ __forceinline__ int cross_only_fast(const xs_plan &pl, int bin, const double *__restrict__ col, int n, double s,
                                               double dsig) {
    const double t = (s - pl.cr_vlo[bin]) * pl.cr_vscale[bin];
    const int b = t > 0.0 ? (int)fmin(t, (double)(kCrInvBuckets - 1)) : 0;
    const unsigned short *inv = pl.cr_inv + (int64_t)bin * (kCrInvBuckets + 1);
    int lo = inv[b], hi = inv[b + 1];
    const int lo0 = lo, hi0 = hi;
    while (lo < hi) {  // first w in the bracket with L[w] >= s
        const int mid = (lo + hi) >> 1;
        if (col[mid] < s)
            lo = mid + 1;
        else
            hi = mid;
    }
    const int k = lo;
    const double lm1 = k > 0 ? col[k - 1] : -CUDART_INF, l0 = k < n ? col[k] : CUDART_INF;
    if ((k == lo0 && !(lm1 < s)) || (k == hi0 && l0 < s)) return -2;  // the bracket missed: general search
    const double xa = k > 0 ? fabs(__dsub_rn(lm1, s)) : CUDART_INF, xb = k < n ? fabs(__dsub_rn(l0, s)) : CUDART_INF;
    const double xm = fmin(xa, xb);
    if (!(dsig >= 1e-100 && dsig <= 1e100 && xm >= 1e-100 * dsig && xm <= 1e100 * dsig)) return -2;
    if (xa > xb * (1.0 + 4e-16)) return k;
    if (xa <= xb) {
        if (k - 1 == 0) return 0;
        if (fabs(__dsub_rn(col[k - 2], s)) > xa * (1.0 + 4e-16)) return k - 1;
    }
    return -2;
}

// (not inlined: the hot path of the cross-pol-only pass is `cross_only_fast`, which keeps the kernel's register count low)
__device__ __noinline__ int cross_interval_search(const xs_plan &pl, int bin, const double *__restrict__ col,
                                                  const double *__restrict__ wg, int n, double s, double dsig, double mag,
                                                  bool hc) {
    auto num_at = [=](int w) { return __dsub_rn(col[w], s); };  // ts = num/dsig has the sign of num (dsig > 0)
    auto tw_at = [=](int w) { return __dmul_rn(__dsub_rn(wg[w], mag), 0.5); };
    auto cost = [=](int w) { return exact_cost_cr(col[w], s, dsig, wg[w], mag, hc); };
    // first index of a non-decreasing sequence that satisfies `ge`, searched from a guessed bracket [lo, hi] that is widened
    // by bisection over everything if the final check fails (so a wrong guess costs time, never correctness)
    auto first_ge = [=](int lo, int hi, auto ge) {
        const int lo0 = lo, hi0 = hi;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (!ge(mid))
                lo = mid + 1;
            else
                hi = mid;
        }
        if ((lo == lo0 && lo > 0 && ge(lo - 1)) || (lo == hi0 && lo < n && !ge(lo))) {
            lo = 0, hi = n;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (!ge(mid))
                    lo = mid + 1;
                else
                    hi = mid;
            }
        }
        return lo;
    };
    // k = first w with num(w) >= 0, i.e. col[w] >= s: bracket from the row's inverse index
    int k;
    {
        const double t = (s - pl.cr_vlo[bin]) * pl.cr_vscale[bin];
        const int b = t > 0.0 ? (int)fmin(t, (double)(kCrInvBuckets - 1)) : 0;
        const unsigned short *inv = pl.cr_inv + (int64_t)bin * (kCrInvBuckets + 1);
        k = first_ge(inv[b], inv[b + 1], [=](int w) { return !(num_at(w) < 0.0); });
    }
    if (!hc) {
        // Cross-pol only: J(w) = g(|x_w|), x_w = fl(L[w] - s), with g(x) = fl(fl(x/dsig)^2) non-decreasing (IEEE rounding is
        // monotone).  |x_w| falls up to k - 1 and rises from k on, so the minimum is J(k-1) or J(k); np.argmin's first
        // minimum is k if J(k) < J(k-1), else the start of the run of costs equal to J(k-1) that ends at k - 1.  Whether two
        // costs are equal is decided on the |x| themselves: g(xa) > g(xb) whenever xa > xb (1 + 4e-16) (relative error of
        // g below 3.4e-16 in the normal range).  The slivers in between, and magnitudes outside the normal range, take the
        // general search below.
        const double xa = k > 0 ? fabs(num_at(k - 1)) : CUDART_INF, xb = k < n ? fabs(num_at(k)) : CUDART_INF;
        const double xm = fmin(xa, xb);
        const bool normal = dsig >= 1e-100 && dsig <= 1e100 && xm >= 1e-100 * dsig && xm <= 1e100 * dsig;
        if (normal) {
            if (xa > xb * (1.0 + 4e-16)) return k;  // J(k) < J(k-1) (also k == 0)
            if (xa <= xb) {                         // J(k-1) <= J(k): the left run
                if (k - 1 == 0) return 0;
                if (fabs(num_at(k - 2)) > xa * (1.0 + 4e-16)) return k - 1;
            }
        }
    }
    double m0 = CUDART_INF;
    if (k < n) m0 = cost(k);
    if (k > 0) m0 = fmin(m0, cost(k - 1));
    int j = 0;
    if (hc) {  // j = first w with tw(w) >= 0, i.e. wg[w] >= mag
        int lo = 0, hi = n;
        if (pl.wspd_cr_uniform) {
            const double t = (mag - pl.wspd_cr_g0) * pl.wspd_cr_inv_step;
            const int g = t <= 0.0 ? 0 : (t >= (double)n ? n : (int)t);
            lo = max(g - 1, 0);
            hi = min(g + 2, n);
        }
        j = first_ge(lo, hi, [=](int w) { return !(tw_at(w) < 0.0); });
        if (j < n) m0 = fmin(m0, cost(j));
        if (j > 0) m0 = fmin(m0, cost(j - 1));
    }
    if (!(m0 < CUDART_INF)) return -1;
    // a(w) > m0 ?  a = fl(fl(num/dsig)^2) = (num/dsig)^2 (1+d1)^2 (1+d2), |d| <= 2^-53, and q = fl(fl(sqrt(m0))*dsig) =
    // dsig*sqrt(m0)(1+d3)(1+d4): |num| outside q(1 -+ 1e-14) decides the comparison without the division (the software
    // FP64 division dominated this kernel); inside that sliver, or where the magnitudes leave the normal range so that
    // the relative-error model does not hold, the reference's own operations are evaluated.
    const double q = sqrt(m0) * dsig;
    const bool cheap = m0 >= 1e-290 && m0 <= 1e290 && q >= 1e-290 && q <= 1e290;
    const double q_hi = q * (1.0 + 1e-14), q_lo = q * (1.0 - 1e-14);
    auto a_gt_m0 = [=](int w) {
        const double num = num_at(w), an = fabs(num);
        if (cheap && an > q_hi) return true;
        if (cheap && an < q_lo) return false;
        const double t = __ddiv_rn(num, dsig);
        return __dmul_rn(t, t) > m0;
    };
    auto b_gt_m0 = [=](int w) {
        const double t = tw_at(w);
        return __dmul_rn(t, t) > m0;
    };
    // Ends of {w : x(w) <= m0} around the valley bottom c (x non-increasing left of c, non-decreasing from c on): the set
    // is a handful of nodes, so its ends are found by galloping outwards from c (1, 2, 4, ... nodes) and bisecting the last
    // stride, instead of bisecting [0, c) and [c, n).
    auto left_end = [=](int c, auto gt) {  // smallest w in [0, c] with !gt on [w, c)
        int ok = c, bad = -1, step = 1;
        for (int pos = c - 1; pos >= 0; pos -= step, step <<= 1) {
            if (gt(pos)) {
                bad = pos;
                break;
            }
            ok = pos;
        }
        int lo = bad + 1, hi = ok;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (gt(mid))
                lo = mid + 1;
            else
                hi = mid;
        }
        return lo;
    };
    auto right_end = [=](int c, auto gt) {  // smallest w in [c, n] with gt(w) (n if none)
        int ok = c - 1, bad = n, step = 1;
        for (int pos = c; pos < n; pos += step, step <<= 1) {
            if (gt(pos)) {
                bad = pos;
                break;
            }
            ok = pos;
        }
        int lo = ok + 1, hi = bad;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (!gt(mid))
                lo = mid + 1;
            else
                hi = mid;
        }
        return lo;
    };
    int first = left_end(k, a_gt_m0), last = right_end(k, a_gt_m0);
    if (hc) {  // intersect with the candidates with b(w) <= m0
        first = max(first, left_end(j, b_gt_m0));
        last = min(last, right_end(j, b_gt_m0));
    }
    double best = CUDART_INF;
    int res = -1;
    for (int w = first; w < last; ++w) {
        const double J = cost(w);
        if (J < best) {
            best = J;
            res = w;
        }
    }
    return res;
}

// Cooperative scan of one pixel's cross-pol candidates by the whole warp (all lanes call it with the same arguments): the
// path of non-monotone or non-finite LUT rows, dsig <= 0, costs that overflow, and XS_FLAG_CR_FULL_SCAN.  Not inlined: it
// is rare, and its registers would otherwise be charged to every pixel of the pass.
__device__ __noinline__ int cross_coop_scan(const xs_plan &pl, double s_cr, double dsig, double mg, int b, bool hc, bool fok,
                                            int lane) {
    const unsigned full = 0xffffffffu;
    const double *col = pl.cr_lut + (int64_t)b * pl.n_wspd_cr;
    int res = -1;
    bool settled = false;
    if (fok) {
        // FP32 filter pass: J32 = ((L32 - s32) * r32)^2 + (w32/2 - mag32/2)^2.  E bounds |J32 - J| (J = the
        // reference's FP64 cost) for every candidate whose cost is within the band of the minimum, so the true
        // argmin has J32 <= m + 2E: a single candidate in the band is the argmin, several are re-evaluated in
        // FP64 with the reference's operation order.
        const float *colf = pl.cr_scan + (int64_t)b * pl.n_wspd_cr;
        const float s32 = (float)s_cr, r32 = (float)(1.0 / dsig), mg2 = hc ? (float)(0.5 * mg) : 0.f;
        float best = CUDART_INF_F, second = CUDART_INF_F;
        int bidx = -1;
        for (int w = lane; w < pl.n_wspd_cr; w += 32) {
            const float ts = (colf[w] - s32) * r32;
            float J = ts * ts;
            if (hc) {
                const float tw = pl.wspd_cr_half[w] - mg2;
                J = fmaf(tw, tw, J);
            }
            second = fminf(second, fmaxf(best, J));
            if (J < best) {
                best = J;
                bidx = w;
            }
        }
        float m = best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(full, m, o));
        const float u = 5.9604645e-8f;
        const float R2 = 1.01f * m + 1.0f, Rp = sqrtf(R2);
        const float E = 1.5f * u * (2.f * Rp * fabsf(r32) * (pl.cr_absmax[b] + fabsf(s32)) * 1.0000002f +
                                    2.f * Rp * ((float)(0.5 * pl.w_cr_absmax) + fabsf(mg2)) * 1.0000002f + 8.f * R2 + 2.f * m);
        const float thr = m + 2.f * E;
        const bool sane = isfinite(m) && isfinite(E) && (2.f * E <= 0.01f * m + 1.0f) && isfinite(r32);
        const unsigned cont = __ballot_sync(full, best <= thr);
        const unsigned wide = __ballot_sync(full, second <= thr);
        if (!sane) {
            // magnitudes outside the range of the bound: leave it to the exhaustive pass
        } else if (wide == 0 && __popc(cont) == 1) {
            res = __shfl_sync(full, bidx, __ffs(cont) - 1);
            settled = true;
        } else if (cont != 0) {
            ArgMin am;
            am.init();
            if (second <= thr) {  // several contenders in this lane: all of the lane's candidates
                for (int w = lane; w < pl.n_wspd_cr; w += 32)
                    am.feed(exact_cost_cr(col[w], s_cr, dsig, pl.wspd_cr_grid[w], mg, hc), w);
            } else if (best <= thr && bidx >= 0) {
                am.feed(exact_cost_cr(col[bidx], s_cr, dsig, pl.wspd_cr_grid[bidx], mg, hc), bidx);
            }
            am.warp_reduce();
            res = am.result();
            settled = true;
        }
    }
    if (!settled) {  // exhaustive, reference order (non-finite inputs, flat or tied costs)
        ArgMin am;
        am.init();
        for (int w = lane; w < pl.n_wspd_cr; w += 32)
            am.feed(exact_cost_cr(col[w], s_cr, dsig, pl.wspd_cr_grid[w], mg, hc), w);
        am.warp_reduce();
        res = am.result();
    }
    return res;
}

// What the cross-pol pass leaves for one pixel: the co-pol result where the scan did not write it (no co-pol inversion:
// NaN, :250), the cross-pol index, and the dual / merged / abs() result (:422-428).
__device__ __forceinline__ void emit_cross(const OutSpec &out, unsigned flags, int64_t px, bool has_co_solution, double2 co,
                                           double2 dual, int ix) {
    double2 *const out_co = reinterpret_cast<double2 *>(out.co);
    if (out.flags & XS_FLAG_OUT_SPEED_DIR) {
        if (out_co) store_wind(out, out_co, px, co);
        if (!has_co_solution && out.idx_co) out.idx_co[px] = -1;
    } else if (!has_co_solution && out_co) {
        out_co[px] = co;
        if (out.idx_co) out.idx_co[px] = -1;
    }
    if (out.idx_cr) out.idx_cr[px] = ix;
    if (out.cr) {
        double2 o = dual;
        if (flags & XS_FLAG_MERGE_DUAL) {
            const double aco = hypot(co.x, co.y), adu = hypot(dual.x, dual.y);
            if (aco < 5.0 || adu < 5.0) o = co;
        }
        if (flags & XS_FLAG_CR_ABS)
            reinterpret_cast<double *>(out.cr)[px] = o.y == 0.0 ? fabs(o.x) : hypot(o.x, o.y);
        else
            store_wind(out, out.cr, px, o);
    }
}

// ---- cross-pol / dual-pol pass + NaN classes + merge --------------------------------------------------------
// windspeed.py:198-207 (NaN classes), :250 (no co-pol), :252-279 (cross-pol argmin), :422-428 (abs / merge).
// A warp takes 32 consecutive pixels: every lane does the per-pixel scalar work of its own pixel (dB prologue,
// incidence bin, |wind_co|), then the warp scans the wspd grid of one pixel after the other cooperatively
// (parameters broadcast by shuffle), and finally every lane writes its own pixel (coalesced).
// `list` / `n_list` (optional): the pass covers only the listed pixels (what k_cross_only left over).
__global__ void __launch_bounds__(256, 3) k_cross(const __grid_constant__ xs_plan pl, RasterArgs a, int64_t n_px, OutSpec out,
                                                  const unsigned *__restrict__ list, const u64 *__restrict__ n_list) {
    double2 *const out_co = reinterpret_cast<double2 *>(out.co);
    int *const idx_co = out.idx_co;
    const bool planes = out.flags & XS_FLAG_OUT_SPEED_DIR;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double nan = CUDART_NAN;
    const unsigned full = 0xffffffffu;
    const int64_t n_items = list ? (int64_t)*n_list : n_px;
    for (int64_t base = warp * 32; base < n_items; base += n_warps * 32) {
        const bool valid = base + lane < n_items;
        const int64_t px = !valid ? 0 : (list ? (int64_t)list[base + lane] : base + lane);
        Pixel p;
        p.cls = 0;
        p.co = 0;
        p.s_cr = p.dsig_cr = p.inc = nan;
        if (valid) p = load_pixel(pl, a, px, false);
        double2 co = make_double2(nan, 0.0), dual = make_double2(nan, 0.0);
        bool scan = false, has_co = false, filter_ok = false;
        int bin = 0, ix = -1;
        double mag = nan;
        if (valid && p.cls != 0) {
            // the co-pol solution: the complex raster the scan wrote, or -- plane outputs -- rebuilt from its index
            co = !p.co ? make_double2(nan, nan) : (planes ? co_from_idx(pl, idx_co[px], p.anc.y < 0.0) : out_co[px]);
            dual = make_double2(nan, nan);
            if (!isnan(p.s_cr) && !isnan(p.dsig_cr) && pl.n_inc_cr > 0) {
                scan = true;
                bin = pl.inc_cr_uniform ? nearest_bin_uniform(pl.inc_cr_grid, pl.n_inc_cr, p.inc, pl.inc_cr_g0, pl.inc_cr_inv_step)
                                        : nearest_bin(pl.inc_cr_grid, pl.n_inc_cr, p.inc, pl.inc_cr_sorted);
                mag = p.co ? hypot(co.x, co.y) : nan;
                has_co = !isnan(mag);
                filter_ok = isfinite(p.s_cr) && isfinite(p.dsig_cr) && p.dsig_cr != 0.0 && (!has_co || isfinite(mag)) &&
                            pl.cr_finite[bin];
            }
        }
        // Pixels of monotone LUT rows are settled lane by lane by the exact interval search (some tens of FP64 cost
        // evaluations instead of n_wspd_cr); the cooperative scan below remains for the other rows and for degenerate
        // inputs.  XS_FLAG_CR_FULL_SCAN forces the cooperative scan (tests).
        bool settled_own = false;
        if (scan && filter_ok && p.dsig_cr > 0.0 && (pl.cr_finite[bin] & 2) && pl.wspd_cr_sorted &&
            !(a.flags & XS_FLAG_CR_FULL_SCAN)) {
            const double *col = pl.cr_lut + (int64_t)bin * pl.n_wspd_cr;
            int r = has_co ? -2 : cross_only_fast(pl, bin, col, pl.n_wspd_cr, p.s_cr, p.dsig_cr);
            if (r == -2) r = cross_interval_search(pl, bin, col, pl.wspd_cr_grid, pl.n_wspd_cr, p.s_cr, p.dsig_cr, mag, has_co);
            if (r >= 0) {
                ix = r;
                settled_own = true;
            }
        }
        unsigned todo = __ballot_sync(full, scan && !settled_own);
        while (todo) {  // rare: the warp scans these pixels' wspd grids cooperatively
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int res = cross_coop_scan(pl, __shfl_sync(full, p.s_cr, src), __shfl_sync(full, p.dsig_cr, src),
                                            __shfl_sync(full, mag, src), __shfl_sync(full, bin, src),
                                            __shfl_sync(full, (int)has_co, src), __shfl_sync(full, (int)filter_ok, src), lane);
            if (lane == src) ix = res;
        }
        if (!valid) continue;
        if (scan) {
            const double wd = pl.wspd_cr_grid[ix];
            if (has_co && mag > 0.0 && !isinf(mag))
                dual = make_double2(wd * (co.x / mag), wd * (co.y / mag));
            else if (has_co && isinf(mag)) {
                const double ang = atan2(co.y, co.x);
                dual = make_double2(wd * cos(ang), wd * sin(ang));
            } else
                dual = make_double2(wd, 0.0);  // angle(0) = 0, and phi_dual = 0 without co-pol
        }
        emit_cross(out, a.flags, px, p.co, co, dual, ix);
    }
}


// ---- cross-pol only (config 4): the argmin as a step function of sigma0 ------------------------------------------------
// Without a co-pol solution the cost is J(w) = fl(fl(fl(L[w] - s)/dsig)^2) (windspeed.py:252-279 with the wind term absent):
// a monotone function of |L[w] - s|.  On a strictly increasing LUT row np.argmin is therefore the number of midpoints
// t_w = (L[w-1] + L[w])/2, w >= 1, below s -- unless s lies within rounding distance of a midpoint, where the rounded costs
// decide and may tie.  This kernel counts the midpoints below sigma0 in the LINEAR domain (10^(t_w/10) against sigma0 +
// 1e-15: no log10 per pixel) inside the bracket the row's inverse index gives, and settles every pixel that is clear of
// the midpoints by a relative guard of 1e-11 (the dB value the reference computes is within 3 ulp, < 1e-13 dB, of the
// exact one; the midpoint table within 1e-14 dB; the guard is 4e-11 dB wide -- and the neighbouring |L - s| then differ by
// far more than the 4e-16 relative sliver in which rounded costs can tie, see cross_only_fast).  Everything else -- guard
// hits, bracket misses, NaN / non-finite / non-positive inputs, NaN incidence, rows that are not strictly increasing --
// goes onto a list for k_cross.  One pixel per thread and iteration; all table loads of a pixel are independent.
constexpr int kStepProbe = 6;  // midpoints inspected per pixel (a bracket of the 1024-bucket inverse index holds 0 - 2)

__global__ void __launch_bounds__(256) k_cross_only(const __grid_constant__ xs_plan pl, RasterArgs a, int64_t n_px, OutSpec out,
                                                    unsigned *__restrict__ slow, u64 *__restrict__ n_slow) {
    const bool db = a.flags & XS_FLAG_SIGMA0_DB;
    const double *const steps = db ? pl.cr_step_db : pl.cr_step_lin;
    const int n = pl.n_wspd_cr;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t px = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; px < n_px; px += stride) {
        const double inc = load_real(a.inc, px, a.dtype);
        const double x = load_real(a.s_cr, px, a.dtype);
        const double dsig = a.dsig_cr ? load_real(a.dsig_cr, px, a.dtype) : a.dsig_cr_scalar;
        const double y = db ? x : x + 1e-15;  // the argument of the reference's log10
        int kk = -1;
        if (!isnan(inc) && isfinite(y) && (db || y > 0.0) && dsig >= 1e-100 && dsig <= 1e100) {
            const int bin = pl.inc_cr_uniform ? nearest_bin_uniform(pl.inc_cr_grid, pl.n_inc_cr, inc, pl.inc_cr_g0, pl.inc_cr_inv_step)
                                              : nearest_bin(pl.inc_cr_grid, pl.n_inc_cr, inc, pl.inc_cr_sorted);
            if (pl.cr_finite[bin] & 4) {
                // bucket of the inverse index from a float estimate of the dB value (a wrong bucket is caught below)
                const double sdb = db ? y : (double)(3.0102999566f * __log2f((float)y));
                const double t = (sdb - pl.cr_vlo[bin]) * pl.cr_vscale[bin];
                const int b = t > 0.0 ? (int)fmin(t, (double)(kCrInvBuckets - 1)) : 0;
                const unsigned short *inv = pl.cr_inv + (int64_t)bin * (kCrInvBuckets + 1);
                // first node with L >= s is in [inv[b], inv[b + 1]]; the count of midpoints below s is that node or the one before
                const int k_lo = max((int)inv[b] - 1, 0), k_hi = min((int)inv[b + 1], n - 1);
                const int w0 = max(k_lo, 1), w1 = min(k_hi + 1, n - 1);  // midpoints w0 .. w1 are inspected
                const double *row = steps + (int64_t)bin * n;
                if (w1 - w0 < kStepProbe) {
                    int c_hi = 0, c_lo = 0;
                    bool first_below = true, last_below = false;
#pragma unroll
                    for (int q = 0; q < kStepProbe; ++q) {
                        const bool on = w0 + q <= w1;
                        const double v = on ? row[w0 + q] : CUDART_INF;
                        // guard band of the midpoint: relative in the linear domain, absolute in dB
                        const bool below_hi = (db ? v + 1e-10 : v * (1.0 + 1e-11)) < y;  // certainly below sigma0
                        const bool below_lo = (db ? v - 1e-10 : v * (1.0 - 1e-11)) < y;  // possibly below sigma0
                        c_hi += below_hi ? 1 : 0;
                        c_lo += below_lo ? 1 : 0;
                        if (q == 0) first_below = below_hi;
                        if (on && w0 + q == w1) last_below = below_lo;
                    }
                    // the bracket is confirmed when the midpoint below it is below sigma0 and the one above it is not
                    const bool ok_lo = w0 == 1 || first_below;
                    const bool ok_hi = w1 == n - 1 ? (k_hi == n - 1 || !last_below) : !last_below;
                    if (w1 < w0)
                        kk = 0;  // a single-node row
                    else if (c_hi == c_lo && ok_lo && ok_hi)
                        kk = w0 - 1 + c_hi;
                }
            }
        }
        if (kk < 0) {
            slow[atomicAdd(n_slow, 1ull)] = (unsigned)px;
            continue;
        }
        // cls 1, no co-pol inversion: wind_co = NaN + NaN j (:250), wind_dual = wspd + 0 j (phi_dual = 0 without co-pol)
        emit_cross(out, a.flags, px, false, make_double2(CUDART_NAN, CUDART_NAN), make_double2(pl.wspd_cr_grid[kk], 0.0), kk);
    }
}
}  // namespace xs

using namespace xs;

// ---- plan ----------------------------------------------------------------------------------------------------
static int upload(double **dst, const double *src, int n, cudaStream_t st) {
    *dst = nullptr;
    if (n <= 0) return XS_OK;
    XS_CUDA(cudaMalloc(dst, sizeof(double) * (size_t)n));
    XS_CUDA(cudaMemcpyAsync(*dst, src, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    return XS_OK;
}
static bool strictly_ascending(const double *g, int n) {
    for (int i = 0; i < n; ++i)
        if (!(g[i] == g[i]) || (i > 0 && !(g[i] > g[i - 1]))) return false;
    return true;
}

extern "C" void xs_plan_destroy(xs_plan *pl) {
    if (!pl) return;
    cudaFree(pl->inc_grid);
    cudaFree(pl->wspd_grid);
    cudaFree(pl->phi_grid);
    cudaFree(pl->cos_phi);
    cudaFree(pl->sin_phi);
    cudaFree(pl->scan);
    cudaFree(pl->cell);
    cudaFree(pl->rowtab);
    cudaFree(pl->first_nan);
    cudaFree(pl->slab_absmax);
    cudaFree(pl->slab_range);
    cudaFree(pl->chunk_lo);
    cudaFree(pl->chunk_hi);
    cudaFree(pl->chunk_wlo);
    cudaFree(pl->chunk_whi);
    cudaFree(pl->seed_rmax);
    cudaFree(pl->inc_cr_grid);
    cudaFree(pl->wspd_cr_grid);
    cudaFree(pl->wspd_cr_half);
    cudaFree(pl->cr_scan);
    cudaFree(pl->cr_absmax);
    cudaFree(pl->cr_finite);
    cudaFree(pl->cr_inv);
    cudaFree(pl->cr_vlo);
    cudaFree(pl->cr_vscale);
    cudaFree(pl->cr_step_db);
    cudaFree(pl->cr_step_lin);
    delete pl;
}

extern "C" int xs_plan_create(const xs_plan_desc *d, void *stream, xs_plan **out) {
    if (!d || !out) {
        set_error("xs_plan_create: null argument");
        return XS_E_INVALID;
    }
    *out = nullptr;
    const bool has_co = d->co_lut_db_dev != nullptr, has_cr = d->cr_lut_db_dev != nullptr;
    if (!has_co && !has_cr) {
        set_error("xs_plan_create: neither a co-pol nor a cross-pol model given");
        return XS_E_INVALID;
    }
    if (has_co && (!d->inc_grid_host || !d->wspd_grid_host || !d->phi_grid_host || !d->cos_phi_host || !d->sin_phi_host ||
                   d->n_inc <= 0 || d->n_wspd <= 0 || d->n_phi <= 0)) {
        set_error("xs_plan_create: incomplete co-pol model description");
        return XS_E_INVALID;
    }
    if (has_cr && (!d->inc_cr_grid_host || !d->wspd_cr_grid_host || d->n_inc_cr <= 0 || d->n_wspd_cr <= 0)) {
        set_error("xs_plan_create: incomplete cross-pol model description");
        return XS_E_INVALID;
    }
    if (has_co && (int64_t)d->n_wspd * d->n_phi >= 0x7fffffffLL) {
        set_error("xs_plan_create: wspd x phi grid too large");
        return XS_E_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    xs_plan *pl = new xs_plan();
    memset(pl, 0, sizeof(*pl));
    cudaGetDevice(&pl->device);
    int rc = XS_OK;
    auto fail = [&](int code) {
        xs_plan_destroy(pl);
        return code;
    };
    if (has_co) {
        pl->n_inc = d->n_inc;
        pl->n_wspd = d->n_wspd;
        pl->n_phi = d->n_phi;
        pl->dsig_co = d->dsig_co;
        pl->co_lut = d->co_lut_db_dev;
        pl->phi_180 = (180.0 - (d->phi_grid_host[d->n_phi - 1] - d->phi_grid_host[0])) < 2.0;  // windspeed.py:152
        pl->inc_sorted = strictly_ascending(d->inc_grid_host, d->n_inc);
        double wmax = 0;
        for (int i = 0; i < d->n_wspd; ++i) wmax = fmax(wmax, fabs(d->wspd_grid_host[i]));
        pl->w_absmax = wmax;
        if ((rc = upload(&pl->inc_grid, d->inc_grid_host, d->n_inc, st)) != XS_OK) return fail(rc);
        if ((rc = upload(&pl->wspd_grid, d->wspd_grid_host, d->n_wspd, st)) != XS_OK) return fail(rc);
        if ((rc = upload(&pl->phi_grid, d->phi_grid_host, d->n_phi, st)) != XS_OK) return fail(rc);
        if ((rc = upload(&pl->cos_phi, d->cos_phi_host, d->n_phi, st)) != XS_OK) return fail(rc);
        if ((rc = upload(&pl->sin_phi, d->sin_phi_host, d->n_phi, st)) != XS_OK) return fail(rc);
        if ((rc = xs::check(cudaMalloc(&pl->first_nan, sizeof(int) * (size_t)d->n_inc), "cudaMalloc")) != XS_OK) return fail(rc);
        // FP32 scan image: only for grids the scan kernel is instantiated for and a sane dsig_co
        const int kp = (d->n_phi + 63) / 64;
        const bool kp_ok = kp == 1 || kp == 2 || kp == 3 || kp == 4 || kp == 6;
        pl->kp = kp;
        pl->nph_pad = 64 * kp;
        pl->n_wspd_pad = (d->n_wspd + kRowPad - 1) / kRowPad * kRowPad;
        // the scan's ring streams kStages chunks ahead, at most into the next tile: a slab needs at least kStages chunks
        pl->fast_ok = kp_ok && std::isfinite(d->dsig_co) && d->dsig_co != 0.0 && std::isfinite(wmax) &&
                      d->n_inc <= kMaxIncBins && pl->n_wspd_pad <= 256 * kChunkRows && pl->n_wspd_pad > (kStages - 1) * kChunkRows;
        if (pl->fast_ok) {
            const size_t n_scan = (size_t)d->n_inc * pl->n_wspd_pad * pl->nph_pad;
            if ((rc = xs::check(cudaMalloc(&pl->scan, sizeof(float) * n_scan), "cudaMalloc scan image")) != XS_OK) return fail(rc);
            if ((rc = xs::check(cudaMalloc(&pl->rowtab, sizeof(float2) * (size_t)pl->n_wspd_pad), "cudaMalloc")) != XS_OK) return fail(rc);
            if ((rc = xs::check(cudaMalloc(&pl->slab_absmax, sizeof(float) * (size_t)d->n_inc), "cudaMalloc")) != XS_OK) return fail(rc);
            if ((rc = xs::check(cudaMalloc(&pl->slab_range, sizeof(int) * 2 * (size_t)d->n_inc), "cudaMalloc")) != XS_OK) return fail(rc);
            pl->n_chunks = (pl->n_wspd_pad + kChunkRows - 1) / kChunkRows;
            while ((pl->n_chunks + (1 << pl->mask_sh) - 1) >> pl->mask_sh > 32) ++pl->mask_sh;
            if ((rc = xs::check(cudaMalloc(&pl->cell, sizeof(float) * (size_t)d->n_inc * pl->n_chunks * kChunkRows * pl->nph_pad), "cudaMalloc cell image")) != XS_OK) return fail(rc);
            pl->n_groups = kp < kPlanGroups ? kp : kPlanGroups;
            const size_t n_cr = (size_t)d->n_inc * pl->n_chunks * pl->n_groups;
            if ((rc = xs::check(cudaMalloc(&pl->chunk_lo, sizeof(double) * n_cr), "cudaMalloc")) != XS_OK) return fail(rc);
            if ((rc = xs::check(cudaMalloc(&pl->chunk_hi, sizeof(double) * n_cr), "cudaMalloc")) != XS_OK) return fail(rc);
            if ((rc = xs::check(cudaMalloc(&pl->chunk_wlo, sizeof(double) * (size_t)pl->n_chunks), "cudaMalloc")) != XS_OK) return fail(rc);
            if ((rc = xs::check(cudaMalloc(&pl->chunk_whi, sizeof(double) * (size_t)pl->n_chunks), "cudaMalloc")) != XS_OK) return fail(rc);
            if ((rc = xs::check(cudaMalloc(&pl->seed_rmax, sizeof(unsigned short) * 64 * (size_t)d->n_inc), "cudaMalloc")) != XS_OK) return fail(rc);
            pl->seed_stride = (d->n_phi + 63) / 64;
        }
    }
    if (has_cr) {
        pl->n_inc_cr = d->n_inc_cr;
        pl->n_wspd_cr = d->n_wspd_cr;
        pl->cr_lut = d->cr_lut_db_dev;
        pl->inc_cr_sorted = strictly_ascending(d->inc_cr_grid_host, d->n_inc_cr);
        pl->wspd_cr_sorted = strictly_ascending(d->wspd_cr_grid_host, d->n_wspd_cr);
        if ((rc = upload(&pl->inc_cr_grid, d->inc_cr_grid_host, d->n_inc_cr, st)) != XS_OK) return fail(rc);
        if ((rc = upload(&pl->wspd_cr_grid, d->wspd_cr_grid_host, d->n_wspd_cr, st)) != XS_OK) return fail(rc);
        if ((rc = xs::check(cudaMalloc(&pl->wspd_cr_half, sizeof(float) * (size_t)d->n_wspd_cr), "cudaMalloc")) != XS_OK) return fail(rc);
        if ((rc = xs::check(cudaMalloc(&pl->cr_scan, sizeof(float) * (size_t)d->n_wspd_cr * d->n_inc_cr), "cudaMalloc")) != XS_OK) return fail(rc);
        if ((rc = xs::check(cudaMalloc(&pl->cr_absmax, sizeof(float) * (size_t)d->n_inc_cr), "cudaMalloc")) != XS_OK) return fail(rc);
        double wcmax = 0;
        for (int i = 0; i < d->n_wspd_cr; ++i) wcmax = fmax(wcmax, fabs(d->wspd_cr_grid_host[i]));
        pl->w_cr_absmax = wcmax;
        if ((rc = xs::check(cudaMalloc(&pl->cr_finite, sizeof(int) * (size_t)d->n_inc_cr), "cudaMalloc")) != XS_OK) return fail(rc);
        if (d->n_wspd_cr >= 65535) {
            set_error("xs_plan_create: cross-pol wspd grid too long");
            return fail(XS_E_UNSUPPORTED);
        }
        if ((rc = xs::check(cudaMalloc(&pl->cr_inv, sizeof(unsigned short) * (size_t)d->n_inc_cr * (kCrInvBuckets + 1)), "cudaMalloc")) != XS_OK) return fail(rc);
        if ((rc = xs::check(cudaMalloc(&pl->cr_vlo, sizeof(double) * (size_t)d->n_inc_cr), "cudaMalloc")) != XS_OK) return fail(rc);
        if ((rc = xs::check(cudaMalloc(&pl->cr_vscale, sizeof(double) * (size_t)d->n_inc_cr), "cudaMalloc")) != XS_OK) return fail(rc);
        if ((rc = xs::check(cudaMalloc(&pl->cr_step_db, sizeof(double) * (size_t)d->n_inc_cr * d->n_wspd_cr), "cudaMalloc")) != XS_OK) return fail(rc);
        if ((rc = xs::check(cudaMalloc(&pl->cr_step_lin, sizeof(double) * (size_t)d->n_inc_cr * d->n_wspd_cr), "cudaMalloc")) != XS_OK) return fail(rc);
        auto uniform = [](const double *g, int n, double *g0, double *inv_step) {  // ascending and within 1 % of a constant step
            if (n < 2 || !strictly_ascending(g, n)) return 0;
            const double step = (g[n - 1] - g[0]) / (n - 1);
            for (int i = 1; i < n; ++i)
                if (fabs((g[i] - g[i - 1]) - step) > 0.01 * step) return 0;
            *g0 = g[0];
            *inv_step = 1.0 / step;
            return 1;
        };
        pl->inc_cr_uniform = uniform(d->inc_cr_grid_host, d->n_inc_cr, &pl->inc_cr_g0, &pl->inc_cr_inv_step);
        pl->wspd_cr_uniform = uniform(d->wspd_cr_grid_host, d->n_wspd_cr, &pl->wspd_cr_g0, &pl->wspd_cr_inv_step);
    }
    int *n_step_rows_dev = nullptr;
    if (has_cr && (rc = xs::check(cudaMalloc(&n_step_rows_dev, sizeof(int)), "cudaMalloc")) != XS_OK) return fail(rc);
    auto build = [&]() -> int {
        if (has_cr) {
            const int nmax = pl->n_wspd_cr > pl->n_inc_cr ? pl->n_wspd_cr : pl->n_inc_cr;
            XS_CUDA(cudaMemsetAsync(n_step_rows_dev, 0, sizeof(int), st));
            XS_LAUNCH(k_build_cr_tables, (int)ceil_div(nmax, 128), 128, 0, st, *pl, n_step_rows_dev);
            XS_CUDA(cudaMemcpyAsync(&pl->cr_step_rows, n_step_rows_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
        }
        if (!has_co) return XS_OK;
        XS_CUDA(cudaMemsetAsync(pl->first_nan, 0x7f, sizeof(int) * (size_t)pl->n_inc, st));  // 0x7f7f7f7f > any index
        if (pl->fast_ok) {
            XS_CUDA(cudaMemsetAsync(pl->slab_absmax, 0, sizeof(float) * (size_t)pl->n_inc, st));
            XS_LAUNCH(k_init_slab_range, (int)ceil_div(pl->n_inc, 256), 256, 0, st, *pl);
            XS_LAUNCH(k_build_scan, kNumSMs * 8, 256, 0, st, *pl);
            XS_LAUNCH(k_build_cell, kNumSMs * 8, 256, 0, st, *pl);
            XS_LAUNCH(k_build_rowtab, (int)ceil_div(pl->n_wspd_pad, 256), 256, 0, st, *pl);
            XS_LAUNCH(k_build_seed_rmax, (int)ceil_div((int64_t)pl->n_inc * 64, 256), 256, 0, st, *pl);
            XS_LAUNCH(k_build_chunk_ranges, (int)ceil_div((int64_t)pl->n_inc * pl->n_chunks * pl->n_groups * 32, 256), 256, 0, st, *pl);
        } else {
            XS_LAUNCH(k_find_first_nan, kNumSMs * 8, 256, 0, st, *pl);
        }
        XS_LAUNCH(k_fix_first_nan, (int)ceil_div(pl->n_inc, 256), 256, 0, st, *pl);
        return XS_OK;
    };
    rc = build();
    if (rc == XS_OK) rc = xs::check(cudaStreamSynchronize(st), "xs_plan_create sync");
    cudaFree(n_step_rows_dev);
    if (rc != XS_OK) return fail(rc);
    *out = pl;
    return XS_OK;
}

extern "C" size_t xs_invert_workspace_bytes(const xs_plan *pl, int64_t n_px, uint32_t flags) {
    if (!pl || n_px < 0) return 0;
    return ws_layout(pl->fast_ok ? pl->n_inc : 0, n_px, flags, pl->fast_ok ? sort_temp_bytes(n_px) : 0, pl->cr_step_rows > 0, nullptr, nullptr);
}

extern "C" int xs_timer_create(xs_timer **out) {
    if (!out) {
        set_error("xs_timer_create: null argument");
        return XS_E_INVALID;
    }
    xs_timer *t = new xs_timer();
    memset(t, 0, sizeof(*t));
    for (int i = 0; i < 3; ++i) {
        const int rc = xs::check(cudaEventCreate(&t->ev[i]), "cudaEventCreate");
        if (rc != XS_OK) {
            xs_timer_destroy(t);
            return rc;
        }
    }
    *out = t;
    return XS_OK;
}
extern "C" void xs_timer_destroy(xs_timer *t) {
    if (!t) return;
    for (int i = 0; i < 3; ++i)
        if (t->ev[i]) cudaEventDestroy(t->ev[i]);
    delete t;
}
extern "C" int xs_timer_elapsed_ms(xs_timer *t, float ms[2]) {
    if (!t || !ms || !t->recorded) {
        set_error("xs_timer_elapsed_ms: the timer has not been recorded by an xs_invert with a co-pol scan");
        return XS_E_INVALID;
    }
    XS_CUDA(cudaEventSynchronize(t->ev[2]));
    XS_CUDA(cudaEventElapsedTime(&ms[0], t->ev[0], t->ev[1]));
    XS_CUDA(cudaEventElapsedTime(&ms[1], t->ev[1], t->ev[2]));
    return XS_OK;
}

extern "C" int xs_invert(const xs_plan *pl, const xs_invert_args *ar, void *stream) {
    if (!pl || !ar) {
        set_error("xs_invert: null argument");
        return XS_E_INVALID;
    }
    const int64_t n = ar->n_px;
    if (n < 0 || n > 0x7fffffffLL) {
        set_error("xs_invert: n_px out of range (0 .. 2^31-1)");
        return XS_E_INVALID;
    }
    if (n == 0) return XS_OK;
    if (!ar->inc || (ar->dtype != XS_F64 && ar->dtype != XS_F32) || (ar->mode != XS_MODE_FAST && ar->mode != XS_MODE_FP64)) {
        set_error("xs_invert: invalid argument (inc/dtype/mode)");
        return XS_E_INVALID;
    }
    const bool co_run = pl->n_inc > 0 && ar->sigma0_co && ar->ancillary;
    if (pl->n_inc > 0 && ar->sigma0_co && !ar->out_co) {
        set_error("xs_invert: out_co is required when a co-pol raster is inverted");
        return XS_E_INVALID;
    }
    if ((ar->flags & XS_FLAG_CR_ABS) && (ar->flags & XS_FLAG_MERGE_DUAL)) {
        set_error("xs_invert: XS_FLAG_CR_ABS and XS_FLAG_MERGE_DUAL are exclusive");
        return XS_E_INVALID;
    }
    if ((ar->flags & (XS_FLAG_DIR_METEO | XS_FLAG_OUT_F32)) && !(ar->flags & XS_FLAG_OUT_SPEED_DIR)) {
        set_error("xs_invert: XS_FLAG_DIR_METEO / XS_FLAG_OUT_F32 need XS_FLAG_OUT_SPEED_DIR");
        return XS_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    RasterArgs ra;
    ra.inc = ar->inc;
    ra.s_co = pl->n_inc > 0 ? ar->sigma0_co : nullptr;
    ra.s_cr = pl->n_inc_cr > 0 ? ar->sigma0_cr : nullptr;
    ra.dsig_cr = ar->dsig_cr;
    ra.anc = ar->ancillary;
    ra.dsig_cr_scalar = ar->dsig_cr_scalar;
    ra.dtype = ar->dtype;
    ra.flags = ar->flags;
    int sms = kNumSMs;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);

    const size_t sort_bytes = pl->fast_ok ? sort_temp_bytes(n) : 0;  // host-side query of CUB, no launch
    const size_t need = ws_layout(pl->fast_ok ? pl->n_inc : 0, n, ar->flags, sort_bytes, pl->cr_step_rows > 0, nullptr, nullptr);
    if (!ar->workspace || ar->workspace_bytes < need) {
        set_error("xs_invert: workspace too small (%zu < %zu)", ar->workspace_bytes, need);
        return XS_E_WORKSPACE;
    }
    if (((uintptr_t)ar->workspace & 255) != 0) {
        set_error("xs_invert: workspace must be 256-byte aligned");
        return XS_E_INVALID;
    }
    Workspace ws;
    ws_layout(pl->fast_ok ? pl->n_inc : 0, n, ar->flags, sort_bytes, pl->cr_step_rows > 0, (char *)ar->workspace, &ws);
    OutSpec out;
    out.co = ar->out_co;
    out.cr = ar->out_cr;
    out.idx_co = ar->idx_co;
    out.idx_cr = ar->idx_cr;
    out.gh = ar->ground_heading;
    out.gh_scalar = ar->ground_heading_scalar;
    out.n_px = n;
    out.flags = ar->flags;
    out.dtype = ar->dtype;
    // plane outputs: the co-pol pass leaves only its argmin, k_cross writes both outputs from it
    if ((ar->flags & XS_FLAG_OUT_SPEED_DIR) && !out.idx_co) out.idx_co = ws.idx_tmp;

    // counters + hist are contiguous at the start of the workspace
    XS_CUDA(cudaMemsetAsync(ws.counters, 0, (char *)ws.bin_start - (char *)ws.counters, st));
    if (co_run) {
        const bool fast = ar->mode == XS_MODE_FAST && pl->fast_ok;
        if (fast) {
            const int tile_px = scan_tile_px(pl->kp);
            const int bin_grid = (int)ceil_div(n, kBinPxPerCta);
            int bin_bits = 1;
            while ((1 << bin_bits) < pl->n_inc + 1) ++bin_bits;  // + 1: the all-ones key of unlisted pixels stays the largest
            XS_LAUNCH(k_bin_keys, bin_grid, kBinThreads, sizeof(unsigned) * pl->n_inc, st, *pl, ra, n, ws, bin_bits);
            XS_LAUNCH(k_bin_offsets, 1, 1024, 0, st, pl->n_inc, tile_px, ws);
            int which = 0;
            int rc = sort_pairs_u32(ws.key, ws.val, n, ws.sort_temp, ws.sort_temp_bytes, st, &which);
            if (rc != XS_OK) return rc;
            rc = launch_scan_pipeline(pl, ra, ws, ws.val[which], out, n, ar->scan_timer, st);
            if (rc != XS_OK) return rc;
            XS_LAUNCH(k_exact, sms * 8, 256, 0, st, *pl, ra, n, ws.fallback, ws.counters + 1, out);
        } else {
            XS_LAUNCH(k_exact, sms * 8, 256, 0, st, *pl, ra, n, (const unsigned *)nullptr, (const u64 *)nullptr, out);
        }
    }
    {
        const int64_t warps_needed = ceil_div(n, 32);
        int64_t grid = ceil_div(warps_needed * 32, 256);
        const int64_t cap = (int64_t)sms * 16;
        if (grid > cap) grid = cap;
        // cross-pol only on strictly increasing LUT rows: the step-function kernel settles almost every pixel, k_cross
        // takes the rest from a list (XS_FLAG_CR_FULL_SCAN keeps the cooperative scan for all of them: tests)
        const bool step = !ra.s_co && ra.s_cr && pl->cr_step_rows > 0 && !(ar->flags & XS_FLAG_CR_FULL_SCAN) && ws.fallback;
        if (step) {
            int64_t g1 = ceil_div(n, 256);
            if (g1 > (int64_t)sms * 64) g1 = (int64_t)sms * 64;
            XS_LAUNCH(k_cross_only, (int)g1, 256, 0, st, *pl, ra, n, out, ws.fallback, ws.counters + 4);
            if (grid > (int64_t)sms * 4) grid = (int64_t)sms * 4;
            XS_LAUNCH(k_cross, (int)grid, 256, 0, st, *pl, ra, n, out, (const unsigned *)ws.fallback, (const u64 *)(ws.counters + 4));
        } else
            XS_LAUNCH(k_cross, (int)grid, 256, 0, st, *pl, ra, n, out, (const unsigned *)nullptr, (const u64 *)nullptr);
    }
    if (ar->counters_dev)
        XS_CUDA(cudaMemcpyAsync(ar->counters_dev, ws.counters, XS_N_COUNTERS * sizeof(u64), cudaMemcpyDeviceToDevice, st));
    return XS_OK;
}
