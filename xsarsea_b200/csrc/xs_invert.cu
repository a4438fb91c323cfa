// K1: wind inversion (reference windspeed/windspeed.py:132-331) for B200 / sm_100a.
//
// Per pixel the reference evaluates a cost J over the whole wspd x phi grid of the LUT slab of the pixel's
// incidence bin and takes np.argmin (windspeed.py:212-232).  J is a squared distance in 3-D:
//     J(w,phi) = ((w cos phi - a)/2)^2 + ((w sin phi - b)/2)^2 + ((L[inc][w][phi] - s)/dsig_co)^2
// Dropping the per-pixel constant (a^2+b^2)/4 leaves
//     J'(w,phi) = d^2 + t,   d = L/dsig_co - s/dsig_co,   t = w^2/4 - (w/2) g(phi),   g = a cos phi + b sin phi
// which costs three FP32 FMA-pipe operations per candidate (FADD, FFMA, FFMA) and, with the packed
// f32x2 forms of sm_100 (FADD2/FFMA2) plus the 3-input FMNMX3, two issue slots per candidate.
//
// Pipeline of one xs_invert call (all on the caller's stream, no host synchronisation):
//   k_bin_count / k_bin_offsets / k_bin_scatter   counting sort of the co-pol pixels by incidence bin
//   k_scan_co      persistent CTAs (4 per SM, 4 warps each) taking tiles from an atomic counter; a tile = 32 pixels
//                  of one bin; the bin's slab (scan image, FP32) is streamed through a 3-stage shared-memory ring by
//                  bulk-async (TMA) copies, 16 rows at a time; lane l owns the phi pairs {2(l+32j), 2(l+32j)+1}; each
//                  warp scans 8 pixels at once keeping per lane and pixel the best 16-row chunk and the runner-up
//                  chunk minimum; then per pixel: warp-shuffle min, rigorous FP32 error band, the candidates inside
//                  the band are found by re-creating the FP32 costs of the contending (lane, chunk) cells; a single
//                  member settles the pixel, several are re-evaluated in FP64 (reference operation order) and reduced
//                  by a warp-shuffle lexicographic (J, index) argmin.
//   k_exact        exhaustive FP64 scan (warp per pixel) of the pixels the fast path cannot handle (non-finite
//                  inputs, magnitudes outside the error bound's range) and of XS_MODE_FP64
//   k_cross        cross-pol / dual-pol pass (windspeed.py:252-279), merge (:426-428), NaN classes
#include <stdlib.h>
#include <string.h>

#include <cmath>

#include "xs_invert.cuh"

namespace xs {

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

// ---- workspace layout ----------------------------------------------------------------------------------------
// counters (u64): [0] n_tiles, [1] fallback count, [2] pixels scanned by k_scan_co, [3] chunks re-evaluated in
// FP64, [4..7] phase timers of the instrumented variant, [8] dynamic tile counter
struct Workspace {
    u64 *counters;         // [16]
    unsigned *hist;        // [n_inc]
    unsigned *bin_start;   // [n_inc + 1]
    unsigned *cursor;      // [n_inc]
    unsigned *tile_start;  // [n_inc + 1]
    unsigned *list;        // [n_px] pixel indices grouped by bin
    unsigned *fallback;    // [n_px] pixels for k_exact_list
};
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static size_t ws_layout(int n_inc, int64_t n_px, char *base, Workspace *w) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char *p = base ? base + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    char *c = take(16 * sizeof(u64));
    char *h = take(sizeof(unsigned) * (size_t)(n_inc + 1));
    char *bs = take(sizeof(unsigned) * (size_t)(n_inc + 1));
    char *cu = take(sizeof(unsigned) * (size_t)(n_inc + 1));
    char *ts = take(sizeof(unsigned) * (size_t)(n_inc + 1));
    char *li = take(sizeof(unsigned) * (size_t)n_px);
    char *fb = take(sizeof(unsigned) * (size_t)n_px);
    if (w) {
        w->counters = (u64 *)c;
        w->hist = (unsigned *)h;
        w->bin_start = (unsigned *)bs;
        w->cursor = (unsigned *)cu;
        w->tile_start = (unsigned *)ts;
        w->list = (unsigned *)li;
        w->fallback = (unsigned *)fb;
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
            else if (!isinf(v))
                atomicMax(reinterpret_cast<unsigned *>(&pl.slab_absmax[bin]), __float_as_uint(fabsf(v)));
        }
        pl.scan[i] = v;
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
// plans without a scan image still need the first-NaN table
__global__ void k_find_first_nan(xs_plan pl) {
    const int64_t per_slab = (int64_t)pl.n_wspd * pl.n_phi;
    const int64_t n = per_slab * pl.n_inc;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (isnan(pl.co_lut[i])) atomicMin(&pl.first_nan[i / per_slab], (int)(i % per_slab));
}
__global__ void k_build_cr_tables(xs_plan pl) {
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
        pl.cr_finite[i] = ok | ((ok & mono) << 1);
        pl.cr_absmax[i] = amax;
    }
}
__global__ void k_fix_first_nan(xs_plan pl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pl.n_inc && pl.first_nan[i] == 0x7f7f7f7f) pl.first_nan[i] = -1;
}

// ---- counting sort of the co-pol pixels by incidence bin -----------------------------------------------------
constexpr int kBinThreads = 256;
constexpr int kBinPxPerCta = 256 * 32;

__global__ void __launch_bounds__(kBinThreads) k_bin_count(xs_plan pl, RasterArgs a, int64_t n_px, Workspace ws) {
    extern __shared__ unsigned sh_hist[];
    for (int b = threadIdx.x; b < pl.n_inc; b += blockDim.x) sh_hist[b] = 0;
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * kBinPxPerCta;
    const int64_t hi = min(lo + (int64_t)kBinPxPerCta, n_px);
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        int bin;
        if (pixel_co_bin(pl, a, i, &bin)) atomicAdd(&sh_hist[bin], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < pl.n_inc; b += blockDim.x)
        if (sh_hist[b]) atomicAdd(&ws.hist[b], sh_hist[b]);
}

// single CTA: exclusive scans of the bin counts (pixels and tiles)
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
            const unsigned ep = carry_px + sh_px[threadIdx.x] - cnt, et = carry_tiles + sh_tl[threadIdx.x] - tiles;
            ws.bin_start[b] = ep;
            ws.cursor[b] = ep;
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
        ws.bin_start[n_inc] = carry_px;
        ws.tile_start[n_inc] = carry_tiles;
        ws.counters[0] = carry_tiles;
    }
}

__global__ void __launch_bounds__(kBinThreads) k_bin_scatter(xs_plan pl, RasterArgs a, int64_t n_px, Workspace ws) {
    extern __shared__ unsigned sh[];  // [n_inc] counts -> bases, [n_inc] local cursors
    unsigned *sh_base = sh, *sh_cur = sh + pl.n_inc;
    for (int b = threadIdx.x; b < pl.n_inc; b += blockDim.x) {
        sh_base[b] = 0;
        sh_cur[b] = 0;
    }
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * kBinPxPerCta;
    const int64_t hi = min(lo + (int64_t)kBinPxPerCta, n_px);
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        int bin;
        if (pixel_co_bin(pl, a, i, &bin)) atomicAdd(&sh_base[bin], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < pl.n_inc; b += blockDim.x)
        if (sh_base[b]) sh_base[b] = atomicAdd(&ws.cursor[b], sh_base[b]);
    __syncthreads();
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        int bin;
        if (pixel_co_bin(pl, a, i, &bin)) ws.list[sh_base[bin] + atomicAdd(&sh_cur[bin], 1u)] = (unsigned)i;
    }
}

// ---- local sort of the pixel list by sigma0 (centred scan flavour only) ----------------------------------------
// The centred flavour of k_scan_co shares (L - c)^2 between the 8 pixels of a warp, which keeps its error band tight only
// if those pixels have similar sigma0.  Every run of 8 consecutive tiles (<= 256 list entries, possibly across a bin
// boundary) is therefore sorted by (incidence bin, sigma0) in shared memory: bins stay contiguous and in order, and a
// warp's 8 pixels span 1/32 of the run's sigma0 range.
constexpr int kSortRun = 256;
__global__ void __launch_bounds__(kSortRun) k_list_localsort(xs_plan pl, RasterArgs a, Workspace ws, int tile_px) {
    __shared__ unsigned long long key[kSortRun];
    __shared__ unsigned val[kSortRun];
    __shared__ unsigned range[2];
    const unsigned n_tiles = (unsigned)ws.counters[0];
    const unsigned tiles_per_run = kSortRun / tile_px;
    const unsigned t0 = blockIdx.x * tiles_per_run;
    if (t0 >= n_tiles) return;
    if (threadIdx.x < 2) {
        const unsigned t = threadIdx.x == 0 ? t0 : min(t0 + tiles_per_run, n_tiles);
        unsigned pos;
        if (t >= n_tiles)
            pos = ws.bin_start[pl.n_inc];
        else {
            int lo = 0, hi = pl.n_inc;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (ws.tile_start[mid] <= t)
                    lo = mid;
                else
                    hi = mid;
            }
            pos = ws.bin_start[lo] + (t - ws.tile_start[lo]) * tile_px;
        }
        range[threadIdx.x] = pos;
    }
    __syncthreads();
    const unsigned first = range[0], count = range[1] - range[0];  // count <= kSortRun
    unsigned long long k = ~0ull;
    unsigned v = 0;
    if (threadIdx.x < count) {
        const unsigned e = first + threadIdx.x;
        v = ws.list[e];
        int lo = 0, hi = pl.n_inc;  // bin of list position e: last b with bin_start[b] <= e
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (ws.bin_start[mid] <= e)
                lo = mid;
            else
                hi = mid;
        }
        const float sf = (float)load_real(a.s_co, v, a.dtype);  // linear or dB: monotone either way
        unsigned b = __float_as_uint(sf);
        b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // order-preserving map of the float bits
        k = ((unsigned long long)lo << 32) | b;
    }
    key[threadIdx.x] = k;
    val[threadIdx.x] = v;
    __syncthreads();
    for (int size = 2; size <= kSortRun; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const int i = threadIdx.x, j = i ^ stride;
            if (j > i) {
                const bool up = (i & size) == 0;
                const unsigned long long ki = key[i], kj = key[j];
                if ((ki > kj) == up) {
                    key[i] = kj;
                    key[j] = ki;
                    const unsigned t = val[i];
                    val[i] = val[j];
                    val[j] = t;
                }
            }
            __syncthreads();
        }
    if (threadIdx.x < count) ws.list[first + threadIdx.x] = val[threadIdx.x];
}

// ---- co-pol result of one pixel ----------------------------------------------------------------------------
// windspeed.py:231-247.  The reference picks +phi or -phi by comparing |angle(anc/sol)| and |angle(anc/sol2)|
// (ties keep +phi); for phi in [0,180] that is Im(anc) >= 0 ("direction sign").
__device__ __forceinline__ void write_co(const xs_plan &pl, int idx, double2 anc, int64_t px, double2 *out_co,
                                         int *idx_co) {
    const int iw = idx / pl.n_phi, ip = idx - iw * pl.n_phi;
    const double w = pl.wspd_grid[iw];
    double re = w * pl.cos_phi[ip], im = w * pl.sin_phi[ip];
    if (pl.phi_180 && anc.y < 0.0) im = -im;
    out_co[px] = make_double2(re, im);
    if (idx_co) idx_co[px] = idx;
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
                                               const u64 *list_count, double2 *out_co, int *idx_co) {
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
        if (lane == 0) write_co(pl, idx, p.anc, px, out_co, idx_co);
    }
}

// ---- the FP32 scan --------------------------------------------------------------------------------------------
struct PixelSlot {  // per-pixel state kept in shared memory during a tile
    double qa, qb, s;  // m_antenna, m_azi (|.| if phi_180), sigma0 dB
    double anc_im;
    unsigned px;
    int state;  // 0: empty slot, 1: scan, 2: result known (NaN-slab shortcut), 3: exhaustive FP64 needed
    int idx;
    float amag;  // |ancillary| rounded up (error-bound input)
};

template <int P, int NT, bool ON>
struct BookSmem {  // per-thread argmin bookkeeping in shared memory (experimental variant)
    float best[P][NT];
    float second[P][NT];
    int bchunk[P][NT];
};
template <int P, int NT>
struct BookSmem<P, NT, false> {
    float best[1][1];
    float second[1][1];
    int bchunk[1][1];
};

template <int KP, int P, int NW, int NS, bool BK>
struct ScanSmem {
    static constexpr int kRowFloats = 64 * KP;
    static constexpr int kChunkBytes = kChunkRows * kRowFloats * 4;
    alignas(128) float ring[NS][kChunkRows * kRowFloats];
    alignas(16) uint64_t full[NS];
    alignas(16) uint64_t empty[NS];
    PixelSlot px[NW * P];
    unsigned next_tile[2];
    BookSmem<P, NW * 32, BK> book;
};

// kMath == 4: instrumented build (clock64 per phase, summed over warps into ws.counters[4..7]); math as flavour 0
#define XS_TICK(slot)                                              \
    do {                                                           \
        if (kMath == 4 || kMath == 8) {                            \
            const long long _now = clock64();                      \
            if (lane == 0) t_acc[slot] += (u64)(_now - t_last);    \
            t_last = _now;                                         \
        }                                                          \
    } while (0)

template <int KP, int P, int NW, int MB, int kMath = 0, bool kBookSmem = false, int NS = kStages>
__global__ void __launch_bounds__(NW * 32, MB)
k_scan_co(xs_plan pl, RasterArgs a, Workspace ws, double2 *out_co, int *idx_co) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using Smem = ScanSmem<KP, P, NW, NS, kBookSmem>;
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    float2 *rowtab_s = reinterpret_cast<float2 *>(smem_raw + sizeof(Smem));  // [n_wspd_pad]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int TP = NW * P;  // pixels per tile
    constexpr bool kCentred = kMath == 5 || kMath == 6 || kMath == 7 || kMath == 8;  // 8: centred + phase timers  // two-FFMA2 centred cost (see the settle section)
    constexpr bool kWarpOwn = kMath == 6;
    constexpr bool kJOuter = kMath == 7;  // experiment: phi-pair loop outside the pixel loop (shorter live ranges of lambda, M)  // every warp loads and writes its own P pixels: one CTA barrier per tile

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < pl.n_wspd_pad; i += blockDim.x) rowtab_s[i] = pl.rowtab[i];
    __syncthreads();

    const unsigned n_tiles = (unsigned)ws.counters[0];
    const int n_chunks = (pl.n_wspd_pad + kChunkRows - 1) / kChunkRows;  // the last chunk may be shorter
    unsigned it = 0;  // chunks consumed so far by this CTA (ring position, continues across tiles)
    u64 n_scanned = 0, n_refined = 0;

    u64 t_acc[6] = {0, 0, 0, 0, 0, 0};  // prologue (incl. barriers), main loop, refinement C, write-out (incl. barrier), refinement A, B
    long long t_last = clock64();
    // Tiles are handed out dynamically (one atomic per tile): CTAs do not all progress at the same speed, and a static
    // round-robin left ~13 % of the SM time idle at the end of the kernel.
    if (kWarpOwn) {
        if (threadIdx.x == 0) sm.next_tile[0] = (unsigned)atomicAdd(&ws.counters[8], 1ull);
        __syncthreads();
    }
    for (unsigned tile_it = 0;; ++tile_it) {
        unsigned tile;
        if (kWarpOwn) {
            // the tile index was fetched during the previous tile (double-buffered), the slots are per warp: the single
            // barrier at the end of the loop body is the only CTA-wide synchronisation of a tile
            tile = sm.next_tile[tile_it & 1];
            if (tile >= n_tiles) break;
            if (threadIdx.x == 0) sm.next_tile[(tile_it + 1) & 1] = (unsigned)atomicAdd(&ws.counters[8], 1ull);
        } else {
            __syncthreads();  // previous tile's slots and sm.next_tile are no longer read
            if (threadIdx.x == 0) sm.next_tile[0] = (unsigned)atomicAdd(&ws.counters[8], 1ull);
            __syncthreads();
            tile = sm.next_tile[0];
            if (tile >= n_tiles) break;
        }
        // tile -> (bin, pixel range): last bin with tile_start[bin] <= tile
        int lo = 0, hi = pl.n_inc;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (ws.tile_start[mid] <= tile)
                lo = mid;
            else
                hi = mid;
        }
        const int bin = lo;
        const unsigned first = ws.bin_start[bin] + (tile - ws.tile_start[bin]) * TP;
        const unsigned last = min(first + TP, ws.bin_start[bin + 1]);
        const int nan_idx = pl.first_nan[bin];

        // ---- load the tile's pixels (one thread per pixel; with kWarpOwn lanes 0..P-1 of every warp load its own) ----
        const int my_slot = kWarpOwn ? warp * P + lane : (int)threadIdx.x;
        if (kWarpOwn ? lane < P : threadIdx.x < TP) {
            PixelSlot sl;
            sl.state = 0;
            sl.px = 0;
            sl.idx = -1;
            sl.qa = sl.qb = sl.s = sl.anc_im = 0.0;
            sl.amag = 0.f;
            const unsigned e = first + my_slot;
            if (e < last) {
                const unsigned px = ws.list[e];
                // only what the co-pol scan needs (load_pixel would also convert the cross-pol sigma0 to dB)
                const double2 anc = load_cplx(a.anc, px, a.dtype);
                const double s_raw = load_real(a.s_co, px, a.dtype);
                sl.px = px;
                sl.qa = anc.x;
                sl.anc_im = anc.y;
                sl.qb = pl.phi_180 ? fabs(anc.y) : anc.y;
                sl.s = (a.flags & XS_FLAG_SIGMA0_DB) ? s_raw : to_db(s_raw);
                sl.amag = (float)hypot(sl.qa, sl.qb) * 1.0000002f;
                const bool finite_q = isfinite(sl.qa) && isfinite(sl.qb) && isfinite(sl.s);
                if (!finite_q)
                    sl.state = 3;
                else if (nan_idx >= 0) {
                    sl.state = 2;
                    sl.idx = nan_idx;
                } else
                    sl.state = 1;
            }
            sm.px[my_slot] = sl;
        }
        if (kWarpOwn)
            __syncwarp();
        else
            __syncthreads();

        if (nan_idx < 0) {
            // ---- per-lane per-pixel query constants ----
            u64 g[P][KP];   // {g(phi_even), g(phi_odd)} as packed FP32
            float nqs[P];   // -s/dsig_co
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                const int ip0 = 2 * (lane + 32 * j);
                double c0 = 0, s0 = 0, c1 = 0, s1 = 0;
                if (ip0 < pl.n_phi) {
                    c0 = pl.cos_phi[ip0];
                    s0 = pl.sin_phi[ip0];
                }
                if (ip0 + 1 < pl.n_phi) {
                    c1 = pl.cos_phi[ip0 + 1];
                    s1 = pl.sin_phi[ip0 + 1];
                }
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const PixelSlot &sl = sm.px[warp * P + p];
                    const bool on = sl.state == 1;
                    const float g0 = on ? g32(sl.qa, sl.qb, c0, s0) : 0.f;
                    const float g1 = on ? g32(sl.qa, sl.qb, c1, s1) : 0.f;
                    g[p][j] = pack2(g0, g1);
                }
            }
            // kCentred (centred flavour): cs = centre of the warp's sigma0/dsig values; nqs[p] then holds k_p = -2 (s_p/dsig
            // - cs) and scabs[p] an upper bound of |s_p/dsig - cs|
            float cs = 0.f, scabs[P];
            if (kCentred) {
                double smin = CUDART_INF, smax = -CUDART_INF;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const PixelSlot &sl = sm.px[warp * P + p];
                    if (sl.state == 1) {
                        const double v = sl.s / pl.dsig_co;
                        smin = fmin(smin, v);
                        smax = fmax(smax, v);
                    }
                }
                cs = smin <= smax ? (float)(0.5 * (smin + smax)) : 0.f;
            }
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const PixelSlot &sl = sm.px[warp * P + p];
                scabs[p] = 0.f;
                if (kCentred) {
                    const double sc = sl.state == 1 ? sl.s / pl.dsig_co - (double)cs : 0.0;
                    nqs[p] = (float)(-2.0 * sc);
                    scabs[p] = (float)fabs(sc) * 1.0000002f;
                } else
                    nqs[p] = sl.state == 1 ? (float)(-(sl.s / pl.dsig_co)) : 0.f;
            }

            float m[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                m[p] = CUDART_INF_F;
                if (kBookSmem) {
                    sm.book.best[kBookSmem ? p : 0][kBookSmem ? threadIdx.x : 0] = CUDART_INF_F;
                    sm.book.second[kBookSmem ? p : 0][kBookSmem ? threadIdx.x : 0] = CUDART_INF_F;
                    sm.book.bchunk[kBookSmem ? p : 0][kBookSmem ? threadIdx.x : 0] = 0;
                }
            }
            float best[P], second[P];
            int bchunk[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                best[p] = CUDART_INF_F;
                second[p] = CUDART_INF_F;
                bchunk[p] = 0;
            }

            const float *slab = pl.scan + (int64_t)bin * pl.n_wspd_pad * pl.nph_pad;
            // ---- producer prologue: fill the ring ----
            if (threadIdx.x == 0) {
                const int pre = min(NS - 1, n_chunks);
                for (int c = 0; c < pre; ++c) {
                    const unsigned g_it = it + c;
                    const int s = g_it % NS;
                    if (g_it >= NS) mbar_wait(&sm.empty[s], ((g_it / NS) - 1) & 1);
                    const uint32_t bytes = (uint32_t)min(kChunkRows, pl.n_wspd_pad - c * kChunkRows) * Smem::kRowFloats * 4;
                    mbar_expect_tx(&sm.full[s], bytes);
                    bulk_g2s(sm.ring[s], slab + (int64_t)c * kChunkRows * Smem::kRowFloats, bytes, &sm.full[s]);
                }
            }
            XS_TICK(0);
            // ---- main loop over 16-row chunks ----
            for (int c = 0; c < n_chunks; ++c) {
                const unsigned g_it = it + c;
                const int s = g_it % NS;
                if (threadIdx.x == 0 && c + NS - 1 < n_chunks) {  // refill the slot consumed last iteration
                    const unsigned n_it = g_it + NS - 1;
                    const int ns = n_it % NS;
                    if (n_it >= NS) mbar_wait(&sm.empty[ns], ((n_it / NS) - 1) & 1);
                    const int nc = c + NS - 1;
                    const uint32_t bytes = (uint32_t)min(kChunkRows, pl.n_wspd_pad - nc * kChunkRows) * Smem::kRowFloats * 4;
                    mbar_expect_tx(&sm.full[ns], bytes);
                    bulk_g2s(sm.ring[ns], slab + (int64_t)nc * kChunkRows * Smem::kRowFloats, bytes, &sm.full[ns]);
                }
                __syncwarp();
                mbar_wait(&sm.full[s], (g_it / NS) & 1);
                const u64 *rows = reinterpret_cast<const u64 *>(sm.ring[s]);
                const int rows_here = min(kChunkRows, pl.n_wspd_pad - c * kChunkRows);  // even (n_wspd_pad is a multiple of 8)
#pragma unroll 2
                for (int r = 0; r < rows_here; ++r) {
                    const float2 rt = rowtab_s[c * kChunkRows + r];
                    const u64 nwh = pack2(rt.x, rt.x), w2q = pack2(rt.y, rt.y);
                    u64 L[KP];
#pragma unroll
                    for (int j = 0; j < KP; ++j) L[j] = rows[r * (32 * KP) + lane + 32 * j];
                    if (kJOuter) {
                        const u64 ncs2 = pack2(-cs, -cs);
#pragma unroll
                        for (int j = 0; j < KP; ++j) {
                            const u64 lc = fadd2(L[j], ncs2);
                            const u64 mj = ffma2(lc, lc, w2q);
#pragma unroll
                            for (int p = 0; p < P; ++p) {
                                const u64 aa = ffma2(pack2(nqs[p], nqs[p]), lc, mj);
                                const u64 J = ffma2(nwh, g[p][j], aa);
                                float j0, j1;
                                unpack2(J, j0, j1);
                                m[p] = fmin3(m[p], j0, j1);
                            }
                        }
                        continue;
                    }
                    u64 M[KP];
                    if (kCentred) {  // shared by the warp's pixels: Lc = L - cs, M = Lc^2 + w^2/4
                        const u64 ncs2 = pack2(-cs, -cs);
#pragma unroll
                        for (int j = 0; j < KP; ++j) {
                            L[j] = fadd2(L[j], ncs2);
                            M[j] = ffma2(L[j], L[j], w2q);
                        }
                    }
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const u64 q2 = pack2(nqs[p], nqs[p]);
#pragma unroll
                        for (int j = 0; j < KP; ++j) {
                            float j0, j1;
                            if (kCentred) {  // J'' = k_p Lc + M + (-w/2) g: two FFMA2 per candidate pair
                                const u64 aa = ffma2(q2, L[j], M[j]);
                                const u64 J = ffma2(nwh, g[p][j], aa);
                                unpack2(J, j0, j1);
                            } else if (kMath == 1) {  // experiment: scalar FADD/FFMA instead of the packed f32x2 forms
                                float l0, l1, g0, g1;
                                unpack2(L[j], l0, l1);
                                unpack2(g[p][j], g0, g1);
                                const float d0 = l0 + nqs[p], d1 = l1 + nqs[p];
                                j0 = fmaf(d0, d0, fmaf(rt.x, g0, rt.y));
                                j1 = fmaf(d1, d1, fmaf(rt.x, g1, rt.y));
                            } else if (kMath == 2) {
                                // t by two scalar FFMA: their row constants (-w/2, w^2/4) are the same registers for
                                // every pixel and phi pair of the row, so they are served by the operand-reuse cache
                                // instead of the register file, which is what bounds this loop
                                float g0, g1;
                                unpack2(g[p][j], g0, g1);
                                const u64 t = pack2(fmaf(rt.x, g0, rt.y), fmaf(rt.x, g1, rt.y));
                                const u64 d = fadd2(L[j], q2);
                                const u64 J = ffma2(d, d, t);
                                unpack2(J, j0, j1);
                            } else {
                                const u64 d = fadd2(L[j], q2);
                                const u64 t = ffma2(nwh, g[p][j], w2q);
                                const u64 J = ffma2(d, d, t);
                                unpack2(J, j0, j1);
                            }
                            m[p] = fmin3(m[p], j0, j1);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty[s]);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    if (kBookSmem) {
                        const float b = sm.book.best[kBookSmem ? p : 0][kBookSmem ? threadIdx.x : 0];
                        sm.book.second[kBookSmem ? p : 0][kBookSmem ? threadIdx.x : 0] = fminf(sm.book.second[kBookSmem ? p : 0][kBookSmem ? threadIdx.x : 0], fmaxf(b, m[p]));
                        if (m[p] < b) {
                            sm.book.best[kBookSmem ? p : 0][kBookSmem ? threadIdx.x : 0] = m[p];
                            sm.book.bchunk[kBookSmem ? p : 0][kBookSmem ? threadIdx.x : 0] = c;
                        }
                    } else {
                        const bool lt = m[p] < best[p];
                        second[p] = fminf(second[p], fmaxf(best[p], m[p]));
                        best[p] = fminf(best[p], m[p]);
                        bchunk[p] = lt ? c : bchunk[p];
                    }
                    m[p] = CUDART_INF_F;
                }
            }
            it += n_chunks;
            XS_TICK(1);
            if (kBookSmem) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    best[p] = sm.book.best[kBookSmem ? p : 0][kBookSmem ? threadIdx.x : 0];
                    second[p] = sm.book.second[kBookSmem ? p : 0][kBookSmem ? threadIdx.x : 0];
                    bchunk[p] = sm.book.bchunk[kBookSmem ? p : 0][kBookSmem ? threadIdx.x : 0];
                }
            }

            // ---- settle the warp's pixels ---------------------------------------------------------------------
            // m32 = warp-shuffle min of the FP32 costs; E bounds |J'_fp32 - J'_exact| for every candidate that can still
            // win, so the reference's FP64 argmin lies in S = {c : J'_fp32(c) <= m32 + 2E}.  S is collected by
            // re-creating the FP32 costs (bit-identical operations) of the (lane, chunk) cells whose minimum is inside
            // the band.  |S| = 1 settles the pixel with no FP64 work (9 pixels in 10); otherwise the members of S are
            // evaluated in FP64 with the reference's operation order and reduced lexicographically on (J, flat index).
            // The work is organised in phases across the P pixels so that their dependent loads overlap.
            const double *slab64 = pl.co_lut + (int64_t)bin * pl.n_wspd * pl.n_phi;
            const float *slab32 = pl.scan + (int64_t)bin * pl.n_wspd_pad * pl.nph_pad;
            const float lmax = pl.slab_absmax[bin];
            constexpr int kCand = kChunkRows * 2 * KP;  // candidates of one (lane, chunk) cell
            constexpr int kIter = (kCand + 31) / 32;

            // FP32 cost of candidate k of cell (L, row0) for pixel slot sl, exactly as the scan computed it
            auto member = [&](const PixelSlot &sl, float nq, float thr, int L, int row0, int k, int n_cand, int &flat) {
                const int iw = row0 + k / (2 * KP);
                const int slot = k % (2 * KP);
                const int ip = 2 * (L + 32 * (slot >> 1)) + (slot & 1);
                flat = iw * pl.n_phi + ip;
                if (k >= n_cand || iw >= pl.n_wspd || ip >= pl.n_phi) return false;
                const float2 rt = rowtab_s[iw];
                if (kCentred) {
                    const float lc = __fadd_rn(slab32[(int64_t)iw * pl.nph_pad + ip], -cs);
                    const float mm = __fmaf_rn(lc, lc, rt.y);
                    const float aa = __fmaf_rn(nq, lc, mm);
                    return __fmaf_rn(rt.x, g32(sl.qa, sl.qb, pl.cos_phi[ip], pl.sin_phi[ip]), aa) <= thr;
                }
                const float d = __fadd_rn(slab32[(int64_t)iw * pl.nph_pad + ip], nq);
                const float t = __fmaf_rn(rt.x, g32(sl.qa, sl.qb, pl.cos_phi[ip], pl.sin_phi[ip]), rt.y);
                return __fmaf_rn(d, d, t) <= thr;
            };

            // phase A: band of every pixel, contender masks
            float thr[P];
            unsigned cont[P], wide[P];
            bool act[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                PixelSlot &sl = sm.px[warp * P + p];
                act[p] = sl.state == 1;  // warp-uniform
                cont[p] = wide[p] = 0;
                thr[p] = 0.f;
                if (!act[p]) continue;
                if (kMath == 3) {  // measurement-only variant: no refinement (results are NOT exact)
                    if (lane == 0) {
                        sl.idx = bchunk[p] * kChunkRows * pl.n_phi;
                        sl.state = 2;
                    }
                    act[p] = false;
                    continue;
                }
                float m32 = best[p];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m32 = fminf(m32, __shfl_xor_sync(0xffffffffu, m32, o));
                // rigorous bound E
                const float A = sl.amag;
                const float W = (float)pl.w_absmax * 1.0000002f;
                const float T = W * A + 0.25f * W * W;
                float D, E;
                if (kCentred) {
                    // J'' = J' - sc^2 (sc = s/dsig - cs): candidates that can still win have |L/dsig - s/dsig| <= D and
                    // |L/dsig - cs| <= Lam = D + |sc|.  Error terms (u = 2^-24): image value and Lc roundings 2 Lam (lmax + Lam)
                    // through Lc^2 and 2 |sc| (lmax + 2 Lam) through k_p Lc; M, a and J roundings Lam^2 + W^2/4,
                    // Lam^2 + W^2/4 + 2 |sc| Lam and D^2 + sc^2 + T; row-table and g roundings W^2/4 + W A; the W terms
                    // add up to W^2 + 2 W A <= 4 T (derivation in DESIGN.md 4.1).
                    const float SC = scabs[p];
                    D = sqrtf(fmaxf(m32 + SC * SC * 1.0000002f, 0.f) + 0.25f * A * A + 1.0f);
                    const float Lam = D + SC;
                    E = 5.9604645e-8f * 1.5f * (2.f * Lam * lmax + 2.f * SC * lmax + 4.f * Lam * Lam + 6.f * SC * Lam + SC * SC + D * D + 4.f * T);
                } else {
                    D = sqrtf(fmaxf(m32, 0.f) + 0.25f * A * A + 1.0f);
                    const float Q = fabsf(nqs[p]);
                    E = 5.9604645e-8f * 1.5f * (3.f * T + 2.f * D * (lmax + Q + D) + (fabsf(m32) + 0.25f * A * A + T));
                }
                thr[p] = m32 + 2.f * E;
                const bool sane = (E < 0.25f) && (m32 < CUDART_INF_F);
                if (!sane) {  // warp-uniform: magnitudes outside the range the error bound was derived for
                    if (lane == 0) sl.state = 3;
                    act[p] = false;
                    continue;
                }
                cont[p] = __ballot_sync(0xffffffffu, best[p] <= thr[p]);
                // lanes holding two or more chunks inside the band: all of the lane's candidates are looked at
                wide[p] = __ballot_sync(0xffffffffu, second[p] <= thr[p]);
                cont[p] &= ~wide[p];
            }
            XS_TICK(4);
            // phase B: membership in S of the candidates of the first contender cell of every pixel (loads batched)
            bool in0[P][kIter];
            int flat0[P][kIter];
            unsigned rest[P];  // contender cells not looked at yet
#pragma unroll
            for (int p = 0; p < P; ++p) {
                rest[p] = cont[p];
#pragma unroll
                for (int q = 0; q < kIter; ++q) {
                    in0[p][q] = false;
                    flat0[p][q] = 0;
                }
                if (act[p] && cont[p]) {  // warp-uniform
                    const int L = __ffs(cont[p]) - 1;
                    rest[p] &= rest[p] - 1;
                    const int row0 = __shfl_sync(0xffffffffu, bchunk[p], L) * kChunkRows;
#pragma unroll
                    for (int q = 0; q < kIter; ++q)
                        in0[p][q] = member(sm.px[warp * P + p], nqs[p], thr[p], L, row0, lane + 32 * q, kCand, flat0[p][q]);
                    ++n_refined;
                }
            }
            XS_TICK(5);
            // phase C: count the members, look at the remaining cells (rare), settle
#pragma unroll
            for (int p = 0; p < P; ++p) {
                if (!act[p]) continue;
                PixelSlot &sl = sm.px[warp * P + p];
                // members of S seen by this lane so far (count, and the flat index of one of them)
                int n_loc = 0, one_loc = -1;
#pragma unroll
                for (int q = 0; q < kIter; ++q)
                    if (in0[p][q]) {
                        ++n_loc;
                        one_loc = flat0[p][q];
                    }
                ArgMin am;
                am.init();
                // generic walk over cell (L, row0, n_cand): note members (exact == false) or FP64 argmin (true);
                // no warp synchronisation inside, so the loads of successive iterations overlap
                auto visit = [&](int L, int row0, int n_cand, bool exact) {
                    for (int k0 = 0; k0 < n_cand; k0 += 32) {
                        int flat;
                        const bool in = member(sl, nqs[p], thr[p], L, row0, k0 + lane, n_cand, flat);
                        if (!in) continue;
                        if (!exact) {
                            ++n_loc;
                            one_loc = flat;
                        } else {
                            const int iw = flat / pl.n_phi, ip = flat - iw * pl.n_phi;
                            am.feed(exact_cost_co(pl.wspd_grid[iw], pl.cos_phi[ip], pl.sin_phi[ip], slab64[flat], sl.qa, sl.qb,
                                                  sl.s, pl.dsig_co), flat);
                        }
                    }
                };
                auto sweep = [&](unsigned cells, unsigned lanes, bool exact) {
                    while (cells) {
                        const int L = __ffs(cells) - 1;
                        cells &= cells - 1;
                        visit(L, __shfl_sync(0xffffffffu, bchunk[p], L) * kChunkRows, kCand, exact);
                        if (!exact) ++n_refined;
                    }
                    while (lanes) {
                        const int L = __ffs(lanes) - 1;
                        lanes &= lanes - 1;
                        visit(L, 0, pl.n_wspd * 2 * KP, exact);
                        if (!exact) n_refined += n_chunks;
                    }
                };
                if (rest[p] | wide[p]) sweep(rest[p], wide[p], false);
                const int n_in = __reduce_add_sync(0xffffffffu, n_loc);
                int result = __reduce_max_sync(0xffffffffu, one_loc);  // the member itself when n_in == 1
                if (n_in > 1) {
                    sweep(cont[p], wide[p], true);
                    am.warp_reduce();
                    result = am.result();
                }
                if (lane == 0) {
                    if (n_in >= 1) {
                        sl.idx = result;
                        sl.state = 2;
                    } else
                        sl.state = 3;  // cannot happen if the re-created costs equal the scan's; be safe
                }
                ++n_scanned;
            }
        }
        XS_TICK(2);
        if (kWarpOwn)
            __syncwarp();
        else
            __syncthreads();
        // ---- write results / queue leftovers ----
        if (kWarpOwn ? lane < P : threadIdx.x < TP) {
            const PixelSlot &sl = sm.px[my_slot];
            if (sl.state == 2)
                write_co(pl, sl.idx, make_double2(sl.qa, sl.anc_im), sl.px, out_co, idx_co);
            else if (sl.state == 3)
                ws.fallback[atomicAdd(&ws.counters[1], 1ull)] = sl.px;
        }
        XS_TICK(3);
        if (kWarpOwn) __syncthreads();  // next tile index visible; every warp is done with this tile's ring traffic
    }
    if (lane == 0) {
        if (n_scanned) atomicAdd(&ws.counters[2], n_scanned);
        if (n_refined) atomicAdd(&ws.counters[3], n_refined);
        if (kMath == 4 || kMath == 8)
        {
            for (int k = 0; k < 4; ++k) atomicAdd(&ws.counters[4 + k], t_acc[k]);
            atomicAdd(&ws.counters[9], t_acc[4]);
            atomicAdd(&ws.counters[10], t_acc[5]);
        }
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
__device__ __forceinline__ int cross_interval_search(const double *__restrict__ col, const double *__restrict__ wg, int n,
                                                     double s, double dsig, double mag, bool hc) {
    auto num_at = [=](int w) { return __dsub_rn(col[w], s); };  // ts = num/dsig has the sign of num (dsig > 0)
    auto tw_at = [=](int w) { return __dmul_rn(__dsub_rn(wg[w], mag), 0.5); };
    auto cost = [=](int w) { return exact_cost_cr(col[w], s, dsig, wg[w], mag, hc); };
    int lo = 0, hi = n;  // k = first w with num(w) >= 0
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (num_at(mid) < 0.0)
            lo = mid + 1;
        else
            hi = mid;
    }
    const int k = lo;
    double m0 = CUDART_INF;
    if (k < n) m0 = cost(k);
    if (k > 0) m0 = fmin(m0, cost(k - 1));
    int j = 0;
    if (hc) {
        lo = 0, hi = n;  // j = first w with tw(w) >= 0
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (tw_at(mid) < 0.0)
                lo = mid + 1;
            else
                hi = mid;
        }
        j = lo;
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
    // [first, last): candidates with a(w) <= m0 (a is non-increasing left of k, non-decreasing from k on)
    lo = 0, hi = k;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a_gt_m0(mid))
            lo = mid + 1;
        else
            hi = mid;
    }
    int first = lo;
    lo = k, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (!a_gt_m0(mid))
            lo = mid + 1;
        else
            hi = mid;
    }
    int last = lo;
    if (hc) {  // intersect with the candidates with b(w) <= m0
        lo = 0, hi = j;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const double t = tw_at(mid);
            if (__dmul_rn(t, t) > m0)
                lo = mid + 1;
            else
                hi = mid;
        }
        first = max(first, lo);
        lo = j, hi = n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const double t = tw_at(mid);
            if (__dmul_rn(t, t) <= m0)
                lo = mid + 1;
            else
                hi = mid;
        }
        last = min(last, lo);
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

// ---- cross-pol / dual-pol pass + NaN classes + merge --------------------------------------------------------
// windspeed.py:198-207 (NaN classes), :250 (no co-pol), :252-279 (cross-pol argmin), :422-428 (abs / merge).
// A warp takes 32 consecutive pixels: every lane does the per-pixel scalar work of its own pixel (dB prologue,
// incidence bin, |wind_co|), then the warp scans the wspd grid of one pixel after the other cooperatively
// (parameters broadcast by shuffle), and finally every lane writes its own pixel (coalesced).
__global__ void __launch_bounds__(256) k_cross(xs_plan pl, RasterArgs a, int64_t n_px, double2 *out_co, void *out_cr,
                                               int *idx_co, int *idx_cr) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double nan = CUDART_NAN;
    const unsigned full = 0xffffffffu;
    for (int64_t base = warp * 32; base < n_px; base += n_warps * 32) {
        const int64_t px = base + lane;
        const bool valid = px < n_px;
        Pixel p;
        p.cls = 0;
        p.co = 0;
        p.s_cr = p.dsig_cr = p.inc = nan;
        if (valid) p = load_pixel(pl, a, px);
        double2 co = make_double2(nan, 0.0), dual = make_double2(nan, 0.0);
        bool scan = false, has_co = false, filter_ok = false;
        int bin = 0, ix = -1;
        double mag = nan;
        if (valid && p.cls != 0) {
            co = p.co ? out_co[px] : make_double2(nan, nan);
            dual = make_double2(nan, nan);
            if (!isnan(p.s_cr) && !isnan(p.dsig_cr) && pl.n_inc_cr > 0) {
                scan = true;
                bin = nearest_bin(pl.inc_cr_grid, pl.n_inc_cr, p.inc, pl.inc_cr_sorted);
                mag = hypot(co.x, co.y);
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
            const int r = cross_interval_search(pl.cr_lut + (int64_t)bin * pl.n_wspd_cr, pl.wspd_cr_grid, pl.n_wspd_cr, p.s_cr,
                                                p.dsig_cr, mag, has_co);
            if (r >= 0) {
                ix = r;
                settled_own = true;
            }
        }
        unsigned todo = __ballot_sync(full, scan && !settled_own);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const double s_cr = __shfl_sync(full, p.s_cr, src), dsig = __shfl_sync(full, p.dsig_cr, src);
            const double mg = __shfl_sync(full, mag, src);
            const int b = __shfl_sync(full, bin, src);
            const bool hc = __shfl_sync(full, (int)has_co, src), fok = __shfl_sync(full, (int)filter_ok, src);
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
        if (!p.co && out_co) {
            out_co[px] = co;
            if (idx_co) idx_co[px] = -1;
        }
        if (idx_cr) idx_cr[px] = ix;
        if (out_cr) {
            double2 o = dual;
            if (a.flags & XS_FLAG_MERGE_DUAL) {
                const double aco = hypot(co.x, co.y), adu = hypot(dual.x, dual.y);
                if (aco < 5.0 || adu < 5.0) o = co;
            }
            if (a.flags & XS_FLAG_CR_ABS)
                reinterpret_cast<double *>(out_cr)[px] = hypot(o.x, o.y);
            else
                reinterpret_cast<double2 *>(out_cr)[px] = o;
        }
    }
}

template <int KP, int P, int NW, int MB, int SC = 0, bool BK = false, int NS = kStages>
static int launch_scan(const xs_plan *pl, const RasterArgs &ra, const Workspace &ws, double2 *out_co, int *idx_co,
                       void *stream) {
    const size_t smem = sizeof(ScanSmem<KP, P, NW, NS, BK>) + sizeof(float2) * (size_t)pl->n_wspd_pad;
    // per launch, not once per process: the attribute belongs to the current device's context
    XS_CUDA(cudaFuncSetAttribute(k_scan_co<KP, P, NW, MB, SC, BK, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (smem > 200 * 1024) {
        set_error("xs_invert: wspd grid too long for the shared-memory row table");
        return XS_E_UNSUPPORTED;
    }
    int per_sm = 1;
    XS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scan_co<KP, P, NW, MB, SC, BK, NS>, NW * 32, smem));
    if (per_sm < 1) per_sm = 1;
    int sms = kNumSMs;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
    XS_LAUNCH((k_scan_co<KP, P, NW, MB, SC, BK, NS>), sms * per_sm, NW * 32, smem, stream, *pl, ra, ws, out_co, idx_co);
    return XS_OK;
}

// Scan configuration: pixels per warp P, warps per CTA NW, CTAs per SM MB (register budget = 64K/(NW*32*MB)), math
// flavour (0 packed f32x2, 1 scalar, 2 scalar t + packed d/J, 3 packed without refinement = measurement only) and
// where the per-lane argmin bookkeeping lives.  XS_SCAN_VARIANT (environment) selects one of the experimental
// configurations for KP == 3 that DESIGN.md section 4.1 reports on; 0 (default) is the shipped one (centred flavour,
// math 5), 99 the direct three-operation form that was the default before.  The other shapes DESIGN.md lists as
// measured (P = 4..16, 2-12 warps per CTA, scalar / hybrid math, shared-memory bookkeeping, ring depths) were removed
// from the dispatch after measurement to keep the build short; their code paths (kMath 1, 2, 7, kBookSmem) remain.
struct ScanConfig {
    int p, nw;
};
static int scan_variant() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("XS_SCAN_VARIANT");
        v = e ? atoi(e) : 0;
    }
    return v;
}
// does dispatch_scan pick a centred (kCentred) instantiation?  (then the pixel list is sorted locally by sigma0)
static bool scan_is_centred(int kp) {
    if (kp == 1 || kp == 2) return scan_variant() != 99;
    if (kp != 3) return false;
    const int v = scan_variant();
    return !(v == 30 || v == 41 || v == 99);
}
static ScanConfig scan_config(int kp) {
    if (kp >= 4) return {4, 8};
    if (kp == 3 && scan_variant() == 30) return {8, 8};
    return {8, 4};
}
static int dispatch_scan(const xs_plan *pl, const RasterArgs &ra, const Workspace &ws, double2 *out_co, int *idx_co,
                         void *stream) {
    switch (pl->kp) {
        case 1:
            if (scan_variant() == 99) return launch_scan<1, 8, 4, 4, 0, false, 3>(pl, ra, ws, out_co, idx_co, stream);
            return launch_scan<1, 8, 4, 3, 5, false, 3>(pl, ra, ws, out_co, idx_co, stream);
        case 2:
            if (scan_variant() == 99) return launch_scan<2, 8, 4, 4, 0, false, 3>(pl, ra, ws, out_co, idx_co, stream);
            return launch_scan<2, 8, 4, 3, 5, false, 3>(pl, ra, ws, out_co, idx_co, stream);
        case 3:
            switch (scan_variant()) {
                case 30: return launch_scan<3, 8, 8, 2, 3>(pl, ra, ws, out_co, idx_co, stream);            // NOT exact: no refinement (measurement)
                case 41: return launch_scan<3, 8, 4, 4, 4, false, 3>(pl, ra, ws, out_co, idx_co, stream);  // direct form with phase timers
                case 42: return launch_scan<3, 8, 4, 3, 8, false, 3>(pl, ra, ws, out_co, idx_co, stream);  // shipped form with phase timers
                case 70: return launch_scan<3, 8, 4, 4, 5, false, 3>(pl, ra, ws, out_co, idx_co, stream);  // centred, 4 CTAs (128 regs, spills)
                case 80: return launch_scan<3, 8, 4, 3, 6, false, 3>(pl, ra, ws, out_co, idx_co, stream);  // centred, per-warp slots, 1 barrier/tile
                case 99: return launch_scan<3, 8, 4, 4, 0, false, 3>(pl, ra, ws, out_co, idx_co, stream);  // direct form (the former default)
                default: return launch_scan<3, 8, 4, 3, 5, false, 3>(pl, ra, ws, out_co, idx_co, stream);  // shipped: centred, 3 CTAs x 4 warps
            }
        case 4: return launch_scan<4, 4, 8, 1>(pl, ra, ws, out_co, idx_co, stream);
        default: return launch_scan<6, 4, 8, 1>(pl, ra, ws, out_co, idx_co, stream);
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
    cudaFree(pl->rowtab);
    cudaFree(pl->first_nan);
    cudaFree(pl->slab_absmax);
    cudaFree(pl->inc_cr_grid);
    cudaFree(pl->wspd_cr_grid);
    cudaFree(pl->wspd_cr_half);
    cudaFree(pl->cr_scan);
    cudaFree(pl->cr_absmax);
    cudaFree(pl->cr_finite);
    cudaFree(pl->stats);
    if (pl->ev_scan0) cudaEventDestroy(pl->ev_scan0);
    if (pl->ev_scan1) cudaEventDestroy(pl->ev_scan1);
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
    if ((rc = xs::check(cudaMalloc(&pl->stats, 16 * sizeof(unsigned long long)), "cudaMalloc stats")) != XS_OK) return fail(rc);
    cudaMemsetAsync(pl->stats, 0, 16 * sizeof(unsigned long long), st);
    if ((rc = xs::check(cudaEventCreate(&pl->ev_scan0), "cudaEventCreate")) != XS_OK) return fail(rc);
    if ((rc = xs::check(cudaEventCreate(&pl->ev_scan1), "cudaEventCreate")) != XS_OK) return fail(rc);
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
        pl->fast_ok = kp_ok && std::isfinite(d->dsig_co) && d->dsig_co != 0.0 && std::isfinite(wmax) &&
                      d->n_inc <= kMaxIncBins && pl->n_wspd_pad <= 16384;
        if (pl->fast_ok) {
            const size_t n_scan = (size_t)d->n_inc * pl->n_wspd_pad * pl->nph_pad;
            if ((rc = xs::check(cudaMalloc(&pl->scan, sizeof(float) * n_scan), "cudaMalloc scan image")) != XS_OK) return fail(rc);
            if ((rc = xs::check(cudaMalloc(&pl->rowtab, sizeof(float2) * (size_t)pl->n_wspd_pad), "cudaMalloc")) != XS_OK) return fail(rc);
            if ((rc = xs::check(cudaMalloc(&pl->slab_absmax, sizeof(float) * (size_t)d->n_inc), "cudaMalloc")) != XS_OK) return fail(rc);
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
    }
    auto build = [&]() -> int {
        if (has_cr) {
            const int nmax = pl->n_wspd_cr > pl->n_inc_cr ? pl->n_wspd_cr : pl->n_inc_cr;
            XS_LAUNCH(k_build_cr_tables, (int)ceil_div(nmax, 128), 128, 0, st, *pl);
        }
        if (!has_co) return XS_OK;
        XS_CUDA(cudaMemsetAsync(pl->first_nan, 0x7f, sizeof(int) * (size_t)pl->n_inc, st));  // 0x7f7f7f7f > any index
        if (pl->fast_ok) {
            XS_CUDA(cudaMemsetAsync(pl->slab_absmax, 0, sizeof(float) * (size_t)pl->n_inc, st));
            XS_LAUNCH(k_build_scan, kNumSMs * 8, 256, 0, st, *pl);
            XS_LAUNCH(k_build_rowtab, (int)ceil_div(pl->n_wspd_pad, 256), 256, 0, st, *pl);
        } else {
            XS_LAUNCH(k_find_first_nan, kNumSMs * 8, 256, 0, st, *pl);
        }
        XS_LAUNCH(k_fix_first_nan, (int)ceil_div(pl->n_inc, 256), 256, 0, st, *pl);
        return XS_OK;
    };
    if ((rc = build()) != XS_OK) return fail(rc);
    if ((rc = xs::check(cudaStreamSynchronize(st), "xs_plan_create sync")) != XS_OK) return fail(rc);
    *out = pl;
    return XS_OK;
}

extern "C" size_t xs_invert_workspace_bytes(const xs_plan *pl, int64_t n_px) {
    if (!pl || n_px < 0) return 0;
    return ws_layout(pl->n_inc, n_px, nullptr, nullptr);
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
    double2 *out_co = (double2 *)ar->out_co;
    int sms = kNumSMs;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);

    XS_CUDA(cudaMemsetAsync(pl->stats, 0, 16 * sizeof(unsigned long long), st));
    if (co_run) {
        const bool fast = ar->mode == XS_MODE_FAST && pl->fast_ok;
        if (fast) {
            const size_t need = ws_layout(pl->n_inc, n, nullptr, nullptr);
            if (!ar->workspace || ar->workspace_bytes < need) {
                set_error("xs_invert: workspace too small (%zu < %zu)", ar->workspace_bytes, need);
                return XS_E_WORKSPACE;
            }
            if (((uintptr_t)ar->workspace & 255) != 0) {
                set_error("xs_invert: workspace must be 256-byte aligned");
                return XS_E_INVALID;
            }
            Workspace ws;
            ws_layout(pl->n_inc, n, (char *)ar->workspace, &ws);
            // counters + hist are contiguous at the start of the workspace
            XS_CUDA(cudaMemsetAsync(ws.counters, 0, (char *)ws.bin_start - (char *)ws.counters, st));
            const ScanConfig sc = scan_config(pl->kp);
            const int tile_px = sc.nw * sc.p;
            const int bin_grid = (int)ceil_div(n, kBinPxPerCta);
            XS_LAUNCH(k_bin_count, bin_grid, kBinThreads, sizeof(unsigned) * pl->n_inc, st, *pl, ra, n, ws);
            XS_LAUNCH(k_bin_offsets, 1, 1024, 0, st, pl->n_inc, tile_px, ws);
            XS_LAUNCH(k_bin_scatter, bin_grid, kBinThreads, 2 * sizeof(unsigned) * pl->n_inc, st, *pl, ra, n, ws);
            if (scan_is_centred(pl->kp)) {  // centred flavour: the pixels of a warp need similar sigma0
                const int64_t max_tiles = ceil_div(n, tile_px) + pl->n_inc;
                XS_LAUNCH(k_list_localsort, (unsigned)ceil_div(max_tiles, kSortRun / tile_px), kSortRun, 0, st, *pl, ra, ws, tile_px);
            }
            xs_plan *mpl = const_cast<xs_plan *>(pl);  // timing events are bookkeeping, not plan state
            XS_CUDA(cudaEventRecord(mpl->ev_scan0, st));
            const int rc = dispatch_scan(pl, ra, ws, out_co, ar->idx_co, stream);
            XS_CUDA(cudaEventRecord(mpl->ev_scan1, st));
            mpl->scan_timed = 1;
            if (rc != XS_OK) return rc;
            XS_LAUNCH(k_exact, sms * 8, 256, 0, st, *pl, ra, n, ws.fallback, ws.counters + 1, out_co, ar->idx_co);
            XS_CUDA(cudaMemcpyAsync(pl->stats, ws.counters, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
        } else {
            XS_LAUNCH(k_exact, sms * 8, 256, 0, st, *pl, ra, n, (const unsigned *)nullptr, (const u64 *)nullptr, out_co,
                      ar->idx_co);
        }
    }
    {
        const int64_t warps_needed = ceil_div(n, 32);
        int64_t grid = ceil_div(warps_needed * 32, 256);
        const int64_t cap = (int64_t)sms * 16;
        if (grid > cap) grid = cap;
        XS_LAUNCH(k_cross, (int)grid, 256, 0, st, *pl, ra, n, out_co, ar->out_cr, ar->idx_co, ar->idx_cr);
    }
    return XS_OK;
}

extern "C" int xs_plan_last_scan_ms(const xs_plan *pl, float *ms) {
    if (!pl || !ms || !pl->scan_timed) {
        set_error("xs_plan_last_scan_ms: no scan has been launched on this plan");
        return XS_E_INVALID;
    }
    XS_CUDA(cudaEventSynchronize(pl->ev_scan1));
    XS_CUDA(cudaEventElapsedTime(ms, pl->ev_scan0, pl->ev_scan1));
    return XS_OK;
}

// Raw device counters of the last xs_invert (development aid; layout in the Workspace comment above).
extern "C" int xs_plan_debug_counters(const xs_plan *pl, unsigned long long out[16]) {
    if (!pl || !out) return XS_E_INVALID;
    XS_CUDA(cudaMemcpy(out, pl->stats, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return XS_OK;
}

extern "C" int xs_plan_last_stats(const xs_plan *pl, int64_t stats[4]) {
    if (!pl || !stats) {
        set_error("xs_plan_last_stats: null argument");
        return XS_E_INVALID;
    }
    unsigned long long h[8];
    XS_CUDA(cudaMemcpy(h, pl->stats, sizeof(h), cudaMemcpyDeviceToHost));
    stats[0] = (int64_t)h[2];  // pixels settled by the FP32 scan
    stats[1] = (int64_t)h[3];  // chunks re-evaluated in FP64
    stats[2] = (int64_t)h[1];  // pixels sent to the exhaustive FP64 scan
    stats[3] = (int64_t)h[0];  // tiles
    return XS_OK;
}
