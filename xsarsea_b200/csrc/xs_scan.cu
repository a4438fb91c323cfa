// The co-pol argmin of K1 (reference windspeed/windspeed.py:212-232) for B200 / sm_100a: FP32 scan + exact refinement.
//
// Per pixel the reference evaluates a cost J over the whole wspd x phi grid of the LUT slab of the pixel's incidence bin
// and takes np.argmin.  J is a squared distance in 3-D:
//     J(w,phi) = ((w cos phi - a)/2)^2 + ((w sin phi - b)/2)^2 + ((L[inc][w][phi] - s)/dsig_co)^2
// Dropping the per-pixel constant (a^2+b^2)/4 and centring on a constant c close to the s/dsig_co of the 8 pixels a warp
// scans together (lambda = L/dsig_co - c, sigma = s/dsig_co - c):
//     J'' = J - (a^2+b^2)/4 - sigma^2 = M + k lambda + (-w/2) g(phi),   M = lambda^2 + w^2/4 (shared by the 8 pixels),
//     k = -2 sigma (per pixel),   g = a cos phi + b sin phi (per pixel and phi node)
// i.e. per pixel and candidate pair two FFMA2 and one FMNMX3.
//
// Shared-sigma0 mode.  The listed pixels arrive sorted by (bin, sigma0) (xs_sort.cu), so the 8 pixels of a warp usually have
// |sigma| of a few 1e-3; then the warp leaves k lambda out of the scanned cost altogether (J'' ~ M + (-w/2) g: ONE FFMA2
// and one FMNMX3 per pixel and candidate pair) and widens the error band by the bound 2 |sigma| Lam of the omitted term;
// the refinement filters the members of the wider band once more with the full FP32 cost before any FP64 work.
//
// Exact pruning.  Most of the slab cannot hold the argmin: k_tile_plan bounds the cost of every (16-row chunk, phi group) cell
// from below and keeps only the cells whose bound does not exceed the cost of a seed candidate (see the comment there).
//
// Kernels, all on the caller's stream, no host synchronisation:
//   k_list_prepare  one 32-byte PixRec per record position from the sorted pixel indices (bins padded to whole tiles)
//   k_tile_plan     one warp per run of consecutive tiles: seeds, lower bounds, the 64-byte plan of every tile (the chunks
//                   the CTA streams; per scan warp and phi group the chunks it computes on)
//   k_scan_co       persistent CTAs (4 per SM, 4 warps, 128 registers): a tile = 32 record positions of one bin; the tile's
//                   PixRecs, its plan and the planned chunks of the bin's slab of the FP32 scan image arrive by bulk-async
//                   (TMA) copies -- records double-buffered one tile ahead, chunks of 16 wspd rows through a 4-stage ring
//                   that keeps streaming across tile boundaries; there is no producer thread and no CTA barrier: the warp
//                   that is the last to finish a chunk refills its stage (shared-memory arrival counter), so no warp ever
//                   waits for another one to release a stage.  Lane l owns the phi pairs {2(l+32j), 2(l+32j)+1}; per lane
//                   and pixel only the running minimum and a bit mask of the chunks whose minimum came within kBandMargin
//                   of it are kept.  After the tile's chunks: warp minima, the rigorous error band (evaluated once per
//                   pixel, in the pixel's lane), and one RefRec per pixel.
//   k_refine_easy   eight lanes per pixel: re-creates the FP32 costs (bit-identical operations) of the recorded (lane,
//                   chunk) cells from the cell image; a single band member settles the pixel, several are evaluated in FP64
//                   with the reference's operation order, lexicographic (J, index) minimum = numpy's first minimum
#include <stdio.h>
#include <stdlib.h>

#include "xs_invert.cuh"

namespace xs {

constexpr unsigned kTileEnd = 0xffffffffu;
constexpr int kTileBatch = 8;  // tiles a CTA takes per global atomic

// ---- pixel records --------------------------------------------------------------------------------------------------------
// One PixRec per record position: position e of bin b (bin_start[b] <= e < bin_start[b + 1]) holds the (e - bin_start[b])-th
// pixel of the bin in sigma0 order, or padding behind the bin's last pixel.
__global__ void __launch_bounds__(256) k_list_prepare(xs_plan pl, RasterArgs a, Workspace ws, const unsigned *__restrict__ sorted_px,
                                                      int tile_px) {
    const int64_t n_pos = (int64_t)ws.counters[0] * tile_px;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_pos; e += stride) {
        int lo = 0, hi = pl.n_inc;  // bin of record position e: last b with bin_start[b] <= e
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (ws.bin_start[mid] <= (unsigned)e)
                lo = mid;
            else
                hi = mid;
        }
        const int bin = lo;
        const unsigned r = (unsigned)e - ws.bin_start[bin];
        PixRec rec;
        rec.qa = rec.qb = rec.s = 0.0;
        rec.px = 0xffffffffu;
        rec.bin = (unsigned short)bin;
        rec.state = 0;
        rec.neg = 0;
        if (r < ws.hist[bin]) {
            const unsigned px = sorted_px[ws.ubase[bin] + r];
            const double2 anc = load_cplx(a.anc, px, a.dtype);
            const double s_raw = load_real(a.s_co, px, a.dtype);
            rec.px = px;
            rec.qa = anc.x;
            rec.qb = pl.phi_180 ? fabs(anc.y) : anc.y;
            rec.neg = anc.y < 0.0;
            rec.s = (a.flags & XS_FLAG_SIGMA0_DB) ? s_raw : to_db(s_raw);
            const bool finite_q = isfinite(rec.qa) && isfinite(rec.qb) && isfinite(rec.s);
            rec.state = !finite_q ? 3 : (pl.first_nan[bin] >= 0 ? 2 : 1);
        }
        ws.pix[e] = rec;
    }
}

// ---- exact chunk pruning ---------------------------------------------------------------------------------------------
// J(w, phi) = Jwind + Jsig is a sum of two non-negative terms, and each has a lower bound that holds for every candidate
// of a 16-row chunk of the slab:
//   Jsig  >= (dist(s, [lo_c, hi_c]) / dsig_co)^2    [lo_c, hi_c] = value range of the chunk over all phi nodes (plan table)
//   Jwind >= (dist(A, [wlo_c, whi_c]) / 2)^2        A = |ancillary|:  (w cos - a)^2 + (w sin - b)^2 >= (|w| - A)^2
// With U = the reference's FP64 cost of ANY candidate (the seed), a chunk whose bound exceeds U cannot hold the argmin (nor a
// tie with it), so the scan may skip it: the argmin c* is scanned, and the error-band argument of k_scan_co --
// J32(c*) <= J(c*) + E <= J(c^) + E <= m32 + 2E with c^ the scanned FP32 minimum -- needs nothing else.  The comparison
// keeps a margin of 1e-6 relative + 1e-6 (1 + A^2 + W^2) absolute, ten orders of magnitude above the FP64 rounding of
// either side, so no rounding argument is needed.  The pixels of a tile are neighbours in sigma0 order and share the
// slab, so they keep nearly the same chunks: the CTA streams the union (word 0 of the tile's plan) and every warp
// computes only on the chunks its own pixels keep (words 1 ..).
// Seed: on up to 64 phi nodes (constant stride) the row where the slab crosses the tile's median sigma0 (bisection; any
// row is a valid seed, monotonicity only makes it a good one); every pixel takes the cheapest of them by an FP32 estimate
// and evaluates that one with the reference's FP64 operations.
constexpr int kSeedMax = 64;

// chunks covered by a mask (a bit covers 2^sh chunks, the last bit possibly fewer)
__device__ __forceinline__ int mask_chunks(unsigned mask, int sh, int n_chunks) {
    if (sh == 0) return __popc(mask);
    int n = 0;
    while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        n += min(1 << sh, n_chunks - (b << sh));
    }
    return n;
}
// smallest chunk > c whose mask bit is set, or n_chunks
__device__ __forceinline__ int next_chunk(unsigned mask, int c, int sh, int n_chunks) {
    const int c1 = c + 1;
    if (c1 >= n_chunks) return n_chunks;
    const int b = c1 >> sh;
    if ((mask >> b) & 1u) return c1;
    const unsigned rest = b >= 31 ? 0u : (mask >> (b + 1)) << (b + 1);
    if (!rest) return n_chunks;
    const int c2 = (__ffs(rest) - 1) << sh;
    return c2 < n_chunks ? c2 : n_chunks;
}

constexpr int kPlanMaxChunks = 256;  // plan creation: n_wspd_pad <= 256 * kChunkRows

__global__ void __launch_bounds__(256) k_tile_plan(xs_plan pl, Workspace ws, int tile_px, int nw, int min_items, int prune) {
    __shared__ float4 seed_s[8][kSeedMax];       // {w cos phi, w sin phi, L / dsig_co, flat index} of the warp's seeds
    extern __shared__ __align__(16) unsigned char plan_smem[];
    double2 *wr_s = reinterpret_cast<double2 *>(plan_smem);  // [n_chunks] {wlo, whi} of every chunk
    // sigma0 part of the tile's lower bounds, [warp][chunk][group]
    double *lbs_all = reinterpret_cast<double *>(plan_smem + sizeof(double2) * pl.n_chunks);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_gw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_tiles = (int64_t)ws.counters[0];
    const int sh = pl.mask_sh, n_chunks = pl.n_chunks;
    const int nb = (n_chunks + (1 << sh) - 1) >> sh;
    const unsigned all = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
    const int P = tile_px / nw;
    const int ng = pl.n_groups;
    double *lbs_s = lbs_all + (size_t)wid * n_chunks * ng;
    const int stride = pl.seed_stride, n_seed = (pl.n_phi + stride - 1) / stride;
    static_assert(kSeedMax == 64, "xs_plan::seed_rmax holds 64 seed nodes per slab");
    const double inv_d = 1.0 / fabs(pl.dsig_co);
    for (int c = threadIdx.x; c < n_chunks; c += blockDim.x) wr_s[c] = make_double2(pl.chunk_wlo[c], pl.chunk_whi[c]);
    __syncthreads();
    // phi nodes per group (statistics: candidates the scan evaluates)
    int gphi[kPlanGroups];
#pragma unroll
    for (int g = 0; g < kPlanGroups; ++g) {
        int cnt = 0;
        for (int j = 0; j < pl.kp; ++j)
            if (g < ng && j * ng / pl.kp == g) cnt += max(min(64 * (j + 1), pl.n_phi) - 64 * j, 0);
        gphi[g] = cnt;
    }
    // lanes of the scan warp w of a tile (pixels [w P, (w + 1) P))
    unsigned grp[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) grp[w] = (w < nw && lane >= w * P && lane < (w + 1) * P) ? 0xffffffffu : 0u;
    unsigned n_items = 0;
    u64 n_warp_items = 0;
    // a warp plans a contiguous range of tiles: neighbouring tiles of a bin have nearly the same sigma0, so the rows where the
    // slab crosses it are those of the previous tile almost always (two loads verify that; a bisection otherwise)
    const int64_t per = (n_tiles + n_gw - 1) / n_gw;
    const int64_t t_end = min((gw + 1) * per, n_tiles);
    int lo_prev[2] = {-1, -1}, bin_prev = -1;
    for (int64_t t = gw * per; t < t_end; ++t) {
        unsigned uni = all, wm[4][kPlanGroups];  // [scan warp][phi group]
#pragma unroll
        for (int w = 0; w < 4; ++w)
#pragma unroll
            for (int g = 0; g < kPlanGroups; ++g) wm[w][g] = g < ng ? all : 0u;
        PixRec rec;
        rec.qa = rec.qb = rec.s = 0.0;
        rec.bin = 0;
        rec.state = 0;
        if (lane < tile_px) rec = ws.pix[t * tile_px + lane];
        const bool on = rec.state == 1;
        const unsigned on_mask = __ballot_sync(0xffffffffu, on);
        const int bin = __shfl_sync(0xffffffffu, (int)rec.bin, 0);  // position 0 of a tile is never padding
        if (prune && on_mask) {
            const double s_mid = __shfl_sync(0xffffffffu, rec.s, __fns(on_mask, 0, __popc(on_mask) / 2 + 1));
            const double *slab = pl.co_lut + (size_t)bin * pl.n_wspd * pl.n_phi;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int j = lane + 32 * k;
                if (j >= n_seed) continue;
                const int ip = j * stride;
                int lo = lo_prev[k];  // first row whose value reaches s_mid
                bool known = false;
                double v_lo = 0.0, v_hi = 0.0;  // slab[lo - 1], slab[lo]
                if (bin == bin_prev && lo >= 0) {
                    if (lo > 0) v_lo = slab[(size_t)(lo - 1) * pl.n_phi + ip];
                    if (lo < pl.n_wspd) v_hi = slab[(size_t)lo * pl.n_phi + ip];
                    known = (lo == 0 || v_lo < s_mid) && (lo == pl.n_wspd || !(v_hi < s_mid));
                    // sigma0 grows from tile to tile: the crossing has moved up by one row far more often than anywhere else
                    if (!known && lo + 1 <= pl.n_wspd && lo < pl.n_wspd && v_hi < s_mid) {
                        const double v_up = lo + 1 < pl.n_wspd ? slab[(size_t)(lo + 1) * pl.n_phi + ip] : 0.0;
                        if (lo + 1 == pl.n_wspd || !(v_up < s_mid)) {
                            v_lo = v_hi;
                            v_hi = v_up;
                            ++lo;
                            known = true;
                        }
                    }
                }
                if (!known) {
                    int hi = pl.n_wspd;
                    lo = 0;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (slab[(size_t)mid * pl.n_phi + ip] < s_mid)
                            lo = mid + 1;
                        else
                            hi = mid;
                    }
                    if (lo > 0) v_lo = slab[(size_t)(lo - 1) * pl.n_phi + ip];
                    if (lo < pl.n_wspd) v_hi = slab[(size_t)lo * pl.n_phi + ip];
                }
                lo_prev[k] = lo;
                int r = min(lo, pl.n_wspd - 1);
                double v = lo < pl.n_wspd ? v_hi : v_lo;
                if (lo > 0 && lo < pl.n_wspd && fabs(v_lo - s_mid) <= fabs(v_hi - s_mid)) {
                    r = lo - 1;
                    v = v_lo;
                }
                if (lo == pl.n_wspd) {
                    // sigma0 above the column's last value: GMFs saturate, the column's largest value may sit at a lower wind
                    // speed -- seed there, or where the rising part of the column crosses sigma0
                    const int rm = pl.seed_rmax[bin * 64 + j];
                    const double vm = slab[(size_t)rm * pl.n_phi + ip];
                    if (vm > v) {
                        r = rm;
                        v = vm;
                        if (vm >= s_mid) {
                            int a = 0, b = rm;  // first row of [0, rm] whose value reaches s_mid
                            while (a < b) {
                                const int mid = (a + b) >> 1;
                                if (slab[(size_t)mid * pl.n_phi + ip] < s_mid)
                                    a = mid + 1;
                                else
                                    b = mid;
                            }
                            r = a;
                            v = slab[(size_t)a * pl.n_phi + ip];
                        }
                    }
                }
                const double w = pl.wspd_grid[r];
                seed_s[wid][j] = make_float4((float)(w * pl.cos_phi[ip]), (float)(w * pl.sin_phi[ip]), (float)(v * inv_d),
                                             __int_as_float(r * pl.n_phi + ip));
            }
            bin_prev = bin;
            // sigma0 part of the bounds, once per tile: dist([s_lo, s_hi], [lo_c, hi_c]) with [s_lo, s_hi] enclosing the sigma0
            // of every pixel of the tile (FP32 minimum / maximum widened by one ulp), lane <-> chunk
            {
                const float sf = (float)rec.s;
                const float f_lo = float_from_order_key(__reduce_min_sync(0xffffffffu, on ? float_order_key(sf) : 0x7fffffff));
                const float f_hi = float_from_order_key(__reduce_max_sync(0xffffffffu, on ? float_order_key(sf) : (int)0x80000000));
                const double s_lo = (double)nextafterf(f_lo, -CUDART_INF_F), s_hi = (double)nextafterf(f_hi, CUDART_INF_F);
                const double *clo = pl.chunk_lo + (size_t)bin * n_chunks * ng, *chi = pl.chunk_hi + (size_t)bin * n_chunks * ng;
                for (int c = lane; c < n_chunks * ng; c += 32) {  // (chunk, group) cells
                    const double ds = fmax(fmax(clo[c] - s_hi, s_lo - chi[c]), 0.0) * inv_d;
                    lbs_s[c] = ds * ds;
                }
            }
            __syncwarp();
            float best = CUDART_INF_F;
            int bflat = __float_as_int(seed_s[wid][0].w);
            {
                const float fa = (float)rec.qa, fb = (float)rec.qb, fs = (float)(rec.s * inv_d);
                for (int j = 0; j < n_seed; ++j) {
                    const float4 v = seed_s[wid][j];
                    const float ta = 0.5f * (v.x - fa), tz = 0.5f * (v.y - fb), ts = v.z - fs;
                    const float J = ta * ta + tz * tz + ts * ts;
                    if (J < best) {
                        best = J;
                        bflat = __float_as_int(v.w);
                    }
                }
            }
            double thr = CUDART_INF;
            const double A = hypot(rec.qa, rec.qb);
            if (on) {
                const int iw = bflat / pl.n_phi, ip = bflat - iw * pl.n_phi;
                const double U = exact_cost_co(pl.wspd_grid[iw], pl.cos_phi[ip], pl.sin_phi[ip], slab[bflat], rec.qa, rec.qb, rec.s, pl.dsig_co);
                // the margin: 1e-6 relative and absolute (in units of the magnitudes involved), plus the (1 - 1e-9) of the bound
                thr = (U * (1.0 + 1e-6) + 1e-6 * (1.0 + A * A + pl.w_absmax * pl.w_absmax)) * (1.0 + 2e-9);
            }
            unsigned km[kPlanGroups] = {0u, 0u, 0u};  // chunk bits this lane's pixel keeps, per phi group
            // only the chunks whose sigma0 bound alone does not already exceed the largest threshold of the tile are examined
            // per pixel (a handful of the slab's chunks): float image of the threshold rounded up, lane <-> chunk
            const float thr_f = __uint_as_float(__reduce_max_sync(0xffffffffu, on ? __float_as_uint(fabsf(__double2float_ru(thr))) : 0u));
            for (int c0 = 0; c0 < n_chunks; c0 += 32) {
                bool cand = false;
                if (c0 + lane < n_chunks)
                    for (int g = 0; g < ng; ++g) cand |= !(lbs_s[(c0 + lane) * ng + g] > (double)thr_f);  // NaN / inf thresholds keep everything
                unsigned todo = __ballot_sync(0xffffffffu, cand);
                while (todo) {
                    const int c = c0 + __ffs(todo) - 1;
                    todo &= todo - 1;
                    const double2 wr = wr_s[c];
                    const double dw = fmax(fmax(wr.x - A, A - wr.y), 0.0) * 0.5;
                    const double dw2 = dw * dw;
                    const unsigned bit = on ? (1u << (c >> sh)) : 0u;
#pragma unroll
                    for (int g = 0; g < kPlanGroups; ++g)
                        if (g < ng) km[g] |= !(dw2 + lbs_s[c * ng + g] > thr) ? bit : 0u;
                }
            }
            __syncwarp();  // the next tile overwrites the warp's shared-memory tables
            uni = __reduce_or_sync(0xffffffffu, km[0] | km[1] | km[2]);
#pragma unroll
            for (int w = 0; w < 4; ++w)
#pragma unroll
                for (int g = 0; g < kPlanGroups; ++g) wm[w][g] = __reduce_or_sync(0xffffffffu, km[g] & grp[w]);
            // the scan's ring looks kStages chunks ahead, at most into the next tile
            while (mask_chunks(uni, sh, n_chunks) < min_items) {
                unsigned ext = ((uni << 1) | (uni >> 1)) & ~uni & all;
                if (!ext) ext = ~uni & all;
                if (!ext) break;
                uni |= ext & (0u - ext);
            }
        } else if (prune && !on_mask) {
#pragma unroll
            for (int w = 0; w < 4; ++w) wm[w][0] = wm[w][1] = wm[w][2] = 0u;  // nothing to scan in this tile (NaN slab / exhaustive pixels only)
            uni = 0u;
            for (int c = 0, k = 0; c < n_chunks && k < min_items; ++c, ++k) uni |= 1u << (c >> sh);
        }
        if (lane == 0) {
            uint4 *dst = reinterpret_cast<uint4 *>(ws.tile_plan + (size_t)t * kPlanWords);
            dst[0] = make_uint4(uni, wm[0][0], wm[0][1], wm[0][2]);
            dst[1] = make_uint4(wm[1][0], wm[1][1], wm[1][2], wm[2][0]);
            dst[2] = make_uint4(wm[2][1], wm[2][2], wm[3][0], wm[3][1]);
            dst[3] = make_uint4(wm[3][2], 0u, 0u, 0u);
            n_items += (unsigned)mask_chunks(uni, sh, n_chunks);
            for (int w = 0; w < nw && w < 4; ++w)
#pragma unroll
                for (int g = 0; g < kPlanGroups; ++g) n_warp_items += (u64)mask_chunks(wm[w][g] & uni, sh, n_chunks) * gphi[g];
        }
    }
    if (lane == 0) {
        if (n_items) atomicAdd(&ws.counters[5], (u64)n_items);
        if (n_warp_items) atomicAdd(&ws.counters[6], n_warp_items);
    }
}

// ---- the FP32 scan ----------------------------------------------------------------------------------------------------
template <int KP, int P, int NW>
struct ScanSmem {
    static constexpr int kRowFloats = 64 * KP;
    alignas(128) float ring[kStages][kChunkRows * kRowFloats];
    alignas(16) PixRec pix[2][NW * P];
    alignas(16) unsigned plan[2][kPlanWords];  // chunk masks of the tile (k_tile_plan)
    alignas(8) uint64_t full[kStages];    // a chunk has landed in the stage
    uint64_t pix_full[2];                 // the tile's records (or the end-of-work mark) have landed
    unsigned done[kStages];               // warps that have finished with the stage's chunk (monotonic)
    unsigned tile_of[2];
    unsigned batch_next, batch_end;
};

template <int KP, int P, int NW, int MB, bool SC = false>  // SC: scalar FFMA per candidate instead of FFMA2 per pair
__global__ void __launch_bounds__(NW * 32, MB) k_scan_co(xs_plan pl, Workspace ws, float share_tau) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using Smem = ScanSmem<KP, P, NW>;
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    float2 *rowtab_s = reinterpret_cast<float2 *>(smem_raw + sizeof(Smem));  // [n_wspd_pad]
    constexpr int TP = NW * P;                                               // list positions per tile
    constexpr int NS = kStages;
    constexpr uint32_t kPixBytes = TP * sizeof(PixRec);
    constexpr uint32_t kPlanBytes = kPlanWords * sizeof(unsigned);
    static_assert(NW <= 4, "a tile plan holds four warp masks");

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned n_tiles = (unsigned)ws.counters[0];
    const int n_chunks = pl.n_chunks;  // >= NS (plan creation checks); the last may be shorter
    const size_t slab_floats = (size_t)pl.n_wspd_pad * Smem::kRowFloats;
    const int mask_sh = pl.mask_sh;  // chunks per bit of the chunk masks: 2^mask_sh (32 bits cover the slab)

    // one chunk of a slab into its ring stage
    auto load_chunk = [&](int bin, int c, int stage) {
        const uint32_t bytes = (uint32_t)min(kChunkRows, pl.n_wspd_pad - c * kChunkRows) * Smem::kRowFloats * 4;
        mbar_expect_tx(&sm.full[stage], bytes);
        bulk_g2s(sm.ring[stage], pl.scan + (size_t)bin * slab_floats + (size_t)c * kChunkRows * Smem::kRowFloats, bytes, &sm.full[stage]);
    };
    // next tile of this CTA (batches of kTileBatch per global atomic) -> records into pix[buf]
    auto fetch_tile = [&](int buf) {
        unsigned t;
        if (sm.batch_next < sm.batch_end)
            t = sm.batch_next++;
        else {
            t = (unsigned)atomicAdd(&ws.counters[8], (u64)kTileBatch);
            sm.batch_next = t + 1;
            sm.batch_end = min(t + (unsigned)kTileBatch, n_tiles);
        }
        if (t >= n_tiles) t = kTileEnd;
        sm.tile_of[buf] = t;
        if (t != kTileEnd) {
            fence_proxy_async();  // the buffer's previous records were read through the generic proxy
            mbar_expect_tx(&sm.pix_full[buf], kPixBytes + kPlanBytes);
            bulk_g2s(sm.pix[buf], ws.pix + (size_t)t * TP, kPixBytes, &sm.pix_full[buf]);
            bulk_g2s(sm.plan[buf], ws.tile_plan + (size_t)t * kPlanWords, kPlanBytes, &sm.pix_full[buf]);
        } else
            mbar_arrive(&sm.pix_full[buf]);
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&sm.full[s], 1);
            sm.done[s] = 0;
        }
        mbar_init(&sm.pix_full[0], 1);
        mbar_init(&sm.pix_full[1], 1);
        sm.batch_next = sm.batch_end = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < pl.n_wspd_pad; i += blockDim.x) rowtab_s[i] = pl.rowtab[i];
    __syncthreads();
    if (threadIdx.x == 0) {  // first tile and the first NS chunks of its plan (every plan lists at least NS)
        fetch_tile(0);
        if (sm.tile_of[0] != kTileEnd) {
            mbar_wait(&sm.pix_full[0], 0);
            const int bin0 = sm.pix[0][0].bin;
            const unsigned mk0 = sm.plan[0][0];
            for (int i = 0, c = -1; i < NS; ++i) {
                c = next_chunk(mk0, c, mask_sh, n_chunks);
                load_chunk(bin0, c, i);
            }
        }
    }

    int stage = 0;
    uint32_t phase = 0;  // parity of full[stage] for the chunk this warp consumes next
    unsigned n_shared = 0;  // record positions scanned in shared-sigma0 mode (statistics)
    for (unsigned u = 0;; ++u) {
        const int b = u & 1;
        mbar_wait(&sm.pix_full[b], (u >> 1) & 1);
        const unsigned tile = sm.tile_of[b];
        if (tile == kTileEnd) break;
        const PixRec *mine = &sm.pix[b][warp * P];
        const int bin = sm.pix[b][0].bin;  // position 0 of a tile is never padding

        // ---- per-warp centre and per-pixel query constants: lane p < P works for the warp's pixel p ----
        // (the per-pixel scalars are the same in every lane: computed once in the pixel's lane and broadcast where needed)
        const PixRec *own = &mine[lane < P ? lane : 0];
        const bool on_l = lane < P && own->state == 1;
        const double v_l = own->s / pl.dsig_co;           // s/dsig_co of this lane's pixel
        const float sp_l = (float)v_l;
        const unsigned on_lanes = __ballot_sync(0xffffffffu, on_l);
        const bool any = on_lanes != 0u;
        // centre: midpoint of the FP32 images of the warp's s/dsig_co (any constant will do; the refinement gets it in the RefRec)
        const float cs = !any ? 0.f
                              : 0.5f * (float_from_order_key(__reduce_min_sync(0xffffffffu, on_l ? float_order_key(sp_l) : 0x7fffffff)) +
                                        float_from_order_key(__reduce_max_sync(0xffffffffu, on_l ? float_order_key(sp_l) : (int)0x80000000)));
        // Shared-sigma0 mode: the list is in sigma0 order, so the warp's pixels usually differ from the centre by a tiny
        // |sigma|; then k lambda = -2 sigma lambda is left out of the scanned cost altogether -- one FFMA2 per pixel and
        // candidate pair instead of two -- and the band is widened by what the omission can change BETWEEN two candidates
        // that can win, 2 |sigma| |lambda(c1) - lambda(c2)| (see the band section below).  The mode is chosen from an
        // estimate of that quantity (performance only -- the band uses the rigorous bound).
        const float slab_lo = float_from_order_key(pl.slab_range[2 * bin]), slab_hi = float_from_order_key(pl.slab_range[2 * bin + 1]);
        float need_l = 0.f;
        if (on_l) {
            const float dmin = fmaxf(fmaxf(slab_lo - sp_l, sp_l - slab_hi), 0.f);
            // |d| of a winning candidate lies in [dmin, ~sqrt(dmin^2 + 6)]: outside the slab's range the spread is small
            const float spread = dmin > 0.f ? sqrtf(dmin * dmin + 6.f) - dmin : 5.f;
            need_l = fabsf(sp_l - cs) * spread;
        }
        // non-negative floats order like their bit patterns; a NaN estimate reads as a huge number (exact-k mode)
        const bool shared = any && __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(need_l) & 0x7fffffffu)) <= share_tau;
        const float nq_l = (on_l && !shared) ? (float)(-2.0 * (v_l - (double)cs)) : 0.f;  // k_p = -2 (s_p/dsig - cs); 0 in shared mode
        float nqs[P];
        u64 g[P][KP];   // {g(phi_even), g(phi_odd)} as packed FP32
#pragma unroll
        for (int p = 0; p < P; ++p) nqs[p] = __shfl_sync(0xffffffffu, nq_l, p);
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            // (cos, sin) of the lane's phi nodes: 12 L1-resident loads per tile; 0 in the padding (+inf image values there)
            const int ip = 2 * (lane + 32 * j);
            const double c0x = ip < pl.n_phi ? pl.cos_phi[ip] : 0.0, c0y = ip < pl.n_phi ? pl.sin_phi[ip] : 0.0;
            const double c1x = ip + 1 < pl.n_phi ? pl.cos_phi[ip + 1] : 0.0, c1y = ip + 1 < pl.n_phi ? pl.sin_phi[ip + 1] : 0.0;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const double qa = mine[p].qa, qb = mine[p].qb;
                g[p][j] = ((on_lanes >> p) & 1u) ? pack2(g32(qa, qb, c0x, c0y), g32(qa, qb, c1x, c1y)) : 0ull;
            }
        }
        // per lane and pixel: the running minimum and the set of chunks (bit c >> mask_sh) whose minimum came within
        // kBandMargin of it.  Any band this kernel accepts is narrower than kBandMargin (2 E < 0.5), so at the end the set
        // holds every chunk of the lane that can contain a band member (a chunk is only dropped when a later minimum is
        // lower by more than the margin, i.e. when it is outside every acceptable band).
        float m[P], best[P];
        unsigned cmask[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            m[p] = best[p] = CUDART_INF_F;
            cmask[p] = 0u;
        }
        const u64 ncs2 = pack2(-cs, -cs);

        // ---- the chunks of the tile's plan, 16 wspd rows at a time ----
        const unsigned tmask = sm.plan[b][0];
        const int c_first = next_chunk(tmask, -1, mask_sh, n_chunks);
        for (int c = c_first; c < n_chunks; c = next_chunk(tmask, c, mask_sh, n_chunks)) {
            mbar_wait(&sm.full[stage], phase);
            // phi groups of the chunk that one of this warp's pixels keeps (bit g: the float2 slots j with j * NG / KP == g)
            constexpr int NG = KP < kPlanGroups ? KP : kPlanGroups;
            unsigned jm = 0u;
#pragma unroll
            for (int gI = 0; gI < NG; ++gI) jm |= ((sm.plan[b][1 + kPlanGroups * warp + gI] >> (c >> mask_sh)) & 1u) << gI;
            const bool mine_on = jm == (1u << NG) - 1u;  // every group: the fused loops below
            if (!mine_on && jm && any) {
                // some groups only: one loop per float2 slot (the same operations per candidate as the fused loops)
                const u64 *rows = reinterpret_cast<const u64 *>(sm.ring[stage]);
                const int rows_here = min(kChunkRows, pl.n_wspd_pad - c * kChunkRows);
#pragma unroll
                for (int j = 0; j < KP; ++j) {
                    if (!((jm >> (j * NG / KP)) & 1u)) continue;
                    if (shared) {
#pragma unroll 2
                        for (int r = 0; r < rows_here; ++r) {
                            const float2 rt = rowtab_s[c * kChunkRows + r];
                            const u64 nwh = pack2(rt.x, rt.x), w2q = pack2(rt.y, rt.y);
                            const u64 lam = fadd2(rows[r * (32 * KP) + lane + 32 * j], ncs2);
                            const u64 Mj = ffma2(lam, lam, w2q);
#pragma unroll
                            for (int p = 0; p < P; ++p) {
                                float j0, j1;
                                if constexpr (SC) {
                                    float g0, g1, M0, M1;
                                    unpack2(g[p][j], g0, g1);
                                    unpack2(Mj, M0, M1);
                                    j0 = __fmaf_rn(rt.x, g0, M0);
                                    j1 = __fmaf_rn(rt.x, g1, M1);
                                } else
                                    unpack2(ffma2(nwh, g[p][j], Mj), j0, j1);
                                m[p] = fmin3(m[p], j0, j1);
                            }
                        }
                    } else {
#pragma unroll 2
                        for (int r = 0; r < rows_here; ++r) {
                            const float2 rt = rowtab_s[c * kChunkRows + r];
                            const u64 nwh = pack2(rt.x, rt.x), w2q = pack2(rt.y, rt.y);
                            const u64 Lj = fadd2(rows[r * (32 * KP) + lane + 32 * j], ncs2);
                            const u64 Mj = ffma2(Lj, Lj, w2q);
#pragma unroll
                            for (int p = 0; p < P; ++p) {
                                const u64 q2 = pack2(nqs[p], nqs[p]);
                                float j0, j1;
                                unpack2(ffma2(nwh, g[p][j], ffma2(q2, Lj, Mj)), j0, j1);
                                m[p] = fmin3(m[p], j0, j1);
                            }
                        }
                    }
                }
            } else if (mine_on && shared) {
                const u64 *rows = reinterpret_cast<const u64 *>(sm.ring[stage]);
                const int rows_here = min(kChunkRows, pl.n_wspd_pad - c * kChunkRows);
#pragma unroll 2
                for (int r = 0; r < rows_here; ++r) {
                    const float2 rt = rowtab_s[c * kChunkRows + r];
                    const u64 nwh = pack2(rt.x, rt.x), w2q = pack2(rt.y, rt.y);
                    u64 M[KP];
#pragma unroll
                    for (int j = 0; j < KP; ++j) {  // M = lambda^2 + w^2/4, shared by the warp's pixels
                        const u64 lam = fadd2(rows[r * (32 * KP) + lane + 32 * j], ncs2);
                        M[j] = ffma2(lam, lam, w2q);
                    }
#pragma unroll
                    for (int p = 0; p < P; ++p) {
#pragma unroll
                        for (int j = 0; j < KP; ++j) {  // J'' ~ M + (-w/2) g: one FFMA2 per candidate pair
                            float j0, j1;
                            if constexpr (SC) {  // the same two roundings as two scalar FFMA
                                float g0, g1, M0, M1;
                                unpack2(g[p][j], g0, g1);
                                unpack2(M[j], M0, M1);
                                j0 = __fmaf_rn(rt.x, g0, M0);
                                j1 = __fmaf_rn(rt.x, g1, M1);
                            } else
                                unpack2(ffma2(nwh, g[p][j], M[j]), j0, j1);
                            m[p] = fmin3(m[p], j0, j1);
                        }
                    }
                }
            } else if (mine_on && any) {
                const u64 *rows = reinterpret_cast<const u64 *>(sm.ring[stage]);
                const int rows_here = min(kChunkRows, pl.n_wspd_pad - c * kChunkRows);  // even (n_wspd_pad is a multiple of 8)
#pragma unroll 2
                for (int r = 0; r < rows_here; ++r) {
                    const float2 rt = rowtab_s[c * kChunkRows + r];
                    const u64 nwh = pack2(rt.x, rt.x), w2q = pack2(rt.y, rt.y);
                    u64 L[KP], M[KP];
#pragma unroll
                    for (int j = 0; j < KP; ++j) {  // shared by the warp's pixels: lambda = L - cs, M = lambda^2 + w^2/4
                        L[j] = fadd2(rows[r * (32 * KP) + lane + 32 * j], ncs2);
                        M[j] = ffma2(L[j], L[j], w2q);
                    }
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const u64 q2 = pack2(nqs[p], nqs[p]);
#pragma unroll
                        for (int j = 0; j < KP; ++j) {  // J'' = k_p lambda + M + (-w/2) g: two FFMA2 per candidate pair
                            const u64 J = ffma2(nwh, g[p][j], ffma2(q2, L[j], M[j]));
                            float j0, j1;
                            unpack2(J, j0, j1);
                            m[p] = fmin3(m[p], j0, j1);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                // the warp that is the last to finish with a stage refills it (stream chunk + NS); at a tile's first
                // chunk it also fetches the CTA's next tile: every warp has left the previous tile by then, so the
                // other record buffer is free
                // (the warp's reads of the stage are complete: their values have been consumed; the arrival itself is relaxed)
                const unsigned old = atom_add_shared(&sm.done[stage], 1u);
                if (old % NW == NW - 1) {
                    __threadfence_block();  // the other warps' arrivals (and what they wrote before them) are visible
                    if (c == c_first) fetch_tile(b ^ 1);
                    // the chunk NS places ahead in the CTA's stream: of this tile's plan, or -- every plan lists at least
                    // NS chunks -- of the next tile's
                    int c2 = c, bin2 = bin;
                    unsigned mk = tmask;
                    bool ok = true;
                    for (int i = 0; i < NS && ok; ++i) {
                        c2 = next_chunk(mk, c2, mask_sh, n_chunks);
                        if (c2 >= n_chunks) {
                            mbar_wait(&sm.pix_full[b ^ 1], ((u + 1) >> 1) & 1);
                            ok = sm.tile_of[b ^ 1] != kTileEnd;
                            if (ok) {
                                bin2 = sm.pix[b ^ 1][0].bin;
                                mk = sm.plan[b ^ 1][0];
                                c2 = next_chunk(mk, -1, mask_sh, n_chunks);
                            }
                        }
                    }
                    if (ok) {
                        fence_proxy_async();
                        load_chunk(bin2, c2, stage);
                    }
                }
            }
            __syncwarp();
            const unsigned cbit = 1u << (c >> mask_sh);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const float dm = m[p] - best[p];  // -inf while best is still +inf; NaN (no bit) when the chunk is all padding too
                cmask[p] = (dm < -kBandMargin) ? 0u : cmask[p];
                cmask[p] |= (dm <= kBandMargin) ? cbit : 0u;
                best[p] = fminf(best[p], m[p]);
                m[p] = CUDART_INF_F;
            }
            if (++stage == NS) {
                stage = 0;
                phase ^= 1u;
            }
        }
        if (!any) continue;
        if (shared) n_shared += P;

        // ---- band of every pixel -> RefRec ---------------------------------------------------------------------------
        // m32 = warp minimum of the FP32 costs; E bounds |J''_fp32 - J''_exact| for every candidate that can still
        // win (derivation: DESIGN.md 4.1, checked on the CPU by tests/test_error_bound.py), so the reference's FP64 argmin
        // lies in S = {c : J''_fp32(c) <= m32 + 2E}; k_refine_easy collects S from the cells recorded here.
        // The bound of pixel p is evaluated once, in lane p.
        float m32_l = CUDART_INF_F;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float mp = float_from_order_key(__reduce_min_sync(0xffffffffu, float_order_key(best[p])));
            if (lane == p) m32_l = mp;
        }
        float thr_l = -CUDART_INF_F, efp_l = 0.f;  // thr = -inf: no lane contends (the scan could not bound its error)
        if (on_l) {
            const float lmax = pl.slab_absmax[bin];
            const float W = (float)pl.w_absmax * 1.0000002f;
            const float m32 = m32_l;
            const float fa = fabsf((float)own->qa), fb = fabsf((float)own->qb);
            const float A = sqrtf(fa * fa + fb * fb) * 1.000001f;  // >= |ancillary|
            const float SC = fabsf((float)(v_l - (double)cs)) * 1.0000002f;
            const float T = W * A + 0.25f * W * W;
            // J'' = J' - sc^2: candidates that can still win have |L/dsig - s/dsig| <= D and |L/dsig - cs| <= Lam = D + |sc|.
            // Error terms (u = 2^-24): image value and lambda roundings 2 Lam (lmax + Lam) through lambda^2 and
            // 2 |sc| (lmax + 2 Lam) through k lambda; M, a and J roundings Lam^2 + W^2/4, Lam^2 + W^2/4 + 2 |sc| Lam and
            // D^2 + sc^2 + T; row-table and g roundings W^2/4 + W A; the W terms add up to W^2 + 2 W A <= 4 T.
            // Shared mode scans J_a = J'' - k lambda.  With c^ the scanned minimum and c* the true argmin: lambda(c^)^2 <=
            // J_a(c^) + A^2/4 <= m32 + A^2/4 + 1 =: R0^2 (the "+ 1" covers the FP32 error, below 0.25), J''(c*) <= J''(c^) <=
            // m32 + E_fp + 2 |sc| R0, hence |d(c*)| <= D with the extra 2 |sc| R0 under the root.
            const float R0 = sqrtf(fmaxf(m32, 0.f) + 0.25f * A * A + 1.0f);
            const float D = sqrtf(fmaxf(m32 + SC * SC * 1.0000002f, 0.f) + 0.25f * A * A + 1.0f + (shared ? 2.f * SC * R0 * 1.000001f : 0.f));
            const float Lam = D + SC;
            float E = 5.9604645e-8f * 1.5f * (2.f * Lam * lmax + 2.f * SC * lmax + 4.f * Lam * Lam + 6.f * SC * Lam + SC * SC + D * D + 4.f * T);
            efp_l = E;  // bound of the full centred form's FP32 error (the refinement's second filter in shared mode)
            if (shared) {
                // J32_a(c*) <= J_a(c*) + E_fp = J''(c*) - k lambda(c*) + E_fp <= J''(c^) - k lambda(c*) + E_fp
                //           <= J32_a(c^) + 2 E_fp + |k| |lambda(c^) - lambda(c*)|,   lambda(c^) - lambda(c*) = d(c^) - d(c*).
                // Both |d| are at most Lam; when sigma0 lies outside the slab's value range by dmin, every d has the same
                // sign and magnitude >= dmin, so the difference is at most Lam - dmin (a pixel far outside the LUT has a
                // large lambda but a small spread); otherwise it is at most 2 Lam.
                const float dist = fmaxf(slab_lo - sp_l, sp_l - slab_hi);  // scan-image values; the exact L/dsig differ by <= 1 ulp
                const float dmin = fmaxf(dist * 0.999999f - 1e-6f * lmax - 1e-6f * fabsf(sp_l), 0.f);
                const float dl = dmin > 0.f ? fmaxf(Lam - dmin, 0.f) : 2.f * Lam;
                E += SC * dl * 1.000001f;  // half of |k| dl: the band is m32 + 2E
            }
            // 2 E < kBandMargin keeps the chunk masks complete; a larger bound means magnitudes outside the range the
            // bound was derived for
            const bool sane = (E < 0.5f * kBandMargin) && (m32 < CUDART_INF_F);
            if (sane) thr_l = m32 + 2.f * E;
        }
        unsigned cont_l = 0u, m0_l = 0u, m1_l = 0u, m2_l = 0u;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            if (!((on_lanes >> p) & 1u)) continue;  // warp-uniform
            const float thr = __shfl_sync(0xffffffffu, thr_l, p);
            const unsigned cont = __ballot_sync(0xffffffffu, best[p] <= thr);  // lanes holding band members
            // the record carries the chunk masks of the first two contending lanes (one or two in 99 % of the pixels) and
            // the union of the masks of all further ones (a superset for each of them)
            const bool mine_c = (cont >> lane) & 1u;
            const int rank = __popc(cont & ((1u << lane) - 1u));
            const unsigned m0 = __reduce_or_sync(0xffffffffu, (mine_c && rank == 0) ? cmask[p] : 0u);
            const unsigned m1 = __reduce_or_sync(0xffffffffu, (mine_c && rank == 1) ? cmask[p] : 0u);
            const unsigned m2 = __reduce_or_sync(0xffffffffu, (mine_c && rank >= 2) ? cmask[p] : 0u);
            if (lane == p) {
                cont_l = cont;
                m0_l = m0;
                m1_l = m1;
                m2_l = m2;
            }
        }
        if (on_l) {
            RefRec rr;
            rr.thr = thr_l;
            rr.cs = cs;
            rr.nq = nq_l;
            rr.efp = efp_l;
            rr.cont = cont_l;
            rr.mask[0] = m0_l;
            rr.mask[1] = m1_l;
            rr.mask[2] = m2_l;
            ws.rec[(size_t)tile * TP + warp * P + lane] = rr;
        }
    }
    if (lane == 0 && n_shared) atomicAdd(&ws.counters[13], (u64)n_shared);
}

// ---- exact refinement ---------------------------------------------------------------------------------------------------
// Eight lanes per record position (four positions per warp at a time, so four times as many pixels are in flight
// as with a warp per pixel -- this pass is bound by the latency of its dependent loads and by instruction issue, not by
// arithmetic).  Settles padding, NaN slabs (answer = first NaN), pixels for the exhaustive kernel, and every pixel whose
// band touches only "cont" cells (one 16-row chunk of one lane each, at most 8): lane `sub` of the group owns rows sub and
// sub + 8 of a cell and all 2 KP phi slots, so g(phi) is evaluated once per slot; a single band member settles the pixel,
// several are evaluated in FP64 with the reference's operation order and reduced to the lexicographic (J, flat index)
// minimum = numpy's first minimum.
template <int KP, int G, int MB = 2>  // G lanes per record position (32 / G positions per warp in flight); MB CTAs per SM
__global__ void __launch_bounds__(256, MB) k_refine_easy(xs_plan pl, Workspace ws, OutSpec out, int tile_px) {
    constexpr int PW = 32 / G;  // positions per warp
    const int lane = threadIdx.x & 31, sub = lane & (G - 1), grp = lane / G;
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (G * grp);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_pos = (int64_t)ws.counters[0] * tile_px;
    static_assert(kChunkRows % G == 0, "whole rows per lane of a group");
    const int n_chunks = pl.n_chunks, mask_sh = pl.mask_sh;
    unsigned n_settled = 0, n_cells = 0, n_fp64 = 0, n_many = 0;
    for (int64_t e0 = warp * PW; e0 < n_pos; e0 += n_warps * PW) {
        // The four groups of a warp stay in lockstep from iteration to iteration (a group that `continue`d out of an
        // iteration used to run ahead for good, and the warp then issued every group's instructions separately): no early
        // exits, one structured block per case, explicit reconvergence here.
        __syncwarp();
        const int64_t e = e0 + grp;
        const bool in_range = e < n_pos;
        if (in_range && sub < 2 && e + n_warps * PW < n_pos) {  // the next iteration's records: into L2 while this one computes
            const void *nxt = sub == 0 ? (const void *)&ws.pix[e + n_warps * PW] : (const void *)&ws.rec[e + n_warps * PW];
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
        }
        PixRec px;
        RefRec rc;
        px.qa = px.qb = px.s = 0.0;
        px.px = 0u;
        px.bin = 0;
        px.state = 0;
        px.neg = 0;
        rc.thr = rc.cs = rc.nq = rc.efp = 0.f;
        rc.cont = rc.mask[0] = rc.mask[1] = rc.mask[2] = 0u;
        if (in_range) {
            px = ws.pix[e];
            rc = ws.rec[e];  // loaded together with the pixel (only meaningful for state 1)
        }
        // padding (state 0) needs nothing; NaN slabs and pixels for the exhaustive kernel are settled by the group's first lane
        if (px.state != 0 && (px.state != 1 || rc.cont == 0u) && sub == 0) {
            if (px.state == 2)
                write_co(pl, out, pl.first_nan[px.bin], px.neg, px.px);  // J is NaN exactly where L is NaN
            else  // non-finite inputs, or the scan could not bound its error for this pixel
                ws.fallback[atomicAdd(&ws.counters[1], 1ull)] = px.px;
        }
        const bool work = px.state == 1 && rc.cont != 0u;  // uniform within the group
        if (work && sub == 0 && __popc(rc.cont) > 2) ++n_many;
        const float *cell_img = pl.cell + (size_t)px.bin * pl.n_chunks * kChunkRows * pl.nph_pad;
        const double *slab64 = pl.co_lut + (size_t)px.bin * pl.n_wspd * pl.n_phi;
        // Shared-sigma0 records (nq == 0): the scanned cost left k lambda out and the band is wider by 2 |sigma| Lam.  The
        // members of that wide band are filtered a second time with the full centred FP32 cost (k recomputed as the scan's
        // exact mode would): the true argmin c* is a member, so full32(c*) <= J(c*) + efp <= J(c) + efp <= full32(c) + 2 efp for
        // every member c -- only members within 2 efp of the smallest full32 can be the argmin.  FP64 is then needed as
        // rarely as without sharing.
        const bool two_stage = rc.nq == 0.f;
        const float kfull = (float)(-2.0 * (px.s / pl.dsig_co - (double)rc.cs));
        float thr2 = CUDART_INF_F;
        // every lane keeps its two best members (second-filter cost, flat index) and the value of its third best: almost
        // always that is all of them, and the pixel is settled without walking the cells again
        float b1 = CUDART_INF_F, b2 = CUDART_INF_F, b3 = CUDART_INF_F;
        int f1 = -1, f2 = -1, n_loc = 0, one_loc = -1;
        double bj = CUDART_INF;  // FP64: lexicographic (J, flat index) minimum of this lane's members
        int bi = 0x7fffffff;
        auto fp64_feed = [&](int flat) {  // J is never NaN here: finite inputs, NaN-free slab
            const int iw = flat / pl.n_phi, ip = flat - iw * pl.n_phi;
            const double J = exact_cost_co(pl.wspd_grid[iw], pl.cos_phi[ip], pl.sin_phi[ip], slab64[flat], px.qa, px.qb, px.s, pl.dsig_co);
            if (J < bj || (J == bj && flat < bi)) {
                bj = J;
                bi = flat;
            }
        };
        // one recorded cell = (scan lane L, chunk c): re-create its FP32 costs, collect the band members.
        // stage 0: keep the lane's best members; 1: count the members that pass the second filter; 2: evaluate them in FP64
        auto do_cell = [&](int stage, int L, int c, const float (&gq)[2 * KP]) {
            if (sub == 0 && stage == 0) ++n_cells;
            // the cell's kp lines first (independent loads: their latencies overlap), then the costs, branch-free; only a
            // lane that holds a band member (rare) goes on to the bookkeeping below
            constexpr int RH = kChunkRows / G;
            static_assert(G == 8 && RH == 2, "the cell image is laid out for 8 row-lanes with 2 rows each");
            float2 rtt[RH];
            float val[RH][2 * KP];
            const float4 *cp = reinterpret_cast<const float4 *>(cell_img) + (((size_t)c * 32 + L) * KP) * 8 + sub;
            {
                float4 q[KP];
#pragma unroll
                for (int m = 0; m < KP; ++m) q[m] = cp[m * 8];
#pragma unroll
                for (int h = 0; h < RH; ++h) rtt[h] = pl.rowtab[min(c * kChunkRows + sub + G * h, pl.n_wspd_pad - 1)];
#pragma unroll
                for (int idx = 0; idx < 4 * KP; ++idx) {
                    const float4 t = q[idx >> 2];
                    val[idx / (2 * KP)][idx % (2 * KP)] = (idx & 3) == 0 ? t.x : ((idx & 3) == 1 ? t.y : ((idx & 3) == 2 ? t.z : t.w));
                }
            }
            unsigned hm = 0u;  // band members among this lane's 4 kp candidates of the cell (bit = h * 2 kp + slot)
#pragma unroll
            for (int h = 0; h < RH; ++h) {
                const bool row_ok = c * kChunkRows + sub + G * h < pl.n_wspd;
#pragma unroll
                for (int k = 0; k < 2 * KP; ++k) {  // exactly the scan's operations (padding slots give NaN or +inf)
                    const float lc = __fadd_rn(val[h][k], -rc.cs);
                    const float mm = __fmaf_rn(lc, lc, rtt[h].y);
                    const float J = __fmaf_rn(rtt[h].x, gq[k], __fmaf_rn(rc.nq, lc, mm));
                    hm |= (row_ok && J <= rc.thr) ? (1u << (h * 2 * KP + k)) : 0u;
                }
            }
            // the members (one per cell as a rule, in one lane of the group): values re-read and re-computed with the
            // same operations rather than selected out of the registers above
            while (hm) {
                const int idx = __ffs(hm) - 1;
                hm &= hm - 1;
                const int h = idx / (2 * KP), k = idx - h * 2 * KP;
                const int iw = c * kChunkRows + sub + G * h;
                const int ip = 2 * (L + 32 * (k >> 1)) + (k & 1);
                if (ip >= pl.n_phi) continue;
                const float v = reinterpret_cast<const float *>(cp)[(idx >> 2) * 32 + (idx & 3)];
                const float2 rt = h ? rtt[RH - 1] : rtt[0];
                const float gk = g32(px.qa, px.qb, pl.cos_phi[ip], pl.sin_phi[ip]);
                const float lc = __fadd_rn(v, -rc.cs);
                const float mm = __fmaf_rn(lc, lc, rt.y);
                const float J = __fmaf_rn(rt.x, gk, __fmaf_rn(rc.nq, lc, mm));
                if (!(J <= rc.thr)) continue;
                const int flat = iw * pl.n_phi + ip;
                const float jf = two_stage ? __fmaf_rn(rt.x, gk, __fmaf_rn(kfull, lc, mm)) : J;
                if (stage == 0) {
                    if (jf < b1) {
                        b3 = b2;
                        b2 = b1;
                        f2 = f1;
                        b1 = jf;
                        f1 = flat;
                    } else if (jf < b2) {
                        b3 = b2;
                        b2 = jf;
                        f2 = flat;
                    } else
                        b3 = fminf(b3, jf);
                } else if (jf <= thr2) {
                    if (stage == 1) {
                        ++n_loc;
                        one_loc = flat;
                    } else
                        fp64_feed(flat);
                }
            }
        };
        auto lane_g = [&](int L, float (&gq)[2 * KP]) {  // g(phi) of the scan lane's phi slots, bit-identical to the scan's
#pragma unroll
            for (int sl = 0; sl < 2 * KP; ++sl) {
                const int ip = 2 * (L + 32 * (sl >> 1)) + (sl & 1);
                gq[sl] = ip < pl.n_phi ? g32(px.qa, px.qb, pl.cos_phi[ip], pl.sin_phi[ip]) : 0.f;
            }
        };
        // walk the recorded cells.  stage 0: collect the band members; 1: count the members that pass the second filter;
        // 2: evaluate them in FP64 (stages 1 and 2 only when a lane holds more than two candidates)
        auto walk = [&](int stage) {
            unsigned lanes = rc.cont;
#pragma unroll 1
            for (int li = 0; lanes; ++li) {  // uniform within the group
                const int L = __ffs(lanes) - 1;
                lanes &= lanes - 1;
                float gq[2 * KP];
                lane_g(L, gq);
                // the chunks of this lane that came within the margin of its minimum (a bit covers 2^mask_sh chunks)
                unsigned cm = li == 0 ? rc.mask[0] : (li == 1 ? rc.mask[1] : rc.mask[2]);
                int c = 0, c_end = 0;
#pragma unroll 1
                for (;;) {
                    if (c == c_end) {
                        if (!cm) break;
                        c = (__ffs(cm) - 1) << mask_sh;
                        c_end = min(c + (1 << mask_sh), n_chunks);
                        cm &= cm - 1;
                    }
                    do_cell(stage, L, c, gq);
                    ++c;
                }
            }
        };
        auto group_sum = [&](int v) {
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
            return v;
        };
        auto group_max = [&](int v) {
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(gmask, v, o));
            return v;
        };
        if (work) {
            walk(0);
            if (two_stage) {
                float jmin = b1;
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) jmin = fminf(jmin, __shfl_xor_sync(gmask, jmin, o));
                thr2 = jmin + 2.f * rc.efp;
            }
            else
                thr2 = rc.thr;  // exact-k records: the members are the candidates
            const bool in1 = f1 >= 0 && b1 <= thr2, in2 = f2 >= 0 && b2 <= thr2;
            const bool overflow = group_max((b3 < CUDART_INF_F && b3 <= thr2) ? 1 : 0) != 0;  // a lane with more than two candidates
            int n_in, result;
            if (!overflow) {
                n_in = group_sum((in1 ? 1 : 0) + (in2 ? 1 : 0));
                result = group_max(in1 ? f1 : -1);  // the candidate itself when there is exactly one
                if (n_in > 1) {
                    if (in1) fp64_feed(f1);
                    if (in2) fp64_feed(f2);
                }
            } else {
                walk(1);
                n_in = group_sum(n_loc);
                result = group_max(one_loc);
                if (n_in > 1) walk(2);
            }
            if (n_in > 1) {  // FP64 with the reference's operation order over the candidates, first minimum wins
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) {
                    const double oj = __shfl_xor_sync(gmask, bj, o);
                    const int oi = __shfl_xor_sync(gmask, bi, o);
                    if (oj < bj || (oj == bj && oi < bi)) {
                        bj = oj;
                        bi = oi;
                    }
                }
                result = bi;
                if (sub == 0) ++n_fp64;
            }
            if (sub == 0) {
                if (n_in >= 1) {
                    write_co(pl, out, result, px.neg, px.px);
                    ++n_settled;
                } else  // cannot happen if the re-created costs equal the scan's; be safe
                    ws.fallback[atomicAdd(&ws.counters[1], 1ull)] = px.px;
            }
        }
    }
    __syncwarp();
    n_settled = __reduce_add_sync(0xffffffffu, n_settled);
    n_cells = __reduce_add_sync(0xffffffffu, n_cells);
    n_fp64 = __reduce_add_sync(0xffffffffu, n_fp64);
    n_many = __reduce_add_sync(0xffffffffu, n_many);
    if (lane == 0) {
        if (n_many) atomicAdd(&ws.counters[12], (u64)n_many);
        if (n_settled) atomicAdd(&ws.counters[2], (u64)n_settled);
        if (n_cells) atomicAdd(&ws.counters[3], (u64)n_cells);
        if (n_fp64) atomicAdd(&ws.counters[11], (u64)n_fp64);
    }
}

// ---- launch ----------------------------------------------------------------------------------------------------------
// pixels per warp P and CTAs per SM by phi pairs per lane KP (register budget: g[P][KP] packed pairs live across the slab).
// XS_SCAN_SHAPE="P,MB" (environment, development aid) selects another instantiation for KP == 3.
struct Shape {
    int p, nw, mb;
};
static Shape scan_shape(int kp) {
    if (kp > 3) return {4, 4, kp == 4 ? 3 : 2};
    Shape s = {8, 4, 4};
    if (kp == 3) {
        static int ep = -1, em = -1;
        if (ep < 0) {
            const char *e = getenv("XS_SCAN_SHAPE");
            ep = 0;
            if (e) sscanf(e, "%d,%d", &ep, &em);
        }
        if (ep == 8 && em == 3) s = {8, 4, 3};
        if (ep == 7 && em == 4) s = {7, 4, 4};
        if (ep == 6 && em == 4) s = {6, 4, 4};
    }
    return s;
}
int scan_tile_px(int kp) {
    const Shape s = scan_shape(kp);
    return s.p * s.nw;
}

template <int KP, int P, int NW, int MB>
static int launch_shape(const xs_plan *pl, const RasterArgs &ra, const Workspace &ws, const unsigned *sorted_px,
                        const OutSpec &out, int64_t n_px, xs_timer *timer, cudaStream_t st) {
    constexpr int TP = P * NW;
    static_assert(TP <= kTilePad, "tile size");
    int sms = kNumSMs;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
    XS_LAUNCH(k_list_prepare, sms * 16, 256, 0, st, *pl, ra, ws, sorted_px, TP);
    static int no_prune = -1;  // XS_NO_PRUNE: development aid (same as XS_FLAG_NO_PRUNE on every call)
    if (no_prune < 0) {
        const char *e = getenv("XS_NO_PRUNE");
        no_prune = e ? atoi(e) : 0;
    }
    const size_t plan_smem = sizeof(double2) * pl->n_chunks + sizeof(double) * 8 * (size_t)pl->n_chunks * pl->n_groups;
    if (plan_smem > 40 * 1024) XS_CUDA(cudaFuncSetAttribute(k_tile_plan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan_smem));
    XS_LAUNCH(k_tile_plan, sms * 16, 256, plan_smem, st, *pl, ws, TP, NW, kStages, (ra.flags & XS_FLAG_NO_PRUNE) || no_prune ? 0 : 1);

    static int scalar = -1;  // XS_SCAN_SCALAR: development aid (KP 3, P 8 only)
    if (scalar < 0) {
        const char *e = getenv("XS_SCAN_SCALAR");
        scalar = e ? atoi(e) : 0;
    }
    auto kern = k_scan_co<KP, P, NW, MB>;
    if constexpr (KP == 3 && P == 8 && NW == 4 && MB == 4)
        if (scalar) kern = k_scan_co<KP, P, NW, MB, true>;
    // largest estimated 2 |sigma| Lam for which a warp scans in shared-sigma0 mode (XS_SHARE_BUDGET: development aid; 0 = never)
    static float share_tau = -1.f;
    if (share_tau < 0.f) {
        const char *e = getenv("XS_SHARE_BUDGET");
        share_tau = e ? (float)atof(e) : 0.03f;
    }
    const size_t smem = sizeof(ScanSmem<KP, P, NW>) + sizeof(float2) * (size_t)pl->n_wspd_pad;
    if (smem > 200 * 1024) {
        set_error("xs_invert: wspd grid too long for the shared-memory row table");
        return XS_E_UNSUPPORTED;
    }
    // per launch, not once per process: the attribute belongs to the current device's context
    XS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int per_sm = 1;
    XS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem));
    if (per_sm < 1) per_sm = 1;
    if (timer) XS_CUDA(cudaEventRecord(timer->ev[0], st));
    XS_LAUNCH(kern, sms * per_sm, NW * 32, smem, st, *pl, ws, share_tau);
    if (timer) XS_CUDA(cudaEventRecord(timer->ev[1], st));
    static int refine_mb = -1;  // CTAs per SM of k_refine_easy (XS_REFINE_MB: development aid)
    if (refine_mb < 0) {
        const char *e = getenv("XS_REFINE_MB");
        refine_mb = e ? atoi(e) : 3;
    }
    if (refine_mb == 4)
        XS_LAUNCH((k_refine_easy<KP, 8, 4>), sms * 8, 256, 0, st, *pl, ws, out, TP);
    else if (refine_mb == 3)
        XS_LAUNCH((k_refine_easy<KP, 8, 3>), sms * 9, 256, 0, st, *pl, ws, out, TP);
    else
        XS_LAUNCH((k_refine_easy<KP, 8, 2>), sms * 8, 256, 0, st, *pl, ws, out, TP);
    if (timer) {
        XS_CUDA(cudaEventRecord(timer->ev[2], st));
        timer->recorded = 1;
    }
    return XS_OK;
}

int launch_scan_pipeline(const xs_plan *pl, const RasterArgs &ra, const Workspace &ws, const unsigned *sorted_px,
                         const OutSpec &out, int64_t n_px, xs_timer *timer, cudaStream_t st) {
    const Shape s = scan_shape(pl->kp);
#define XS_SHAPE_NW(KP_, P_, NW_, MB_) \
    if (pl->kp == KP_ && s.p == P_ && s.nw == NW_ && s.mb == MB_) \
    return launch_shape<KP_, P_, NW_, MB_>(pl, ra, ws, sorted_px, out, n_px, timer, st)
#define XS_SHAPE(KP_, P_, MB_) XS_SHAPE_NW(KP_, P_, 4, MB_)
    XS_SHAPE(1, 8, 4);
    XS_SHAPE(2, 8, 4);
    XS_SHAPE(3, 8, 4);
    XS_SHAPE(3, 8, 3);
    XS_SHAPE(3, 7, 4);
    XS_SHAPE(3, 6, 4);
    XS_SHAPE(4, 4, 3);
    XS_SHAPE(6, 4, 2);
#undef XS_SHAPE
#undef XS_SHAPE_NW
    set_error("xs_invert: no scan instantiation for kp=%d", pl->kp);
    return XS_E_UNSUPPORTED;
}

}  // namespace xs
