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
// Three kernels, all on the caller's stream, no host synchronisation:
//   k_list_prepare  orders runs of the bin-grouped pixel list by sigma0 (so that a warp's pixels have a small |sigma|) and
//                   materialises one 32-byte PixRec per list position
//   k_scan_co       persistent CTAs (4 per SM, 4 warps, 128 registers): a tile = 32 list positions of one bin; the tile's
//                   PixRecs and the bin's slab of the FP32 scan image arrive by bulk-async (TMA) copies -- records
//                   double-buffered one tile ahead, slab chunks of 16 wspd rows through a 3-stage ring that keeps
//                   streaming across tile boundaries; there is no producer thread and no CTA barrier: the warp that is
//                   the last to finish a chunk refills its stage (shared-memory arrival counter), so no warp ever waits
//                   for another one to release a stage.  Lane l owns the phi pairs {2(l+32j), 2(l+32j)+1}; per lane and
//                   pixel only the best 16-row chunk, its index and the runner-up chunk minimum are kept.  After the
//                   slab: warp-shuffle min, the rigorous FP32 error band E, and one RefRec per pixel.
//   k_refine_co     a warp per pixel: re-creates the FP32 costs (bit-identical operations) of the (lane, chunk) cells
//                   inside the band; a single member settles the pixel, several are evaluated in FP64 with the reference's
//                   operation order and reduced by a warp-shuffle lexicographic (J, index) argmin = numpy's first minimum.
#include <stdio.h>
#include <stdlib.h>

#include "xs_invert.cuh"

namespace xs {

constexpr unsigned kTileEnd = 0xffffffffu;
constexpr int kTileBatch = 8;   // tiles a CTA takes per global atomic
constexpr int kSortRun = 256;   // list positions ordered together by k_list_prepare (a whole number of tiles fits)

// ---- list preparation: local order by sigma0 + pixel records ----------------------------------------------------------
// The centred scan shares (L - c)^2 between the 8 pixels of a warp, which keeps its error band tight only if those pixels
// have similar sigma0.  Every run of list positions (<= 256, a whole number of tiles, possibly across bin boundaries) is
// sorted by (incidence bin, sigma0) in shared memory: bins stay contiguous and in order, padding stays at the end of its
// bin's segment, and a warp's 8 pixels span 1/32 of the run's sigma0 range.
__global__ void __launch_bounds__(kSortRun) k_list_prepare(xs_plan pl, RasterArgs a, Workspace ws, int tile_px) {
    __shared__ unsigned long long key[kSortRun];
    __shared__ unsigned val[kSortRun];
    const unsigned n_tiles = (unsigned)ws.counters[0];
    const unsigned tiles_per_run = kSortRun / tile_px;
    const unsigned t0 = blockIdx.x * tiles_per_run;
    if (t0 >= n_tiles) return;
    const unsigned first = t0 * tile_px;
    const unsigned count = (min(t0 + tiles_per_run, n_tiles) - t0) * tile_px;  // <= kSortRun
    unsigned long long k = ~0ull;
    unsigned v = 0xffffffffu;
    int bin = 0;
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
        bin = lo;
        unsigned b = 0xffffffffu;  // padding sorts behind every pixel of its bin
        if (v != 0xffffffffu) {
            const float sf = (float)load_real(a.s_co, v, a.dtype);  // linear or dB: monotone either way
            b = __float_as_uint(sf);
            b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // order-preserving map of the float bits
            if (b == 0xffffffffu) b = 0xfffffffeu;
        }
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
    if (threadIdx.x >= count) return;
    // the bin of a position does not change (segments keep their extent); its pixel does
    const unsigned px = val[threadIdx.x];
    PixRec r;
    r.qa = r.qb = r.s = 0.0;
    r.px = px;
    r.bin = (unsigned short)bin;
    r.state = 0;
    r.neg = 0;
    if (px != 0xffffffffu) {
        const double2 anc = load_cplx(a.anc, px, a.dtype);
        const double s_raw = load_real(a.s_co, px, a.dtype);
        r.qa = anc.x;
        r.qb = pl.phi_180 ? fabs(anc.y) : anc.y;
        r.neg = anc.y < 0.0;
        r.s = (a.flags & XS_FLAG_SIGMA0_DB) ? s_raw : to_db(s_raw);
        const bool finite_q = isfinite(r.qa) && isfinite(r.qb) && isfinite(r.s);
        r.state = !finite_q ? 3 : (pl.first_nan[bin] >= 0 ? 2 : 1);
    }
    ws.list[first + threadIdx.x] = px;
    ws.pix[first + threadIdx.x] = r;
}

// ---- the FP32 scan ----------------------------------------------------------------------------------------------------
template <int KP, int P, int NW>
struct ScanSmem {
    static constexpr int kRowFloats = 64 * KP;
    alignas(128) float ring[kStages][kChunkRows * kRowFloats];
    alignas(16) PixRec pix[2][NW * P];
    alignas(16) double2 cs_phi[64 * KP];  // (cos, sin) of the phi node, (0, 0) in the padding
    alignas(8) uint64_t full[kStages];    // a chunk has landed in the stage
    uint64_t pix_full[2];                 // the tile's records (or the end-of-work mark) have landed
    unsigned done[kStages];               // warps that have finished with the stage's chunk (monotonic)
    unsigned tile_of[2];
    unsigned batch_next, batch_end;
};

template <int KP, int P, int NW, int MB>
__global__ void __launch_bounds__(NW * 32, MB) k_scan_co(xs_plan pl, Workspace ws) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using Smem = ScanSmem<KP, P, NW>;
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    float2 *rowtab_s = reinterpret_cast<float2 *>(smem_raw + sizeof(Smem));  // [n_wspd_pad]
    constexpr int TP = NW * P;                                               // list positions per tile
    constexpr int NS = kStages;
    constexpr uint32_t kPixBytes = TP * sizeof(PixRec);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned n_tiles = (unsigned)ws.counters[0];
    const int n_chunks = (pl.n_wspd_pad + kChunkRows - 1) / kChunkRows;  // >= NS (plan creation checks); the last may be shorter
    const size_t slab_floats = (size_t)pl.n_wspd_pad * Smem::kRowFloats;

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
            mbar_expect_tx(&sm.pix_full[buf], kPixBytes);
            bulk_g2s(sm.pix[buf], ws.pix + (size_t)t * TP, kPixBytes, &sm.pix_full[buf]);
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
    for (int i = threadIdx.x; i < 64 * KP; i += blockDim.x)
        sm.cs_phi[i] = i < pl.n_phi ? make_double2(pl.cos_phi[i], pl.sin_phi[i]) : make_double2(0.0, 0.0);
    __syncthreads();
    if (threadIdx.x == 0) {  // first tile and the first NS chunks of its slab
        fetch_tile(0);
        if (sm.tile_of[0] != kTileEnd) {
            mbar_wait(&sm.pix_full[0], 0);
            const int bin0 = sm.pix[0][0].bin;
            for (int c = 0; c < NS; ++c) load_chunk(bin0, c, c);
        }
    }

    int stage = 0;
    uint32_t phase = 0;  // parity of full[stage] for the chunk this warp consumes next
    for (unsigned u = 0;; ++u) {
        const int b = u & 1;
        mbar_wait(&sm.pix_full[b], (u >> 1) & 1);
        const unsigned tile = sm.tile_of[b];
        if (tile == kTileEnd) break;
        const PixRec *mine = &sm.pix[b][warp * P];
        const int bin = sm.pix[b][0].bin;  // position 0 of a tile is never padding

        // ---- per-warp centre and per-lane per-pixel query constants ----
        bool any = false;
        float cs = 0.f;
        {
            double smin = CUDART_INF, smax = -CUDART_INF;
#pragma unroll
            for (int p = 0; p < P; ++p)
                if (mine[p].state == 1) {
                    const double v = mine[p].s / pl.dsig_co;
                    smin = fmin(smin, v);
                    smax = fmax(smax, v);
                    any = true;
                }
            if (any) cs = (float)(0.5 * (smin + smax));
        }
        float nqs[P];   // k_p = -2 (s_p/dsig - cs)
        u64 g[P][KP];   // {g(phi_even), g(phi_odd)} as packed FP32
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const bool on = mine[p].state == 1;
            nqs[p] = on ? (float)(-2.0 * (mine[p].s / pl.dsig_co - (double)cs)) : 0.f;
            const double qa = mine[p].qa, qb = mine[p].qb;
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                const double2 c0 = sm.cs_phi[2 * (lane + 32 * j)], c1 = sm.cs_phi[2 * (lane + 32 * j) + 1];
                g[p][j] = on ? pack2(g32(qa, qb, c0.x, c0.y), g32(qa, qb, c1.x, c1.y)) : 0ull;
            }
        }
        float m[P], best[P], second[P];
        unsigned bch[(P + 3) / 4];  // index of the best chunk, one byte per pixel (n_chunks <= 256)
#pragma unroll
        for (int p = 0; p < P; ++p) m[p] = best[p] = second[p] = CUDART_INF_F;
#pragma unroll
        for (int w = 0; w < (P + 3) / 4; ++w) bch[w] = 0u;
        const u64 ncs2 = pack2(-cs, -cs);

        // ---- the slab, 16 wspd rows at a time ----
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(&sm.full[stage], phase);
            if (any) {
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
                const unsigned old = atom_add_acq_rel_shared(&sm.done[stage], 1u);
                if (old % NW == NW - 1) {
                    if (c == 0) fetch_tile(b ^ 1);
                    int c2 = c + NS, bin2 = bin;
                    bool ok = true;
                    if (c2 >= n_chunks) {  // belongs to the next tile
                        c2 -= n_chunks;
                        mbar_wait(&sm.pix_full[b ^ 1], ((u + 1) >> 1) & 1);
                        ok = sm.tile_of[b ^ 1] != kTileEnd;
                        if (ok) bin2 = sm.pix[b ^ 1][0].bin;
                    }
                    if (ok) {
                        fence_proxy_async();
                        load_chunk(bin2, c2, stage);
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const bool lt = m[p] < best[p];
                second[p] = fminf(second[p], fmaxf(best[p], m[p]));
                best[p] = fminf(best[p], m[p]);
                constexpr unsigned kSel[4] = {0x3214u, 0x3240u, 0x3410u, 0x4210u};  // byte p & 3 <- c
                bch[p >> 2] = lt ? __byte_perm(bch[p >> 2], (unsigned)c, kSel[p & 3]) : bch[p >> 2];
                m[p] = CUDART_INF_F;
            }
            if (++stage == NS) {
                stage = 0;
                phase ^= 1u;
            }
        }
        if (!any) continue;

        // ---- band of every pixel -> RefRec ---------------------------------------------------------------------------
        // m32 = warp-shuffle min of the FP32 costs; E bounds |J''_fp32 - J''_exact| for every candidate that can still
        // win (derivation: DESIGN.md 4.1, checked on the CPU by tests/test_error_bound.py), so the reference's FP64 argmin
        // lies in S = {c : J''_fp32(c) <= m32 + 2E}; k_refine_co collects S from the cells recorded here.
        const float lmax = pl.slab_absmax[bin];
        const float W = (float)pl.w_absmax * 1.0000002f;
        constexpr int cap = 8, bits = 8;  // a record holds the best-chunk index of 8 lanes
#pragma unroll
        for (int p = 0; p < P; ++p) {
            if (mine[p].state != 1) continue;  // warp-uniform
            float m32 = best[p];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m32 = fminf(m32, __shfl_xor_sync(0xffffffffu, m32, o));
            const float fa = fabsf((float)mine[p].qa), fb = fabsf((float)mine[p].qb);
            const float A = sqrtf(fa * fa + fb * fb) * 1.000001f;  // >= |ancillary|
            const float SC = fabsf((float)(mine[p].s / pl.dsig_co - (double)cs)) * 1.0000002f;
            const float T = W * A + 0.25f * W * W;
            // J'' = J' - sc^2: candidates that can still win have |L/dsig - s/dsig| <= D and |L/dsig - cs| <= Lam = D + |sc|.
            // Error terms (u = 2^-24): image value and lambda roundings 2 Lam (lmax + Lam) through lambda^2 and
            // 2 |sc| (lmax + 2 Lam) through k lambda; M, a and J roundings Lam^2 + W^2/4, Lam^2 + W^2/4 + 2 |sc| Lam and
            // D^2 + sc^2 + T; row-table and g roundings W^2/4 + W A; the W terms add up to W^2 + 2 W A <= 4 T.
            const float D = sqrtf(fmaxf(m32 + SC * SC * 1.0000002f, 0.f) + 0.25f * A * A + 1.0f);
            const float Lam = D + SC;
            const float E = 5.9604645e-8f * 1.5f * (2.f * Lam * lmax + 2.f * SC * lmax + 4.f * Lam * Lam + 6.f * SC * Lam + SC * SC + D * D + 4.f * T);
            const float thr = m32 + 2.f * E;
            const bool sane = (E < 0.25f) && (m32 < CUDART_INF_F);  // else: magnitudes outside the range of the bound
            unsigned cont = __ballot_sync(0xffffffffu, sane && best[p] <= thr);
            // lanes holding two or more chunks inside the band: all of the lane's candidates are looked at
            unsigned wide = __ballot_sync(0xffffffffu, sane && second[p] <= thr);
            cont &= ~wide;
            // the record holds the best-chunk index of `cap` cont lanes; further ones are treated like wide lanes
            const bool mine_c = (cont >> lane) & 1u;
            const int rank = __popc(cont & ((1u << lane) - 1u));
            const unsigned over = __ballot_sync(0xffffffffu, mine_c && rank >= cap);
            cont &= ~over;
            wide |= over;
            const unsigned bchunk = (bch[p >> 2] >> (8 * (p & 3))) & 0xffu;
            const u64 idv = (mine_c && rank < cap) ? ((u64)bchunk << (bits * rank)) : 0ull;
            const unsigned ch_lo = __reduce_or_sync(0xffffffffu, (unsigned)idv);
            const unsigned ch_hi = __reduce_or_sync(0xffffffffu, (unsigned)(idv >> 32));
            if (lane == 0) {
                RefRec rr;
                rr.thr = thr;
                rr.cs = cs;
                rr.nq = nqs[p];
                rr.cont = cont;
                rr.wide = wide;
                rr.ch_lo = ch_lo;
                rr.ch_hi = ch_hi;
                rr.spare = 0;
                ws.rec[(size_t)tile * TP + warp * P + p] = rr;
            }
        }
    }
}

// ---- exact refinement ---------------------------------------------------------------------------------------------------
// FP32 cost of candidate k of cell (lane L, rows row0...) of a pixel, exactly as the scan computed it; true if it is inside
// the pixel's band.  flat = w * n_phi + phi index of the candidate.
template <int KP>
__device__ __forceinline__ bool band_member(const xs_plan &pl, const PixRec &px, const RefRec &rc, int L, int row0, int k,
                                            int n_cand, int &flat) {
    const int iw = row0 + k / (2 * KP);
    const int slot = k % (2 * KP);
    const int ip = 2 * (L + 32 * (slot >> 1)) + (slot & 1);
    flat = iw * pl.n_phi + ip;
    if (k >= n_cand || iw >= pl.n_wspd || ip >= pl.n_phi) return false;
    const float2 rt = pl.rowtab[iw];
    const float *slab32 = pl.scan + (size_t)px.bin * pl.n_wspd_pad * pl.nph_pad;
    const float lc = __fadd_rn(slab32[(size_t)iw * pl.nph_pad + ip], -rc.cs);
    const float mm = __fmaf_rn(lc, lc, rt.y);
    const float aa = __fmaf_rn(rc.nq, lc, mm);
    return __fmaf_rn(rt.x, g32(px.qa, px.qb, pl.cos_phi[ip], pl.sin_phi[ip]), aa) <= rc.thr;
}

// Pass 1: eight lanes per list position (four positions per warp at a time, so four times as many pixels are in flight
// as with a warp per pixel -- this pass is bound by the latency of its dependent loads and by instruction issue, not by
// arithmetic).  Settles padding, NaN slabs (answer = first NaN), pixels for the exhaustive kernel, and every pixel whose
// band touches only "cont" cells (one 16-row chunk of one lane each, at most 8): lane `sub` of the group owns rows sub and
// sub + 8 of a cell and all 2 KP phi slots, so g(phi) is evaluated once per slot; a single band member settles the pixel,
// several are evaluated in FP64 with the reference's operation order and reduced to the lexicographic (J, flat index)
// minimum = numpy's first minimum.  Pixels with "wide" lanes (several chunks of one lane in the band) go to pass 2.
template <int KP>
__global__ void __launch_bounds__(256, 3) k_refine_easy(xs_plan pl, Workspace ws, OutSpec out, int tile_px) {
    const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
    const unsigned gmask = 0xffu << (8 * grp);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_pos = (int64_t)ws.counters[0] * tile_px;
    static_assert(kChunkRows == 16, "two rows per lane of a group");
    unsigned n_settled = 0, n_cells = 0, n_fp64 = 0;
    for (int64_t e0 = warp * 4; e0 < n_pos; e0 += n_warps * 4) {
        const int64_t e = e0 + grp;
        if (e >= n_pos) continue;
        const PixRec px = ws.pix[e];
        const RefRec rc = ws.rec[e];  // loaded together with the pixel (only meaningful for state 1)
        if (px.state == 0) continue;  // uniform within the group; no warp-wide synchronisation below
        if (px.state != 1 || (rc.cont | rc.wide) == 0u) {
            if (sub == 0) {
                if (px.state == 2)
                    write_co(pl, out, pl.first_nan[px.bin], px.neg, px.px);  // J is NaN exactly where L is NaN
                else  // non-finite inputs, or the scan could not bound its error for this pixel
                    ws.fallback[atomicAdd(&ws.counters[1], 1ull)] = px.px;
            }
            continue;
        }
        if (rc.wide != 0u) {
            if (sub == 0) ws.hard[atomicAdd(&ws.counters[12], 1ull)] = (unsigned)e;
            continue;
        }
        const float *slab32 = pl.scan + (size_t)px.bin * pl.n_wspd_pad * pl.nph_pad;
        const double *slab64 = pl.co_lut + (size_t)px.bin * pl.n_wspd * pl.n_phi;
        int n_loc = 0, one_loc = -1;
        double bj = CUDART_INF;  // FP64 pass: lexicographic (J, flat index) minimum of this lane's members
        int bi = 0x7fffffff;
        // walk the cont cells: count the band members (exact == false) or evaluate them in FP64 (true)
        auto walk = [&](bool exact) {
            unsigned cells = rc.cont;
            u64 ids = ((u64)rc.ch_hi << 32) | rc.ch_lo;
#pragma unroll 1
            while (cells) {  // uniform within the group
                const int L = __ffs(cells) - 1;
                cells &= cells - 1;
                const int row0 = (int)(ids & 0xffu) * kChunkRows + sub;
                ids >>= 8;
                float gq[2 * KP];
#pragma unroll
                for (int sl = 0; sl < 2 * KP; ++sl) {
                    const int ip = 2 * (L + 32 * (sl >> 1)) + (sl & 1);
                    gq[sl] = ip < pl.n_phi ? g32(px.qa, px.qb, pl.cos_phi[ip], pl.sin_phi[ip]) : 0.f;
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int iw = row0 + 8 * h;
                    if (iw >= pl.n_wspd) continue;
                    const float2 rt = pl.rowtab[iw];
                    const float2 *rowp = reinterpret_cast<const float2 *>(slab32 + (size_t)iw * pl.nph_pad) + L;
#pragma unroll
                    for (int j = 0; j < KP; ++j) {
                        const float2 v = rowp[32 * j];
#pragma unroll
                        for (int o = 0; o < 2; ++o) {  // exactly the scan's operations
                            const int ip = 2 * (L + 32 * j) + o;
                            const float lc = __fadd_rn(o ? v.y : v.x, -rc.cs);
                            const float aa = __fmaf_rn(rc.nq, lc, __fmaf_rn(lc, lc, rt.y));
                            if (ip >= pl.n_phi || !(__fmaf_rn(rt.x, gq[2 * j + o], aa) <= rc.thr)) continue;
                            const int flat = iw * pl.n_phi + ip;
                            if (!exact) {
                                ++n_loc;
                                one_loc = flat;
                            } else {  // J is never NaN here: finite inputs, NaN-free slab
                                const double J = exact_cost_co(pl.wspd_grid[iw], pl.cos_phi[ip], pl.sin_phi[ip], slab64[flat], px.qa,
                                                               px.qb, px.s, pl.dsig_co);
                                if (J < bj || (J == bj && flat < bi)) {
                                    bj = J;
                                    bi = flat;
                                }
                            }
                        }
                    }
                }
                if (sub == 0 && !exact) ++n_cells;
            }
        };
        walk(false);
        int n_in = n_loc, result = one_loc;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            n_in += __shfl_xor_sync(gmask, n_in, o);
            result = max(result, __shfl_xor_sync(gmask, result, o));  // the member itself when there is exactly one
        }
        if (n_in > 1) {  // FP64 with the reference's operation order over the members, first minimum wins
            walk(true);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
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
    n_settled = __reduce_add_sync(0xffffffffu, n_settled);
    n_cells = __reduce_add_sync(0xffffffffu, n_cells);
    n_fp64 = __reduce_add_sync(0xffffffffu, n_fp64);
    if (lane == 0) {
        if (n_settled) atomicAdd(&ws.counters[2], (u64)n_settled);
        if (n_cells) atomicAdd(&ws.counters[3], (u64)n_cells);
        if (n_fp64) atomicAdd(&ws.counters[11], (u64)n_fp64);
    }
}

// Pass 2: a warp per hard pixel (RP at a time so that their dependent loads overlap): collects the band members of all
// contending cells; a single member settles the pixel, several are evaluated in FP64 with the reference's operation order
// and reduced by a warp-shuffle lexicographic (J, flat index) argmin = numpy's first minimum.
template <int KP, int RP>
__global__ void __launch_bounds__(256, 3) k_refine_co(xs_plan pl, Workspace ws, OutSpec out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_hard = (int64_t)ws.counters[12];
    const int n_chunks = (pl.n_wspd_pad + kChunkRows - 1) / kChunkRows;
    constexpr int kCand = kChunkRows * 2 * KP;
    constexpr int kIter = (kCand + 31) / 32;
    u64 n_scanned = 0, n_cells = 0, n_fp64 = 0;

    for (int64_t h0 = warp * RP; h0 < n_hard; h0 += n_warps * RP) {
        PixRec px[RP];
        RefRec rc[RP];
        bool act[RP];
#pragma unroll
        for (int i = 0; i < RP; ++i) {
            act[i] = h0 + i < n_hard;
            if (act[i]) {
                const unsigned e = ws.hard[h0 + i];
                px[i] = ws.pix[e];
                rc[i] = ws.rec[e];
            }
        }
        auto chunk_of = [&](int i, int L) {  // best chunk of cont lane L
            const int rank = __popc(rc[i].cont & ((1u << L) - 1u));
            const u64 ids = ((u64)rc[i].ch_hi << 32) | rc[i].ch_lo;
            return (int)((ids >> (8 * rank)) & 0xffu);
        };
        // membership in S of the candidates of the first contender cell of every pixel (loads batched)
        bool in0[RP][kIter];
        int flat0[RP][kIter];
        unsigned rest[RP];  // contender cells not looked at yet
#pragma unroll
        for (int i = 0; i < RP; ++i) {
            rest[i] = act[i] ? rc[i].cont : 0u;
#pragma unroll
            for (int q = 0; q < kIter; ++q) {
                in0[i][q] = false;
                flat0[i][q] = 0;
            }
            if (rest[i]) {  // warp-uniform
                const int L = __ffs(rest[i]) - 1;
                rest[i] &= rest[i] - 1;
                const int row0 = chunk_of(i, L) * kChunkRows;
#pragma unroll
                for (int q = 0; q < kIter; ++q) in0[i][q] = band_member<KP>(pl, px[i], rc[i], L, row0, lane + 32 * q, kCand, flat0[i][q]);
                ++n_cells;
            }
        }
        // count the members, look at the remaining cells, settle
#pragma unroll
        for (int i = 0; i < RP; ++i) {
            if (!act[i]) continue;
            int n_loc = 0, one_loc = -1;  // members of S seen by this lane so far (count, and the flat index of one of them)
#pragma unroll
            for (int q = 0; q < kIter; ++q)
                if (in0[i][q]) {
                    ++n_loc;
                    one_loc = flat0[i][q];
                }
            ArgMin am;
            am.init();
            const double *slab64 = pl.co_lut + (size_t)px[i].bin * pl.n_wspd * pl.n_phi;
            // generic walk over cell (L, row0, n_cand): note members (exact == false) or FP64 argmin (true)
            auto visit = [&](int L, int row0, int n_cand, bool exact) {
                for (int k0 = 0; k0 < n_cand; k0 += 32) {
                    int flat;
                    if (!band_member<KP>(pl, px[i], rc[i], L, row0, k0 + lane, n_cand, flat)) continue;
                    if (!exact) {
                        ++n_loc;
                        one_loc = flat;
                    } else {
                        const int iw = flat / pl.n_phi, ip = flat - iw * pl.n_phi;
                        am.feed(exact_cost_co(pl.wspd_grid[iw], pl.cos_phi[ip], pl.sin_phi[ip], slab64[flat], px[i].qa, px[i].qb,
                                              px[i].s, pl.dsig_co), flat);
                    }
                }
            };
            auto sweep = [&](unsigned cells, unsigned lanes, bool exact) {
                while (cells) {
                    const int L = __ffs(cells) - 1;
                    cells &= cells - 1;
                    visit(L, chunk_of(i, L) * kChunkRows, kCand, exact);
                    if (!exact) ++n_cells;
                }
                while (lanes) {
                    const int L = __ffs(lanes) - 1;
                    lanes &= lanes - 1;
                    visit(L, 0, pl.n_wspd * 2 * KP, exact);
                    if (!exact) n_cells += n_chunks;
                }
            };
            if (rest[i] | rc[i].wide) sweep(rest[i], rc[i].wide, false);
            const int n_in = __reduce_add_sync(0xffffffffu, n_loc);
            int result = __reduce_max_sync(0xffffffffu, one_loc);  // the member itself when n_in == 1
            if (n_in > 1) {
                sweep(rc[i].cont, rc[i].wide, true);
                am.warp_reduce();
                result = am.result();
                ++n_fp64;
            }
            if (lane == 0) {
                if (n_in >= 1)
                    write_co(pl, out, result, px[i].neg, px[i].px);
                else  // cannot happen if the re-created costs equal the scan's; be safe
                    ws.fallback[atomicAdd(&ws.counters[1], 1ull)] = px[i].px;
            }
            ++n_scanned;
        }
    }
    if (lane == 0) {
        if (n_scanned) atomicAdd(&ws.counters[2], n_scanned);
        if (n_cells) atomicAdd(&ws.counters[3], n_cells);
        if (n_fp64) atomicAdd(&ws.counters[11], n_fp64);
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
        if (ep == 6 && em == 5) s = {6, 4, 5};
    }
    return s;
}
int scan_tile_px(int kp) {
    const Shape s = scan_shape(kp);
    return s.p * s.nw;
}

template <int KP, int P, int NW, int MB>
static int launch_shape(const xs_plan *pl, const RasterArgs &ra, const Workspace &ws, const OutSpec &out, int64_t n_px,
                        xs_timer *timer, cudaStream_t st) {
    constexpr int TP = P * NW;
    static_assert(TP <= kTilePad && TP <= kSortRun, "tile size");
    int sms = kNumSMs;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
    const int64_t max_tiles = ceil_div(n_px, TP) + pl->n_inc;
    XS_LAUNCH(k_list_prepare, (unsigned)ceil_div(max_tiles, kSortRun / TP), kSortRun, 0, st, *pl, ra, ws, TP);

    auto kern = k_scan_co<KP, P, NW, MB>;
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
    XS_LAUNCH(kern, sms * per_sm, NW * 32, smem, st, *pl, ws);
    if (timer) XS_CUDA(cudaEventRecord(timer->ev[1], st));
    XS_LAUNCH(k_refine_easy<KP>, sms * 6, 256, 0, st, *pl, ws, out, TP);
    XS_LAUNCH((k_refine_co<KP, 2>), sms * 6, 256, 0, st, *pl, ws, out);
    if (timer) {
        XS_CUDA(cudaEventRecord(timer->ev[2], st));
        timer->recorded = 1;
    }
    return XS_OK;
}

int launch_scan_pipeline(const xs_plan *pl, const RasterArgs &ra, const Workspace &ws, const OutSpec &out, int64_t n_px,
                         xs_timer *timer, cudaStream_t st) {
    const Shape s = scan_shape(pl->kp);
#define XS_SHAPE(KP_, P_, MB_) \
    if (pl->kp == KP_ && s.p == P_ && s.mb == MB_) return launch_shape<KP_, P_, 4, MB_>(pl, ra, ws, out, n_px, timer, st)
    XS_SHAPE(1, 8, 4);
    XS_SHAPE(2, 8, 4);
    XS_SHAPE(3, 8, 4);
    XS_SHAPE(3, 8, 3);
    XS_SHAPE(3, 7, 4);
    XS_SHAPE(3, 6, 4);
    XS_SHAPE(3, 6, 5);
    XS_SHAPE(4, 4, 3);
    XS_SHAPE(6, 4, 2);
#undef XS_SHAPE
    set_error("xs_invert: no scan instantiation for kp=%d", pl->kp);
    return XS_E_UNSUPPORTED;
}

}  // namespace xs
