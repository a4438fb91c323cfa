/*
 * xsarsea_b200.h -- C ABI of libxsarsea_b200.so: the B200 (sm_100a) implementation of xsarsea's
 * wind-inversion hot path.  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * Each entry point names the reference interface it replaces (paths relative to
 * /root/reference/src/xsarsea/).  The reference has no FFI of its own (it is pure Python + numba);
 * the operator boundary a native library sits behind is the numpy-level gufunc layer
 * (SURVEY.md section 8 B2), so that is what these functions mirror.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error (XS_E_*); xs_last_error() gives the text;
 *    nothing throws, nothing is printed.
 *  - "dev" pointers are CUDA device pointers on the current device; "host" pointers are ordinary host
 *    memory (small grids / tables that the host computes with numpy so they carry numpy's rounding).
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *    asynchronous with respect to the host unless stated otherwise.
 *  - no function allocates its outputs; xs_invert takes a caller-provided workspace.
 *  - missing data is NaN, never an error (windspeed/windspeed.py:198-207).
 */
#ifndef XSARSEA_B200_H
#define XSARSEA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XS_ABI_VERSION 2

/* error codes */
#define XS_OK 0
#define XS_E_INVALID (-1)   /* bad argument */
#define XS_E_CUDA (-2)      /* CUDA runtime error (text in xs_last_error) */
#define XS_E_BOUNDS (-3)    /* interpolation target outside source grid (scipy bounds_error=True) */
#define XS_E_WORKSPACE (-4) /* workspace too small */
#define XS_E_UNSUPPORTED (-5)

/* Built-in analytic GMFs, in registration order of windspeed/gmfs_impl.py:207-210,213,325,388,451,517,555,593,632,671 */
enum xs_gmf_id {
    XS_GMF_CMOD5 = 0,
    XS_GMF_CMOD5N = 1,
    XS_GMF_CMOD5N_PR_ZHANGA = 2,
    XS_GMF_CMOD5N_PR_MOUCHE1 = 3,
    XS_GMF_CMODIFR2 = 4,
    XS_GMF_RS2_V2 = 5,
    XS_GMF_S1_V2 = 6,
    XS_GMF_RCM_NOAA = 7,
    XS_GMF_S1_V3_EW_REC = 8,
    XS_GMF_RS2_V3 = 9,
    XS_GMF_RCM_V3 = 10,
    XS_GMF_RCM_V4 = 11,
    XS_GMF_RS2_V4 = 12,
    XS_GMF_COUNT = 13
};

/* element types of rasters */
#define XS_F64 0 /* float64 rasters, complex128 ancillary (the reference gufunc types, windspeed.py:309-317) */
#define XS_F32 1 /* float32 rasters, complex64 ancillary (promoted to f64 on load) */

int xs_abi_version(void);
const char *xs_last_error(void);
/* Number of kernel launches issued by this library in this process so far (bench.py's gpu_launches). */
int64_t xs_launch_count(void);

/* Measurement aid (bench.py): FP32 FMA-pipe peak of the current device in TFLOP/s, measured with a register-resident
 * fma.rn.f32x2 loop (best of three ~10 ms launches on `stream`, synchronous). */
int xs_bench_fp32_peak(double *tflops, void *stream);

/* A pair of CUDA events owned by the library, handed to xs_invert through xs_invert_args.scan_timer to time the co-pol
 * scan on the launching stream (bench.py's roofline of the dominant kernel).  One timer per in-flight call: timers, like
 * workspaces, belong to the call and not to the plan, so concurrent calls on one plan do not share any mutable state. */
typedef struct xs_timer xs_timer;
int xs_timer_create(xs_timer **timer_out);
void xs_timer_destroy(xs_timer *timer);
/* Waits for the second event; XS_E_INVALID if the timer has not been recorded by an xs_invert yet. */
int xs_timer_elapsed_ms(xs_timer *timer, float *ms);

/* ---- GMF evaluation ------------------------------------------------------------------------- */

/* K3/K4: element-wise sigma0_linear = gmf(inc, wspd, phi) over n already-broadcast elements.
 * Replaces the numba `vectorize` kernels of windspeed/gmfs.py:210-214 (signatures ddd->d and ffd->f:
 * dtype selects the type of inc/wspd/out; phi is always float64; phi_dev may be NULL for cross-pol GMFs). */
int xs_gmf_eval(int model_id, int dtype, const void *inc_dev, const void *wspd_dev, const double *phi_dev,
                void *out_dev, int64_t n, void *stream);

/* K2: outer-product LUT out[n_inc][n_wspd][n_phi] (n_phi = 0 and phi NULL for cross-pol: out[n_inc][n_wspd]).
 * Replaces the numba `guvectorize "(n),(m),(p)->(n,m,p)"` kernel of windspeed/gmfs.py:218-230 as used by
 * GmfModel._raw_lut (gmfs.py:350-395).  Grids are host arrays built with the reference's np.linspace
 * expression (gmfs.py:385-390). */
int xs_lut_build(int model_id, const double *inc_grid_host, int n_inc, const double *wspd_grid_host, int n_wspd,
                 const double *phi_grid_host, int n_phi, double *out_dev, void *stream);

/* K5: linear interpolation along one axis of src viewed as [outer][n_src][inner] -> dst [outer][n_dst][inner],
 * with the arithmetic of scipy.interpolate.interp1d(kind="linear", bounds_error=True) that
 * xarray.DataArray.interp applies per dimension at windspeed/models.py:167.  XS_E_BOUNDS if a target lies
 * outside [x_src[0], x_src[n_src-1]]. */
int xs_lut_interp_axis(const double *src_dev, int64_t outer, int n_src, int64_t inner, const double *x_src_host,
                       const double *x_dst_host, int n_dst, double *dst_dev, void *stream);

/* K6: unit conversion of windspeed/models.py:215 (10*log10(x+1e-15)) and :221 (10**(x/10)). In place allowed. */
int xs_lut_to_db(const double *src_dev, double *dst_dev, int64_t n, void *stream);
int xs_lut_to_linear(const double *src_dev, double *dst_dev, int64_t n, void *stream);

/* ---- inversion ------------------------------------------------------------------------------ */

/* An inversion plan holds what windspeed.py:139-181 sets up per call in the reference (LUTs in dB, their
 * grids, the wspd*cos/sin(phi) tables, phi_180, dsig_co) plus the device-side scan image of the co-pol LUT.
 * The f64 LUT buffers stay owned by the caller and must outlive the plan. */
typedef struct xs_plan xs_plan;

typedef struct xs_plan_desc {
    /* co-pol model (all NULL/0 for cross-pol-only inversion) */
    const double *co_lut_db_dev;   /* [n_inc][n_wspd][n_phi], dB, model-native order (models.py:99) */
    const double *inc_grid_host;   /* [n_inc] */
    const double *wspd_grid_host;  /* [n_wspd] */
    const double *phi_grid_host;   /* [n_phi] degrees */
    const double *cos_phi_host;    /* [n_phi] np.cos(np.radians(phi)) computed by the host */
    const double *sin_phi_host;    /* [n_phi] np.sin(np.radians(phi)) */
    int32_t n_inc, n_wspd, n_phi;
    /* cross-pol model (all NULL/0 for co-pol-only inversion) */
    const double *cr_lut_db_dev;      /* [n_inc_cr][n_wspd_cr], dB */
    const double *inc_cr_grid_host;   /* [n_inc_cr] */
    const double *wspd_cr_grid_host;  /* [n_wspd_cr] */
    int32_t n_inc_cr, n_wspd_cr;
    double dsig_co; /* windspeed.py:24 */
} xs_plan_desc;

/* Synchronous (returns after the device-side preparation finished). */
int xs_plan_create(const xs_plan_desc *desc, void *stream, xs_plan **plan_out);
void xs_plan_destroy(xs_plan *plan);

/* flags of xs_invert_args.flags */
#define XS_FLAG_SIGMA0_DB 1u    /* sigma0 rasters are already 10*log10(s+1e-15) (the B2 boundary, windspeed.py:132);
                                   otherwise the kernel applies windspeed.py:126-128 itself */
#define XS_FLAG_MERGE_DUAL 2u   /* out_cr receives where(|co|<5 or |dual|<5, co, dual) (windspeed.py:426-428) */
#define XS_FLAG_CR_ABS 4u       /* out_cr is float64 |wind| instead of complex128 (windspeed.py:422-423) */
#define XS_FLAG_CR_FULL_SCAN 8u /* verification: scan every cross-pol candidate (FP32 filter + FP64 refinement) even where
                                   the exact interval search applies (LUT row non-decreasing in wspd); same results */
#define XS_FLAG_OUT_SPEED_DIR 16u /* row F2 epilogue: out_co / out_cr are two planes [speed m/s | direction deg], n_px
                                   elements each (float64, or float32 with XS_FLAG_OUT_F32), instead of complex128:
                                   np.abs(wind) and np.angle(wind, deg=True), the post-processing every caller of the
                                   reference applies (docs/examples/windspeed_retrieval_L1.ipynb cell 33).  Halves (f64)
                                   or quarters (f32) the device->host copy.  With XS_FLAG_CR_ABS out_cr stays the single
                                   float64 speed plane. */
#define XS_FLAG_DIR_METEO 32u   /* with OUT_SPEED_DIR: direction = (90 - angle + ground_heading) % 360, i.e.
                                   dir_sample_to_meteo (detrend.py:114-130) wrapped to [0, 360) */
#define XS_FLAG_OUT_F32 64u     /* with OUT_SPEED_DIR: float32 planes */
#define XS_FLAG_NO_PRUNE 128u   /* co-pol scan: evaluate every candidate of the slab (brute force, what the reference does)
                                   instead of skipping the 16-row chunks whose rigorous lower bound of the cost exceeds the
                                   cost of a seed candidate (k_tile_plan); same results, the roofline figure is quoted on it */

/* scan modes */
#define XS_MODE_FAST 0  /* FP32 FFMA2 scan (k_scan_co) + exact refinement of every candidate inside the error band (k_refine_easy) */
#define XS_MODE_FP64 1  /* exhaustive FP64 evaluation of every candidate (verification / fallback) */

typedef struct xs_invert_args {
    /* inputs, device, n_px elements each; dtype XS_F64 or XS_F32 */
    const void *inc;          /* incidence angle, degrees */
    const void *sigma0_co;    /* NULL = no co-pol (all-NaN raster in the reference, windspeed.py:71) */
    const void *sigma0_cr;    /* NULL = no cross-pol */
    const void *dsig_cr;      /* NULL = use dsig_cr_scalar (windspeed.py:122-123) */
    const void *ancillary;    /* complex, antenna convention; NULL = all NaN */
    double dsig_cr_scalar;
    int32_t dtype;
    uint32_t flags;
    int32_t mode;
    int32_t reserved;
    int64_t n_px;
    /* outputs, device */
    void *out_co;             /* complex128[n_px] or NULL */
    void *out_cr;             /* complex128[n_px] (float64[n_px] with XS_FLAG_CR_ABS) or NULL */
    int32_t *idx_co;          /* optional: flat argmin w*n_phi+p, -1 where no co-pol inversion happened */
    int32_t *idx_cr;          /* optional: argmin over the cross-pol wspd grid, -1 where none */
    /* scratch */
    void *workspace;
    size_t workspace_bytes;   /* >= xs_invert_workspace_bytes(plan, n_px) */
    /* per-call bookkeeping (ABI 2; all optional): nothing mutable lives on the plan, so one plan may serve any number
     * of concurrent xs_invert calls (different host threads and streams), each with its own workspace */
    uint64_t *counters_dev;   /* device, XS_N_COUNTERS words: the call's counters, copied on the stream at the end */
    xs_timer *scan_timer;     /* recorded around the co-pol scan (k_scan_co + k_refine_easy) of this call */
    /* F2 epilogue (XS_FLAG_OUT_SPEED_DIR): planes instead of complex128 */
    const void *ground_heading; /* device, n_px of `dtype`, degrees; NULL = directions stay in the antenna convention */
    double ground_heading_scalar; /* used when ground_heading is NULL and XS_FLAG_DIR_METEO is set */
} xs_invert_args;

#define XS_N_COUNTERS 16

/* Workspace an xs_invert call with these flags needs (256-byte aligned device memory, owned by the call). */
size_t xs_invert_workspace_bytes(const xs_plan *plan, int64_t n_px, uint32_t flags);

/* K1: replaces _invert_from_model_numpy / __invert_from_model_1d, windspeed/windspeed.py:132-331
 * (gufunc "(n),(n),(n),(n),(n)->(n),(n)"), with the dB prologue (:126-128) and the dual-pol merge
 * epilogue (:426-428) optionally fused. */
int xs_invert(const xs_plan *plan, const xs_invert_args *args, void *stream);

/* Layout of the counters an xs_invert call leaves in args.counters_dev (read them with any device->host copy after
 * the stream has reached the end of the call): [0] co-pol tiles, [1] pixels sent to the exhaustive FP64 scan,
 * [2] co-pol pixels settled by the FP32 scan (+ refinement), [3] (lane, chunk) cells re-examined by the refinement,
 * [4] pixels of a cross-pol-only call that the step-function kernel (k_cross_only) left to the general cross-pol pass,
 * [5] 16-row chunks the scan CTAs streamed (summed over tiles; tiles x chunks per slab without pruning), [6] (chunk, phi
 * node) pairs the scan warps computed on, summed over the warps of every tile (x 16 rows x 8 pixels = candidates evaluated;
 * tiles x warps per tile x chunks per slab x n_phi without pruning), [7], [9], [10] unused (0),
 * [8] tile hand-out cursor, [11] pixels the refinement settled in FP64 (more than one candidate inside the band),
 * [12] pixels with more than two contending lanes, [13] record positions scanned in shared-sigma0 mode. */

/* ---- detrend -------------------------------------------------------------------------------- */

/* K7: out[l][s] = sigma0[l][s] / (gmf_line[s] / nanmean(gmf_line)), windspeed-independent part of
 * detrend.py:63-64.  gmf_line is the GMF profile of the first image line (xs_gmf_eval). dtype as above. */
int xs_detrend(const void *sigma0_dev, const double *gmf_line_dev, int64_t n_lines, int64_t n_samples, int dtype,
               void *out_dev, void *stream);

/* ---- dsig_cr pre-processors --------------------------------------- */

/* dsig_cr formulas of windspeed/utils.py:66-86, selected by the model name the reference switches on */
enum xs_dsig_id {
    XS_DSIG_GMF_S1_V2 = 0,    /* "gmf_s1_v2": 1/sqrt((s/n)**c(inc)), utils.py:66-76 */
    XS_DSIG_GMF_RS2_V2 = 1,   /* "gmf_rs2_v2": 1/sqrt((s/n)**8), utils.py:78-81 */
    XS_DSIG_CMODMS1AHW = 2,   /* "sarwing_lut_cmodms1ahw" / "nc_lut_cmodms1ahw": (1.25/(s/n))**4, utils.py:83-87 */
    XS_DSIG_COUNT = 3
};
/* co/cross blending weights of windspeed/utils.py:26-42 */
enum xs_dsig_wspd_id {
    XS_DSIG_WSPD_RS2_V3 = 0,
    XS_DSIG_WSPD_S1_EW_REC_V3 = 1,
    XS_DSIG_WSPD_RCM_V3 = 2,
    XS_DSIG_WSPD_COUNT = 3
};

/* Replaces get_dsig(name, inc, sigma0_cr, nesz_cr), windspeed/utils.py:47-91: element-wise over n already-broadcast
 * device elements of type `dtype` (promoted to f64 on load); out is float64.  inc_dev may be NULL unless
 * dsig_id == XS_DSIG_GMF_S1_V2. */
int xs_dsig(int dsig_id, int dtype, const void *inc_dev, const void *sigma0_cr_dev, const void *nesz_cr_dev,
            double *out_dev, int64_t n, void *stream);

/* Replaces get_dsig_wspd(name, U_crosspol, SNR_cr), windspeed/utils.py:18-44.  out is float64 in [0, 1] (NaN kept). */
int xs_dsig_wspd(int dsig_wspd_id, int dtype, const void *u_crosspol_dev, const void *snr_cr_dev, double *out_dev,
                 int64_t n, void *stream);

/* Replaces nesz_flattening(noise, inc), windspeed/utils.py:94-163: noise [n_lines][n_samples] linear NESZ, NaN filled
 * with the column nanmean, converted to dB, fitted per line by an order-1 polynomial of the column-mean incidence
 * (np.polyfit over the finite points), out = 10**((inc_mean*a + b - 1)/10) float64 [n_lines][n_samples]; a line with
 * no finite point is NaN.  workspace >= xs_nesz_flatten_workspace_bytes(). */
size_t xs_nesz_flatten_workspace_bytes(int64_t n_lines, int64_t n_samples);
int xs_nesz_flatten(const void *noise_dev, const void *inc_dev, int64_t n_lines, int64_t n_samples, int dtype,
                    double *out_dev, void *workspace, size_t workspace_bytes, void *stream);

/* ---- local gradients ---------------------------------------------- */

/* Replaces local_gradients(image), gradients.py:588-634 (with R2, :676-722): Scharr gradient (cv2.Scharr, 3x3,
 * BORDER_REFLECT_101) as a complex number, squared, reduced by 2 (5x5 binomial pre-smoothing with scipy's 'symm'
 * boundary, NaN-skipping 2x2 mean with an odd trailing line/sample trimmed, 3x3 binomial post-smoothing).
 * image [n_lines][n_samples] of `dtype`; outputs are [n_lines/2][n_samples/2]:
 *   g2_dev complex128 = sqrt(R2(grad**2)), g3_dev float64 = R2(|grad**2|),
 *   c_dev float64 = |R2(grad**2)| / (g3 + 1e-5) with values > 1 or NaN replaced by 0.
 * workspace >= xs_local_gradients_workspace_bytes() (three half-size float64 planes). */
size_t xs_local_gradients_workspace_bytes(int64_t n_lines, int64_t n_samples);
int xs_local_gradients(const void *image_dev, int64_t n_lines, int64_t n_samples, int dtype, void *g2_dev,
                       double *g3_dev, double *c_dev, void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* XSARSEA_B200_H */
