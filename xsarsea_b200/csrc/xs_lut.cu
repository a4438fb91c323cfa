// LUT post-processing on the device: per-axis linear interpolation (reference K5, windspeed/models.py:154-168
// through xarray.DataArray.interp -> scipy interp1d), unit conversion (K6, models.py:215 / :221) and the
// state shared by the library (error text, launch counter).
#include <stdarg.h>

#include <atomic>
#include <vector>

#include "xs_common.cuh"

namespace xs {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

// The stream-ordered allocations of this library (small scratch buffers) stay in the device's default pool instead of going
// back to the OS at every synchronisation: once per device.
void keep_async_pool() {
    static std::atomic<unsigned long long> done{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return;
    if (done.load() & (1ull << dev)) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done.fetch_or(1ull << dev);
}

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// One thread per output element of dst[outer][n_dst][inner]; `inner` is the fastest index so both the two
// source reads and the store are coalesced.  Bracket indices and the two weights are precomputed on the host
// in FP64 with the expressions of the installed scipy (1.18) interp1d._call_linear:
//   w_hi = (x_new - x_lo)/(x_hi - x_lo), w_lo = (x_hi - x_new)/(x_hi - x_lo), y = w_hi*y_hi + w_lo*y_lo
// __dmul_rn/__dadd_rn keep the two products and the sum separately rounded (no FMA contraction), which is
// what numpy does.
__global__ void k_interp_axis(const double *__restrict__ src, int64_t outer, int n_src, int64_t inner,
                              const int *__restrict__ hi_idx, const double *__restrict__ w_hi,
                              const double *__restrict__ w_lo, int n_dst, double *__restrict__ dst) {
    const int64_t n = outer * n_dst * inner;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t q = i % inner;
        const int64_t r = i / inner;
        const int t = (int)(r % n_dst);
        const int64_t o = r / n_dst;
        const int h = hi_idx[t];
        const double yl = src[(o * n_src + (h - 1)) * inner + q];
        const double yh = src[(o * n_src + h) * inner + q];
        dst[i] = __dadd_rn(__dmul_rn(w_hi[t], yh), __dmul_rn(w_lo[t], yl));
    }
}

__global__ void k_to_db(const double *__restrict__ src, double *__restrict__ dst, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = 10.0 * log10(src[i] + 1e-15);
}

__global__ void k_to_linear(const double *__restrict__ src, double *__restrict__ dst, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = exp10(src[i] / 10.0);
}

static int grid_for(int64_t n, int block) {
    int64_t g = ceil_div(n, block);
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace xs

extern "C" int xs_abi_version(void) { return XS_ABI_VERSION; }
extern "C" const char *xs_last_error(void) { return xs::g_err; }
extern "C" int64_t xs_launch_count(void) { return xs::g_launches.load(); }

extern "C" int xs_lut_interp_axis(const double *src, int64_t outer, int n_src, int64_t inner, const double *x_src,
                                  const double *x_dst, int n_dst, double *dst, void *stream) {
    using namespace xs;
    if (!src || !dst || !x_src || !x_dst || outer <= 0 || inner <= 0 || n_src < 2 || n_dst <= 0) {
        set_error("xs_lut_interp_axis: invalid argument");
        return XS_E_INVALID;
    }
    std::vector<int> hi(n_dst);
    std::vector<double> wh(n_dst), wl(n_dst);
    for (int t = 0; t < n_dst; ++t) {
        const double xn = x_dst[t];
        if (!(xn >= x_src[0]) || !(xn <= x_src[n_src - 1])) {
            set_error("A value in x_new is outside the interpolation range.");
            return XS_E_BOUNDS;
        }
        int lo = 0, up = n_src;  // searchsorted side='left'
        while (lo < up) {
            const int mid = (lo + up) / 2;
            if (x_src[mid] < xn)
                lo = mid + 1;
            else
                up = mid;
        }
        int h = lo < 1 ? 1 : (lo > n_src - 1 ? n_src - 1 : lo);
        hi[t] = h;
        const double x_lo = x_src[h - 1], x_hi = x_src[h];
        wh[t] = (xn - x_lo) / (x_hi - x_lo);
        wl[t] = (x_hi - xn) / (x_hi - x_lo);
    }
    cudaStream_t st = (cudaStream_t)stream;
    char *buf = nullptr;
    const size_t nb_i = sizeof(int) * n_dst, nb_d = sizeof(double) * n_dst;
    const size_t off_w = (nb_i + 15) / 16 * 16;
    keep_async_pool();
    XS_CUDA(cudaMallocAsync(&buf, off_w + 2 * nb_d, st));
    XS_CUDA(cudaMemcpyAsync(buf, hi.data(), nb_i, cudaMemcpyHostToDevice, st));
    XS_CUDA(cudaMemcpyAsync(buf + off_w, wh.data(), nb_d, cudaMemcpyHostToDevice, st));
    XS_CUDA(cudaMemcpyAsync(buf + off_w + nb_d, wl.data(), nb_d, cudaMemcpyHostToDevice, st));
    const int64_t n = outer * n_dst * inner;
    XS_LAUNCH(k_interp_axis, grid_for(n, 256), 256, 0, stream, src, outer, n_src, inner, (const int *)buf,
              (const double *)(buf + off_w), (const double *)(buf + off_w + nb_d), n_dst, dst);
    XS_CUDA(cudaFreeAsync(buf, st));
    return XS_OK;
}

extern "C" int xs_lut_to_db(const double *src, double *dst, int64_t n, void *stream) {
    using namespace xs;
    if (!src || !dst || n < 0) {
        set_error("xs_lut_to_db: invalid argument");
        return XS_E_INVALID;
    }
    if (n) XS_LAUNCH(k_to_db, grid_for(n, 256), 256, 0, stream, src, dst, n);
    return XS_OK;
}

extern "C" int xs_lut_to_linear(const double *src, double *dst, int64_t n, void *stream) {
    using namespace xs;
    if (!src || !dst || n < 0) {
        set_error("xs_lut_to_linear: invalid argument");
        return XS_E_INVALID;
    }
    if (n) XS_LAUNCH(k_to_linear, grid_for(n, 256), 256, 0, stream, src, dst, n);
    return XS_OK;
}
