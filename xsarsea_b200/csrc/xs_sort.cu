// Pre-pass of the co-pol scan: order the listed pixels by (incidence bin, sigma0) with CUB's radix sort, so that the 8 pixels
// a warp of k_scan_co scans together have (nearly) the same sigma0 and can share the sigma0-dependent part of the cost
// (xs_scan.cu).  Library code for a plain sort (like cuBLAS for a plain GEMM): 32-bit keys and 32-bit pixel indices, 4 digit
// passes, about 0.3 % of a step; everything that is specific to this problem lives in the hand-written kernels around it.
#include <cub/device/device_radix_sort.cuh>

#include "xs_invert.cuh"

namespace xs {

size_t sort_temp_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DoubleBuffer<unsigned> k(nullptr, nullptr), v(nullptr, nullptr);
    if (n <= 0) return 0;
    if (cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, n, 0, 32, (cudaStream_t)0) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return bytes + 256;
}

// sorts (keys[0], vals[0]) using (keys[1], vals[1]) as the alternate buffers; *which = index of the buffers holding the result
int sort_pairs_u32(unsigned *keys[2], unsigned *vals[2], int64_t n, void *temp, size_t temp_bytes, cudaStream_t st, int *which) {
    cub::DoubleBuffer<unsigned> k(keys[0], keys[1]), v(vals[0], vals[1]);
    size_t need = 0;
    XS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, k, v, n, 0, 32, st));
    if (need > temp_bytes) {
        set_error("xs_invert: sort workspace too small (%zu < %zu)", temp_bytes, need);
        return XS_E_WORKSPACE;
    }
    XS_CUDA(cub::DeviceRadixSort::SortPairs(temp, need, k, v, n, 0, 32, st));
    g_launches += 5;  // histogram + 4 digit passes (CUB's own kernels)
    *which = k.selector;
    return XS_OK;
}

}  // namespace xs
