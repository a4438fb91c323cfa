// Analytic geophysical model functions in FP64 on the device, and the two operators built on them:
//   xs_gmf_eval  -- element-wise evaluation (reference K3, windspeed/gmfs.py:210-214)
//   xs_lut_build -- outer-product LUT generation (reference K2, windspeed/gmfs.py:218-230)
// Formulas: windspeed/gmfs_impl.py (CMOD5 family :117-201, CMOD-IFR2 :214-303, VH power-law/sigmoid
// families :326-707).  All math is FP64 CUDA-core; device libm differs from the host's by <= 2 ulp, so
// parity with the oracle is by tolerance (tests/test_gpu_lut.py), not bits.
#include "xs_common.cuh"

namespace xs {

// ---- coefficient tables (constant memory; indices follow the published CMOD5 numbering, 0 unused) ----
__constant__ double kCmod5[2][29] = {
    {0.0, -0.688, -0.793, 0.338, -0.173, 0.0, 0.004, 0.111, 0.0162, 6.34, 2.57, -2.18, 0.4, -0.6, 0.045, 0.007,
     0.33, 0.012, 22.0, 1.95, 3.0, 8.39, -3.44, 1.36, 5.35, 1.99, 0.29, 3.80, 1.53},
    {0.0, -0.6878, -0.7957, 0.338, -0.1728, 0.0, 0.004, 0.1103, 0.0159, 6.7329, 2.7713, -2.2885, 0.4971, -0.725,
     0.045, 0.0066, 0.3222, 0.012, 22.7, 2.0813, 3.0, 8.3659, -3.3428, 1.3236, 6.2437, 2.3893, 0.3249, 4.159,
     1.693}};

__constant__ double kIfr2[26] = {0.0,       -2.437597, -1.5670307, 0.3708242, -0.040590, 0.404678,  0.188397,
                                 -0.027262, 0.064650,  0.054500,   0.086350,  0.055100,  -0.058450, -0.096100,
                                 0.412754,  0.121785,  -0.024333,  0.072163,  -0.062954, 0.015958,  -0.069514,
                                 -0.062945, 0.035538,  0.023049,   0.074654,  -0.014713};

// VH models: {a0_Z1, b0_Z1, b1_Z1, a0_Z2, a1_Z2, a2_Z2, b0_Z2, b1_Z2, b2_Z2, c0, c1, c2, c3}
struct XpolCoef {
    double p[13];
    int blend_db;     // 0: sigmoid blend of linear sigma0 (v2 family); 1: blend of dB values (v3/v4)
    double b0z2_mul;  // 1.01 for the v4 variants
};
__constant__ XpolCoef kXpol[8] = {
    // rs2_v2
    {{6.55519203e-06, 2.49753154e00, -1.35734881e-02, 1.47342197e-04, -4.07334797e-06, 3.43593382e-08,
      1.10188639e00, 1.40782758e-02, -1.53748743e-04, -0.18675905, 24.48859492, 0.19185442, 25.38275738},
     0, 1.0},
    // s1_v2
    {{2.13755392e-06, 2.47395267e00, -2.85775085e-03, 6.54058552e-05, -2.43845137e-06, 2.87698338e-08,
      1.14509104e00, 3.41828829e-02, -4.79715441e-04, -0.23257086, 12.39717002, 0.21667263, 12.22862991},
     0, 1.0},
    // rcm_noaa
    {{2.2309436836414871e-12, 8.3374911282878728, -0.033443488982800210, 7.7945050373193260e-05,
      -2.4425748662769216e-06, 2.7625550632547159e-08, 1.2524896108831316, 0.019203092214131894,
      -0.00028408046502692580, -0.34498737004629487, 12.558975188752012, 0.12713502524515713,
      4.2806865431046752},
     0, 1.0},
    // s1_v3_ew_rec
    {{3.5033427638479895e-06, 2.5486758595982275, -0.009042529888607539, 4.142689709809047e-05,
      -1.6620917447744406e-06, 2.4331104610101826e-08, 1.277314996198736, 0.03813903872809897,
      -0.0006506765114704733, -0.2522916645939956, 15.3393676653533, 0.24259895576004784, 15.203063214062643},
     1, 1.0},
    // rs2_v3
    {{8.423384272498706e-06, 2.4351127340627374, -0.01450322326682606, 0.00014955206131320428,
      -4.737691852310481e-06, 3.813107432709729e-08, 1.524883207000445, -0.01322253424944054,
      0.00037527120092119504, -0.2222881984904166, 13.118282628673661, 0.21426139278646567, 12.768845054319682},
     1, 1.0},
    // rcm_v3
    {{7.093964676135241e-06, 2.3722948391886542, -0.009516840375089524, 6.689451099284358e-05,
      -1.3956325894252652e-06, 9.227949977841212e-09, 1.4687699534267797, 0.005735224541037088,
      -7.164130353316848e-05, -0.2454472887447197, 15.537961353644508, 0.24011368010838255, 15.332883245452303},
     1, 1.0},
    // rcm_v4 = rcm_v3 with b0_Z2 * 1.01
    {{7.093964676135241e-06, 2.3722948391886542, -0.009516840375089524, 6.689451099284358e-05,
      -1.3956325894252652e-06, 9.227949977841212e-09, 1.4687699534267797, 0.005735224541037088,
      -7.164130353316848e-05, -0.2454472887447197, 15.537961353644508, 0.24011368010838255, 15.332883245452303},
     1, 1.01},
    // rs2_v4 = rs2_v3 with b0_Z2 * 1.01
    {{8.423384272498706e-06, 2.4351127340627374, -0.01450322326682606, 0.00014955206131320428,
      -4.737691852310481e-06, 3.813107432709729e-08, 1.524883207000445, -0.01322253424944054,
      0.00037527120092119504, -0.2222881984904166, 13.118282628673661, 0.21426139278646567, 12.768845054319682},
     1, 1.01},
};

// CMOD5 / CMOD5.N (+ HH polarisation ratios), gmfs_impl.py:117-201
template <int SET, int PR>
__device__ double gmf_cmod5(double inc, double wspd, double phi) {
    const double *c = kCmod5[SET];
    const double y0 = c[19], pn = c[20];
    const double a = y0 - (y0 - 1.0) / pn;
    const double b = 1.0 / (pn * pow(y0 - 1.0, pn - 1.0));
    const double cphi = cos(deg2rad(phi));
    const double x = (inc - 40.0) / 25.0;
    const double x2 = x * x;

    const double a0 = c[1] + c[2] * x + c[3] * x2 + c[4] * x * x2;
    const double a1 = c[5] + c[6] * x;
    const double a2 = c[7] + c[8] * x;
    const double gam = c[9] + c[10] * x + c[11] * x2;
    const double s0 = c[12] + c[13] * x;
    const double s = a2 * wspd;
    double a3 = 1.0 / (1.0 + exp(-s0));
    if (s < s0)
        a3 = a3 * pow(s / s0, s0 * (1.0 - a3));
    else
        a3 = 1.0 / (1.0 + exp(-s));
    const double b0 = pow(a3, gam) * exp10(a0 + a1 * wspd);

    double b1 = c[15] * wspd * (0.5 + x - tanh(4.0 * (x + c[16] + c[17] * wspd)));
    b1 = (c[14] * (1.0 + x) - b1) / (exp(0.34 * (wspd - c[18])) + 1.0);

    const double v0 = c[21] + c[22] * x + c[23] * x2;
    const double d1 = c[24] + c[25] * x + c[26] * x2;
    const double d2 = c[27] + c[28] * x;
    double v2 = wspd / v0 + 1.0;
    if (v2 < y0) v2 = a + b * pow(v2 - 1.0, pn);
    const double b2 = (-d1 + d2 * v2) * exp(-v2);

    double sig = b0 * pow(1.0 + b1 * cphi + b2 * (2.0 * cphi * cphi - 1.0), 1.6);
    if (PR == 1) {  // Zhang A ratio, f(inc, wspd), :165-172
        const double ars = 1.3794 + (-3.19e-2 + 1.4e-3 * inc) * inc;
        const double brs = -0.1711 + 2.6e-3 * inc;
        sig = sig / (ars * pow(wspd, brs));
    } else if (PR == 2) {  // Mouche et al. ratio, f(inc, phi), :174-199
        const double p0 = 0.00650704 * exp(0.128983 * inc) + 0.992839;
        const double ph = 0.00782194 * exp(0.121405 * inc) + 0.992839;
        const double pp = 0.00598416 * exp(0.140952 * inc) + 0.992885;
        const double k0 = (p0 + pp + 2 * ph) / 4, k1 = (p0 - pp) / 2, k2 = (p0 + pp - 2 * ph) / 4;
        sig = sig / (k0 + k1 * cos(deg2rad(phi)) + k2 * cos(2 * deg2rad(phi)));
    }
    return sig;
}

// CMOD-IFR2, gmfs_impl.py:214-303
__device__ double gmf_ifr2(double inc, double wspd, double phi) {
    const double *C = kIfr2;
    const double ti = (inc - 36.0) / 19.0, tq = ti * ti;
    const double P2 = (3.0 * tq - 1.0) / 2.0, P3 = (5.0 * tq - 3.0) * ti / 2.0;
    const double alph = C[1] + C[2] * ti + C[3] * P2 + C[4] * P3;
    const double beta = C[5] + C[6] * ti + C[7] * P2;
    const double ci = cos(deg2rad(phi)), c2i = 2.0 * ci * ci - 1.0;
    const double tn = (2.0 * inc - 76.0) / 40.0;   // (2T - (18+58)) / (58-18)
    const double vn = (2.0 * wspd - 28.0) / 22.0;  // (2v - (25+3)) / (25-3)
    const double pv1 = vn, pv2 = 2 * vn * pv1 - 1.0, pv3 = 2 * vn * pv2 - pv1;
    const double pt1 = tn, pt2 = 2 * tn * pt1 - 1.0;
    const double b1 = C[8] + C[9] * pv1 + (C[10] + C[11] * pv1) * pt1 + (C[12] + C[13] * pv1) * pt2;
    const double b2 = C[14] + C[15] * pt1 + C[16] * pt2 + (C[17] + C[18] * pt1 + C[19] * pt2) * pv1 +
                      (C[20] + C[21] * pt1 + C[22] * pt2) * pv2 + (C[23] + C[24] * pt1 + C[25] * pt2) * pv3;
    const double b0 = exp10(alph + beta * sqrt(wspd));
    return b0 * (1.0 + b1 * ci + tanh(b2) * c2i);
}

// VH families, gmfs_impl.py:326-707
__device__ double gmf_xpol(int k, double inc, double u) {
    const XpolCoef &m = kXpol[k];
    const double *p = m.p;
    const double z1 = p[0] * pow(u, p[1] + p[2] * inc);
    const double a2 = p[3] + p[4] * inc + p[5] * (inc * inc);
    const double z2 = a2 * pow(u, p[6] * m.b0z2_mul + p[7] * inc + p[8] * (inc * inc));
    const double g1 = 1 / (1 + exp(-p[9] * (u - p[10])));
    const double g2 = 1 / (1 + exp(-p[11] * (u - p[12])));
    if (!m.blend_db) return z1 * g1 + z2 * g2;
    return exp10((10 * log10(z1) * g1 + 10 * log10(z2) * g2) / 10);
}

__device__ double gmf_dispatch(int model, double inc, double wspd, double phi) {
    switch (model) {
        case XS_GMF_CMOD5: return gmf_cmod5<0, 0>(inc, wspd, phi);
        case XS_GMF_CMOD5N: return gmf_cmod5<1, 0>(inc, wspd, phi);
        case XS_GMF_CMOD5N_PR_ZHANGA: return gmf_cmod5<1, 1>(inc, wspd, phi);
        case XS_GMF_CMOD5N_PR_MOUCHE1: return gmf_cmod5<1, 2>(inc, wspd, phi);
        case XS_GMF_CMODIFR2: return gmf_ifr2(inc, wspd, phi);
        default: return gmf_xpol(model - XS_GMF_RS2_V2, inc, wspd);
    }
}

// ---- K3: element-wise -------------------------------------------------------------------------------
template <typename T>
__global__ void k_gmf_eval(int model, const T *__restrict__ inc, const T *__restrict__ wspd,
                           const double *__restrict__ phi, T *__restrict__ out, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double ph = phi ? phi[i] : 0.0;
        out[i] = (T)gmf_dispatch(model, (double)inc[i], (double)wspd[i], ph);
    }
}

// ---- K2: outer product; one thread per LUT node, phi fastest so stores coalesce ------------------------
__global__ void k_lut_build(int model, const double *__restrict__ inc, int n_inc, const double *__restrict__ wspd,
                            int n_wspd, const double *__restrict__ phi, int n_phi, double *__restrict__ out) {
    const int64_t n = (int64_t)n_inc * n_wspd * n_phi;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int ip = (int)(i % n_phi);
        const int64_t r = i / n_phi;
        const int iw = (int)(r % n_wspd), ii = (int)(r / n_wspd);
        out[i] = gmf_dispatch(model, inc[ii], wspd[iw], phi ? phi[ip] : 0.0);
    }
}

static int grid_for(int64_t n, int block) {
    int64_t g = ceil_div(n, block);
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace xs

extern "C" int xs_gmf_eval(int model_id, int dtype, const void *inc, const void *wspd, const double *phi, void *out,
                           int64_t n, void *stream) {
    using namespace xs;
    if (model_id < 0 || model_id >= XS_GMF_COUNT || n < 0 || (n > 0 && (!inc || !wspd || !out))) {
        set_error("xs_gmf_eval: invalid argument");
        return XS_E_INVALID;
    }
    if (model_id <= XS_GMF_CMODIFR2 && !phi) {
        set_error("xs_gmf_eval: co-pol GMF needs phi");
        return XS_E_INVALID;
    }
    if (n == 0) return XS_OK;
    const int block = 256, grid = grid_for(n, block);
    if (dtype == XS_F64)
        XS_LAUNCH(k_gmf_eval<double>, grid, block, 0, stream, model_id, (const double *)inc, (const double *)wspd, phi,
                  (double *)out, n);
    else if (dtype == XS_F32)
        XS_LAUNCH(k_gmf_eval<float>, grid, block, 0, stream, model_id, (const float *)inc, (const float *)wspd, phi,
                  (float *)out, n);
    else {
        set_error("xs_gmf_eval: bad dtype %d", dtype);
        return XS_E_INVALID;
    }
    return XS_OK;
}

extern "C" int xs_lut_build(int model_id, const double *inc_h, int n_inc, const double *wspd_h, int n_wspd,
                            const double *phi_h, int n_phi, double *out, void *stream) {
    using namespace xs;
    if (model_id < 0 || model_id >= XS_GMF_COUNT || !inc_h || !wspd_h || !out || n_inc <= 0 || n_wspd <= 0) {
        set_error("xs_lut_build: invalid argument");
        return XS_E_INVALID;
    }
    const bool copol = model_id <= XS_GMF_CMODIFR2;
    if (copol && (!phi_h || n_phi <= 0)) {
        set_error("xs_lut_build: co-pol GMF needs a phi grid");
        return XS_E_INVALID;
    }
    const int np = (phi_h && n_phi > 0) ? n_phi : 1;
    cudaStream_t st = (cudaStream_t)stream;
    double *grids = nullptr;
    const size_t ng = (size_t)n_inc + n_wspd + np;
    keep_async_pool();
    XS_CUDA(cudaMallocAsync(&grids, ng * sizeof(double), st));
    XS_CUDA(cudaMemcpyAsync(grids, inc_h, n_inc * sizeof(double), cudaMemcpyHostToDevice, st));
    XS_CUDA(cudaMemcpyAsync(grids + n_inc, wspd_h, n_wspd * sizeof(double), cudaMemcpyHostToDevice, st));
    if (phi_h && n_phi > 0)
        XS_CUDA(cudaMemcpyAsync(grids + n_inc + n_wspd, phi_h, n_phi * sizeof(double), cudaMemcpyHostToDevice, st));
    const int64_t n = (int64_t)n_inc * n_wspd * np;
    const int block = 256, grid = grid_for(n, block);
    XS_LAUNCH(k_lut_build, grid, block, 0, stream, model_id, grids, n_inc, grids + n_inc, n_wspd,
              (phi_h && n_phi > 0) ? grids + n_inc + n_wspd : nullptr, np, out);
    XS_CUDA(cudaFreeAsync(grids, st));
    // host grids may be freed by the caller right after return: the H2D copies above are from pageable
    // memory and therefore already staged when cudaMemcpyAsync returns.
    return XS_OK;
}
