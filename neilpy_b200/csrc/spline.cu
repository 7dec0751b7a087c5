// Interpolating bicubic (not-a-knot) spline of the provisional DTM / slope raster and the
// fused point gather + classification (neilpy.py:1768-1795).
//
// Prefilter: the collocation system is banded and depends only on the axis length; the
// host factors it once (neilpy_b200/spline.py: notaknot_factors) and the device does the
// forward / backward substitutions.  Influence along a line decays by ~0.268 per cell, so
// every line is cut into chunks of kChunk cells that start kWarm cells early from a zero
// state (0.268^40 ~ 1e-23: exact to rounding); chunks that reach the line start are exact
// by construction.  Threads run along the fast axis (coalesced), so the row-direction
// solve is done on a transposed copy.
//
// Classify: one thread per point, 16-tap gathers from the two coefficient grids (L2).
#include <stdlib.h>

#include "common.cuh"

namespace smrf {
namespace spline {

constexpr int kChunk = 256;
constexpr int kWarm = 40;

struct Factors {   // device arrays of length n
    const double *l1, *l2, *dinv, *u1, *u2;
};

static Factors split(const double* f, int64_t n) {
    return Factors{f, f + n, f + 2 * n, f + 3 * n, f + 4 * n};
}

// forward substitution along axis 0 (length n0) of a [n0][n1] array: y = L^-1 f
template <typename TI, typename TO>
__global__ void __launch_bounds__(128) forward_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t n0,
                                                      int64_t n1, Factors fa) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n1) return;
    const int64_t s = (int64_t)blockIdx.y * kChunk;
    const int64_t e = s + kChunk < n0 ? s + kChunk : n0;
    int64_t i = s - kWarm;
    if (i < 0) i = 0;
    double y1 = 0.0, y2 = 0.0;   // y[i-1], y[i-2]
    for (; i < e; ++i) {
        const double f = (double)in[i * n1 + x];
        const double y = f - fa.l1[i] * y1 - fa.l2[i] * y2;
        y2 = y1; y1 = y;
        if (i >= s) out[i * n1 + x] = (TO)y;
    }
}

// backward substitution: c = U^-1 y
template <typename TI, typename TO>
__global__ void __launch_bounds__(128) backward_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t n0,
                                                       int64_t n1, Factors fa) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n1) return;
    const int64_t s = (int64_t)blockIdx.y * kChunk;
    const int64_t e = s + kChunk < n0 ? s + kChunk : n0;
    int64_t i = e - 1 + kWarm;
    if (i > n0 - 1) i = n0 - 1;
    double c1 = 0.0, c2 = 0.0;   // c[i+1], c[i+2]
    for (; i >= s; --i) {
        const double y = (double)in[i * n1 + x];
        const double c = (y - fa.u1[i] * c1 - fa.u2[i] * c2) * fa.dinv[i];
        c2 = c1; c1 = c;
        if (i < e) out[i * n1 + x] = (TO)c;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t n0,
                                                        int64_t n1) {
    __shared__ T tile[32][33];
    const int64_t bx = (int64_t)blockIdx.x * 32, by = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int64_t y = by + ty + k, x = bx + tx;
        if (y < n0 && x < n1) tile[ty + k][tx] = in[y * n1 + x];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int64_t x = bx + ty + k, y = by + tx;
        if (y < n0 && x < n1) out[x * n0 + y] = tile[tx][ty + k];
    }
}

template <typename T>
static int prefilter_t(const T* grid, T* coef, double* s1, double* s2, int64_t ny, int64_t nx, const double* rowf,
                       const double* colf, cudaStream_t st) {
    // intermediates are kept in float64 whatever the grid type
    dim3 b(128);
    {   // along y (axis 0 of [ny][nx]) with the row-axis factors
        Factors fa = split(rowf, ny);
        dim3 g((unsigned)((nx + 127) / 128), (unsigned)((ny + kChunk - 1) / kChunk));
        forward_kernel<T, double><<<g, b, 0, st>>>(grid, s1, ny, nx, fa);
        backward_kernel<double, double><<<g, b, 0, st>>>(s1, s2, ny, nx, fa);
    }
    {
        dim3 g((unsigned)((nx + 31) / 32), (unsigned)((ny + 31) / 32));
        transpose_kernel<double><<<g, 256, 0, st>>>(s2, s1, ny, nx);   // s1 = [nx][ny]
    }
    {   // along x, now axis 0 of [nx][ny], with the column-axis factors
        Factors fa = split(colf, nx);
        dim3 g((unsigned)((ny + 127) / 128), (unsigned)((nx + kChunk - 1) / kChunk));
        forward_kernel<double, double><<<g, b, 0, st>>>(s1, s2, nx, ny, fa);
        backward_kernel<double, double><<<g, b, 0, st>>>(s2, s1, nx, ny, fa);
    }
    {
        // transpose back, narrowing to the grid type on the way
        dim3 g((unsigned)((ny + 31) / 32), (unsigned)((nx + 31) / 32));
        transpose_kernel<double><<<g, 256, 0, st>>>(s1, s2, nx, ny);   // s2 = [ny][nx]
    }
    return 0;
}

// float64 intermediates -> the coefficient array, optionally interleaved with another spline's
// coefficients (out[i * stride + offset]): the gather kernel then fetches both splines' taps from
// the same sectors
template <typename T>
__global__ void __launch_bounds__(256) narrow_kernel(const double* __restrict__ in, T* __restrict__ out, int64_t n,
                                                     int64_t stride, int64_t offset) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i * stride + offset] = (T)in[i];
}

// ---- evaluation ---------------------------------------------------------------------------
__device__ __forceinline__ double knot(int j, int n) {
    return j <= 3 ? 0.5 : (j >= n ? (double)n - 0.5 : (double)j - 1.5);
}

// FITPACK fpbspl for k = 3 on the arithmetic knot vector; returns the interval index l
__device__ __forceinline__ int bspline4(double x, int n, double (&h)[4]) {
    const double hi = (double)n - 0.5;
    x = x < 0.5 ? 0.5 : (x > hi ? hi : x);   // bispeu clamps the argument to [t[3], t[n]]
    int l = 3;
    if (x >= 2.5) {
        l = (int)floor(x + 1.5);
        if (l > n - 1) l = n - 1;
    }
    // Interior intervals (knots l-2 .. l+3 all of the arithmetic kind, spacing 1): the recurrence below collapses to
    // the uniform cubic B-spline polynomials.  Same values to rounding (~1e-16 relative; the spline is held to 1e-9 m),
    // without the recurrence's six divisions per axis -- the kernel is bound by its float64 arithmetic.
    if (l >= 6 && l <= n - 4) {
        const double u = x - ((double)l - 1.5), v = 1.0 - u;
        const double u2 = u * u, u3 = u2 * u;
        h[0] = v * v * v * (1.0 / 6.0);
        h[1] = (3.0 * u3 - 6.0 * u2 + 4.0) * (1.0 / 6.0);
        h[2] = (-3.0 * u3 + 3.0 * u2 + 3.0 * u + 1.0) * (1.0 / 6.0);
        h[3] = u3 * (1.0 / 6.0);
        return l;
    }
    h[0] = 1.0; h[1] = 0.0; h[2] = 0.0; h[3] = 0.0;
#pragma unroll
    for (int j = 1; j <= 3; ++j) {
        double hh[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) hh[i] = h[i];
        h[0] = 0.0;
#pragma unroll
        for (int i = 1; i <= 3; ++i) {
            if (i <= j) {
                const double tli = knot(l + i, n), tlj = knot(l + i - j, n);
                const double f = __ddiv_rn(hh[i - 1], __dsub_rn(tli, tlj));
                h[i - 1] = __dadd_rn(h[i - 1], __dmul_rn(f, __dsub_rn(tli, x)));
                h[i] = __dmul_rn(f, __dsub_rn(x, tlj));
            }
        }
    }
    return l;
}

template <typename T, int FMT>
__global__ void __launch_bounds__(256) classify_kernel(PointLoader<FMT> pts, int64_t n, Inv6 inv,
                                                       const T* __restrict__ cz, const T* __restrict__ cs, int ny,
                                                       int nx, double et, double es, uint8_t* __restrict__ is_object,
                                                       double* __restrict__ elev_out, double* __restrict__ slope_out,
                                                       const uint8_t* __restrict__ drop_raster,
                                                       uint8_t* __restrict__ when_pt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double x, y, z;
        pts.xyz(i, x, y, z);
        double c, r;
        affine_apply(inv, x, y, c, r);
        double hr[4], hc[4];
        const int lr = bspline4(r, ny, hr);
        const int lc = bspline4(c, nx, hc);
        // cs == nullptr: cz holds (z, slope) coefficient pairs, [ny][nx][2]
        const int64_t st = cs ? 1 : 2;
        const T* pz = cz + ((int64_t)(lr - 3) * nx + (lc - 3)) * st;
        const T* ps = cs ? cs + (int64_t)(lr - 3) * nx + (lc - 3) : pz + 1;
        double ez = 0.0, sl = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                // FITPACK fpbisp: sp = sp + c*h(i1)*w(j1), left to right
                const int64_t o = (a * (int64_t)nx + b) * st;
                ez = __dadd_rn(ez, __dmul_rn(__dmul_rn((double)__ldg(pz + o), hr[a]), hc[b]));
                sl = __dadd_rn(sl, __dmul_rn(__dmul_rn((double)__ldg(ps + o), hr[a]), hc[b]));
            }
        }
        const double required = __dadd_rn(et, __dmul_rn(es, sl));
        is_object[i] = fabs(__dsub_rn(ez, z)) > required ? 1 : 0;
        if (elev_out) elev_out[i] = ez;
        if (slope_out) slope_out[i] = sl;
        if (when_pt) {
            // drop_raster[np.round(r), np.round(c)] (round half to even), neilpy.py:1780
            int64_t rr = (int64_t)rint(r), cc = (int64_t)rint(c);
            rr = rr < 0 ? rr + ny : rr;   // numpy negative index wrap; anything else would have raised
            cc = cc < 0 ? cc + nx : cc;
            when_pt[i] = (rr >= 0 && rr < ny && cc >= 0 && cc < nx) ? drop_raster[rr * nx + cc] : 0;
        }
    }
}

}  // namespace spline
}  // namespace smrf

using namespace smrf;
using namespace smrf::spline;

template <typename T, int FMT>
static void classify_launch(const void* x, const void* y, const void* z, int64_t n, Inv6 inv, const void* cz,
                            const void* cs, int ny, int nx, double et, double es, uint8_t* is_object, double* elev,
                            double* slope, const uint8_t* drop, uint8_t* when_pt, cudaStream_t st) {
    int g = (int)((n + 255) / 256);
    int cap = num_sms() * 16;
    if (g > cap) g = cap;
    PointLoader<FMT> pl;
    if constexpr (FMT == SMRF_PTS_SOA_F64) pl = PointLoader<FMT>{(const double*)x, (const double*)y, (const double*)z};
    else if constexpr (FMT == SMRF_PTS_XYZW_F32) pl = PointLoader<FMT>{(const float4*)x, nullptr, nullptr};
    else pl = PointLoader<FMT>{(const float*)x, (const float*)y, (const float*)z};
    classify_kernel<T, FMT><<<g, 256, 0, st>>>(pl, n, inv, (const T*)cz, (const T*)cs, ny, nx, et, es, is_object, elev,
                                               slope, drop, when_pt);
}

static int classify_any(const void* x, const void* y, const void* z, int64_t n, int point_fmt, const double* inv6_host,
                        const void* coef_z, const void* coef_s, int64_t ny, int64_t nx, int dtype,
                        double elevation_threshold, double elevation_scaler, uint8_t* is_object, double* elevation,
                        double* slope_out, const uint8_t* drop_raster, uint8_t* when_dropped_pt, void* stream);

extern "C" {

size_t smrf_spline_workspace_bytes(int64_t ny, int64_t nx) {
    size_t plane = ((size_t)ny * (size_t)nx * 8 + 255) & ~(size_t)255;
    return 2 * plane;
}

int smrf_spline_prefilter(const void* grid, void* coef, int64_t coef_stride, int64_t coef_offset, int64_t ny,
                          int64_t nx, int dtype, const double* row_factors, const double* col_factors,
                          void* workspace, size_t workspace_bytes, void* stream) {
    SMRF_CHECK_ARG(grid && coef && row_factors && col_factors && workspace, "null pointer");
    SMRF_CHECK_ARG(ny >= 4 && nx >= 4, "the interpolating cubic spline needs at least 4 grid rows and columns");
    SMRF_CHECK_ARG(dtype == SMRF_F32 || dtype == SMRF_F64, "bad dtype");
    SMRF_CHECK_ARG(coef_stride >= 1 && coef_offset >= 0 && coef_offset < coef_stride, "bad coefficient stride / offset");
    SMRF_CHECK_ARG(coef_stride == 1 || coef != grid, "an interleaved coefficient array cannot alias the grid");
    if (workspace_bytes < smrf_spline_workspace_bytes(ny, nx)) {
        set_error("smrf_spline_prefilter: workspace %zu < %zu bytes", workspace_bytes, smrf_spline_workspace_bytes(ny, nx));
        return SMRF_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    size_t plane = smrf_spline_workspace_bytes(ny, nx) / 2;
    double* s1 = (double*)workspace;
    double* s2 = (double*)((char*)workspace + plane);
    const int64_t n = ny * nx;
    int g = (int)((n + 255) / 256);
    int cap = num_sms() * 16;
    if (g > cap) g = cap;
    if (dtype == SMRF_F32) {
        prefilter_t<float>((const float*)grid, (float*)coef, s1, s2, ny, nx, row_factors, col_factors, st);
        narrow_kernel<float><<<g, 256, 0, st>>>(s2, (float*)coef, n, coef_stride, coef_offset);
    } else {
        prefilter_t<double>((const double*)grid, (double*)coef, s1, s2, ny, nx, row_factors, col_factors, st);
        narrow_kernel<double><<<g, 256, 0, st>>>(s2, (double*)coef, n, coef_stride, coef_offset);
    }
    SMRF_LAUNCH_CHECK();
    count_launches(7);
    return 0;
}

int smrf_classify(const void* x, const void* y, const void* z, int64_t n, int point_fmt, const double* inv6_host,
                  const void* coef_z, const void* coef_s, int64_t ny, int64_t nx, int dtype,
                  double elevation_threshold, double elevation_scaler, uint8_t* is_object, double* elevation,
                  double* slope_out, const uint8_t* drop_raster, uint8_t* when_dropped_pt, void* stream) {
    SMRF_CHECK_ARG(x && inv6_host && coef_z && is_object, "null pointer");
    SMRF_CHECK_ARG(point_fmt == SMRF_PTS_XYZW_F32 || (y && z), "y/z null");
    SMRF_CHECK_ARG(ny >= 4 && nx >= 4 && ny < (1LL << 30) && nx < (1LL << 30), "bad grid size");
    SMRF_CHECK_ARG(dtype == SMRF_F32 || dtype == SMRF_F64, "bad dtype");
    SMRF_CHECK_ARG(!when_dropped_pt || drop_raster, "when_dropped_pt needs drop_raster");
    return classify_any(x, y, z, n, point_fmt, inv6_host, coef_z, coef_s, ny, nx, dtype, elevation_threshold, elevation_scaler,
                        is_object, elevation, slope_out, drop_raster, when_dropped_pt, stream);
}

/* Row-band sharding: `coef_band` holds rows [row0, row0 + rows) of the interleaved (z, slope) coefficient grid; every
 * point must be interpolated from those rows only (points of the band's own rows need two rows of margin on each
 * side: the 4 x 4 taps of a point in cell row r span rows r-2 .. r+2). */
int smrf_classify_band(const void* x, const void* y, const void* z, int64_t n, int point_fmt, const double* inv6_host,
                       const void* coef_band, int64_t ny, int64_t nx, int64_t row0, int64_t rows, int dtype,
                       double elevation_threshold, double elevation_scaler, uint8_t* is_object, void* stream) {
    SMRF_CHECK_ARG(inv6_host && coef_band, "null pointer");
    SMRF_CHECK_ARG(row0 >= 0 && rows > 0 && row0 + rows <= ny, "bad row band");
    SMRF_CHECK_ARG(dtype == SMRF_F32 || dtype == SMRF_F64, "bad dtype");
    // the kernel addresses coefficients by global row: hand it the address row 0 would have
    const size_t es = dtype == SMRF_F32 ? 4 : 8;
    const char* virt = (const char*)coef_band - (size_t)row0 * (size_t)nx * 2 * es;
    return classify_any(x, y, z, n, point_fmt, inv6_host, virt, nullptr, ny, nx, dtype, elevation_threshold, elevation_scaler,
                        is_object, nullptr, nullptr, nullptr, nullptr, stream);
}

}  // extern "C"

static int classify_any(const void* x, const void* y, const void* z, int64_t n, int point_fmt, const double* inv6_host,
                        const void* coef_z, const void* coef_s, int64_t ny, int64_t nx, int dtype,
                        double elevation_threshold, double elevation_scaler, uint8_t* is_object, double* elevation,
                        double* slope_out, const uint8_t* drop_raster, uint8_t* when_dropped_pt, void* stream) {
    if (n <= 0) return 0;
    SMRF_CHECK_ARG(x && is_object, "null pointer");
    SMRF_CHECK_ARG(point_fmt == SMRF_PTS_XYZW_F32 || (y && z), "y/z null");
    SMRF_CHECK_ARG(ny >= 4 && nx >= 4 && ny < (1LL << 30) && nx < (1LL << 30), "bad grid size");
    Inv6 inv{inv6_host[0], inv6_host[1], inv6_host[2], inv6_host[3], inv6_host[4], inv6_host[5]};
    cudaStream_t st = (cudaStream_t)stream;
#define SMRF_GO(T, F)                                                                                         \
    classify_launch<T, F>(x, y, z, n, inv, coef_z, coef_s, (int)ny, (int)nx, elevation_threshold, elevation_scaler, \
                          is_object, elevation, slope_out, drop_raster, when_dropped_pt, st)
    if (dtype == SMRF_F32) {
        if (point_fmt == SMRF_PTS_SOA_F64) SMRF_GO(float, SMRF_PTS_SOA_F64);
        else if (point_fmt == SMRF_PTS_XYZW_F32) SMRF_GO(float, SMRF_PTS_XYZW_F32);
        else if (point_fmt == SMRF_PTS_SOA_F32) SMRF_GO(float, SMRF_PTS_SOA_F32);
        else SMRF_CHECK_ARG(false, "bad point_fmt");
    } else {
        if (point_fmt == SMRF_PTS_SOA_F64) SMRF_GO(double, SMRF_PTS_SOA_F64);
        else if (point_fmt == SMRF_PTS_XYZW_F32) SMRF_GO(double, SMRF_PTS_XYZW_F32);
        else if (point_fmt == SMRF_PTS_SOA_F32) SMRF_GO(double, SMRF_PTS_SOA_F32);
        else SMRF_CHECK_ARG(false, "bad point_fmt");
    }
#undef SMRF_GO
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}
