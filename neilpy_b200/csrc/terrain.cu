// Raster products of the DTM, the step after the SMRF path:
//   slope      neilpy.py:456-467     aspect     neilpy.py:471-484
//   hillshade  neilpy.py:814-824     pssm       neilpy.py:846-867
// One pass over the grid: the two np.gradient components from the four neighbours, then the
// per-cell formula of the requested product in float64, in the reference's order of
// operations.  sqrt, products and sums are correctly rounded and therefore identical to
// numpy's; atan / atan2 / sin / cos are CUDA's (<= 2 ulp), numpy's are the host libm's: float
// outputs agree to ~1e-15 relative, the uint8 outputs except where 255*H lands within that of
// a rounding tie.  4 or 8 bytes read and 1 to 32 bytes written per cell; the float64
// transcendentals make it FP64-bound rather than HBM-bound, which is irrelevant at
// 25 M cells (one launch, well under a millisecond).
#include "common.cuh"

namespace smrf {
namespace terrain {

constexpr double kHalfPi = 1.5707963267948966;      // np.pi / 2
constexpr double kTwoPi = 6.283185307179586;        // 2 * np.pi
constexpr double kRadToDeg = 57.29577951308232;     // 180 / np.pi, the factor np.rad2deg multiplies by

struct Params {
    int mode;             // SMRF_TERRAIN_*
    int return_as;        // slope / aspect: 0 percent (slope only), 1 radians, 2 degrees
    int out_u8;           // hillshade: uint8 output (else float64)
    int flat_is_nan;      // aspect: flat cells become NaN (else flat_value)
    double spacing;       // gradient spacing of slope / hillshade / pssm (cellsize / z_factor)
    double flat_value;
    double cos_zenith, sin_zenith, azimuth;      // hillshade, radians; computed by the caller with numpy
    double ve;            // pssm vertical exaggeration
};

// np.gradient(Z, h) at (y, x): central differences, one-sided at the edges; axis 0 = rows
template <typename T>
__device__ __forceinline__ void gradient(const T* __restrict__ z, int64_t ny, int64_t nx, int64_t y, int64_t x, double h,
                                         double& gy, double& gx) {
    const int64_t i = y * nx + x;
    const double h2 = __dmul_rn(2.0, h);
    if (ny == 1) gy = 0.0;
    else if (y == 0) gy = __ddiv_rn(__dsub_rn((double)z[i + nx], (double)z[i]), h);
    else if (y == ny - 1) gy = __ddiv_rn(__dsub_rn((double)z[i], (double)z[i - nx]), h);
    else gy = __ddiv_rn(__dsub_rn((double)z[i + nx], (double)z[i - nx]), h2);
    if (nx == 1) gx = 0.0;
    else if (x == 0) gx = __ddiv_rn(__dsub_rn((double)z[i + 1], (double)z[i]), h);
    else if (x == nx - 1) gx = __ddiv_rn(__dsub_rn((double)z[i], (double)z[i - 1]), h);
    else gx = __ddiv_rn(__dsub_rn((double)z[i + 1], (double)z[i - 1]), h2);
}

__device__ __forceinline__ double norm2(double gx, double gy) {          // sqrt(gx**2 + gy**2)
    return __dsqrt_rn(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)));
}

// aspect in radians, before the flat-cell override (neilpy.py:476-478)
__device__ __forceinline__ double bearing(double gy, double gx) {
    double a = __dsub_rn(kHalfPi, atan2(gy, -gx));
    if (a < 0) a = __dadd_rn(a, kTwoPi);
    return a;
}

__device__ __forceinline__ uint8_t to_u8(double v) {                     // np.round(v).astype(np.uint8), v in [0, 255]
    if (!(v == v)) return 0;
    const double r = rint(v);                                            // half to even, as np.round
    return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : (int)r));
}

template <typename T>
__global__ void __launch_bounds__(256) terrain_kernel(const T* __restrict__ z, double* __restrict__ out_f64,
                                                      uint8_t* __restrict__ out_u8, double* __restrict__ rgba,
                                                      const double* __restrict__ lut, int64_t ny, int64_t nx, Params p) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nx) return;
    for (int64_t y = blockIdx.y; y < ny; y += gridDim.y) {
        const int64_t i = y * nx + x;
        double gy, gx;
        if (p.mode == SMRF_TERRAIN_ASPECT) {
            gradient(z, ny, nx, y, x, 1.0, gy, gx);
            double a = bearing(gy, gx);
            if (p.return_as == 2) a = __dmul_rn(a, kRadToDeg);
            if (gx == 0 && gy == 0) a = p.flat_is_nan ? quiet_nan<double>() : p.flat_value;
            out_f64[i] = a;
            continue;
        }
        gradient(z, ny, nx, y, x, p.spacing, gy, gx);
        double s = norm2(gx, gy);
        if (p.mode == SMRF_TERRAIN_SLOPE) {
            if (p.return_as >= 1) s = atan(s);
            if (p.return_as == 2) s = __dmul_rn(s, kRadToDeg);
            out_f64[i] = s;
        } else if (p.mode == SMRF_TERRAIN_HILLSHADE) {
            s = atan(s);
            double g1y, g1x;
            gradient(z, ny, nx, y, x, 1.0, g1y, g1x);
            double a = bearing(g1y, g1x);
            if (g1x == 0 && g1y == 0) a = 0.0;                            // aspect(..., flat_as=0)
            double h = __dadd_rn(__dmul_rn(p.cos_zenith, cos(s)),
                                 __dmul_rn(__dmul_rn(p.sin_zenith, sin(s)), cos(__dsub_rn(p.azimuth, a))));
            if (h < 0) h = 0.0;
            if (p.out_u8) out_u8[i] = to_u8(__dmul_rn(255.0, h));
            else out_f64[i] = h;
        } else {                                                          // pssm
            const double deg = __dmul_rn(atan(__dmul_rn(p.ve, s)), kRadToDeg);
            const uint8_t k = to_u8(__dmul_rn(255.0, __ddiv_rn(deg, 90.0)));
            if (out_u8) out_u8[i] = k;
            if (rgba) {
                const double4 c = *reinterpret_cast<const double4*>(lut + 4 * (int)k);
                *reinterpret_cast<double4*>(rgba + 4 * i) = c;
            }
        }
    }
}

}  // namespace terrain
}  // namespace smrf

using namespace smrf;

extern "C" {

int smrf_terrain(const void* grid, int64_t ny, int64_t nx, int dtype, int mode, int return_as, double spacing,
                 int flat_is_nan, double flat_value, double cos_zenith, double sin_zenith, double azimuth_rad,
                 double ve, double* out_f64, uint8_t* out_u8, double* rgba, const double* lut_rgba, void* stream) {
    SMRF_CHECK_ARG(grid, "null grid");
    SMRF_CHECK_ARG(ny > 0 && nx > 0, "empty grid");
    SMRF_CHECK_ARG(mode >= SMRF_TERRAIN_SLOPE && mode <= SMRF_TERRAIN_PSSM, "bad mode");
    SMRF_CHECK_ARG(return_as >= 0 && return_as <= 2, "bad return_as");
    if (mode == SMRF_TERRAIN_SLOPE || mode == SMRF_TERRAIN_ASPECT) SMRF_CHECK_ARG(out_f64, "float64 output required");
    if (mode == SMRF_TERRAIN_HILLSHADE) SMRF_CHECK_ARG((out_u8 != nullptr) != (out_f64 != nullptr), "exactly one of out_u8 / out_f64");
    if (mode == SMRF_TERRAIN_PSSM) {
        SMRF_CHECK_ARG(out_u8 || rgba, "no output");
        SMRF_CHECK_ARG(!rgba || lut_rgba, "rgba output needs the colour table");
        SMRF_CHECK_ARG((((uintptr_t)rgba | (uintptr_t)lut_rgba) & 31) == 0, "rgba and the table must be 32-byte aligned");
    }
    terrain::Params p;
    p.mode = mode; p.return_as = return_as; p.out_u8 = out_u8 != nullptr; p.flat_is_nan = flat_is_nan;
    p.spacing = spacing; p.flat_value = flat_value; p.cos_zenith = cos_zenith; p.sin_zenith = sin_zenith;
    p.azimuth = azimuth_rad; p.ve = ve;
    dim3 g((unsigned)((nx + 255) / 256), (unsigned)(ny < 32768 ? ny : 32768));
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SMRF_F32) terrain::terrain_kernel<float><<<g, 256, 0, st>>>((const float*)grid, out_f64, out_u8, rgba, lut_rgba, ny, nx, p);
    else if (dtype == SMRF_F64) terrain::terrain_kernel<double><<<g, 256, 0, st>>>((const double*)grid, out_f64, out_u8, rgba, lut_rgba, ny, nx, p);
    else SMRF_CHECK_ARG(false, "bad dtype");
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

}  // extern "C"
