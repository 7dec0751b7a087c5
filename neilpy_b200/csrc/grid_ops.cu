// Small fused grid passes of smrf(): mask merge + NaN punch (neilpy.py:1762-1763) and
// the slope raster (neilpy.py:1785-1786).  One read and one write per cell: HBM-bound.
#include "common.cuh"

namespace smrf {

template <typename T>
__global__ void __launch_bounds__(256) merge_punch_kernel(T* __restrict__ grid, const uint8_t* __restrict__ empty,
                                                          const uint8_t* __restrict__ low, const uint8_t* __restrict__ obj,
                                                          uint8_t* __restrict__ object_cells, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint8_t o = 0;
        if (empty) o |= empty[i];
        if (low) o |= low[i];
        if (obj) o |= obj[i];
        o = o ? 1 : 0;
        if (object_cells) object_cells[i] = o;
        if (o) grid[i] = quiet_nan<T>();
    }
}

// np.gradient(Z, cs): interior (f[i+1]-f[i-1])/(2cs), edges (f[1]-f[0])/cs, (f[n-1]-f[n-2])/cs;
// axis 0 = rows.  S = sqrt(gy^2 + gx^2).  Every operation is a separately rounded float64
// operation, in numpy's order.
template <typename T>
__global__ void __launch_bounds__(256) slope_kernel(const T* __restrict__ z, T* __restrict__ s, int64_t ny, int64_t nx,
                                                    double cs) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nx) return;
    const double cs2 = __dmul_rn(2.0, cs);
    for (int64_t y = blockIdx.y; y < ny; y += gridDim.y) {
        const int64_t i = y * nx + x;
        double gy, gx;
        if (ny == 1) gy = 0.0;
        else if (y == 0) gy = __ddiv_rn(__dsub_rn((double)z[i + nx], (double)z[i]), cs);
        else if (y == ny - 1) gy = __ddiv_rn(__dsub_rn((double)z[i], (double)z[i - nx]), cs);
        else gy = __ddiv_rn(__dsub_rn((double)z[i + nx], (double)z[i - nx]), cs2);
        if (nx == 1) gx = 0.0;
        else if (x == 0) gx = __ddiv_rn(__dsub_rn((double)z[i + 1], (double)z[i]), cs);
        else if (x == nx - 1) gx = __ddiv_rn(__dsub_rn((double)z[i], (double)z[i - 1]), cs);
        else gx = __ddiv_rn(__dsub_rn((double)z[i + 1], (double)z[i - 1]), cs2);
        s[i] = (T)__dsqrt_rn(__dadd_rn(__dmul_rn(gy, gy), __dmul_rn(gx, gx)));
    }
}

}  // namespace smrf

using namespace smrf;

extern "C" {

int smrf_merge_punch(void* grid, const uint8_t* empty, const uint8_t* low, const uint8_t* obj, uint8_t* object_cells,
                     int64_t ny, int64_t nx, int dtype, void* stream) {
    SMRF_CHECK_ARG(grid, "null grid");
    SMRF_CHECK_ARG(ny > 0 && nx > 0, "empty grid");
    const int64_t n = ny * nx;
    int g = (int)((n + 255) / 256);
    int cap = num_sms() * 16;
    if (g > cap) g = cap;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SMRF_F32) merge_punch_kernel<float><<<g, 256, 0, st>>>((float*)grid, empty, low, obj, object_cells, n);
    else if (dtype == SMRF_F64) merge_punch_kernel<double><<<g, 256, 0, st>>>((double*)grid, empty, low, obj, object_cells, n);
    else SMRF_CHECK_ARG(false, "bad dtype");
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int smrf_slope(const void* grid, void* slope, int64_t ny, int64_t nx, int dtype, double cellsize, void* stream) {
    SMRF_CHECK_ARG(grid && slope && grid != slope, "null or aliased pointer");
    SMRF_CHECK_ARG(ny > 0 && nx > 0, "empty grid");
    dim3 g((unsigned)((nx + 255) / 256), (unsigned)(ny < 32768 ? ny : 32768));
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SMRF_F32) slope_kernel<float><<<g, 256, 0, st>>>((const float*)grid, (float*)slope, ny, nx, cellsize);
    else if (dtype == SMRF_F64) slope_kernel<double><<<g, 256, 0, st>>>((const double*)grid, (double*)slope, ny, nx, cellsize);
    else SMRF_CHECK_ARG(false, "bad dtype");
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

}  // extern "C"
