// Point -> grid minimum/maximum-surface binning (create_dem, neilpy.py:1110-1166).
//
// HBM-bound streaming kernels: the point stream is read once, coalesced (one float4
// or three doubles per point); each point issues one REDG.MIN/MAX on an
// order-preserving integer key of its z.  The grid holds keys until finalize
// decodes them in place, so there is no second grid-sized buffer.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace smrf {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// ------------------------------------------------------------------ extent
__global__ void extent_init_kernel(long long* keys4, int64_t* nonfinite) {
    keys4[0] = KeyOf<double>::empty_min;  // min x
    keys4[1] = KeyOf<double>::empty_max;  // max x
    keys4[2] = KeyOf<double>::empty_min;  // min y
    keys4[3] = KeyOf<double>::empty_max;  // max y
    nonfinite[0] = 0;
}

template <int FMT>
__global__ void __launch_bounds__(256) extent_kernel(PointLoader<FMT> pts, int64_t n, long long* keys4,
                                                     int64_t* nonfinite) {
    double mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
    int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double x, y;
        pts.xy(i, x, y);
        if (!isfinite(x) || !isfinite(y)) { ++bad; continue; }
        mnx = fmin(mnx, x); mxx = fmax(mxx, x);
        mny = fmin(mny, y); mxy = fmax(mxy, y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
    }
    __shared__ double s[4][8];
    __shared__ int sbad[8];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s[0][w] = mnx; s[1][w] = mxx; s[2][w] = mny; s[3][w] = mxy; sbad[w] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int nb = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
            mnx = fmin(mnx, s[0][k]); mxx = fmax(mxx, s[1][k]);
            mny = fmin(mny, s[2][k]); mxy = fmax(mxy, s[3][k]);
            nb += sbad[k];
        }
        if (mnx <= mxx) {
            atomicMin(&keys4[0], f64_key(mnx)); atomicMax(&keys4[1], f64_key(mxx));
            atomicMin(&keys4[2], f64_key(mny)); atomicMax(&keys4[3], f64_key(mxy));
        }
        if (nb) atomicAdd((unsigned long long*)nonfinite, (unsigned long long)nb);
    }
}

__global__ void extent_final_kernel(const long long* keys4, double* out4) {
    int i = threadIdx.x;
    if (i < 4) {
        long long k = keys4[i];
        bool empty = (i & 1) ? (k == KeyOf<double>::empty_max) : (k == KeyOf<double>::empty_min);
        out4[i] = empty ? quiet_nan<double>() : f64_unkey(k);
    }
}

// ------------------------------------------------------------------ binning
template <typename T>
__global__ void __launch_bounds__(256) bin_init_kernel(typename KeyOf<T>::type* keys, int64_t n, int bin_type) {
    using K = typename KeyOf<T>::type;
    const K e = bin_type == SMRF_BIN_MIN ? KeyOf<T>::empty_min : KeyOf<T>::empty_max;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        keys[i] = e;
}

template <typename T, int FMT>
__global__ void __launch_bounds__(256) bin_accumulate_kernel(PointLoader<FMT> pts, int64_t n, Inv6 inv,
                                                             typename KeyOf<T>::type* keys, int64_t ny, int64_t nx,
                                                             int bin_type, int64_t* out_of_range, int64_t row0,
                                                             int64_t rows) {
    int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double x, y, z;
        pts.xyz(i, x, y, z);
        double c, r;
        affine_apply(inv, x, y, c, r);
        c = floor(c);
        r = floor(r);
        // NaN compares false -> counted as out of range, as np.ravel_multi_index would raise
        if (!(c >= 0.0 && c < (double)nx && r >= 0.0 && r < (double)ny)) { ++bad; continue; }
        // `keys` holds rows [row0, row0 + rows) of the grid (the whole grid unless row-band sharded); a point
        // routed to the wrong band is reported like an out-of-range one
        const int64_t rl = (int64_t)r - row0;
        if (rl < 0 || rl >= rows) { ++bad; continue; }
        if (z != z) continue;  // pandas groupby().min()/max() skips NaN
        int64_t cell = rl * nx + (int64_t)c;
        typename KeyOf<T>::type k = KeyOf<T>::key((T)z);
        if (bin_type == SMRF_BIN_MIN) atomicMin(&keys[cell], k);
        else atomicMax(&keys[cell], k);
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if (bad && (threadIdx.x & 31) == 0) atomicAdd((unsigned long long*)out_of_range, (unsigned long long)bad);
}

template <typename T>
__global__ void __launch_bounds__(256) bin_finalize_kernel(T* grid, uint8_t* empty, int64_t n, int bin_type) {
    using K = typename KeyOf<T>::type;
    const K e = bin_type == SMRF_BIN_MIN ? KeyOf<T>::empty_min : KeyOf<T>::empty_max;
    K* keys = reinterpret_cast<K*>(grid);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        K k = keys[i];
        bool is_empty = (k == e);
        grid[i] = is_empty ? quiet_nan<T>() : KeyOf<T>::unkey(k);
        if (empty) empty[i] = is_empty ? 1 : 0;
    }
}

// row-band sharding: a rank's partial grid keeps +inf (min) / -inf (max) in untouched cells
// so that an elementwise min/max reduce-scatter over ranks is the global binning
template <typename T>
__global__ void __launch_bounds__(256) bin_finalize_inf_kernel(T* grid, int64_t n, int bin_type) {
    using K = typename KeyOf<T>::type;
    const K e = bin_type == SMRF_BIN_MIN ? KeyOf<T>::empty_min : KeyOf<T>::empty_max;
    const T ident = bin_type == SMRF_BIN_MIN ? (T)INFINITY : (T)-INFINITY;
    K* keys = reinterpret_cast<K*>(grid);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        K k = keys[i];
        grid[i] = (k == e) ? ident : KeyOf<T>::unkey(k);
    }
}
template <typename T>
__global__ void __launch_bounds__(256) mark_empty_kernel(T* grid, uint8_t* empty, int64_t n, int bin_type) {
    const T ident = bin_type == SMRF_BIN_MIN ? (T)INFINITY : (T)-INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool is_empty = grid[i] == ident;
        if (is_empty) grid[i] = quiet_nan<T>();
        if (empty) empty[i] = is_empty ? 1 : 0;
    }
}

static inline int grid_for(int64_t n, int per_block = 256, int waves = 8) {
    int64_t b = (n + per_block - 1) / per_block;
    int64_t cap = (int64_t)num_sms() * waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

template <int FMT>
static PointLoader<FMT> make_loader(const void* x, const void* y, const void* z);
template <>
PointLoader<SMRF_PTS_SOA_F64> make_loader<SMRF_PTS_SOA_F64>(const void* x, const void* y, const void* z) {
    return PointLoader<SMRF_PTS_SOA_F64>{(const double*)x, (const double*)y, (const double*)z};
}
template <>
PointLoader<SMRF_PTS_XYZW_F32> make_loader<SMRF_PTS_XYZW_F32>(const void* x, const void*, const void*) {
    return PointLoader<SMRF_PTS_XYZW_F32>{(const float4*)x, nullptr, nullptr};
}
template <>
PointLoader<SMRF_PTS_SOA_F32> make_loader<SMRF_PTS_SOA_F32>(const void* x, const void* y, const void* z) {
    return PointLoader<SMRF_PTS_SOA_F32>{(const float*)x, (const float*)y, (const float*)z};
}

}  // namespace smrf

using namespace smrf;

template <typename T>
static int bin_accumulate_t(const void* x, const void* y, const void* z, int64_t n, int point_fmt, Inv6 inv,
                            void* grid, int64_t ny, int64_t nx, int bin_type, int64_t* oor, cudaStream_t st,
                            int64_t row0, int64_t rows) {
    using K = typename KeyOf<T>::type;
    int g = grid_for(n, 256, 16);
    switch (point_fmt) {
        case SMRF_PTS_SOA_F64:
            bin_accumulate_kernel<T, SMRF_PTS_SOA_F64><<<g, 256, 0, st>>>(make_loader<SMRF_PTS_SOA_F64>(x, y, z), n, inv, (K*)grid, ny, nx, bin_type, oor, row0, rows);
            break;
        case SMRF_PTS_XYZW_F32:
            bin_accumulate_kernel<T, SMRF_PTS_XYZW_F32><<<g, 256, 0, st>>>(make_loader<SMRF_PTS_XYZW_F32>(x, y, z), n, inv, (K*)grid, ny, nx, bin_type, oor, row0, rows);
            break;
        case SMRF_PTS_SOA_F32:
            bin_accumulate_kernel<T, SMRF_PTS_SOA_F32><<<g, 256, 0, st>>>(make_loader<SMRF_PTS_SOA_F32>(x, y, z), n, inv, (K*)grid, ny, nx, bin_type, oor, row0, rows);
            break;
        default:
            set_error("smrf_bin_accumulate: bad point_fmt");
            return SMRF_E_ARG;
    }
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

extern "C" {

int smrf_abi_version(void) { return 1; }
unsigned long long smrf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
const char* smrf_last_error(void) { return g_err; }

int smrf_extent(const void* x, const void* y, int64_t n, int point_fmt, double* out4, int64_t* nonfinite,
                int64_t* scratch4, void* stream) {
    SMRF_CHECK_ARG(x && out4 && nonfinite && scratch4, "null pointer");
    SMRF_CHECK_ARG(n >= 0, "negative n");
    SMRF_CHECK_ARG(point_fmt == SMRF_PTS_XYZW_F32 || y, "y is null");
    cudaStream_t st = (cudaStream_t)stream;
    long long* keys = (long long*)scratch4;
    extent_init_kernel<<<1, 1, 0, st>>>(keys, nonfinite);
    if (n > 0) {
        int g = grid_for(n, 256, 8);
        switch (point_fmt) {
            case SMRF_PTS_SOA_F64:
                extent_kernel<SMRF_PTS_SOA_F64><<<g, 256, 0, st>>>(make_loader<SMRF_PTS_SOA_F64>(x, y, nullptr), n, keys, nonfinite);
                break;
            case SMRF_PTS_XYZW_F32:
                extent_kernel<SMRF_PTS_XYZW_F32><<<g, 256, 0, st>>>(make_loader<SMRF_PTS_XYZW_F32>(x, y, nullptr), n, keys, nonfinite);
                break;
            case SMRF_PTS_SOA_F32:
                extent_kernel<SMRF_PTS_SOA_F32><<<g, 256, 0, st>>>(make_loader<SMRF_PTS_SOA_F32>(x, y, nullptr), n, keys, nonfinite);
                break;
            default:
                SMRF_CHECK_ARG(false, "bad point_fmt");
        }
    }
    extent_final_kernel<<<1, 32, 0, st>>>(keys, out4);
    SMRF_LAUNCH_CHECK();
    count_launches(n > 0 ? 3 : 2);
    return 0;
}

int smrf_bin_init(void* grid, int64_t ny, int64_t nx, int dtype, int bin_type, void* stream) {
    SMRF_CHECK_ARG(grid, "null grid");
    SMRF_CHECK_ARG(ny > 0 && nx > 0, "empty grid");
    SMRF_CHECK_ARG(bin_type == SMRF_BIN_MIN || bin_type == SMRF_BIN_MAX, "This type not supported.");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t n = ny * nx;
    int g = grid_for(n, 256, 16);
    if (dtype == SMRF_F32) bin_init_kernel<float><<<g, 256, 0, st>>>((int*)grid, n, bin_type);
    else if (dtype == SMRF_F64) bin_init_kernel<double><<<g, 256, 0, st>>>((long long*)grid, n, bin_type);
    else SMRF_CHECK_ARG(false, "bad dtype");
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int smrf_bin_accumulate(const void* x, const void* y, const void* z, int64_t n, int point_fmt,
                        const double* inv6_host, void* grid, int64_t ny, int64_t nx, int dtype, int bin_type,
                        int64_t* out_of_range, void* stream) {
    SMRF_CHECK_ARG(x && inv6_host && grid && out_of_range, "null pointer");
    SMRF_CHECK_ARG(point_fmt == SMRF_PTS_XYZW_F32 || (y && z), "y/z null");
    SMRF_CHECK_ARG(ny > 0 && nx > 0 && n >= 0, "bad size");
    SMRF_CHECK_ARG(bin_type == SMRF_BIN_MIN || bin_type == SMRF_BIN_MAX, "This type not supported.");
    if (n == 0) return 0;
    Inv6 inv{inv6_host[0], inv6_host[1], inv6_host[2], inv6_host[3], inv6_host[4], inv6_host[5]};
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SMRF_F32) return bin_accumulate_t<float>(x, y, z, n, point_fmt, inv, grid, ny, nx, bin_type, out_of_range, st, 0, ny);
    if (dtype == SMRF_F64) return bin_accumulate_t<double>(x, y, z, n, point_fmt, inv, grid, ny, nx, bin_type, out_of_range, st, 0, ny);
    SMRF_CHECK_ARG(false, "bad dtype");
}

int smrf_bin_accumulate_band(const void* x, const void* y, const void* z, int64_t n, int point_fmt,
                             const double* inv6_host, void* band, int64_t ny, int64_t nx, int64_t row0, int64_t rows,
                             int dtype, int bin_type, int64_t* out_of_range, void* stream) {
    SMRF_CHECK_ARG(inv6_host && band && out_of_range, "null pointer");
    SMRF_CHECK_ARG(ny > 0 && nx > 0 && n >= 0 && row0 >= 0 && rows > 0 && row0 + rows <= ny, "bad size / row band");
    SMRF_CHECK_ARG(bin_type == SMRF_BIN_MIN || bin_type == SMRF_BIN_MAX, "This type not supported.");
    if (n == 0) return 0;
    SMRF_CHECK_ARG(x && (point_fmt == SMRF_PTS_XYZW_F32 || (y && z)), "null points");
    Inv6 inv{inv6_host[0], inv6_host[1], inv6_host[2], inv6_host[3], inv6_host[4], inv6_host[5]};
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SMRF_F32) return bin_accumulate_t<float>(x, y, z, n, point_fmt, inv, band, ny, nx, bin_type, out_of_range, st, row0, rows);
    if (dtype == SMRF_F64) return bin_accumulate_t<double>(x, y, z, n, point_fmt, inv, band, ny, nx, bin_type, out_of_range, st, row0, rows);
    SMRF_CHECK_ARG(false, "bad dtype");
}

int smrf_bin_finalize(void* grid, uint8_t* empty, int64_t ny, int64_t nx, int dtype, int bin_type, void* stream) {
    SMRF_CHECK_ARG(grid, "null grid");
    SMRF_CHECK_ARG(ny > 0 && nx > 0, "empty grid");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t n = ny * nx;
    int g = grid_for(n, 256, 16);
    if (dtype == SMRF_F32) bin_finalize_kernel<float><<<g, 256, 0, st>>>((float*)grid, empty, n, bin_type);
    else if (dtype == SMRF_F64) bin_finalize_kernel<double><<<g, 256, 0, st>>>((double*)grid, empty, n, bin_type);
    else SMRF_CHECK_ARG(false, "bad dtype");
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int smrf_bin_finalize_partial(void* grid, int64_t ny, int64_t nx, int dtype, int bin_type, void* stream) {
    SMRF_CHECK_ARG(grid, "null grid");
    SMRF_CHECK_ARG(ny > 0 && nx > 0, "empty grid");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t n = ny * nx;
    int g = grid_for(n, 256, 16);
    if (dtype == SMRF_F32) bin_finalize_inf_kernel<float><<<g, 256, 0, st>>>((float*)grid, n, bin_type);
    else if (dtype == SMRF_F64) bin_finalize_inf_kernel<double><<<g, 256, 0, st>>>((double*)grid, n, bin_type);
    else SMRF_CHECK_ARG(false, "bad dtype");
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int smrf_bin_mark_empty(void* grid, uint8_t* empty, int64_t ny, int64_t nx, int dtype, int bin_type, void* stream) {
    SMRF_CHECK_ARG(grid, "null grid");
    SMRF_CHECK_ARG(ny > 0 && nx > 0, "empty grid");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t n = ny * nx;
    int g = grid_for(n, 256, 16);
    if (dtype == SMRF_F32) mark_empty_kernel<float><<<g, 256, 0, st>>>((float*)grid, empty, n, bin_type);
    else if (dtype == SMRF_F64) mark_empty_kernel<double><<<g, 256, 0, st>>>((double*)grid, empty, n, bin_type);
    else SMRF_CHECK_ARG(false, "bad dtype");
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

}  // extern "C"
