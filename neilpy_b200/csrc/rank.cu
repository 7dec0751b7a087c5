// float64 surfaces are opened in RANK SPACE (progressive_filter, neilpy.py:1659-1680, float64 as the
// reference computes it).
//
// A grey-scale opening only ever selects input values, so it commutes with any order-preserving map.
// float64 has no min/max instruction on sm_100a (DSETP + two selects, and twice the registers); instead
// the surface is sorted once (CUB radix sort of order-preserving 64-bit keys), every cell gets its
// rank r, and the rank is stored as the float32 whose BIT PATTERN is r + 0x00800000 -- positive normal
// floats order like their bit patterns, so the float32 marching kernels (opening_march.cuh) run
// unmodified on the rank plane, window after window, at float32 speed, for up to 2.1e9 cells.  The
// slope test needs real elevations: after each window a compare pass looks the two ranks up in the
// sorted table (only where the rank changed: an unchanged rank is an unchanged elevation) and evaluates
// (last - this) > thr in float64 exactly as the float64 kernels would.  Bit-exact for arbitrary float64
// input; NaN cells stay NaN ("no sample") through the float32 NaN.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "opening.cuh"

namespace smrf {
namespace rank {

constexpr uint32_t kBias = 0x00800000u;      // smallest positive normal float
constexpr uint32_t kNaNBits = 0x7fc00000u;
constexpr uint64_t kNaNKey = ~0ull;

__device__ __forceinline__ uint64_t ordered_key(double v) {
    if (v != v) return kNaNKey;
    return (uint64_t)f64_key(v) ^ 0x8000000000000000ull;     // signed order -> unsigned order
}
__device__ __forceinline__ double key_value(uint64_t k) {
    return f64_unkey((long long)(k ^ 0x8000000000000000ull));
}

__global__ void __launch_bounds__(256) keys_kernel(const double* __restrict__ surf, int64_t nx, int64_t pitch, int64_t n,
                                                   uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int64_t y = i / nx, x = i - y * nx;
        keys[i] = ordered_key(surf[y * pitch + x]);
        idx[i] = (uint32_t)i;
    }
}

// rank plane: cell sorted_idx[i] gets the float whose bits are i + kBias (NaN cells: the float NaN)
__global__ void __launch_bounds__(256) scatter_kernel(const uint64_t* __restrict__ sorted_keys,
                                                      const uint32_t* __restrict__ sorted_idx, int64_t n, int64_t nx,
                                                      int64_t pitch, float* __restrict__ plane) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int64_t cell = sorted_idx[i];
        const int64_t y = cell / nx, x = cell - y * nx;
        const uint32_t bits = sorted_keys[i] == kNaNKey ? kNaNBits : (uint32_t)i + kBias;
        plane[y * pitch + x] = __uint_as_float(bits);
    }
}

__device__ __forceinline__ double rank_value(const uint64_t* __restrict__ table, uint32_t bits) {
    if (bits == kNaNBits) return quiet_nan<double>();
    return key_value(__ldg(table + (bits - kBias)));
}

// new_obj = (last - this) > thr in float64 on rows [row_lo, row_hi); mask |= new_obj; when[new_obj] = widx
__global__ void __launch_bounds__(256) threshold_kernel(const float* __restrict__ last, const float* __restrict__ cur,
                                                        const uint64_t* __restrict__ table, uint8_t* __restrict__ mask,
                                                        uint8_t* __restrict__ when, int64_t nx, int64_t pitch, double thr,
                                                        int widx, int64_t row_lo, int64_t row_hi) {
    const int64_t n0 = row_lo * nx, n1 = row_hi * nx;
    for (int64_t i = n0 + (int64_t)blockIdx.x * 256 + threadIdx.x; i < n1; i += (int64_t)gridDim.x * 256) {
        const int64_t y = i / nx, g = y * pitch + (i - y * nx);
        const uint32_t rl = __float_as_uint(last[g]), rt = __float_as_uint(cur[g]);
        bool obj;
        if (rl == rt) obj = (rl != kNaNBits) && (0.0 > thr);
        else obj = __dsub_rn(rank_value(table, rl), rank_value(table, rt)) > thr;
        if (obj) {
            mask[i] = 1;
            if (when) when[i] = (uint8_t)widx;
        }
    }
}

__global__ void __launch_bounds__(256) decode_kernel(const float* __restrict__ plane, const uint64_t* __restrict__ table,
                                                     double* __restrict__ out, int64_t nx, int64_t pitch, int64_t out_pitch,
                                                     int64_t row_lo, int64_t row_hi) {
    const int64_t n0 = row_lo * nx, n1 = row_hi * nx;
    for (int64_t i = n0 + (int64_t)blockIdx.x * 256 + threadIdx.x; i < n1; i += (int64_t)gridDim.x * 256) {
        const int64_t y = i / nx, x = i - y * nx;
        out[y * out_pitch + x] = rank_value(table, __float_as_uint(plane[y * pitch + x]));
    }
}

static inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }
static inline int64_t pitch32(int64_t nx) { return (nx + 3) / 4 * 4; }

struct Layout {
    size_t table, a, b, c, temp, total, temp_bytes, plane;
};

static size_t sort_temp_bytes(int64_t n) {
    size_t bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)n);
    if (e != cudaSuccess) {       // no device (size queries on a CPU box): a generous bound
        cudaGetLastError();
        bytes = (size_t)n / 8 + ((size_t)32 << 20);
    }
    return bytes;
}

// [table 8n][A: two rank planes; the unsorted keys before that][B: the two-pass kernels' tmp plane; the unsorted
// indices before that][C: sorted indices][CUB temp]
static Layout layout(int64_t ny, int64_t nx) {
    Layout L;
    const size_t n = (size_t)ny * (size_t)nx;
    L.plane = up256((size_t)ny * (size_t)pitch32(nx) * 4);
    L.temp_bytes = sort_temp_bytes((int64_t)n);
    size_t off = 0;
    L.table = off; off += up256(n * 8);
    L.a = off; off += (2 * L.plane > up256(n * 8) ? 2 * L.plane : up256(n * 8));
    L.b = off; off += L.plane;
    L.c = off; off += up256(n * 4);
    L.temp = off; off += up256(L.temp_bytes);
    L.total = off;
    return L;
}

static inline int grid_for(int64_t n) {
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace rank

size_t open_f64_workspace_bytes(int64_t ny, int64_t nx) { return rank::layout(ny, nx).total; }

// The whole progressive filter of a float64 surface in rank space.  `surface` rows are `pitch_in` doubles apart.
// Windows run on rows [row_lo, row_hi) (pass 0, ny for the whole grid); window i is reported as window_index0 + i.
// `advance`: the surface advances from window to window (neilpy.py:1675-1676: only with more than one window).
int open_f64_rank(const double* surface, int64_t ny, int64_t nx, int64_t pitch_in, void* ws, size_t ws_bytes, uint8_t* mask,
                  uint8_t* when, const int32_t* windows, const double* thr, int n_windows, int window_index0, int advance,
                  int64_t row_lo, int64_t row_hi, double* last_out, int64_t out_pitch, cudaStream_t st) {
    using namespace rank;
    const int64_t n = ny * nx;
    if (n + (int64_t)kBias >= (int64_t)0x7f800000) {
        set_error("open_f64_rank: %lld cells exceed the %d ranks a float32 bit pattern can carry", (long long)n,
                  (int)(0x7f800000u - kBias));
        return SMRF_E_UNSUPPORTED;
    }
    const Layout L = layout(ny, nx);
    if (ws_bytes < L.total) {
        set_error("open_f64_rank: workspace %zu < %zu bytes", ws_bytes, L.total);
        return SMRF_E_WORKSPACE;
    }
    char* base = (char*)ws;
    uint64_t* table = (uint64_t*)(base + L.table);
    uint64_t* keys_in = (uint64_t*)(base + L.a);
    uint32_t* idx_in = (uint32_t*)(base + L.b);
    uint32_t* idx_out = (uint32_t*)(base + L.c);
    const int64_t pitch = pitch32(nx);
    keys_kernel<<<grid_for(n), 256, 0, st>>>(surface, nx, pitch_in, n, keys_in, idx_in);
    size_t tb = L.temp_bytes;
    SMRF_CUDA(cub::DeviceRadixSort::SortPairs(base + L.temp, tb, (const uint64_t*)keys_in, table, (const uint32_t*)idx_in,
                                              idx_out, n, 0, 64, st));
    // the unsorted keys are dead: their space becomes the two rank planes
    float* pa = (float*)(base + L.a);
    float* pb = (float*)(base + L.a + L.plane);
    float* tmp = (float*)(base + L.b);
    scatter_kernel<<<grid_for(n), 256, 0, st>>>(table, idx_out, n, nx, pitch, pa);
    SMRF_LAUNCH_CHECK();
    count_launches(2);
    float *cur = pa, *nxt = pb;
    for (int i = 0; i < n_windows; ++i) {
        if (int rc = open_window_march(cur, nxt, tmp, nullptr, nullptr, ny, nx, pitch, SMRF_F32, windows[i], 0.0, 0, 0, row_lo,
                                       row_hi, st))
            return rc;
        if (mask) {
            threshold_kernel<<<grid_for((row_hi - row_lo) * nx), 256, 0, st>>>(cur, nxt, table, mask, when, nx, pitch, thr[i],
                                                                               window_index0 + i, row_lo, row_hi);
            SMRF_LAUNCH_CHECK();
            count_launches(1);
        }
        if (advance) {
            float* t = cur; cur = nxt; nxt = t;
        }
    }
    if (last_out && n_windows > 0) {
        const float* res = advance ? cur : nxt;
        decode_kernel<<<grid_for((row_hi - row_lo) * nx), 256, 0, st>>>(res, table, last_out, nx, pitch, out_pitch, row_lo, row_hi);
        SMRF_LAUNCH_CHECK();
        count_launches(1);
    }
    return 0;
}

}  // namespace smrf
