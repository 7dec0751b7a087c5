// Generic (any radius, any dtype) disk erosion / dilation and the progressive-filter
// driver (progressive_filter, neilpy.py:1659-1680).
//
// The kernels in this file evaluate the disk footprint chord by chord straight from
// global memory (L1/L2-cached): |disk(w)| loads per cell.  They are the fallback for
// radii the register-marching kernels (opening_march.cu) are not instantiated for,
// and the in-library cross-check of those kernels.
#include "common.cuh"
#include "opening.cuh"

namespace smrf {

__device__ __forceinline__ int chord_half(int w, int dy) {
    int rem = w * w - dy * dy;
    int h = (int)sqrtf((float)rem);
    while (h * h > rem) --h;
    while ((h + 1) * (h + 1) <= rem) ++h;
    return h;
}

template <typename T, bool IS_MAX>
__device__ __forceinline__ T pick(T a, T b) {
    // fmin/fmax return the non-NaN operand: NaN acts as "no sample"
    if (sizeof(T) == 4) return IS_MAX ? (T)fmaxf((float)a, (float)b) : (T)fminf((float)a, (float)b);
    return IS_MAX ? (T)fmax((double)a, (double)b) : (T)fmin((double)a, (double)b);
}

// out[y][x] = min/max over disk(w) of in (negated on load if `negate`), ignoring
// samples outside the image.  grid: x fastest.
template <typename T, bool IS_MAX>
__global__ void __launch_bounds__(256) morph_direct_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t ny,
                                                           int64_t nx, int64_t pitch, int w, int negate,
                                                           int64_t row_lo, int64_t row_hi) {
    int64_t x = (int64_t)blockIdx.x * 64 + (threadIdx.x & 63);
    int64_t y = row_lo + (int64_t)blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= nx || y >= row_hi) return;
    T acc = IS_MAX ? (T)-INFINITY : (T)INFINITY;
    for (int dy = -w; dy <= w; ++dy) {
        int64_t yy = y + dy;
        if (yy < 0 || yy >= ny) continue;
        int h = chord_half(w, dy);
        int64_t x0 = x - h < 0 ? 0 : x - h;
        int64_t x1 = x + h >= nx ? nx - 1 : x + h;
        const T* row = in + yy * pitch;
        for (int64_t xx = x0; xx <= x1; ++xx) {
            T v = __ldg(row + xx);
            if (negate) v = -v;
            acc = pick<T, IS_MAX>(acc, v);
        }
    }
    out[y * pitch + x] = acc;
}

// new_obj = (last - this) > thr in float64; mask |= new_obj; when[new_obj] = widx
template <typename T>
__global__ void __launch_bounds__(256) threshold_kernel(const T* __restrict__ last, const T* __restrict__ cur,
                                                        uint8_t* __restrict__ mask, uint8_t* __restrict__ when,
                                                        int64_t nx, int64_t pitch, double thr, int widx, int negate,
                                                        int64_t row_lo, int64_t row_hi) {
    int64_t n0 = row_lo * nx, n1 = row_hi * nx;
    for (int64_t i = n0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t y = i / nx, g = y * pitch + (i - y * nx);
        double l = (double)last[g];
        if (negate) l = -l;
        bool obj = __dsub_rn(l, (double)cur[g]) > thr;
        if (obj) {
            mask[i] = 1;
            if (when) when[i] = (uint8_t)widx;
        }
    }
}

int open_window_generic(const void* in, void* out, void* tmp, uint8_t* mask, uint8_t* when, int64_t ny, int64_t nx,
                        int64_t pitch, int dtype, int w, double thr, int widx, int negate, int64_t row_lo,
                        int64_t row_hi, cudaStream_t st) {
    // erosion is needed on rows [row_lo - w, row_hi + w) of the image
    int64_t e_lo = row_lo - w < 0 ? 0 : row_lo - w;
    int64_t e_hi = row_hi + w > ny ? ny : row_hi + w;
    dim3 blk(256);
    dim3 ge((unsigned)((nx + 63) / 64), (unsigned)((e_hi - e_lo + 3) / 4));
    dim3 gd((unsigned)((nx + 63) / 64), (unsigned)((row_hi - row_lo + 3) / 4));
    int gt = (int)(((row_hi - row_lo) * nx + 255) / 256);
    int cap = num_sms() * 16;
    if (gt > cap) gt = cap;
    if (gt < 1) gt = 1;
    if (dtype == SMRF_F32) {
        morph_direct_kernel<float, false><<<ge, blk, 0, st>>>((const float*)in, (float*)tmp, ny, nx, pitch, w, negate, e_lo, e_hi);
        // rows of tmp outside [e_lo, e_hi) are never read: the dilation clips to the image,
        // and rows between the image border and e_lo/e_hi do not exist when e_lo/e_hi clip.
        morph_direct_kernel<float, true><<<gd, blk, 0, st>>>((const float*)tmp, (float*)out, ny, nx, pitch, w, 0, row_lo, row_hi);
        if (mask) threshold_kernel<float><<<gt, 256, 0, st>>>((const float*)in, (const float*)out, mask, when, nx, pitch, thr, widx, negate, row_lo, row_hi);
    } else {
        morph_direct_kernel<double, false><<<ge, blk, 0, st>>>((const double*)in, (double*)tmp, ny, nx, pitch, w, negate, e_lo, e_hi);
        morph_direct_kernel<double, true><<<gd, blk, 0, st>>>((const double*)tmp, (double*)out, ny, nx, pitch, w, 0, row_lo, row_hi);
        if (mask) threshold_kernel<double><<<gt, 256, 0, st>>>((const double*)in, (const double*)out, mask, when, nx, pitch, thr, widx, negate, row_lo, row_hi);
    }
    SMRF_LAUNCH_CHECK();
    count_launches(mask ? 3 : 2);
    return 0;
}

}  // namespace smrf

using namespace smrf;

namespace smrf {
size_t open_f64_workspace_bytes(int64_t ny, int64_t nx);
int open_f64_rank(const double* surface, int64_t ny, int64_t nx, int64_t pitch_in, void* ws, size_t ws_bytes, uint8_t* mask,
                  uint8_t* when, const int32_t* windows, const double* thr, int n_windows, int window_index0, int advance,
                  int64_t row_lo, int64_t row_hi, double* last_out, int64_t out_pitch, cudaStream_t st);
}

// float64 surfaces run in rank space (rank.cu) unless every radius is so small that |disk| loads per cell are cheaper
static bool use_rank_space(int dtype, const int32_t* windows, int n, int negate) {
    if (dtype != SMRF_F64 || negate || n <= 0 || open_force_generic()) return false;
    int mx = 0;
    for (int i = 0; i < n; ++i) {
        if (windows[i] < 1 || windows[i] > SMRF_MARCH_MAX_W) return false;
        if (windows[i] > mx) mx = windows[i];
    }
    return mx >= 3;
}

static inline int64_t aligned_pitch(int64_t nx, int dtype) {
    const int64_t q = dtype == SMRF_F64 ? 2 : 4;   // 16-byte rows
    return (nx + q - 1) / q * q;
}

extern "C" {

size_t smrf_open_workspace_bytes(int64_t ny, int64_t nx, int dtype, int max_window) {
    (void)max_window;
    size_t es = dtype == SMRF_F64 ? 8 : 4;
    size_t plane = ((size_t)ny * (size_t)aligned_pitch(nx, dtype) * es + 255) & ~(size_t)255;
    size_t need = 3 * plane;   // two ping-pong surfaces + the erosion intermediate (generic path, two-pass radii)
    if (dtype == SMRF_F64) {   // rank space: sorted table, two float32 rank planes, sort scratch
        size_t r = open_f64_workspace_bytes(ny, nx);
        if (r > need) need = r;
    }
    return need;
}

const char* smrf_open_variant(int dtype, int window) {
    return open_march_available(dtype, window, 0) ? open_march_name(dtype, window) : "direct_generic";
}

int smrf_open_window(const void* in, void* out, void* tmp, uint8_t* mask, uint8_t* when_dropped, int64_t ny,
                     int64_t nx, int64_t pitch, int dtype, int window, double threshold, int window_index,
                     int negate, int64_t row_lo, int64_t row_hi, void* stream) {
    SMRF_CHECK_ARG(in && out && tmp, "null pointer");
    SMRF_CHECK_ARG(in != out && in != tmp && out != tmp, "in/out/tmp must be distinct");
    SMRF_CHECK_ARG(ny > 0 && nx > 0 && pitch >= nx, "bad grid size / pitch");
    SMRF_CHECK_ARG(dtype == SMRF_F32 || dtype == SMRF_F64, "bad dtype");
    SMRF_CHECK_ARG(window >= 0 && window <= 4096, "bad window radius");
    SMRF_CHECK_ARG(0 <= row_lo && row_lo <= row_hi && row_hi <= ny, "bad row range");
    SMRF_CHECK_ARG(!when_dropped || mask, "when_dropped needs mask");
    if (row_lo == row_hi) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (use_rank_space(dtype, &window, 1, negate)) {
        // float64, one window: rank space with stream-ordered scratch (the one entry point that allocates;
        // smrf_progressive_open carries its own workspace)
        const size_t need = open_f64_workspace_bytes(ny, nx);
        void* scratch = nullptr;
        SMRF_CUDA(cudaMallocAsync(&scratch, need, st));
        int rc = open_f64_rank((const double*)in, ny, nx, pitch, scratch, need, mask, when_dropped, &window, &threshold, 1,
                               window_index, 0, row_lo, row_hi, (double*)out, pitch, st);
        cudaFreeAsync(scratch, st);
        return rc;
    }
    if (open_march_available(dtype, window, negate) && !open_force_generic())
        return open_window_march(in, out, tmp, mask, when_dropped, ny, nx, pitch, dtype, window, threshold, window_index,
                                 negate, row_lo, row_hi, st);
    return open_window_generic(in, out, tmp, mask, when_dropped, ny, nx, pitch, dtype, window, threshold,
                               window_index, negate, row_lo, row_hi, st);
}

int smrf_open_window_bruteforce(const void* in, void* out, void* tmp, int64_t ny, int64_t nx, int dtype, int window,
                                void* stream) {
    SMRF_CHECK_ARG(in && out && tmp, "null pointer");
    SMRF_CHECK_ARG(ny > 0 && nx > 0, "empty grid");
    SMRF_CHECK_ARG(dtype == SMRF_F32 || dtype == SMRF_F64, "bad dtype");
    SMRF_CHECK_ARG(window >= 0 && window <= 4096, "bad window radius");
    return open_window_generic(in, out, tmp, nullptr, nullptr, ny, nx, nx, dtype, window, 0.0, 0, 0, 0, ny,
                               (cudaStream_t)stream);
}

int smrf_progressive_open(const void* surface, void* workspace, size_t workspace_bytes, uint8_t* mask,
                          uint8_t* when_dropped, int64_t ny, int64_t nx, int dtype, const int32_t* windows_host,
                          const double* thresholds_host, int n_windows, int negate, void* last_out, void* stream) {
    SMRF_CHECK_ARG(surface && workspace && mask && windows_host && thresholds_host, "null pointer");
    SMRF_CHECK_ARG(ny > 0 && nx > 0 && n_windows >= 0, "bad size");
    SMRF_CHECK_ARG(dtype == SMRF_F32 || dtype == SMRF_F64, "bad dtype");
    SMRF_CHECK_ARG(!(negate && n_windows > 1), "negate supports a single window (the low-outlier pass)");
    size_t need = smrf_open_workspace_bytes(ny, nx, dtype, 0);
    if (workspace_bytes < need) {
        set_error("smrf_progressive_open: workspace %zu < %zu bytes", workspace_bytes, need);
        return SMRF_E_WORKSPACE;
    }
    if (n_windows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (use_rank_space(dtype, windows_host, n_windows, negate))
        return open_f64_rank((const double*)surface, ny, nx, nx, workspace, workspace_bytes, mask, when_dropped, windows_host,
                             thresholds_host, n_windows, 0, n_windows > 1, 0, ny, (double*)last_out, nx, st);
    const size_t es = dtype == SMRF_F64 ? 8 : 4;
    const size_t plane = (((size_t)ny * (size_t)aligned_pitch(nx, dtype) * es + 255) & ~(size_t)255);
    // The windows ping-pong between two workspace surfaces whose rows are padded to 16 bytes, so the
    // marching kernels take their vector paths whatever nx is; the caller's surface is never written.
    const int64_t pitch = aligned_pitch(nx, dtype);
    char* a = (char*)workspace;
    char* b = (char*)workspace + plane;
    char* tmp = (char*)workspace + 2 * plane;
    const char* cur = (const char*)surface;
    int64_t cur_pitch = nx;
    char* nxt = a;
    if (pitch != nx && n_windows > 1) {   // one strided copy buys aligned rows for every window
        SMRF_CUDA(cudaMemcpy2DAsync(b, (size_t)pitch * es, surface, (size_t)nx * es, (size_t)nx * es, (size_t)ny,
                                    cudaMemcpyDeviceToDevice, st));
        cur = b;
        cur_pitch = pitch;
    }
    for (int i = 0; i < n_windows; ++i) {
        int rc;
        if (cur_pitch == pitch) {
            rc = smrf_open_window(cur, nxt, tmp, mask, when_dropped, ny, nx, pitch, dtype, windows_host[i],
                                  thresholds_host[i], i, negate, 0, ny, stream);
        } else {
            // a single window on an unpadded surface: run it in place of the copy (output rows unpadded too)
            rc = smrf_open_window(cur, nxt, tmp, mask, when_dropped, ny, nx, nx, dtype, windows_host[i],
                                  thresholds_host[i], i, negate, 0, ny, stream);
        }
        if (rc) return rc;
        // neilpy.py:1675-1676: last_surface only advances when there is more than one window
        if (n_windows > 1) {
            cur = nxt;
            nxt = (nxt == a) ? b : a;
        }
    }
    if (last_out) {
        const char* src = n_windows > 1 ? cur : nxt;
        const int64_t sp = (n_windows > 1 || cur_pitch == pitch) ? cur_pitch : nx;
        SMRF_CUDA(cudaMemcpy2DAsync(last_out, (size_t)nx * es, src, (size_t)sp * es, (size_t)nx * es, (size_t)ny,
                                    cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

}  // extern "C"
