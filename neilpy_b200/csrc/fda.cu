// inpaint_nans_by_fda -- neilpy.py:1171-1216 ("finite difference approximation", D'Errico's method 0/1).
//
// The reference writes one equation per grid cell i whose stencil touches a NaN cell,
//     V(i) (u[up] + u[down] - 2 u[i]) + H(i) (u[left] + u[right] - 2 u[i]) = 0,
// V(i) = 1 on rows 1..m-2 (second difference along the columns), H(i) = 1 on columns 1..n-2, moves the known cells
// to the right-hand side and asks LSQR for the least-squares values of the NaN cells.  Its row selection
// (`fda[:, nan].nonzero()[0]`, neilpy.py:1206-1208) lists a row once per NaN cell in its stencil, so equation i enters
// the least-squares problem w_i times, w_i = number of NaN cells its stencil touches: a WEIGHTED problem, reproduced
// here.  The minimiser solves the normal equations A^T W A x = A^T W b with A = the operator L above restricted to the
// rows with w > 0 and to the NaN columns -- a
// fourth-order (squared-Laplacian-like) system, not the 5-point system of inpaint_nans_by_springs, so it has its own
// solver here: CGLS (conjugate gradients on the normal equations without forming them; the Krylov method LSQR is built
// on), every vector in HBM, float64, started from zero like LSQR (so a rank-deficient system gets the same minimum-norm
// answer).  Per iteration: q = A p (one stencil pass), s = A^T r (one stencil pass), two dot products reduced on the
// device into per-iteration slots; the host polls max |A^T r| every kPoll iterations.
// This is the step-after-the-path helper of SURVEY.md 8(f) rank 4; it is not on the SMRF path and is not tuned.
#include <string.h>

#include "common.cuh"

namespace smrf {
namespace fda {

constexpr int kBlock = 256;
constexpr int kPoll = 64;
constexpr int kSlots = 1 << 16;

struct Scalars {
    double gamma[kSlots + 1];    // |s_k|^2
    double qq[kSlots + 1];       // q_k . W q_k
    unsigned long long smax[kSlots + 1];   // bits of max |s_k|
    unsigned long long n_unknown;
};

struct Ws {
    double *x, *r, *p, *q, *s;
    uint8_t *unk, *row;          // NaN mask; w_i = how many NaN cells the stencil of equation i touches (0: not kept)
    Scalars* sc;
};

static inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

static size_t carve(void* ws, int64_t n, Ws* w) {
    char* b = (char*)ws;
    Ws t;
    const size_t plane = up256((size_t)n * 8);
    t.x = (double*)b; b += plane;
    t.r = (double*)b; b += plane;
    t.p = (double*)b; b += plane;
    t.q = (double*)b; b += plane;
    t.s = (double*)b; b += plane;
    t.unk = (uint8_t*)b; b += up256((size_t)n);
    t.row = (uint8_t*)b; b += up256((size_t)n);
    t.sc = (Scalars*)b; b += up256(sizeof(Scalars));
    if (w) *w = t;
    return (size_t)(b - (char*)ws);
}

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sh[kBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = threadIdx.x < kBlock / 32 ? sh[threadIdx.x] : 0.0;
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = kBlock / 64; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;
}
__device__ __forceinline__ double block_max(double v) {
    __shared__ double sh[kBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = threadIdx.x < kBlock / 32 ? sh[threadIdx.x] : 0.0;
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = kBlock / 64; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
    }
    return t;
}

// L v at cell (y, x):  V (v[up] + v[down] - 2 v) + H (v[left] + v[right] - 2 v)
__device__ __forceinline__ double apply_L(const double* __restrict__ v, int64_t y, int64_t x, int64_t ny, int64_t nx) {
    const int64_t i = y * nx + x;
    double a = 0.0;
    if (y >= 1 && y <= ny - 2) a += v[i - nx] + v[i + nx] - 2.0 * v[i];
    if (x >= 1 && x <= nx - 2) a += v[i - 1] + v[i + 1] - 2.0 * v[i];
    return a;
}
// (L^T W r) at cell (y, x): r is zero outside the kept rows, wt holds the weights
__device__ __forceinline__ double apply_Lt(const double* __restrict__ r, const uint8_t* __restrict__ wt, int64_t y, int64_t x,
                                           int64_t ny, int64_t nx) {
    const int64_t i = y * nx + x;
    double a = 0.0;
    // vertical parts of the equations of (y-1, x), (y+1, x) and (y, x) itself
    if (y - 1 >= 1 && y - 1 <= ny - 2) a += (double)wt[i - nx] * r[i - nx];
    if (y + 1 >= 1 && y + 1 <= ny - 2) a += (double)wt[i + nx] * r[i + nx];
    if (y >= 1 && y <= ny - 2) a -= 2.0 * (double)wt[i] * r[i];
    if (x - 1 >= 1 && x - 1 <= nx - 2) a += (double)wt[i - 1] * r[i - 1];
    if (x + 1 >= 1 && x + 1 <= nx - 2) a += (double)wt[i + 1] * r[i + 1];
    if (x >= 1 && x <= nx - 2) a -= 2.0 * (double)wt[i] * r[i];
    return a;
}

// NaN mask; x = 0; p (scratch) = the known part of the grid (NaN -> 0)
template <typename T>
__global__ void __launch_bounds__(kBlock) scan_kernel(const T* __restrict__ grid, Ws w, int64_t n) {
    unsigned long long nu = 0;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
        const T v = grid[i];
        const bool isn = v != v;
        w.unk[i] = isn ? 1 : 0;
        w.p[i] = isn ? 0.0 : (double)v;
        w.x[i] = 0.0;
        nu += isn ? 1 : 0;
    }
    const double t = block_sum((double)nu);
    if (threadIdx.x == 0 && t != 0.0) atomicAdd(&w.sc->n_unknown, (unsigned long long)t);
}

// kept rows, their weights and the right-hand side: row[i] = NaN cells in the stencil of i; r = b = -L(known part) where > 0
__global__ void __launch_bounds__(kBlock) rows_kernel(Ws w, int64_t ny, int64_t nx) {
    const int64_t n = ny * nx;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
        const int64_t y = i / nx, x = i - y * nx;
        const bool V = y >= 1 && y <= ny - 2, H = x >= 1 && x <= nx - 2;
        int touch = 0;
        if (V || H) touch = w.unk[i];
        if (V) touch += w.unk[i - nx] + w.unk[i + nx];
        if (H) touch += w.unk[i - 1] + w.unk[i + 1];
        w.row[i] = (uint8_t)touch;
        w.r[i] = touch ? -apply_L(w.p, y, x, ny, nx) : 0.0;
    }
}

// s = A^T W r on the NaN cells (0 elsewhere); gamma[k] = |s|^2, smax[k] = max |s|; FIRST: p = s
template <bool FIRST>
__global__ void __launch_bounds__(kBlock) at_kernel(Ws w, int64_t ny, int64_t nx, int k) {
    const int64_t n = ny * nx;
    double g = 0.0, mx = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
        double v = 0.0;
        if (w.unk[i]) {
            const int64_t y = i / nx, x = i - y * nx;
            v = apply_Lt(w.r, w.row, y, x, ny, nx);
        }
        w.s[i] = v;
        if (FIRST) w.p[i] = v;
        g += v * v;
        const double a = fabs(v);
        mx = (a < INFINITY) ? fmax(mx, a) : INFINITY;
    }
    g = block_sum(g);
    mx = block_max(mx);
    if (threadIdx.x == 0) {
        if (g != 0.0) atomicAdd(&w.sc->gamma[k], g);
        atomicMax(&w.sc->smax[k], (unsigned long long)__double_as_longlong(mx));
    }
}

// p = s + (gamma[k] / gamma[k-1]) p   (k >= 1), then q = A p on the kept rows, qq[k] = |q|^2 -- two kernels: q needs
// the neighbours' new p.  qq[k] = q . W q
__global__ void __launch_bounds__(kBlock) p_kernel(Ws w, int64_t n, int k) {
    const double g0 = w.sc->gamma[k - 1];
    const double beta = g0 != 0.0 ? w.sc->gamma[k] / g0 : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
        w.p[i] = w.s[i] + beta * w.p[i];
}
__global__ void __launch_bounds__(kBlock) a_kernel(Ws w, int64_t ny, int64_t nx, int k) {
    const int64_t n = ny * nx;
    double qq = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
        double v = 0.0;
        if (w.row[i]) {
            const int64_t y = i / nx, x = i - y * nx;
            v = apply_L(w.p, y, x, ny, nx);
        }
        w.q[i] = v;
        qq += (double)w.row[i] * v * v;
    }
    qq = block_sum(qq);
    if (threadIdx.x == 0 && qq != 0.0) atomicAdd(&w.sc->qq[k], qq);
}
// x += alpha p, r -= alpha q, alpha = gamma[k] / qq[k]
__global__ void __launch_bounds__(kBlock) update_kernel(Ws w, int64_t n, int k) {
    const double qq = w.sc->qq[k];
    const double alpha = qq != 0.0 ? w.sc->gamma[k] / qq : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
        w.x[i] += alpha * w.p[i];
        w.r[i] -= alpha * w.q[i];
    }
}
template <typename T>
__global__ void __launch_bounds__(kBlock) writeback_kernel(T* __restrict__ grid, Ws w, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
        if (w.unk[i]) grid[i] = (T)w.x[i];
}

}  // namespace fda
}  // namespace smrf

using namespace smrf;
using namespace smrf::fda;

static int grid_for(int64_t n) {
    int64_t g = (n + kBlock - 1) / kBlock;
    const int64_t cap = (int64_t)num_sms() * 16;
    return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

extern "C" {

size_t smrf_inpaint_fda_workspace_bytes(int64_t ny, int64_t nx) {
    if (ny <= 0 || nx <= 0) return 0;
    return carve(nullptr, ny * nx, nullptr);
}

int smrf_inpaint_fda(void* grid, int64_t ny, int64_t nx, int dtype, void* workspace, size_t workspace_bytes, double tol,
                     int max_iter, double* info_host, void* stream) {
    SMRF_CHECK_ARG(grid && workspace, "null pointer");
    SMRF_CHECK_ARG(ny > 0 && nx > 0, "empty grid");
    SMRF_CHECK_ARG(dtype == SMRF_F32 || dtype == SMRF_F64, "bad dtype");
    SMRF_CHECK_ARG(tol >= 0.0, "negative tol");
    const int64_t n = ny * nx;
    if (workspace_bytes < carve(nullptr, n, nullptr)) {
        set_error("smrf_inpaint_fda: workspace %zu < %zu bytes", workspace_bytes, carve(nullptr, n, nullptr));
        return SMRF_E_WORKSPACE;
    }
    if (max_iter <= 0 || max_iter > kSlots - 1) max_iter = kSlots - 1;
    Ws w;
    carve(workspace, n, &w);
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(n);
    SMRF_CUDA(cudaMemsetAsync(w.sc, 0, sizeof(Scalars), st));
    if (dtype == SMRF_F32) scan_kernel<float><<<g, kBlock, 0, st>>>((const float*)grid, w, n);
    else scan_kernel<double><<<g, kBlock, 0, st>>>((const double*)grid, w, n);
    rows_kernel<<<g, kBlock, 0, st>>>(w, ny, nx);
    at_kernel<true><<<g, kBlock, 0, st>>>(w, ny, nx, 0);       // s_0 = A^T b, p_0 = s_0
    int launches = 3;
    unsigned long long nu = 0, bits = 0;
    SMRF_CUDA(cudaMemcpyAsync(&nu, &w.sc->n_unknown, 8, cudaMemcpyDeviceToHost, st));
    SMRF_CUDA(cudaMemcpyAsync(&bits, &w.sc->smax[0], 8, cudaMemcpyDeviceToHost, st));
    SMRF_CUDA(cudaStreamSynchronize(st));
    double smax = 0.0;
    memcpy(&smax, &bits, 8);
    int it = 0;
    if (nu > 0) {
        while (smax > tol && it < max_iter) {
            int burst = kPoll;
            if (it + burst > max_iter) burst = max_iter - it;
            for (int j = 0; j < burst; ++j, ++it) {
                const int k = it;
                if (k > 0) p_kernel<<<g, kBlock, 0, st>>>(w, n, k);
                a_kernel<<<g, kBlock, 0, st>>>(w, ny, nx, k);
                update_kernel<<<g, kBlock, 0, st>>>(w, n, k);
                at_kernel<false><<<g, kBlock, 0, st>>>(w, ny, nx, k + 1);
                launches += 4;
            }
            SMRF_LAUNCH_CHECK();
            SMRF_CUDA(cudaMemcpyAsync(&bits, &w.sc->smax[it], 8, cudaMemcpyDeviceToHost, st));
            SMRF_CUDA(cudaStreamSynchronize(st));
            memcpy(&smax, &bits, 8);
            if (!(smax < INFINITY)) break;
        }
        if (dtype == SMRF_F32) writeback_kernel<float><<<g, kBlock, 0, st>>>((float*)grid, w, n);
        else writeback_kernel<double><<<g, kBlock, 0, st>>>((double*)grid, w, n);
        ++launches;
        SMRF_LAUNCH_CHECK();
        SMRF_CUDA(cudaStreamSynchronize(st));
    }
    count_launches(launches);
    if (info_host) {
        info_host[0] = (double)it;
        info_host[1] = smax;
        info_host[2] = (double)nu;
    }
    return 0;
}

}  // extern "C"
