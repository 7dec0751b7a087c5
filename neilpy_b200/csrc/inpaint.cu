// Harmonic ("springs") fill of NaN cells -- inpaint_nans_by_springs, neilpy.py:1227-1271.
//
// The reference builds one spring per 4-neighbour pair that touches a NaN cell and asks
// LSQR for the least-squares displacement; the minimiser satisfies, for every NaN cell i,
//     deg(i) u_i - sum_{NaN nbrs j} u_j = sum_{known nbrs k} a_k,   deg = in-grid neighbours,
// a symmetric positive-definite system (per connected NaN region that touches a known
// cell).  Here it is solved in float64 by preconditioned conjugate gradients whose every
// vector lives in HBM; dot products are reduced on the device (block reduce + one
// atomicAdd(double) per block) into per-iteration slots, so an iteration is three
// stream-ordered kernel launches with no host round trip; the host only polls the
// residual max-norm every kCheckEvery iterations.
//
// HBM-bound stencil kernels: every launch is a 2-D grid over (row, 256-column block);
// neighbours come from L1/L2.
#include <string.h>

#include "common.cuh"

namespace smrf {
namespace inpaint {

constexpr int kMaxIter = 1 << 15;
constexpr int kCheckEvery = 16;
constexpr int kBlock = 256;

struct Scalars {          // device-resident, indexed by iteration
    double rz[kMaxIter + 2];
    double pq[kMaxIter + 2];
    unsigned long long rmax[kMaxIter + 2];   // bits of max |r| (non-negative doubles order as integers)
    double sum_known;
    unsigned long long n_known;
    unsigned long long n_unknown;
};

struct Ws {
    double *u, *r, *z, *p, *q;
    uint8_t* unk;
    Scalars* sc;
};

static inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

static Ws carve(void* workspace, int64_t n) {
    char* b = (char*)workspace;
    Ws w;
    size_t plane = align_up((size_t)n * 8);
    w.u = (double*)b; b += plane;
    w.r = (double*)b; b += plane;
    w.z = (double*)b; b += plane;
    w.p = (double*)b; b += plane;
    w.q = (double*)b; b += plane;
    w.unk = (uint8_t*)b; b += align_up((size_t)n);
    w.sc = (Scalars*)b;
    return w;
}

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double s[kBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();   // protects s across successive calls
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < kBlock / 32) t = s[threadIdx.x];
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = kBlock / 64; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;   // valid in thread 0
}
__device__ __forceinline__ double block_max(double v) {
    __shared__ double s[kBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < kBlock / 32) t = s[threadIdx.x];
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = kBlock / 64; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
    }
    return t;
}

// ---- pass 0: NaN mask, statistics of the known cells ------------------------------------
template <typename T>
__global__ void __launch_bounds__(kBlock) scan_kernel(const T* __restrict__ grid, uint8_t* __restrict__ unk, int64_t n,
                                                      Scalars* sc) {
    double s = 0.0;
    unsigned long long nk = 0, nu = 0;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
        T v = grid[i];
        bool isn = (v != v);
        unk[i] = isn ? 1 : 0;
        if (isn) ++nu;
        else { s += (double)v; ++nk; }
    }
    s = block_sum(s);
    double dk = block_sum((double)nk), du = block_sum((double)nu);
    if (threadIdx.x == 0) {
        atomicAdd(&sc->sum_known, s);
        atomicAdd(&sc->n_known, (unsigned long long)dk);
        atomicAdd(&sc->n_unknown, (unsigned long long)du);
    }
}

// ---- u = known value, or the mean of the known cells as the starting guess ---------------
template <typename T>
__global__ void __launch_bounds__(kBlock) init_u_kernel(const T* __restrict__ grid, const uint8_t* __restrict__ unk,
                                                        double* __restrict__ u, int64_t n, const Scalars* sc) {
    const double mean = sc->n_known ? sc->sum_known / (double)sc->n_known : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
        u[i] = unk[i] ? mean : (double)grid[i];
}

__device__ __forceinline__ int degree(int64_t y, int64_t x, int64_t ny, int64_t nx) {
    return (y > 0) + (y + 1 < ny) + (x > 0) + (x + 1 < nx);
}

// r = b - A u on the unknown cells (u holds the known values at known cells, so the
// right-hand side is implicit); z = r / deg; rz[0] = r.z; rmax[0] = max |r|
__global__ void __launch_bounds__(kBlock) residual0_kernel(Ws w, int64_t ny, int64_t nx) {
    double rz = 0.0, rm = 0.0;
    const int64_t x = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    for (int64_t y = blockIdx.y; y < ny; y += gridDim.y) {
        if (x < nx) {
            const int64_t i = y * nx + x;
            if (w.unk[i]) {
                double s = 0.0;
                if (y > 0) s += w.u[i - nx];
                if (y + 1 < ny) s += w.u[i + nx];
                if (x > 0) s += w.u[i - 1];
                if (x + 1 < nx) s += w.u[i + 1];
                const int d = degree(y, x, ny, nx);
                const double r = d ? s - (double)d * w.u[i] : 0.0;
                const double z = d ? r / (double)d : 0.0;
                w.r[i] = r; w.z[i] = z; w.p[i] = 0.0;
                rz += r * z; rm = fmax(rm, fabs(r));
            }
        }
    }
    rz = block_sum(rz);
    rm = block_max(rm);
    if (threadIdx.x == 0) {
        if (rz != 0.0) atomicAdd(&w.sc->rz[0], rz);
        atomicMax(&w.sc->rmax[0], (unsigned long long)__double_as_longlong(rm));
    }
}

// p = z + beta p,  beta = rz[k] / rz[k-1]
__global__ void __launch_bounds__(kBlock) p_update_kernel(Ws w, int64_t n, int k) {
    const double beta = (k == 0 || w.sc->rz[k - 1] == 0.0) ? 0.0 : w.sc->rz[k] / w.sc->rz[k - 1];
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
        if (w.unk[i]) w.p[i] = w.z[i] + beta * w.p[i];
}

// q = A p (p is zero on known cells by construction: never written there);  pq[k] = p.q
__global__ void __launch_bounds__(kBlock) apply_kernel(Ws w, int64_t ny, int64_t nx, int k) {
    double pq = 0.0;
    const int64_t x = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    for (int64_t y = blockIdx.y; y < ny; y += gridDim.y) {
        if (x < nx) {
            const int64_t i = y * nx + x;
            if (w.unk[i]) {
                double s = 0.0;
                if (y > 0 && w.unk[i - nx]) s += w.p[i - nx];
                if (y + 1 < ny && w.unk[i + nx]) s += w.p[i + nx];
                if (x > 0 && w.unk[i - 1]) s += w.p[i - 1];
                if (x + 1 < nx && w.unk[i + 1]) s += w.p[i + 1];
                const double pi = w.p[i];
                const double q = (double)degree(y, x, ny, nx) * pi - s;
                w.q[i] = q;
                pq += pi * q;
            }
        }
    }
    pq = block_sum(pq);
    if (threadIdx.x == 0 && pq != 0.0) atomicAdd(&w.sc->pq[k], pq);
}

// u += alpha p; r -= alpha q; z = r / deg; rz[k+1] = r.z; rmax[k+1] = max |r|
__global__ void __launch_bounds__(kBlock) update_kernel(Ws w, int64_t ny, int64_t nx, int k) {
    const double pqk = w.sc->pq[k];
    const double alpha = pqk != 0.0 ? w.sc->rz[k] / pqk : 0.0;
    double rz = 0.0, rm = 0.0;
    const int64_t x = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    for (int64_t y = blockIdx.y; y < ny; y += gridDim.y) {
        if (x < nx) {
            const int64_t i = y * nx + x;
            if (w.unk[i]) {
                w.u[i] += alpha * w.p[i];
                const double r = w.r[i] - alpha * w.q[i];
                const int d = degree(y, x, ny, nx);
                const double z = d ? r / (double)d : 0.0;
                w.r[i] = r; w.z[i] = z;
                rz += r * z; rm = fmax(rm, fabs(r));
            }
        }
    }
    rz = block_sum(rz);
    rm = block_max(rm);
    if (threadIdx.x == 0) {
        if (rz != 0.0) atomicAdd(&w.sc->rz[k + 1], rz);
        atomicMax(&w.sc->rmax[k + 1], (unsigned long long)__double_as_longlong(rm));
    }
}

template <typename T>
__global__ void __launch_bounds__(kBlock) writeback_kernel(T* __restrict__ grid, const uint8_t* __restrict__ unk,
                                                           const double* __restrict__ u, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
        if (unk[i]) grid[i] = (T)u[i];
}

}  // namespace inpaint
}  // namespace smrf

using namespace smrf;
using namespace smrf::inpaint;

extern "C" {

size_t smrf_inpaint_workspace_bytes(int64_t ny, int64_t nx) {
    size_t n = (size_t)ny * (size_t)nx;
    return 5 * align_up(n * 8) + align_up(n) + align_up(sizeof(Scalars));
}

int smrf_inpaint(void* grid, int64_t ny, int64_t nx, int dtype, uint8_t* unknown, void* workspace,
                 size_t workspace_bytes, double tol, int max_iter, double* info_host, void* stream) {
    SMRF_CHECK_ARG(grid && workspace, "null pointer");
    SMRF_CHECK_ARG(ny > 0 && nx > 0, "empty grid");
    SMRF_CHECK_ARG(dtype == SMRF_F32 || dtype == SMRF_F64, "bad dtype");
    SMRF_CHECK_ARG(tol >= 0.0, "negative tol");
    if (workspace_bytes < smrf_inpaint_workspace_bytes(ny, nx)) {
        set_error("smrf_inpaint: workspace %zu < %zu bytes", workspace_bytes, smrf_inpaint_workspace_bytes(ny, nx));
        return SMRF_E_WORKSPACE;
    }
    if (max_iter <= 0 || max_iter > kMaxIter) max_iter = kMaxIter;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = ny * nx;
    Ws w = carve(workspace, n);
    SMRF_CUDA(cudaMemsetAsync(w.sc, 0, sizeof(Scalars), st));

    int g1 = (int)((n + kBlock - 1) / kBlock);
    int cap = num_sms() * 16;
    if (g1 > cap) g1 = cap;
    dim3 g2((unsigned)((nx + kBlock - 1) / kBlock), (unsigned)(ny < 32768 ? ny : 32768));

    if (dtype == SMRF_F32) scan_kernel<float><<<g1, kBlock, 0, st>>>((const float*)grid, w.unk, n, w.sc);
    else scan_kernel<double><<<g1, kBlock, 0, st>>>((const double*)grid, w.unk, n, w.sc);
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    unsigned long long counts[2];
    SMRF_CUDA(cudaMemcpyAsync(counts, &w.sc->n_known, sizeof(counts), cudaMemcpyDeviceToHost, st));
    SMRF_CUDA(cudaStreamSynchronize(st));
    const unsigned long long n_unknown = counts[1];
    if (unknown) SMRF_CUDA(cudaMemcpyAsync(unknown, w.unk, (size_t)n, cudaMemcpyDeviceToDevice, st));
    int it = 0;
    double rmax = 0.0;
    if (n_unknown > 0) {
        if (dtype == SMRF_F32) init_u_kernel<float><<<g1, kBlock, 0, st>>>((const float*)grid, w.unk, w.u, n, w.sc);
        else init_u_kernel<double><<<g1, kBlock, 0, st>>>((const double*)grid, w.unk, w.u, n, w.sc);
        residual0_kernel<<<g2, kBlock, 0, st>>>(w, ny, nx);
        SMRF_LAUNCH_CHECK();
        count_launches(2);
        unsigned long long bits = 0;
        SMRF_CUDA(cudaMemcpyAsync(&bits, &w.sc->rmax[0], 8, cudaMemcpyDeviceToHost, st));
        SMRF_CUDA(cudaStreamSynchronize(st));
        memcpy(&rmax, &bits, 8);
        while (rmax > tol && it < max_iter) {
            int burst = kCheckEvery;
            if (it + burst > max_iter) burst = max_iter - it;
            for (int j = 0; j < burst; ++j, ++it) {
                p_update_kernel<<<g1, kBlock, 0, st>>>(w, n, it);
                apply_kernel<<<g2, kBlock, 0, st>>>(w, ny, nx, it);
                update_kernel<<<g2, kBlock, 0, st>>>(w, ny, nx, it);
            }
            SMRF_LAUNCH_CHECK();
            count_launches(3 * burst);
            SMRF_CUDA(cudaMemcpyAsync(&bits, &w.sc->rmax[it], 8, cudaMemcpyDeviceToHost, st));
            SMRF_CUDA(cudaStreamSynchronize(st));
            memcpy(&rmax, &bits, 8);
            if (!(rmax == rmax)) break;   // NaN: give up rather than spin
        }
        if (dtype == SMRF_F32) writeback_kernel<float><<<g1, kBlock, 0, st>>>((float*)grid, w.unk, w.u, n);
        else writeback_kernel<double><<<g1, kBlock, 0, st>>>((double*)grid, w.unk, w.u, n);
        SMRF_LAUNCH_CHECK();
        count_launches(1);
        SMRF_CUDA(cudaStreamSynchronize(st));
    }
    if (info_host) {
        info_host[0] = (double)it;
        info_host[1] = rmax;
        info_host[2] = (double)n_unknown;
    }
    return 0;
}

}  // extern "C"
