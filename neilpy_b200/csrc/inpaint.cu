// Harmonic ("springs") fill of NaN cells -- inpaint_nans_by_springs, neilpy.py:1227-1271.
//
// The reference builds one spring per 4-neighbour pair that touches a NaN cell and asks
// LSQR for the least-squares displacement; the minimiser satisfies, for every NaN cell i,
//     deg(i) u_i - sum_{NaN nbrs j} u_j = sum_{known nbrs k} a_k,   deg = in-grid neighbours,
// a symmetric positive-definite system (per connected NaN region that touches a known
// cell).  Here it is solved by conjugate gradients in float64 whose every vector lives in
// HBM, preconditioned by one geometric multigrid V(3,3)-cycle per iteration:
//   - cell-centred 2x2 coarsening; a coarse cell is unknown only if all its children are
//     (the coarse domains shrink, so Dirichlet data never leaks into a correction);
//   - Jacobi smoothing with Chebyshev weights (three sweeps per leg), piecewise-constant prolongation and its transpose as the
//     restriction, the 5-point operator rediscretised on every level: the cycle is a
//     symmetric positive-definite operator, as CG needs;
//   - the cycle runs in float32 (it only steers the search direction; the residual
//     recurrence that decides convergence is float64).
// Dot products are reduced on the device (block reduce + one atomicAdd(double) per block)
// into per-iteration slots, so an iteration is a fixed sequence of stream-ordered launches
// with no host round trip; the host polls the residual max-norm every kCheckEvery iterations.
// Every kernel is an HBM-bound stencil pass over (row, 256-column) tiles; neighbours come
// from L1/L2.  SMRF_INPAINT_PRECOND=jacobi switches the V-cycle off (diagnostics).
#include <stdlib.h>
#include <string.h>

#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace smrf {
namespace inpaint {

constexpr int kMaxIter = 1 << 15;
constexpr int kCheckEvery = 8;
constexpr int kBlock = 256;
constexpr int kMaxLevels = 20;
constexpr float kOmega = 0.8f;     // damped Jacobi on the coarsest level
// The three sweeps of a V-cycle leg use the degree-3 Chebyshev weights for the high-frequency band
// [0.5, 2] of D^-1 A (smoothing factor 0.074 per leg; two sweeps with 0.8 give 0.36); the up leg
// applies them in reverse order, which keeps the cycle symmetric.
constexpr float kOmegaA = 1.6653f, kOmegaB = 0.8f, kOmegaC = 0.5265f;
constexpr int kCoarsestSweeps = 8;   // even: the coarsest result lands in Level::y

struct Scalars {          // device-resident, indexed by iteration
    double rz[kMaxIter + 2];
    double pq[kMaxIter + 2];
    unsigned long long rmax[kMaxIter + 2];   // bits of max |r| (non-negative doubles order as integers)
    double sum_known;
    unsigned long long n_known;
    unsigned long long n_unknown;
};

struct Level {
    int64_t ny, nx;
    uint8_t* m;       // 1 = unknown on this level
    float *x, *y, *b; // two iterates (ping-pong) and the right-hand side
};

struct Ws {
    double *u, *r, *p, *q;
    int *pos, *idx;          // compact CG: pos[cell] = index of the unknown (-1: known), idx[k] = cell of unknown k
    void* scan_tmp;
    size_t scan_tmp_bytes;
    Scalars* sc;
    int nlev;
    // row-band sharding: does a neighbouring band exist above row 0 / below row ny-1?  Its cells
    // count in the degree; inside the preconditioner they are zero-correction (block Jacobi).
    int has_above, has_below;
    Level lev[kMaxLevels];
};

static inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

static int level_dims(int64_t ny, int64_t nx, int64_t* lny, int64_t* lnx) {
    int n = 0;
    while (n < kMaxLevels) {
        lny[n] = ny; lnx[n] = nx; ++n;
        if (ny <= 2 && nx <= 2) break;
        ny = (ny + 1) / 2; nx = (nx + 1) / 2;
    }
    return n;
}

// scratch of cub::DeviceScan over n flags; a generous bound when no device is present (size queries on a CPU box)
static size_t scan_temp_bytes(size_t n) {
    if (n >= ((size_t)1 << 31)) return 0;        // the compact path is not used for such grids
    size_t bytes = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const uint8_t*)nullptr, (int*)nullptr, (int)n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        bytes = n / 256 + ((size_t)1 << 20);
    }
    return bytes + 4096;
}

static size_t carve(void* workspace, int64_t ny, int64_t nx, Ws* w) {
    char* b = (char*)workspace;
    const size_t n = (size_t)ny * (size_t)nx;
    const size_t plane = align_up(n * 8);
    Ws t;
    t.has_above = t.has_below = 0;
    t.u = (double*)b; b += plane;
    t.r = (double*)b; b += plane;
    t.p = (double*)b; b += plane;
    t.q = (double*)b; b += plane;
    int64_t lny[kMaxLevels], lnx[kMaxLevels];
    t.nlev = level_dims(ny, nx, lny, lnx);
    for (int l = 0; l < t.nlev; ++l) {
        const size_t nl = (size_t)lny[l] * (size_t)lnx[l];
        t.lev[l].ny = lny[l]; t.lev[l].nx = lnx[l];
        t.lev[l].m = (uint8_t*)b; b += align_up(nl);
        t.lev[l].x = (float*)b; b += align_up(nl * 4);
        t.lev[l].y = (float*)b; b += align_up(nl * 4);
        t.lev[l].b = (float*)b; b += align_up(nl * 4);
    }
    t.sc = (Scalars*)b; b += align_up(sizeof(Scalars));
    // appended last so that the offsets smrf_inpaint_layout reports never move: the index maps of the compact CG
    // vectors (single-GPU solver) and the scratch of the prefix sum that builds them
    t.pos = (int*)b; b += align_up(n * 4);
    t.idx = (int*)b; b += align_up(n * 4);
    t.scan_tmp = (void*)b; t.scan_tmp_bytes = scan_temp_bytes(n); b += align_up(t.scan_tmp_bytes);
    if (w) *w = t;
    return (size_t)(b - (char*)workspace);
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

// Tiles of kRows rows x 256 columns, walked with a block stride: a few thousand CTAs whatever
// the grid size, so the per-block atomics of the reductions stay cheap, and one div/mod per
// tile rather than per row.
constexpr int kRows = 4;
struct Tiles {
    int64_t ny, nx, per_row, total;
    __host__ __device__ Tiles(int64_t ny_, int64_t nx_)
        : ny(ny_), nx(nx_), per_row((nx_ + kBlock - 1) / kBlock),
          total(((ny_ + kRows - 1) / kRows) * ((nx_ + kBlock - 1) / kBlock)) {}
};
// body sees: int64_t Y_, X_ (row, column of this thread); bool IN_ (X_ < nx)
__device__ __forceinline__ int64_t tile_first() { return blockIdx.x; }
__device__ __forceinline__ int64_t tile_step() { return gridDim.x; }
__device__ __forceinline__ int64_t tile_lane() { return threadIdx.x; }
#define SMRF_FOR_TILES(T, Y_, X_, IN_)                                                         \
    for (int64_t t__ = tile_first(); t__ < (T).total; t__ += tile_step())                      \
        if (const int64_t ty__ = t__ / (T).per_row; true)                                      \
            if (const int64_t X_ = (t__ - ty__ * (T).per_row) * kBlock + tile_lane(); true)    \
                if (const bool IN_ = X_ < (T).nx; true)                                        \
                    for (int64_t Y_ = ty__ * kRows, ye__ = (Y_ + kRows < (T).ny ? Y_ + kRows : (T).ny); Y_ < ye__; ++Y_)

__device__ __forceinline__ int degree(int64_t y, int64_t x, int64_t ny, int64_t nx, int above = 0, int below = 0) {
    return (y > 0 || above) + (y + 1 < ny || below) + (x > 0) + (x + 1 < nx);
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

// ---- starting guess: u = known value; an unknown cell takes the caller's guess if there is one,
// else the mean of its known 4-neighbours (exact for an isolated empty cell, which is what more
// than half of the empty cells of a minimum surface are), else the mean of all known cells.
template <typename T>
__global__ void __launch_bounds__(kBlock) init_u_kernel(const T* __restrict__ grid, const uint8_t* __restrict__ unk,
                                                        double* __restrict__ u, int64_t ny, int64_t nx,
                                                        const Scalars* sc, const T* __restrict__ guess) {
    const double mean = sc->n_known ? sc->sum_known / (double)sc->n_known : 0.0;
    const Tiles T2(ny, nx);
    SMRF_FOR_TILES(T2, y, x, in) {
        if (in) {
            const int64_t i = y * nx + x;
            double v;
            if (unk[i]) {
                v = mean;
                bool have = false;
                if (guess) {
                    const double g = (double)guess[i];
                    if (g == g && fabs(g) < 1e300) { v = g; have = true; }   // a NaN / inf guess is ignored
                }
                if (!have) {
                    double s = 0.0;
                    int c = 0;
                    if (y > 0 && !unk[i - nx]) { s += (double)grid[i - nx]; ++c; }
                    if (y + 1 < ny && !unk[i + nx]) { s += (double)grid[i + nx]; ++c; }
                    if (x > 0 && !unk[i - 1]) { s += (double)grid[i - 1]; ++c; }
                    if (x + 1 < nx && !unk[i + 1]) { s += (double)grid[i + 1]; ++c; }
                    if (c) v = s / (double)c;
                }
            } else {
                v = (double)grid[i];
            }
            u[i] = v;
        }
    }
}

// r = b - A u on the unknown cells (u holds the known values at known cells, so the
// right-hand side is implicit); b0 = (float) r feeds the preconditioner; rmax[0] = max |r|
__global__ void __launch_bounds__(kBlock) residual0_kernel(Ws w, int64_t ny, int64_t nx, const double* __restrict__ ua,
                                                           const double* __restrict__ ub) {
    const Tiles T(ny, nx);
    const uint8_t* unk = w.lev[0].m;
    double rm = 0.0;
    SMRF_FOR_TILES(T, y, x, in) {
        if (in) {
            const int64_t i = y * nx + x;
            float rf = 0.f;
            if (unk[i]) {
                double s = 0.0;
                if (y > 0) s += w.u[i - nx];
                else if (w.has_above) s += ua[x];
                if (y + 1 < ny) s += w.u[i + nx];
                else if (w.has_below) s += ub[x];
                if (x > 0) s += w.u[i - 1];
                if (x + 1 < nx) s += w.u[i + 1];
                const int d = degree(y, x, ny, nx, w.has_above, w.has_below);
                const double r = d ? s - (double)d * w.u[i] : 0.0;
                w.r[i] = r; w.p[i] = 0.0;
                rf = (float)r;
                rm = (fabs(r) < INFINITY) ? fmax(rm, fabs(r)) : INFINITY;     // a NaN / inf residual must surface
            }
            w.lev[0].b[i] = rf;
        }
    }
    rm = block_max(rm);
    if (threadIdx.x == 0) atomicMax(&w.sc->rmax[0], (unsigned long long)__double_as_longlong(rm));
}

// The four CG kernels below give every thread one column of a kRows-row tile and keep the
// tile's values in registers: all loads of a tile are issued before any arithmetic (the
// kernels are latency-bound otherwise: ncu showed 70-90 % long-scoreboard stalls with one
// dependent load chain per cell).
#define SMRF_FOR_TILE_BLOCKS(T, Y0_, X_, IN_)                                                 \
    for (int64_t t__ = tile_first(); t__ < (T).total; t__ += tile_step())                     \
        if (const int64_t ty__ = t__ / (T).per_row; true)                                     \
            if (const int64_t X_ = (t__ - ty__ * (T).per_row) * kBlock + tile_lane(); true)   \
                if (const bool IN_ = X_ < (T).nx; true)                                       \
                    if (const int64_t Y0_ = ty__ * kRows; true)

// rz[k] = r.z, then p = z + beta p with beta = rz[k] / rz[k-1] needs the finished sum: two kernels.
template <bool JACOBI>
__global__ void __launch_bounds__(kBlock) rz_kernel(Ws w, const float* __restrict__ z, int64_t ny, int64_t nx, int k) {
    const Tiles T(ny, nx);
    const uint8_t* __restrict__ unk = w.lev[0].m;
    double rz = 0.0;
    SMRF_FOR_TILE_BLOCKS(T, y0, x, in) {
        if (in) {
            uint8_t mk[kRows];
            double rv[kRows];
            float zv[kRows];
            const int64_t i0 = y0 * nx + x;
#pragma unroll
            for (int r = 0; r < kRows; ++r) mk[r] = (y0 + r < ny) ? unk[i0 + r * nx] : 0;
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                // the preconditioned product uses the float32 residual the cycle saw (lev[0].b), as the fused up leg does
                rv[r] = mk[r] ? (JACOBI ? w.r[i0 + r * nx] : (double)w.lev[0].b[i0 + r * nx]) : 0.0;
                zv[r] = (!JACOBI && mk[r]) ? z[i0 + r * nx] : 0.f;
            }
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                if (JACOBI) {
                    const int d = degree(y0 + r, x, ny, nx, w.has_above, w.has_below);
                    if (mk[r] && d) rz += rv[r] * (rv[r] / (double)d);
                } else {
                    rz += rv[r] * (double)zv[r];
                }
            }
        }
    }
    rz = block_sum(rz);
    if (threadIdx.x == 0 && rz != 0.0) atomicAdd(&w.sc->rz[k], rz);
}

template <bool JACOBI>
__global__ void __launch_bounds__(kBlock) p_update_kernel(Ws w, const float* __restrict__ z, int64_t ny, int64_t nx,
                                                          int k) {
    const Tiles T(ny, nx);
    const uint8_t* __restrict__ unk = w.lev[0].m;
    const double beta = (k == 0 || w.sc->rz[k - 1] == 0.0) ? 0.0 : w.sc->rz[k] / w.sc->rz[k - 1];
    SMRF_FOR_TILE_BLOCKS(T, y0, x, in) {
        if (in) {
            uint8_t mk[kRows];
            double zv[kRows], pv[kRows];
            const int64_t i0 = y0 * nx + x;
#pragma unroll
            for (int r = 0; r < kRows; ++r) mk[r] = (y0 + r < ny) ? unk[i0 + r * nx] : 0;
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                if (mk[r]) {
                    pv[r] = w.p[i0 + r * nx];
                    if (JACOBI) {
                        const int d = degree(y0 + r, x, ny, nx, w.has_above, w.has_below);
                        zv[r] = d ? w.r[i0 + r * nx] / (double)d : 0.0;
                    } else {
                        zv[r] = (double)z[i0 + r * nx];
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < kRows; ++r)
                if (mk[r]) w.p[i0 + r * nx] = zv[r] + beta * pv[r];
        }
    }
}

// q = A p (p is only ever read on unknown cells);  pq[k] = p.q
__global__ void __launch_bounds__(kBlock) apply_kernel(Ws w, int64_t ny, int64_t nx, int k, const double* __restrict__ pa,
                                                       const double* __restrict__ pb, const uint8_t* __restrict__ ma,
                                                       const uint8_t* __restrict__ mb) {
    const Tiles T(ny, nx);
    const uint8_t* __restrict__ unk = w.lev[0].m;
    const double* __restrict__ p = w.p;
    double pq = 0.0;
    SMRF_FOR_TILE_BLOCKS(T, y0, x, in) {
        if (in) {
            // slot j <-> row y0 + j - 1: the column of the tile plus one row above and below
            uint8_t mk[kRows + 2], ml[kRows], mr[kRows];
            double pv[kRows + 2], pl[kRows], pr[kRows];
            const int64_t i0 = y0 * nx + x;
#pragma unroll
            for (int j = 0; j < kRows + 2; ++j) {
                const int64_t y = y0 + j - 1;
                uint8_t m = 0;
                if (y >= 0 && y < ny) m = unk[i0 + (int64_t)(j - 1) * nx];
                else if (y < 0 && w.has_above) m = ma[x];
                else if (y == ny && w.has_below) m = mb[x];
                mk[j] = m;
            }
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const bool me = mk[r + 1] != 0;
                ml[r] = (me && x > 0) ? unk[i0 + r * nx - 1] : 0;
                mr[r] = (me && x + 1 < nx) ? unk[i0 + r * nx + 1] : 0;
            }
#pragma unroll
            for (int j = 0; j < kRows + 2; ++j) {
                const int64_t y = y0 + j - 1;
                double v = 0.0;
                if (mk[j]) v = (y < 0) ? pa[x] : ((y >= ny) ? pb[x] : p[i0 + (int64_t)(j - 1) * nx]);
                pv[j] = v;
            }
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                pl[r] = ml[r] ? p[i0 + r * nx - 1] : 0.0;
                pr[r] = mr[r] ? p[i0 + r * nx + 1] : 0.0;
            }
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                if (mk[r + 1] && y0 + r < ny) {
                    const double pi = pv[r + 1];
                    const double sum = pv[r] + pv[r + 2] + pl[r] + pr[r];
                    const double q = (double)degree(y0 + r, x, ny, nx, w.has_above, w.has_below) * pi - sum;
                    w.q[i0 + r * nx] = q;
                    pq += pi * q;
                }
            }
        }
    }
    pq = block_sum(pq);
    if (threadIdx.x == 0 && pq != 0.0) atomicAdd(&w.sc->pq[k], pq);
}

// u += alpha p; r -= alpha q; b0 = (float) r; rmax[k+1] = max |r|
__global__ void __launch_bounds__(kBlock) update_kernel(Ws w, int64_t ny, int64_t nx, int k) {
    const Tiles T(ny, nx);
    const uint8_t* __restrict__ unk = w.lev[0].m;
    const double pqk = w.sc->pq[k];
    const double alpha = pqk != 0.0 ? w.sc->rz[k] / pqk : 0.0;
    double rm = 0.0;
    SMRF_FOR_TILE_BLOCKS(T, y0, x, in) {
        if (in) {
            uint8_t mk[kRows];
            double pv[kRows], qv[kRows], uv[kRows], rv[kRows];
            const int64_t i0 = y0 * nx + x;
#pragma unroll
            for (int r = 0; r < kRows; ++r) mk[r] = (y0 + r < ny) ? unk[i0 + r * nx] : 0;
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                if (mk[r]) {
                    pv[r] = w.p[i0 + r * nx]; qv[r] = w.q[i0 + r * nx];
                    uv[r] = w.u[i0 + r * nx]; rv[r] = w.r[i0 + r * nx];
                }
            }
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                if (mk[r]) {
                    const double rn = rv[r] - alpha * qv[r];
                    w.u[i0 + r * nx] = uv[r] + alpha * pv[r];
                    w.r[i0 + r * nx] = rn;
                    w.lev[0].b[i0 + r * nx] = (float)rn;
                    rm = (fabs(rn) < INFINITY) ? fmax(rm, fabs(rn)) : INFINITY;
                }
            }
        }
    }
    rm = block_max(rm);
    if (threadIdx.x == 0) atomicMax(&w.sc->rmax[k + 1], (unsigned long long)__double_as_longlong(rm));
}

// ---- compact CG (single-GPU solver) -----------------------------------------------------------------------
// Only 14 % (first solve) to ~30 % (second solve) of the cells are unknown, yet the grid-layout CG kernels stream
// whole planes of u, r, p, q (the mask only predicates the accesses: at these densities nearly every 32-byte sector
// is still touched).  Here the four float64 vectors hold the unknown cells only, in row-major order of the grid:
// idx[k] is the cell of unknown k, pos[cell] its index (-1 for a known cell).  The V-cycle keeps the grid layout
// (lev[0].b / lev[0].y): the residual is scattered into it by the update, the correction gathered from it by the
// direction update.  The row-band solver keeps the grid-layout kernels (its halo exchanges address rows).
__global__ void __launch_bounds__(kBlock) compact_build_kernel(const uint8_t* __restrict__ unk, int* __restrict__ pos,
                                                               int* __restrict__ idx, int n) {
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
        const int k = pos[i];               // exclusive prefix sum of the flags
        if (unk[i]) idx[k] = i;
        else pos[i] = -1;
    }
}

// u_c = the caller's guess, else the mean of the known 4-neighbours, else the mean of all known cells
template <typename T>
__global__ void __launch_bounds__(kBlock) compact_init_kernel(const T* __restrict__ grid, Ws w, int ny, int nx, int nu,
                                                              const T* __restrict__ guess) {
    const double mean = w.sc->n_known ? w.sc->sum_known / (double)w.sc->n_known : 0.0;
    for (int k = blockIdx.x * kBlock + threadIdx.x; k < nu; k += gridDim.x * kBlock) {
        const int i = w.idx[k];
        const int y = i / nx, x = i - y * nx;
        double v = mean;
        bool have = false;
        if (guess) {
            const double g = (double)guess[i];
            if (g == g && fabs(g) < 1e300) { v = g; have = true; }   // a NaN / inf guess is ignored
        }
        if (!have) {
            double s = 0.0;
            int c = 0;
            if (y > 0 && w.pos[i - nx] < 0) { s += (double)grid[i - nx]; ++c; }
            if (y + 1 < ny && w.pos[i + nx] < 0) { s += (double)grid[i + nx]; ++c; }
            if (x > 0 && w.pos[i - 1] < 0) { s += (double)grid[i - 1]; ++c; }
            if (x + 1 < nx && w.pos[i + 1] < 0) { s += (double)grid[i + 1]; ++c; }
            if (c) v = s / (double)c;
        }
        w.u[k] = v;
    }
}

// value of the neighbour cell j in the current iterate: an unknown's u_c, else the grid value
template <typename T>
__device__ __forceinline__ double compact_val(const T* __restrict__ grid, const Ws& w, int j) {
    const int pj = w.pos[j];
    return pj >= 0 ? w.u[pj] : (double)grid[j];
}

// r_c = b - A u (the right-hand side is implicit in the known neighbours); lev[0].b = (float) r; p_c = 0; rmax[0]
// (row bands: ua / ub = the neighbouring bands' boundary rows of u, known values included)
template <typename T>
__global__ void __launch_bounds__(kBlock) compact_residual0_kernel(const T* __restrict__ grid, Ws w, int ny, int nx, int nu,
                                                                   const double* __restrict__ ua,
                                                                   const double* __restrict__ ub) {
    double rm = 0.0;
    for (int k = blockIdx.x * kBlock + threadIdx.x; k < nu; k += gridDim.x * kBlock) {
        const int i = w.idx[k];
        const int y = i / nx, x = i - y * nx;
        double s = 0.0;
        int d = 0;
        if (y > 0) { s += compact_val(grid, w, i - nx); ++d; }
        else if (w.has_above) { s += ua[x]; ++d; }
        if (y + 1 < ny) { s += compact_val(grid, w, i + nx); ++d; }
        else if (w.has_below) { s += ub[x]; ++d; }
        if (x > 0) { s += compact_val(grid, w, i - 1); ++d; }
        if (x + 1 < nx) { s += compact_val(grid, w, i + 1); ++d; }
        const double r = d ? s - (double)d * w.u[k] : 0.0;
        w.r[k] = r;
        w.p[k] = 0.0;
        w.lev[0].b[i] = (float)r;
        rm = (fabs(r) < INFINITY) ? fmax(rm, fabs(r)) : INFINITY;     // a NaN / inf residual must surface
    }
    rm = block_max(rm);
    if (threadIdx.x == 0) atomicMax(&w.sc->rmax[0], (unsigned long long)__double_as_longlong(rm));
}

// first and last row of a compact vector as dense rows for the halo exchange of the row-band solver: the unknown's
// value, else `grid` (u: the known elevation) or 0 (p: no search direction on known cells)
template <typename T>
__global__ void __launch_bounds__(kBlock) compact_rows_kernel(const double* __restrict__ v, const T* __restrict__ grid, Ws w,
                                                              int ny, int nx, double* __restrict__ first,
                                                              double* __restrict__ last) {
    for (int x = blockIdx.x * kBlock + threadIdx.x; x < nx; x += gridDim.x * kBlock) {
        const int i0 = x, i1 = (ny - 1) * nx + x;
        const int p0 = w.pos[i0], p1 = w.pos[i1];
        if (first) first[x] = p0 >= 0 ? v[p0] : (grid ? (double)grid[i0] : 0.0);
        if (last) last[x] = p1 >= 0 ? v[p1] : (grid ? (double)grid[i1] : 0.0);
    }
}

// p_c = z + (rz[k] / rz[k-1]) p_c, z gathered from the grid-layout result of the cycle
__global__ void __launch_bounds__(kBlock) compact_p_kernel(Ws w, const float* __restrict__ z, int nu, int k) {
    const double beta = (k == 0 || w.sc->rz[k - 1] == 0.0) ? 0.0 : w.sc->rz[k] / w.sc->rz[k - 1];
    for (int j = blockIdx.x * kBlock + threadIdx.x; j < nu; j += gridDim.x * kBlock)
        w.p[j] = (double)z[w.idx[j]] + beta * w.p[j];
}

// q_c = A p_c; pq[k] = p . q
// (row bands: pa / pb = the neighbouring bands' boundary rows of p, zero at their known cells)
__global__ void __launch_bounds__(kBlock) compact_apply_kernel(Ws w, int ny, int nx, int nu, int k,
                                                               const double* __restrict__ pa, const double* __restrict__ pb) {
    double pq = 0.0;
    for (int j = blockIdx.x * kBlock + threadIdx.x; j < nu; j += gridDim.x * kBlock) {
        const int i = w.idx[j];
        const int y = i / nx, x = i - y * nx;
        double s = 0.0;
        int d = 0;
        if (y > 0) { const int t = w.pos[i - nx]; if (t >= 0) s += w.p[t]; ++d; }
        else if (w.has_above) { s += pa[x]; ++d; }
        if (y + 1 < ny) { const int t = w.pos[i + nx]; if (t >= 0) s += w.p[t]; ++d; }
        else if (w.has_below) { s += pb[x]; ++d; }
        if (x > 0) { const int t = w.pos[i - 1]; if (t >= 0) s += w.p[t]; ++d; }
        if (x + 1 < nx) { const int t = w.pos[i + 1]; if (t >= 0) s += w.p[t]; ++d; }
        const double pj = w.p[j];
        const double q = (double)d * pj - s;
        w.q[j] = q;
        pq += pj * q;
    }
    pq = block_sum(pq);
    if (threadIdx.x == 0 && pq != 0.0) atomicAdd(&w.sc->pq[k], pq);
}

// u_c += alpha p_c; r_c -= alpha q_c; lev[0].b = (float) r; rmax[k+1] = max |r|
__global__ void __launch_bounds__(kBlock) compact_update_kernel(Ws w, int nu, int k) {
    const double pqk = w.sc->pq[k];
    const double alpha = pqk != 0.0 ? w.sc->rz[k] / pqk : 0.0;
    double rm = 0.0;
    for (int j = blockIdx.x * kBlock + threadIdx.x; j < nu; j += gridDim.x * kBlock) {
        const double rn = w.r[j] - alpha * w.q[j];
        w.u[j] += alpha * w.p[j];
        w.r[j] = rn;
        w.lev[0].b[w.idx[j]] = (float)rn;
        rm = (fabs(rn) < INFINITY) ? fmax(rm, fabs(rn)) : INFINITY;
    }
    rm = block_max(rm);
    if (threadIdx.x == 0) atomicMax(&w.sc->rmax[k + 1], (unsigned long long)__double_as_longlong(rm));
}

template <typename T>
__global__ void __launch_bounds__(kBlock) compact_writeback_kernel(T* __restrict__ grid, Ws w, int nu) {
    for (int k = blockIdx.x * kBlock + threadIdx.x; k < nu; k += gridDim.x * kBlock) grid[w.idx[k]] = (T)w.u[k];
}

template <typename T>
__global__ void __launch_bounds__(kBlock) writeback_kernel(T* __restrict__ grid, const uint8_t* __restrict__ unk,
                                                           const double* __restrict__ u, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
        if (unk[i]) grid[i] = (T)u[i];
}

// ---- multigrid pieces (float32) -------------------------------------------------------------
// coarse cell unknown <=> every in-grid child unknown
__global__ void __launch_bounds__(kBlock) coarsen_kernel(const uint8_t* __restrict__ mf, uint8_t* __restrict__ mc,
                                                         int64_t fy, int64_t fx, int64_t cy, int64_t cx) {
    const Tiles T(cy, cx);
    SMRF_FOR_TILES(T, Y, X, in) {
        if (in) {
            uint8_t all = 1;
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int64_t y = 2 * Y + a, x = 2 * X + b;
                    if (y < fy && x < fx) all &= mf[y * fx + x];
                }
            mc[Y * cx + X] = all;
        }
    }
}

// ---- fused legs of the V-cycle: each level is two launches and two passes over its vectors ----
// A CTA owns a 26 x 58 tile; everything its three sweeps need from neighbours is recomputed in a
// three-cell halo held in shared memory (temporal blocking of the Jacobi sweeps).  The halo'd
// tile is 32 x 64 so that local indices are shifts and masks; the in-grid test, the mask and
// 1/degree are evaluated once per element while loading.
constexpr int kH = 3;                                // sweeps per leg = halo width
constexpr int kHY = 32, kHX = 64;
constexpr int kTY = kHY - 2 * kH, kTX = kHX - 2 * kH;

struct TileBuf {
    float b[kHY][kHX];
    float s0[kHY][kHX];
    float s1[kHY][kHX];
    float dinv[kHY][kHX];   // 1 / degree on unknown cells, 0 elsewhere (which also zeroes the iterate there)
    float deg[kHY][kHX];
};

// x + omega (b - A x) / deg, with x == 0 off the unknown set (dinv == 0 there keeps it so)
__device__ __forceinline__ float jacobi(const float (*x)[kHX], const TileBuf& t, int ly, int lx, float omega) {
    const float xi = x[ly][lx];
    const float s = x[ly - 1][lx] + x[ly + 1][lx] + x[ly][lx - 1] + x[ly][lx + 1];
    const float dv = t.dinv[ly][lx];
    return dv == 0.f ? 0.f : xi + omega * dv * (t.b[ly][lx] - (t.deg[ly][lx] * xi - s));
}

// Thread mapping of the fused legs: thread t owns column lx = t & 63 of the halo'd tile and the
// rows ly = (t >> 6) + 4 e, e = 0..7, so everything that depends on the column only (in-grid
// test, horizontal part of the degree, base address) is computed once per tile.
template <typename I>
struct TilePos {
    int lx, lyb;        // local column, first local row
    int y_first;        // grid row of local row lyb
    bool col_ok;        // column inside the grid
    int dxh;            // in-grid horizontal neighbours
    I g0;               // flat index of (y_first, column)
};
template <typename I>
__device__ __forceinline__ TilePos<I> tile_pos(I y0, I x0, I nx) {
    TilePos<I> p;
    p.lx = threadIdx.x & 63;
    p.lyb = threadIdx.x >> 6;
    const I xx = x0 + p.lx - kH;
    p.y_first = (int)(y0 - kH + p.lyb);
    p.col_ok = xx >= 0 && xx < nx;
    p.dxh = (xx > 0) + (xx + 1 < nx);
    p.g0 = (I)p.y_first * nx + xx;
    return p;
}

// loads b (and, on the way up, the iterate plus the coarse correction) of the halo'd tile;
// fills t.b, t.dinv, t.deg, t.s0
template <bool UP, typename I>
__device__ __forceinline__ void load_tile(TileBuf& t, const TilePos<I>& p, const float* __restrict__ b,
                                          const uint8_t* __restrict__ m, const float* __restrict__ x,
                                          const float* __restrict__ xc, int ny, I nx, I cx, int above,
                                          int below) {
    uint8_t mm[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int y = p.y_first + 4 * e;
        mm[e] = (p.col_ok && y >= 0 && y < ny) ? m[p.g0 + (I)(4 * e) * nx] : 0;
    }
    float bb[8], xv[8], cv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const I g = p.g0 + (I)(4 * e) * nx;
        bb[e] = mm[e] ? b[g] : 0.f;
        if (UP) {
            const int y = p.y_first + 4 * e;
            xv[e] = mm[e] ? x[g] : 0.f;
            cv[e] = mm[e] ? xc[(I)(y >> 1) * cx + ((g - (I)y * nx) >> 1)] : 0.f;
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int y = p.y_first + 4 * e;
        const int ly = p.lyb + 4 * e;
        float dv = 0.f, dg = 0.f, s = 0.f;
        if (mm[e]) {
            const int d = p.dxh + ((y > 0) || above) + ((y + 1 < ny) || below);
            dg = (float)d;
            dv = __frcp_rn((float)(d > 0 ? d : 1));
            s = UP ? xv[e] + cv[e] : kOmegaA * dv * bb[e];     // down: first sweep (weight A) from the zero vector
        }
        t.b[ly][p.lx] = bb[e]; t.dinv[ly][p.lx] = dv; t.deg[ly][p.lx] = dg; t.s0[ly][p.lx] = s;
    }
}

// dst = jacobi(src, omega) on the tile plus `ring` rings
template <int RING, typename I>
__device__ __forceinline__ void sweep(float (*dst)[kHX], const float (*src)[kHX], const TileBuf& t, const TilePos<I>& p,
                                      float omega) {
    constexpr int lo = kH - RING, hiy = kHY - 1 - (kH - RING), hix = kHX - 1 - (kH - RING);
    if (p.lx >= lo && p.lx <= hix) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int ly = p.lyb + 4 * e;
            if (ly >= lo && ly <= hiy) dst[ly][p.lx] = jacobi(src, t, ly, p.lx, omega);
        }
    }
}

// down leg: x = three Jacobi sweeps (weights A, B, C) from zero on A x = b;  bc = P^T (b - A x)
// I: index type -- int when the level has fewer than 2^31 cells (all index arithmetic in 32 bits: the legs are
// issue-bound and 64-bit multiplies / divides are instruction pairs and sequences), int64_t otherwise.
template <typename I>
__global__ void __launch_bounds__(kBlock) down_kernel(const float* __restrict__ b, const uint8_t* __restrict__ m,
                                                      float* __restrict__ xout, const uint8_t* __restrict__ mc,
                                                      float* __restrict__ bc, int64_t ny_, int64_t nx_, int64_t cy_,
                                                      int64_t cx_, int above, int below) {
    __shared__ TileBuf t;
    const I ny = (I)ny_, nx = (I)nx_, cy = (I)cy_, cx = (I)cx_;
    const I tiles_x = (nx + kTX - 1) / kTX, tiles = tiles_x * ((ny + kTY - 1) / kTY);
    for (I tile = (I)tile_first(); tile < tiles; tile += (I)tile_step()) {
        const I ty = tile / tiles_x;
        const I y0 = ty * kTY, x0 = (tile - ty * tiles_x) * kTX;
        const TilePos<I> p = tile_pos<I>(y0, x0, nx);
        __syncthreads();
        load_tile<false, I>(t, p, b, m, nullptr, nullptr, (int)ny, nx, (I)0, above, below);
        __syncthreads();
        sweep<2>(t.s1, t.s0, t, p, kOmegaB);
        __syncthreads();
        sweep<1>(t.s0, t.s1, t, p, kOmegaC);
        __syncthreads();
        if (p.col_ok && p.lx >= kH && p.lx < kTX + kH) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int ly = p.lyb + 4 * e, y = p.y_first + 4 * e;
                if (ly >= kH && ly < kTY + kH && y < ny) xout[p.g0 + (I)(4 * e) * nx] = t.s0[ly][p.lx];
            }
        }
        // coarse right-hand side: 13 x 29 coarse cells per tile
        for (int i = threadIdx.x; i < (kTY / 2) * 32; i += kBlock) {
            const int cly = i >> 5, clx = i & 31;
            const I Y = y0 / 2 + cly, X = x0 / 2 + clx;
            if (clx < kTX / 2 && Y < cy && X < cx) {
                float acc = 0.f;
                if (mc[Y * cx + X]) {
#pragma unroll
                    for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                        for (int c2 = 0; c2 < 2; ++c2) {
                            const int ly = 2 * cly + a2 + kH, lx = 2 * clx + c2 + kH;
                            if (t.dinv[ly][lx] != 0.f) {
                                const float sum = t.s0[ly - 1][lx] + t.s0[ly + 1][lx] + t.s0[ly][lx - 1] + t.s0[ly][lx + 1];
                                acc += t.b[ly][lx] - (t.deg[ly][lx] * t.s0[ly][lx] - sum);
                            }
                        }
                }
                bc[Y * cx + X] = acc;
            }
        }
    }
}

// up leg: x += P xc, then three Jacobi sweeps (weights C, B, A: the down leg's in reverse, which
// keeps the cycle symmetric); written to `xout` (another buffer: tiles read each other's halo of x)
template <bool RZ, typename I>
__global__ void __launch_bounds__(kBlock) up_kernel(const float* __restrict__ x, const float* __restrict__ xc,
                                                    const float* __restrict__ b, const uint8_t* __restrict__ m,
                                                    float* __restrict__ xout, int64_t ny_, int64_t nx_, int64_t cx_,
                                                    int above, int below, double* rz_slot, int64_t rz_lo, int64_t rz_hi) {
    __shared__ TileBuf t;
    double rz = 0.0;       // RZ (level 0): *rz_slot += b . z over rows [rz_lo, rz_hi) while z is at hand -- one pass less over the fine grid
    const I ny = (I)ny_, nx = (I)nx_, cx = (I)cx_;
    const I tiles_x = (nx + kTX - 1) / kTX, tiles = tiles_x * ((ny + kTY - 1) / kTY);
    for (I tile = (I)tile_first(); tile < tiles; tile += (I)tile_step()) {
        const I ty = tile / tiles_x;
        const I y0 = ty * kTY, x0 = (tile - ty * tiles_x) * kTX;
        const TilePos<I> p = tile_pos<I>(y0, x0, nx);
        __syncthreads();
        load_tile<true, I>(t, p, b, m, x, xc, (int)ny, nx, cx, above, below);
        __syncthreads();
        sweep<2>(t.s1, t.s0, t, p, kOmegaC);
        __syncthreads();
        sweep<1>(t.s0, t.s1, t, p, kOmegaB);
        __syncthreads();
        if (p.col_ok && p.lx >= kH && p.lx < kTX + kH) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int ly = p.lyb + 4 * e, y = p.y_first + 4 * e;
                if (ly >= kH && ly < kTY + kH && y < ny) {
                    const float zv = jacobi(t.s0, t, ly, p.lx, kOmegaA);
                    xout[p.g0 + (I)(4 * e) * nx] = zv;
                    if constexpr (RZ) {
                        if (y >= rz_lo && y < rz_hi) rz += (double)t.b[ly][p.lx] * (double)zv;
                    }
                }
            }
        }
    }
    if constexpr (RZ) {
        rz = block_sum(rz);
        if (threadIdx.x == 0 && rz != 0.0) atomicAdd(rz_slot, rz);
    }
}

// ---- the coarse tail of the cycle in ONE launch ---------------------------------------------------
// Levels of at most kTailCells cells are latency-bound: two five-microsecond launches per level and eight for the
// coarsest sweeps were ~150 us of every iteration.  One CTA walks all of them -- the same sweeps, restriction and
// prolongation as the per-level kernels, on the level arrays in global memory (a few KB: L1 / L2 resident), with
// a block barrier between sweeps.
constexpr int kTailThreads = 1024;
constexpr int64_t kTailCells = 8192;

__device__ __forceinline__ float tail_nbr_sum(const float* __restrict__ x, int64_t i, int64_t y, int64_t xx, int64_t ny,
                                              int64_t nx) {
    float s = 0.f;
    if (y > 0) s += x[i - nx];
    if (y + 1 < ny) s += x[i + nx];
    if (xx > 0) s += x[i - 1];
    if (xx + 1 < nx) s += x[i + 1];
    return s;
}

// out = x + omega (b - A x) / deg on unknown cells, 0 elsewhere, with the arithmetic of jacobi() / load_tile()
// (FIRST: x is the zero vector)
template <bool FIRST>
__device__ __forceinline__ void tail_sweep(const float* __restrict__ x, float* __restrict__ out, const Level& v, int above,
                                           int below, float omega) {
    const int64_t n = v.ny * v.nx;
    for (int64_t i = threadIdx.x; i < n; i += kTailThreads) {
        float r = 0.f;
        if (v.m[i]) {
            const int64_t y = i / v.nx, xx = i - y * v.nx;
            const int d = degree(y, xx, v.ny, v.nx, above, below);
            const float dv = __frcp_rn((float)(d > 0 ? d : 1));
            if (FIRST) r = omega * dv * v.b[i];
            else {
                const float xi = x[i];
                r = xi + omega * dv * (v.b[i] - ((float)d * xi - tail_nbr_sum(x, i, y, xx, v.ny, v.nx)));
            }
        }
        out[i] = r;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kTailThreads) tail_kernel(Ws w, int l0) {
    const int L = w.nlev;
    const int above = w.has_above, below = w.has_below;
    for (int l = l0; l + 1 < L; ++l) {          // down legs: three sweeps from zero, then bc = P^T (b - A x)
        const Level v = w.lev[l], c = w.lev[l + 1];
        tail_sweep<true>(nullptr, v.x, v, above, below, kOmegaA);
        tail_sweep<false>(v.x, v.y, v, above, below, kOmegaB);
        tail_sweep<false>(v.y, v.x, v, above, below, kOmegaC);
        const int64_t nc = c.ny * c.nx;
        for (int64_t j = threadIdx.x; j < nc; j += kTailThreads) {
            float acc = 0.f;
            if (c.m[j]) {
                const int64_t Y = j / c.nx, X = j - Y * c.nx;
#pragma unroll
                for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) {
                        const int64_t y = 2 * Y + a2, xx = 2 * X + c2;
                        if (y < v.ny && xx < v.nx) {
                            const int64_t i = y * v.nx + xx;
                            if (v.m[i]) {
                                const int d = degree(y, xx, v.ny, v.nx, above, below);
                                acc += v.b[i] - ((float)d * v.x[i] - tail_nbr_sum(v.x, i, y, xx, v.ny, v.nx));
                            }
                        }
                    }
            }
            c.b[j] = acc;
        }
        __syncthreads();
    }
    {   // coarsest level: kCoarsestSweeps damped-Jacobi sweeps (smooth_kernel's arithmetic); the result lands in y
        const Level v = w.lev[L - 1];
        const int64_t n = v.ny * v.nx;
        const float* a = nullptr;
        float* o = v.x;
        for (int s2 = 0; s2 < kCoarsestSweeps; ++s2) {
            for (int64_t i = threadIdx.x; i < n; i += kTailThreads) {
                float r = 0.f;
                if (v.m[i]) {
                    const int64_t y = i / v.nx, xx = i - y * v.nx;
                    const float d = (float)degree(y, xx, v.ny, v.nx, above, below);
                    if (s2 == 0) r = d > 0.f ? kOmega * v.b[i] / d : 0.f;
                    else {
                        const float xi = a[i];
                        r = d > 0.f ? xi + kOmega * (v.b[i] - (d * xi - tail_nbr_sum(a, i, y, xx, v.ny, v.nx))) / d : 0.f;
                    }
                }
                o[i] = r;
            }
            __syncthreads();
            a = o;
            o = (o == v.x) ? v.y : v.x;
        }
    }
    for (int l = L - 2; l >= l0; --l) {         // up legs: x += P xc, sweeps C, B, A; the result lands in y
        const Level v = w.lev[l], c = w.lev[l + 1];
        const int64_t n = v.ny * v.nx;
        for (int64_t i = threadIdx.x; i < n; i += kTailThreads) {
            if (v.m[i]) {
                const int64_t y = i / v.nx, xx = i - y * v.nx;
                v.x[i] = v.x[i] + c.y[(y >> 1) * c.nx + (xx >> 1)];
            }
        }
        __syncthreads();
        tail_sweep<false>(v.x, v.y, v, above, below, kOmegaC);
        tail_sweep<false>(v.y, v.x, v, above, below, kOmegaB);
        tail_sweep<false>(v.x, v.y, v, above, below, kOmegaA);
    }
}

// first level of the tail: the first one small enough (never level 0 unless it is the only level: the level-0 up
// leg carries the r.z product of the single-GPU loop)
static int tail_level(const Ws& w) {
    int l = 1;
    while (l < w.nlev - 1 && w.lev[l].ny * w.lev[l].nx > kTailCells) ++l;
    return l < w.nlev ? l : w.nlev - 1;
}

static inline int fused_grid(int64_t ny, int64_t nx) {
    int64_t t = ((nx + kTX - 1) / kTX) * ((ny + kTY - 1) / kTY);
    int64_t cap = (int64_t)num_sms() * 8;
    if (t > cap) t = cap;
    if (t < 1) t = 1;
    return (int)t;
}

static inline int tile_grid(int64_t ny, int64_t nx) {
    int64_t t = ((ny + kRows - 1) / kRows) * ((nx + kBlock - 1) / kBlock);
    int64_t cap = (int64_t)num_sms() * 16;
    if (t > cap) t = cap;
    if (t < 1) t = 1;
    return (int)t;
}

static void launch_down(Ws& w, int l, cudaStream_t st) {
    Level &v = w.lev[l], &c = w.lev[l + 1];
    if (v.ny * v.nx < (int64_t)2000000000)
        down_kernel<int><<<fused_grid(v.ny, v.nx), kBlock, 0, st>>>(v.b, v.m, v.x, c.m, c.b, v.ny, v.nx, c.ny, c.nx, w.has_above,
                                                                    w.has_below);
    else
        down_kernel<int64_t><<<fused_grid(v.ny, v.nx), kBlock, 0, st>>>(v.b, v.m, v.x, c.m, c.b, v.ny, v.nx, c.ny, c.nx,
                                                                        w.has_above, w.has_below);
}
// where the level-0 up leg accumulates b . z (rows [lo, hi) of the level-0 grid); slot == nullptr: nowhere
struct RzTarget {
    double* slot = nullptr;
    int64_t lo = 0, hi = 0;
};
static void launch_up(Ws& w, int l, cudaStream_t st, RzTarget rz = RzTarget()) {
    Level &v = w.lev[l], &c = w.lev[l + 1];
    const bool small = v.ny * v.nx < (int64_t)2000000000;
    const int g = fused_grid(v.ny, v.nx);
#define SMRF_UP(RZ, I, SLOT, LO, HI) \
    up_kernel<RZ, I><<<g, kBlock, 0, st>>>(v.x, c.y, v.b, v.m, v.y, v.ny, v.nx, c.nx, w.has_above, w.has_below, SLOT, LO, HI)
    if (l == 0 && rz.slot) {
        if (small) SMRF_UP(true, int, rz.slot, rz.lo, rz.hi);
        else SMRF_UP(true, int64_t, rz.slot, rz.lo, rz.hi);
    } else {
        if (small) SMRF_UP(false, int, nullptr, 0, 0);
        else SMRF_UP(false, int64_t, nullptr, 0, 0);
    }
#undef SMRF_UP
}

// Parts of one V(3,3) cycle.  part 0: down legs of levels [0, split) (leaves lev[split].b);
// part 1: everything from level `split` down to the coarsest and back (leaves lev[split].y);
// part 2: up legs of levels split-1 .. 0 (leaves lev[0].y).  split == 0 with all three parts
// is the whole cycle.  The row-band solver runs parts 0 and 2 on its band and part 1 on a
// hierarchy of the global coarse grid that every rank holds (neilpy_b200/distributed.py).
// Inside part 1 the levels from tail_level() on run in one launch (tail_kernel).
static void vcycle_part(Ws& w, int split, int part, cudaStream_t st, int* launches, RzTarget rz = RzTarget()) {
    const int L = w.nlev;
    if (split > L - 1) split = L - 1;
    int n = 0;
    if (part == 0) {
        for (int l = 0; l < split; ++l, ++n) launch_down(w, l, st);
    } else if (part == 1) {
        int lt = tail_level(w);
        if (lt < split) lt = split;
        for (int l = split; l < lt; ++l, ++n) launch_down(w, l, st);
        tail_kernel<<<1, kTailThreads, 0, st>>>(w, lt);
        ++n;
        for (int l = lt - 1; l >= split; --l, ++n) launch_up(w, l, st, rz);
    } else {
        for (int l = split - 1; l >= 0; --l, ++n) launch_up(w, l, st, rz);
    }
    *launches += n;
}

// z = M^-1 b0: one whole cycle; the level-0 result is left in lev[0].y.  rz_k >= 0: rz[rz_k] = b0 . z comes with it
// (the level-0 up leg forms it), unless the hierarchy is a single level (then the caller runs rz_kernel).
static const float* vcycle(Ws& w, cudaStream_t st, int* launches, int rz_k = -1) {
    RzTarget rz;
    if (rz_k >= 0) { rz.slot = &w.sc->rz[rz_k]; rz.lo = 0; rz.hi = w.lev[0].ny; }
    vcycle_part(w, 0, 1, st, launches, rz);
    return w.lev[0].y;
}
static bool vcycle_forms_rz(const Ws& w) { return tail_level(w) >= 1 && w.nlev >= 2; }

}  // namespace inpaint
}  // namespace smrf

using namespace smrf;
using namespace smrf::inpaint;

static int g1_for(int64_t n) {
    int g1 = (int)((n + kBlock - 1) / kBlock);
    int cap = num_sms() * 16;
    return g1 > cap ? cap : (g1 < 1 ? 1 : g1);
}
static bool use_jacobi() {
    const char* env = getenv("SMRF_INPAINT_PRECOND");
    return env && strcmp(env, "jacobi") == 0;
}
static int check_ws(const char* fn, void* workspace, size_t bytes, int64_t ny, int64_t nx, int has_above, int has_below,
                    Ws* w) {
    if (!workspace || ny <= 0 || nx <= 0) {
        set_error("%s: null workspace or empty grid", fn);
        return SMRF_E_ARG;
    }
    if (bytes < carve(nullptr, ny, nx, nullptr)) {
        set_error("%s: workspace %zu < %zu bytes", fn, bytes, carve(nullptr, ny, nx, nullptr));
        return SMRF_E_WORKSPACE;
    }
    carve(workspace, ny, nx, w);
    w->has_above = has_above ? 1 : 0;
    w->has_below = has_below ? 1 : 0;
    return 0;
}

extern "C" {

size_t smrf_inpaint_workspace_bytes(int64_t ny, int64_t nx) {
    if (ny <= 0 || nx <= 0) return 0;
    return carve(nullptr, ny, nx, nullptr);
}

int smrf_inpaint_layout(int64_t ny, int64_t nx, int64_t* out8_host) {
    SMRF_CHECK_ARG(out8_host && ny > 0 && nx > 0, "bad argument");
    Ws w;
    carve(nullptr, ny, nx, &w);   // pointers relative to a null base = byte offsets
    out8_host[0] = (int64_t)(uintptr_t)w.u;
    out8_host[1] = (int64_t)(uintptr_t)w.p;
    out8_host[2] = (int64_t)(uintptr_t)w.lev[0].m;
    out8_host[3] = (int64_t)(uintptr_t)&w.sc->rz[0];
    out8_host[4] = (int64_t)(uintptr_t)&w.sc->pq[0];
    out8_host[5] = (int64_t)(uintptr_t)&w.sc->rmax[0];
    out8_host[6] = (int64_t)(uintptr_t)&w.sc->sum_known;   // {sum_known (f64), n_known (u64), n_unknown (u64)}
    out8_host[7] = (int64_t)kMaxIter;
    return 0;
}

int smrf_inpaint_setup(const void* grid, int64_t ny, int64_t nx, int dtype, void* workspace, size_t workspace_bytes,
                       int has_above, int has_below, void* stream) {
    SMRF_CHECK_ARG(grid, "null grid");
    SMRF_CHECK_ARG(dtype == SMRF_F32 || dtype == SMRF_F64, "bad dtype");
    Ws w;
    if (int rc = check_ws("smrf_inpaint_setup", workspace, workspace_bytes, ny, nx, has_above, has_below, &w)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = ny * nx;
    SMRF_CUDA(cudaMemsetAsync(w.sc, 0, sizeof(Scalars), st));
    if (dtype == SMRF_F32) scan_kernel<float><<<g1_for(n), kBlock, 0, st>>>((const float*)grid, w.lev[0].m, n, w.sc);
    else scan_kernel<double><<<g1_for(n), kBlock, 0, st>>>((const double*)grid, w.lev[0].m, n, w.sc);
    int launches = 1;
    if (!use_jacobi()) {
        for (int l = 0; l + 1 < w.nlev; ++l) {
            Level &f = w.lev[l], &c = w.lev[l + 1];
            coarsen_kernel<<<tile_grid(c.ny, c.nx), kBlock, 0, st>>>(f.m, c.m, f.ny, f.nx, c.ny, c.nx);
            ++launches;
        }
    }
    SMRF_LAUNCH_CHECK();
    count_launches(launches);
    return 0;
}

int smrf_inpaint_start(const void* grid, int64_t ny, int64_t nx, int dtype, void* workspace, size_t workspace_bytes,
                       int has_above, int has_below, double guess, const void* guess_grid, int phase,
                       const double* u_above, const double* u_below, void* stream) {
    SMRF_CHECK_ARG(grid, "null grid");
    SMRF_CHECK_ARG(dtype == SMRF_F32 || dtype == SMRF_F64, "bad dtype");
    SMRF_CHECK_ARG(phase == 0 || phase == 1, "bad phase");
    Ws w;
    if (int rc = check_ws("smrf_inpaint_start", workspace, workspace_bytes, ny, nx, has_above, has_below, &w)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = ny * nx;
    if (phase == 0) {   // u = known value or the starting guess (the caller exchanges boundary rows of u next)
        // init_u_kernel reads the guess from the statistics block: store it there as a mean over one cell
        double g = guess;
        unsigned long long one = 1;
        SMRF_CUDA(cudaMemcpyAsync(&w.sc->sum_known, &g, 8, cudaMemcpyHostToDevice, st));
        SMRF_CUDA(cudaMemcpyAsync(&w.sc->n_known, &one, 8, cudaMemcpyHostToDevice, st));
        if (dtype == SMRF_F32) init_u_kernel<float><<<tile_grid(ny, nx), kBlock, 0, st>>>((const float*)grid, w.lev[0].m, w.u, ny, nx, w.sc, (const float*)guess_grid);
        else init_u_kernel<double><<<tile_grid(ny, nx), kBlock, 0, st>>>((const double*)grid, w.lev[0].m, w.u, ny, nx, w.sc, (const double*)guess_grid);
    } else {
        SMRF_CHECK_ARG((!has_above || u_above) && (!has_below || u_below), "missing halo row of u");
        residual0_kernel<<<tile_grid(ny, nx), kBlock, 0, st>>>(w, ny, nx, u_above, u_below);
    }
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int smrf_inpaint_step(int64_t ny, int64_t nx, void* workspace, size_t workspace_bytes, int has_above, int has_below,
                      int k, int phase, const float* z_ext, const double* p_above, const double* p_below,
                      const uint8_t* m_above, const uint8_t* m_below, void* stream) {
    SMRF_CHECK_ARG(k >= 0 && k < kMaxIter, "iteration index out of range");
    Ws w;
    if (int rc = check_ws("smrf_inpaint_step", workspace, workspace_bytes, ny, nx, has_above, has_below, &w)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int g2 = tile_grid(ny, nx);
    const bool jacobi = use_jacobi();
    int launches = 0;
    const float* z = z_ext ? z_ext : w.lev[0].y;   // where vcycle() leaves its result, unless the caller preconditions
    switch (phase) {
        case 0:   // z = M^-1 r (local V-cycle), rz[k] += r.z over this band
            if (jacobi) rz_kernel<true><<<g2, kBlock, 0, st>>>(w, nullptr, ny, nx, k);
            else if (vcycle_forms_rz(w)) {
                z = vcycle(w, st, &launches, k);          // rz[k] accumulated by the level-0 up leg
                --launches;
            } else {
                z = vcycle(w, st, &launches);
                rz_kernel<false><<<g2, kBlock, 0, st>>>(w, z, ny, nx, k);
            }
            ++launches;
            break;
        case 20:  // the caller has preconditioned (z_ext = M^-1 r on this band): rz[k] += r.z only
            SMRF_CHECK_ARG(z_ext, "phase 20 needs z_ext");
            rz_kernel<false><<<g2, kBlock, 0, st>>>(w, z, ny, nx, k);
            ++launches;
            break;
        case 1:   // p = z + (rz[k]/rz[k-1]) p      (rz[k] must be complete: all-reduced)
            if (jacobi) p_update_kernel<true><<<g2, kBlock, 0, st>>>(w, nullptr, ny, nx, k);
            else p_update_kernel<false><<<g2, kBlock, 0, st>>>(w, z, ny, nx, k);
            ++launches;
            break;
        case 2:   // q = A p with the neighbours' boundary rows of p, pq[k] += p.q over this band
            SMRF_CHECK_ARG((!has_above || (p_above && m_above)) && (!has_below || (p_below && m_below)), "missing halo row of p");
            apply_kernel<<<g2, kBlock, 0, st>>>(w, ny, nx, k, p_above, p_below, m_above, m_below);
            ++launches;
            break;
        case 3:   // u += alpha p, r -= alpha q, rmax[k+1] = max|r| over this band   (pq[k] complete)
            update_kernel<<<g2, kBlock, 0, st>>>(w, ny, nx, k);
            ++launches;
            break;
        default:
            SMRF_CHECK_ARG(false, "bad phase");
    }
    SMRF_LAUNCH_CHECK();
    count_launches(launches);
    return 0;
}

// ---- a multigrid hierarchy on its own (the replicated global coarse grid of the band solver) ----
int smrf_mg_level_layout(int64_t ny, int64_t nx, int level, int64_t* out6_host) {
    SMRF_CHECK_ARG(out6_host && ny > 0 && nx > 0 && level >= 0, "bad argument");
    Ws w;
    carve(nullptr, ny, nx, &w);
    SMRF_CHECK_ARG(level < w.nlev, "no such level");
    const Level& v = w.lev[level];
    out6_host[0] = v.ny; out6_host[1] = v.nx;
    out6_host[2] = (int64_t)(uintptr_t)v.m; out6_host[3] = (int64_t)(uintptr_t)v.x;
    out6_host[4] = (int64_t)(uintptr_t)v.y; out6_host[5] = (int64_t)(uintptr_t)v.b;
    return 0;
}

int smrf_mg_setup_mask(const uint8_t* mask, int64_t ny, int64_t nx, void* workspace, size_t workspace_bytes, void* stream) {
    SMRF_CHECK_ARG(mask, "null mask");
    Ws w;
    if (int rc = check_ws("smrf_mg_setup_mask", workspace, workspace_bytes, ny, nx, 0, 0, &w)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    SMRF_CUDA(cudaMemcpyAsync(w.lev[0].m, mask, (size_t)ny * nx, cudaMemcpyDeviceToDevice, st));
    int launches = 0;
    for (int l = 0; l + 1 < w.nlev; ++l) {
        Level &f = w.lev[l], &c = w.lev[l + 1];
        coarsen_kernel<<<tile_grid(c.ny, c.nx), kBlock, 0, st>>>(f.m, c.m, f.ny, f.nx, c.ny, c.nx);
        ++launches;
    }
    SMRF_LAUNCH_CHECK();
    count_launches(launches);
    return 0;
}

int smrf_mg_cycle_part(int64_t ny, int64_t nx, void* workspace, size_t workspace_bytes, int has_above, int has_below,
                       int split, int part, void* stream) {
    SMRF_CHECK_ARG(split >= 0 && part >= 0 && part <= 2, "bad split / part");
    Ws w;
    if (int rc = check_ws("smrf_mg_cycle_part", workspace, workspace_bytes, ny, nx, has_above, has_below, &w)) return rc;
    int launches = 0;
    vcycle_part(w, split, part, (cudaStream_t)stream, &launches);
    SMRF_LAUNCH_CHECK();
    count_launches(launches);
    return 0;
}

int smrf_mg_cycle_up_rz(int64_t ny, int64_t nx, void* workspace, size_t workspace_bytes, int has_above, int has_below,
                        int split, double* rz_slot, int64_t row_lo, int64_t row_hi, void* stream) {
    SMRF_CHECK_ARG(split >= 1 && rz_slot && 0 <= row_lo && row_lo <= row_hi && row_hi <= ny, "bad split / rows / slot");
    Ws w;
    if (int rc = check_ws("smrf_mg_cycle_up_rz", workspace, workspace_bytes, ny, nx, has_above, has_below, &w)) return rc;
    int launches = 0;
    RzTarget rz;
    rz.slot = rz_slot; rz.lo = row_lo; rz.hi = row_hi;
    vcycle_part(w, split, 2, (cudaStream_t)stream, &launches, rz);
    SMRF_LAUNCH_CHECK();
    count_launches(launches);
    return 0;
}

int smrf_mg_vcycle(int64_t ny, int64_t nx, void* workspace, size_t workspace_bytes, void* stream) {
    Ws w;
    if (int rc = check_ws("smrf_mg_vcycle", workspace, workspace_bytes, ny, nx, 0, 0, &w)) return rc;
    int launches = 0;
    vcycle(w, (cudaStream_t)stream, &launches);
    SMRF_LAUNCH_CHECK();
    count_launches(launches);
    return 0;
}

int smrf_inpaint_finish(void* grid, int64_t ny, int64_t nx, int dtype, void* workspace, size_t workspace_bytes,
                        void* stream) {
    SMRF_CHECK_ARG(grid, "null grid");
    SMRF_CHECK_ARG(dtype == SMRF_F32 || dtype == SMRF_F64, "bad dtype");
    Ws w;
    if (int rc = check_ws("smrf_inpaint_finish", workspace, workspace_bytes, ny, nx, 0, 0, &w)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = ny * nx;
    if (dtype == SMRF_F32) writeback_kernel<float><<<g1_for(n), kBlock, 0, st>>>((float*)grid, w.lev[0].m, w.u, n);
    else writeback_kernel<double><<<g1_for(n), kBlock, 0, st>>>((double*)grid, w.lev[0].m, w.u, n);
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

// The compact solver, one phase at a time, for a row band (neilpy_b200/distributed.py); see include/smrf_b200.h.
int smrf_inpaint_compact(int op, const void* grid, int64_t ny, int64_t nx, int dtype, void* workspace, size_t workspace_bytes,
                         int has_above, int has_below, int64_t n_unknown, int k, double guess, const void* guess_grid,
                         const float* z, float* r_plane, const double* row_above, const double* row_below,
                         double* out_first, double* out_last, void* stream) {
    SMRF_CHECK_ARG(dtype == SMRF_F32 || dtype == SMRF_F64, "bad dtype");
    SMRF_CHECK_ARG(ny > 0 && nx > 0 && ny * nx < ((int64_t)1 << 31), "grid too large for the compact solver");
    SMRF_CHECK_ARG(n_unknown >= 0 && n_unknown <= ny * nx && k >= 0 && k < kMaxIter, "bad n_unknown / iteration");
    Ws w;
    if (int rc = check_ws("smrf_inpaint_compact", workspace, workspace_bytes, ny, nx, has_above, has_below, &w)) return rc;
    SMRF_CHECK_ARG(w.scan_tmp_bytes > 0, "no scan scratch");
    if (r_plane) w.lev[0].b = r_plane;   // the float32 residual goes straight into the caller's (ghost-extended) plane
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = ny * nx;
    const int nu = (int)n_unknown, ni = (int)n, iy = (int)ny, ix = (int)nx;
    int64_t gc = ((int64_t)nu + kBlock - 1) / kBlock;
    if (gc > (int64_t)num_sms() * 16) gc = (int64_t)num_sms() * 16;
    const int g = (int)(gc < 1 ? 1 : gc);
    const int gr = (int)((nx + kBlock - 1) / kBlock);
    int launches = 1;
    switch (op) {
        case 0: {   // index maps (after smrf_inpaint_setup) and a zeroed float32 residual plane
            size_t tb = w.scan_tmp_bytes;
            SMRF_CUDA(cub::DeviceScan::ExclusiveSum(w.scan_tmp, tb, (const uint8_t*)w.lev[0].m, w.pos, ni, st));
            compact_build_kernel<<<g1_for(n), kBlock, 0, st>>>(w.lev[0].m, w.pos, w.idx, ni);
            SMRF_CUDA(cudaMemsetAsync(w.lev[0].b, 0, (size_t)n * sizeof(float), st));
            launches = 2;
            break;
        }
        case 1: {   // starting guess (the mean of the known cells travels in the statistics block, as in smrf_inpaint_start)
            SMRF_CHECK_ARG(grid, "null grid");
            unsigned long long one = 1;
            SMRF_CUDA(cudaMemcpyAsync(&w.sc->sum_known, &guess, 8, cudaMemcpyHostToDevice, st));
            SMRF_CUDA(cudaMemcpyAsync(&w.sc->n_known, &one, 8, cudaMemcpyHostToDevice, st));
            if (nu == 0) { launches = 0; break; }
            if (dtype == SMRF_F32) compact_init_kernel<float><<<g, kBlock, 0, st>>>((const float*)grid, w, iy, ix, nu, (const float*)guess_grid);
            else compact_init_kernel<double><<<g, kBlock, 0, st>>>((const double*)grid, w, iy, ix, nu, (const double*)guess_grid);
            break;
        }
        case 2:     // boundary rows of u (known values included) for the neighbours
            SMRF_CHECK_ARG(grid, "null grid");
            if (dtype == SMRF_F32) compact_rows_kernel<float><<<gr, kBlock, 0, st>>>(w.u, (const float*)grid, w, iy, ix, out_first, out_last);
            else compact_rows_kernel<double><<<gr, kBlock, 0, st>>>(w.u, (const double*)grid, w, iy, ix, out_first, out_last);
            break;
        case 3:     // r = b - A u with the neighbours' rows of u, rmax[0]
            SMRF_CHECK_ARG(grid && (!has_above || row_above) && (!has_below || row_below), "missing halo row of u");
            if (nu == 0) { launches = 0; break; }
            if (dtype == SMRF_F32) compact_residual0_kernel<float><<<g, kBlock, 0, st>>>((const float*)grid, w, iy, ix, nu, row_above, row_below);
            else compact_residual0_kernel<double><<<g, kBlock, 0, st>>>((const double*)grid, w, iy, ix, nu, row_above, row_below);
            break;
        case 4:     // p = z + beta p (rz[k] complete), then its boundary rows for the neighbours
            SMRF_CHECK_ARG(z, "null z");
            if (nu) compact_p_kernel<<<g, kBlock, 0, st>>>(w, z, nu, k);
            compact_rows_kernel<float><<<gr, kBlock, 0, st>>>(w.p, nullptr, w, iy, ix, out_first, out_last);
            launches = 2;
            break;
        case 5:     // q = A p with the neighbours' rows of p, pq[k]
            SMRF_CHECK_ARG((!has_above || row_above) && (!has_below || row_below), "missing halo row of p");
            if (nu == 0) { launches = 0; break; }
            compact_apply_kernel<<<g, kBlock, 0, st>>>(w, iy, ix, nu, k, row_above, row_below);
            break;
        case 6:     // u += alpha p, r -= alpha q (pq[k] complete), rmax[k+1]
            if (nu == 0) { launches = 0; break; }
            compact_update_kernel<<<g, kBlock, 0, st>>>(w, nu, k);
            break;
        case 7:     // the solution into the NaN cells of the grid
            SMRF_CHECK_ARG(grid, "null grid");
            if (nu == 0) { launches = 0; break; }
            if (dtype == SMRF_F32) compact_writeback_kernel<float><<<g, kBlock, 0, st>>>((float*)const_cast<void*>(grid), w, nu);
            else compact_writeback_kernel<double><<<g, kBlock, 0, st>>>((double*)const_cast<void*>(grid), w, nu);
            break;
        default:
            SMRF_CHECK_ARG(false, "bad op");
    }
    SMRF_LAUNCH_CHECK();
    count_launches(launches);
    return 0;
}

int smrf_inpaint(void* grid, int64_t ny, int64_t nx, int dtype, uint8_t* unknown, const void* guess, void* workspace,
                 size_t workspace_bytes, double tol, int max_iter, double* info_host, void* stream) {
    SMRF_CHECK_ARG(grid && workspace, "null pointer");
    SMRF_CHECK_ARG(ny > 0 && nx > 0, "empty grid");
    SMRF_CHECK_ARG(dtype == SMRF_F32 || dtype == SMRF_F64, "bad dtype");
    SMRF_CHECK_ARG(tol >= 0.0, "negative tol");
    if (max_iter <= 0 || max_iter > kMaxIter - 1) max_iter = kMaxIter - 1;
    const bool jacobi = use_jacobi();
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = ny * nx;
    if (int rc = smrf_inpaint_setup(grid, ny, nx, dtype, workspace, workspace_bytes, 0, 0, stream)) return rc;
    Ws w;
    carve(workspace, ny, nx, &w);
    struct { double sum; unsigned long long nk, nu; } stats;
    SMRF_CUDA(cudaMemcpyAsync(&stats, &w.sc->sum_known, sizeof(stats), cudaMemcpyDeviceToHost, st));
    SMRF_CUDA(cudaStreamSynchronize(st));
    const unsigned long long n_unknown = stats.nu;
    if (unknown) SMRF_CUDA(cudaMemcpyAsync(unknown, w.lev[0].m, (size_t)n, cudaMemcpyDeviceToDevice, st));
    int it = 0;
    double rmax = 0.0;
    static const bool no_compact = []() { const char* e = getenv("SMRF_INPAINT_COMPACT"); return e && e[0] == '0'; }();
    const bool compact = !jacobi && !no_compact && vcycle_forms_rz(w) && n < ((int64_t)1 << 31) && w.scan_tmp_bytes > 0;
    if (n_unknown > 0 && compact) {
        // ---- compact CG: u, r, p, q hold the unknown cells only (see compact_build_kernel)
        const int nu = (int)n_unknown, ni = (int)n, iy = (int)ny, ix = (int)nx;
        int64_t gc = ((int64_t)nu + kBlock - 1) / kBlock;
        if (gc > (int64_t)num_sms() * 16) gc = (int64_t)num_sms() * 16;
        const int g = (int)(gc < 1 ? 1 : gc), gn = g1_for(n);
        size_t tb = w.scan_tmp_bytes;
        SMRF_CUDA(cub::DeviceScan::ExclusiveSum(w.scan_tmp, tb, (const uint8_t*)w.lev[0].m, w.pos, ni, st));
        compact_build_kernel<<<gn, kBlock, 0, st>>>(w.lev[0].m, w.pos, w.idx, ni);
        SMRF_CUDA(cudaMemsetAsync(w.lev[0].b, 0, (size_t)n * sizeof(float), st));
        if (dtype == SMRF_F32) {
            compact_init_kernel<float><<<g, kBlock, 0, st>>>((const float*)grid, w, iy, ix, nu, (const float*)guess);
            compact_residual0_kernel<float><<<g, kBlock, 0, st>>>((const float*)grid, w, iy, ix, nu, nullptr, nullptr);
        } else {
            compact_init_kernel<double><<<g, kBlock, 0, st>>>((const double*)grid, w, iy, ix, nu, (const double*)guess);
            compact_residual0_kernel<double><<<g, kBlock, 0, st>>>((const double*)grid, w, iy, ix, nu, nullptr, nullptr);
        }
        SMRF_LAUNCH_CHECK();
        int launches = 4;
        unsigned long long bits = 0;
        SMRF_CUDA(cudaMemcpyAsync(&bits, &w.sc->rmax[0], 8, cudaMemcpyDeviceToHost, st));
        SMRF_CUDA(cudaStreamSynchronize(st));
        memcpy(&rmax, &bits, 8);
        double r_prev = rmax;
        int it_prev = 0, burst = 4;
        while (rmax > tol && it < max_iter) {
            if (it + burst > max_iter) burst = max_iter - it;
            for (int j = 0; j < burst; ++j, ++it) {
                const float* z = vcycle(w, st, &launches, it);          // z = M^-1 r, rz[it] from the level-0 up leg
                compact_p_kernel<<<g, kBlock, 0, st>>>(w, z, nu, it);
                compact_apply_kernel<<<g, kBlock, 0, st>>>(w, iy, ix, nu, it, nullptr, nullptr);
                compact_update_kernel<<<g, kBlock, 0, st>>>(w, nu, it);
                launches += 3;
            }
            SMRF_LAUNCH_CHECK();
            SMRF_CUDA(cudaMemcpyAsync(&bits, &w.sc->rmax[it], 8, cudaMemcpyDeviceToHost, st));
            SMRF_CUDA(cudaStreamSynchronize(st));
            memcpy(&rmax, &bits, 8);
            if (!(rmax < INFINITY)) break;   // NaN / inf: give up rather than spin (the caller reports it)
            int next = kCheckEvery;
            if (rmax > tol && rmax > 0.0 && rmax < r_prev && it > it_prev) {
                const double rate = (log(rmax) - log(r_prev)) / (double)(it - it_prev);   // < 0
                const double left = (log(tol) - log(rmax)) / rate;
                next = left < 1.0 ? 1 : (left > (double)kCheckEvery ? kCheckEvery : (int)ceil(left));
            }
            r_prev = rmax; it_prev = it; burst = next;
        }
        if (dtype == SMRF_F32) compact_writeback_kernel<float><<<g, kBlock, 0, st>>>((float*)grid, w, nu);
        else compact_writeback_kernel<double><<<g, kBlock, 0, st>>>((double*)grid, w, nu);
        SMRF_LAUNCH_CHECK();
        count_launches(launches + 1);
        SMRF_CUDA(cudaStreamSynchronize(st));
    } else if (n_unknown > 0) {
        const double mean = stats.nk ? stats.sum / (double)stats.nk : 0.0;
        if (int rc = smrf_inpaint_start(grid, ny, nx, dtype, workspace, workspace_bytes, 0, 0, mean, guess, 0, nullptr, nullptr, stream)) return rc;
        if (int rc = smrf_inpaint_start(grid, ny, nx, dtype, workspace, workspace_bytes, 0, 0, mean, nullptr, 1, nullptr, nullptr, stream)) return rc;
        unsigned long long bits = 0;
        SMRF_CUDA(cudaMemcpyAsync(&bits, &w.sc->rmax[0], 8, cudaMemcpyDeviceToHost, st));
        SMRF_CUDA(cudaStreamSynchronize(st));
        memcpy(&rmax, &bits, 8);
        // The host polls the residual between bursts of iterations; the burst length follows the
        // observed convergence rate so that few iterations run past the tolerance.
        double r_prev = rmax;
        int it_prev = 0, burst = jacobi ? 32 : 4;
        while (rmax > tol && it < max_iter) {
            if (it + burst > max_iter) burst = max_iter - it;
            for (int j = 0; j < burst; ++j, ++it)
                for (int ph = 0; ph < 4; ++ph)
                    if (int rc = smrf_inpaint_step(ny, nx, workspace, workspace_bytes, 0, 0, it, ph, nullptr, nullptr, nullptr, nullptr, nullptr, stream)) return rc;
            SMRF_CUDA(cudaMemcpyAsync(&bits, &w.sc->rmax[it], 8, cudaMemcpyDeviceToHost, st));
            SMRF_CUDA(cudaStreamSynchronize(st));
            memcpy(&rmax, &bits, 8);
            if (!(rmax < INFINITY)) break;   // NaN / inf: give up rather than spin (the caller reports it)
            const int cap = jacobi ? 64 : kCheckEvery;
            int next = cap;
            if (rmax > tol && rmax > 0.0 && rmax < r_prev && it > it_prev) {
                const double rate = (log(rmax) - log(r_prev)) / (double)(it - it_prev);   // < 0
                const double left = (log(tol) - log(rmax)) / rate;
                next = left < 1.0 ? 1 : (left > (double)cap ? cap : (int)ceil(left));
            }
            r_prev = rmax; it_prev = it; burst = next;
        }
        if (int rc = smrf_inpaint_finish(grid, ny, nx, dtype, workspace, workspace_bytes, stream)) return rc;
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
