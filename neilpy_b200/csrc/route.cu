// Row-band sharding: every point travels once to the rank that owns its row band.
//
// With the points of a band resident on its rank, the binning (create_dem, neilpy.py:1142-1161) and the
// interpolation + classification (neilpy.py:1772-1795) are band-local: no rank ever holds a full-grid replica or
// the coefficient grids of another band.  The owner of a point is the band of floor(row), with the row computed by
// exactly the arithmetic the binning uses (affine_apply, two roundings per product/sum), so a routed point always
// lands inside the rows its new owner bins.
//   route_plan   dest[i] = owner band of point i (world = out of grid / non-finite: nobody's; such points make the
//                binning raise, as np.ravel_multi_index does in the reference), counts[d] += 1
//   route_pack   points are written to the send buffer grouped by destination (order inside a group is arbitrary;
//                perm[i] = slot of point i), as one float4 stream or three float64 columns
//   route_unpack out[i] = back[perm[i]]: the classification that came back, in the caller's point order
#include "common.cuh"

namespace smrf {
namespace route {

constexpr int kBlock = 256;
constexpr int kItems = 8;            // points per thread
constexpr int kMaxWorld = 64;

template <int FMT>
__global__ void __launch_bounds__(kBlock) plan_kernel(PointLoader<FMT> pts, int64_t n, Inv6 inv, int64_t ny, int64_t nx,
                                                      int64_t per, int world, uint8_t* __restrict__ dest,
                                                      unsigned long long* __restrict__ counts) {
    __shared__ unsigned int hist[kMaxWorld + 1];
    for (int i = threadIdx.x; i <= world; i += kBlock) hist[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kBlock * kItems;
#pragma unroll
    for (int e = 0; e < kItems; ++e) {
        const int64_t i = base + (int64_t)e * kBlock + threadIdx.x;
        if (i < n) {
            double x, y;
            pts.xy(i, x, y);
            double c, r;
            affine_apply(inv, x, y, c, r);
            c = floor(c);
            r = floor(r);
            int d = world;           // not in the grid (NaN compares false)
            if (c >= 0.0 && c < (double)nx && r >= 0.0 && r < (double)ny) {
                d = (int)((int64_t)r / per);
                if (d >= world) d = world - 1;
            }
            dest[i] = (uint8_t)d;
            atomicAdd(&hist[d], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= world; i += kBlock)
        if (hist[i]) atomicAdd(&counts[i], (unsigned long long)hist[i]);
}

// cursors[d] = next free slot of destination d (initialised by the caller to the exclusive prefix sums of counts)
template <int FMT>
__global__ void __launch_bounds__(kBlock) pack_kernel(PointLoader<FMT> pts, int64_t n, int world,
                                                      const uint8_t* __restrict__ dest,
                                                      unsigned long long* __restrict__ cursors, float4* __restrict__ out4,
                                                      double* __restrict__ ox, double* __restrict__ oy,
                                                      double* __restrict__ oz, int64_t* __restrict__ perm) {
    __shared__ unsigned int hist[kMaxWorld + 1];
    __shared__ unsigned long long first[kMaxWorld + 1];
    for (int i = threadIdx.x; i <= world; i += kBlock) hist[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kBlock * kItems;
    unsigned int rank_in_block[kItems];
    int d[kItems];
#pragma unroll
    for (int e = 0; e < kItems; ++e) {
        const int64_t i = base + (int64_t)e * kBlock + threadIdx.x;
        d[e] = -1;
        if (i < n) {
            d[e] = dest[i];
            rank_in_block[e] = atomicAdd(&hist[d[e]], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= world; i += kBlock)
        first[i] = hist[i] ? atomicAdd(&cursors[i], (unsigned long long)hist[i]) : 0ull;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < kItems; ++e) {
        const int64_t i = base + (int64_t)e * kBlock + threadIdx.x;
        if (d[e] < 0) continue;
        const int64_t slot = (int64_t)(first[d[e]] + rank_in_block[e]);
        perm[i] = slot;
        double x, y, z;
        pts.xyz(i, x, y, z);
        if (out4) out4[slot] = make_float4((float)x, (float)y, (float)z, 0.f);      // float32 inputs: exact
        else { ox[slot] = x; oy[slot] = y; oz[slot] = z; }
    }
}

__global__ void __launch_bounds__(kBlock) unpack_kernel(const uint8_t* __restrict__ back, const int64_t* __restrict__ perm,
                                                        int64_t n, uint8_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
        out[i] = back[perm[i]];
}

}  // namespace route
}  // namespace smrf

using namespace smrf;
using namespace smrf::route;

template <int FMT>
static PointLoader<FMT> loader(const void* x, const void* y, const void* z) {
    if constexpr (FMT == SMRF_PTS_SOA_F64) return PointLoader<FMT>{(const double*)x, (const double*)y, (const double*)z};
    else if constexpr (FMT == SMRF_PTS_XYZW_F32) return PointLoader<FMT>{(const float4*)x, nullptr, nullptr};
    else return PointLoader<FMT>{(const float*)x, (const float*)y, (const float*)z};
}

extern "C" {

int smrf_route_plan(const void* x, const void* y, int64_t n, int point_fmt, const double* inv6_host, int64_t ny, int64_t nx,
                    int64_t rows_per_band, int world, uint8_t* dest, int64_t* counts, void* stream) {
    SMRF_CHECK_ARG(inv6_host && counts, "null pointer");
    SMRF_CHECK_ARG(n == 0 || (x && dest), "null pointer");
    SMRF_CHECK_ARG(point_fmt == SMRF_PTS_XYZW_F32 || n == 0 || y, "y null");
    SMRF_CHECK_ARG(ny > 0 && nx > 0 && rows_per_band > 0 && world >= 1 && world <= kMaxWorld, "bad size");
    cudaStream_t st = (cudaStream_t)stream;
    SMRF_CUDA(cudaMemsetAsync(counts, 0, (size_t)(world + 1) * 8, st));
    if (n == 0) return 0;
    Inv6 inv{inv6_host[0], inv6_host[1], inv6_host[2], inv6_host[3], inv6_host[4], inv6_host[5]};
    const int64_t per_block = (int64_t)kBlock * kItems;
    const unsigned g = (unsigned)((n + per_block - 1) / per_block);
    unsigned long long* c = (unsigned long long*)counts;
    if (point_fmt == SMRF_PTS_SOA_F64)
        plan_kernel<SMRF_PTS_SOA_F64><<<g, kBlock, 0, st>>>(loader<SMRF_PTS_SOA_F64>(x, y, nullptr), n, inv, ny, nx, rows_per_band, world, dest, c);
    else if (point_fmt == SMRF_PTS_XYZW_F32)
        plan_kernel<SMRF_PTS_XYZW_F32><<<g, kBlock, 0, st>>>(loader<SMRF_PTS_XYZW_F32>(x, nullptr, nullptr), n, inv, ny, nx, rows_per_band, world, dest, c);
    else if (point_fmt == SMRF_PTS_SOA_F32)
        plan_kernel<SMRF_PTS_SOA_F32><<<g, kBlock, 0, st>>>(loader<SMRF_PTS_SOA_F32>(x, y, nullptr), n, inv, ny, nx, rows_per_band, world, dest, c);
    else SMRF_CHECK_ARG(false, "bad point_fmt");
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int smrf_route_pack(const void* x, const void* y, const void* z, int64_t n, int point_fmt, int world, const uint8_t* dest,
                    int64_t* cursors, void* out_xyzw, double* out_x, double* out_y, double* out_z, int64_t* perm,
                    void* stream) {
    SMRF_CHECK_ARG(world >= 1 && world <= kMaxWorld, "bad world");
    if (n == 0) return 0;
    SMRF_CHECK_ARG(x && dest && cursors && perm, "null pointer");
    SMRF_CHECK_ARG(point_fmt == SMRF_PTS_XYZW_F32 || (y && z), "y/z null");
    SMRF_CHECK_ARG(out_xyzw || (out_x && out_y && out_z), "no output buffer");
    SMRF_CHECK_ARG(!out_xyzw || point_fmt != SMRF_PTS_SOA_F64, "float64 points travel as three float64 columns");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t per_block = (int64_t)kBlock * kItems;
    const unsigned g = (unsigned)((n + per_block - 1) / per_block);
    unsigned long long* c = (unsigned long long*)cursors;
    float4* o4 = (float4*)out_xyzw;
    if (point_fmt == SMRF_PTS_SOA_F64)
        pack_kernel<SMRF_PTS_SOA_F64><<<g, kBlock, 0, st>>>(loader<SMRF_PTS_SOA_F64>(x, y, z), n, world, dest, c, o4, out_x, out_y, out_z, perm);
    else if (point_fmt == SMRF_PTS_XYZW_F32)
        pack_kernel<SMRF_PTS_XYZW_F32><<<g, kBlock, 0, st>>>(loader<SMRF_PTS_XYZW_F32>(x, nullptr, nullptr), n, world, dest, c, o4, out_x, out_y, out_z, perm);
    else if (point_fmt == SMRF_PTS_SOA_F32)
        pack_kernel<SMRF_PTS_SOA_F32><<<g, kBlock, 0, st>>>(loader<SMRF_PTS_SOA_F32>(x, y, z), n, world, dest, c, o4, out_x, out_y, out_z, perm);
    else SMRF_CHECK_ARG(false, "bad point_fmt");
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int smrf_route_unpack(const uint8_t* back, const int64_t* perm, int64_t n, uint8_t* out, void* stream) {
    if (n == 0) return 0;
    SMRF_CHECK_ARG(back && perm && out, "null pointer");
    int64_t g = (n + kBlock - 1) / kBlock;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (g > cap) g = cap;
    unpack_kernel<<<(unsigned)g, kBlock, 0, (cudaStream_t)stream>>>(back, perm, n, out);
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

}  // extern "C"
