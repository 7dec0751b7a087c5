// LAS point records on the device (the step before the SMRF path):
//   smrf_las_decode       records -> x, y, z float64 columns (+ the raw classification byte)
//                         x = X*scale + offset, product and sum rounded separately
//                         (neilpy.py:1056-1059: int32 column * float + float in pandas)
//   smrf_las_write_class  classification = 2*(1 - is_object_point), written into the records
//                         (examples/smrf/SMRF Classification using laspy to read and write.ipynb, cell 5)
//
// Byte work, HBM-bound: a record is L = 20..67 packed bytes of which 13 are wanted, but every
// sector of the stream is touched, so the algorithmic traffic is L + 24 (+1) bytes per point.
// Records are not 4-byte aligned in general (L = 26, 57, 63, ...), so a CTA stages a tile of
// 512 records (512*L bytes, 16-byte aligned for every L) into shared memory with 16-byte
// cp.async copies, one tile ahead of the one being unpacked, and the unpacking reads shared
// memory bytewise (or as words when L % 4 == 0).  The column stores are fully coalesced.
#include "common.cuh"

namespace smrf {
namespace las {

constexpr int kTile = 512;        // records per tile; a multiple of 16 keeps every tile 16-byte aligned
constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ int32_t le32(const uint8_t* p) {
    return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
}

// global bytes [b0, b0 + nbytes) -> dst; whole 16-byte chunks by cp.async, the ragged end of the
// last tile of the stream bytewise
__device__ __forceinline__ void stage_tile(uint8_t* dst, const uint8_t* __restrict__ rec, int64_t b0, int nbytes) {
    const int chunks = nbytes >> 4;
    for (int c = threadIdx.x; c < chunks; c += kThreads)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst + 16 * c)), "l"(rec + b0 + 16 * c) : "memory");
    for (int b = (chunks << 4) + threadIdx.x; b < nbytes; b += kThreads) dst[b] = rec[b0 + b];
}

template <bool WORDS>
__global__ void __launch_bounds__(kThreads) decode_kernel(const uint8_t* __restrict__ rec, int64_t n, int L, double sx,
                                                          double sy, double sz, double ox, double oy, double oz,
                                                          double* __restrict__ x, double* __restrict__ y,
                                                          double* __restrict__ z, uint8_t* __restrict__ cls,
                                                          int class_offset, int64_t n_tiles) {
    extern __shared__ __align__(16) uint8_t tiles[];
    const int tile_bytes = kTile * L;
    int buf = 0;
    int64_t t = blockIdx.x;
    if (t < n_tiles) {
        const int64_t r0 = t * kTile;
        const int cnt = (int)((n - r0) < kTile ? (n - r0) : kTile);
        stage_tile(tiles, rec, r0 * L, cnt * L);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (; t < n_tiles; t += gridDim.x) {
        const int64_t tn = t + gridDim.x;
        if (tn < n_tiles) {
            const int64_t rn = tn * kTile;
            const int cn = (int)((n - rn) < kTile ? (n - rn) : kTile);
            stage_tile(tiles + (buf ^ 1) * tile_bytes, rec, rn * L, cn * L);
        }
        asm volatile("cp.async.commit_group;\n cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const int64_t r0 = t * kTile;
        const int cnt = (int)((n - r0) < kTile ? (n - r0) : kTile);
        const uint8_t* tile = tiles + buf * tile_bytes;
#pragma unroll
        for (int k = 0; k < kTile / kThreads; ++k) {
            const int r = threadIdx.x + k * kThreads;
            if (r < cnt) {
                const uint8_t* p = tile + r * L;
                int32_t xi, yi, zi;
                if (WORDS) {
                    const int32_t* w = reinterpret_cast<const int32_t*>(p);
                    xi = w[0]; yi = w[1]; zi = w[2];
                } else {
                    xi = le32(p); yi = le32(p + 4); zi = le32(p + 8);
                }
                x[r0 + r] = __dadd_rn(__dmul_rn((double)xi, sx), ox);
                y[r0 + r] = __dadd_rn(__dmul_rn((double)yi, sy), oy);
                z[r0 + r] = __dadd_rn(__dmul_rn((double)zi, sz), oz);
                if (cls) cls[r0 + r] = p[class_offset];
            }
        }
        __syncthreads();          // the tile is free before the next iteration stages over it
        buf ^= 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(256) write_class_kernel(uint8_t* __restrict__ rec, int64_t n, int L, int class_offset,
                                                          int keep_mask, const uint8_t* __restrict__ is_object,
                                                          int ground_code, int object_code) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint8_t* p = rec + i * L + class_offset;
        const int code = is_object[i] ? object_code : ground_code;
        *p = (uint8_t)((*p & keep_mask) | code);
    }
}

}  // namespace las
}  // namespace smrf

using namespace smrf;

extern "C" {

int smrf_las_decode(const uint8_t* records, int64_t n, int record_length, const double* scale3_host,
                    const double* offset3_host, double* x, double* y, double* z, uint8_t* classification,
                    int class_offset, void* stream) {
    SMRF_CHECK_ARG(n >= 0, "negative point count");
    if (n == 0) return 0;
    SMRF_CHECK_ARG(records && x && y && z, "null pointer");
    SMRF_CHECK_ARG(scale3_host && offset3_host, "null scale/offset");
    SMRF_CHECK_ARG(record_length >= 12 && record_length <= 1024, "record length must be 12..1024 bytes");
    SMRF_CHECK_ARG(((uintptr_t)records & 15) == 0, "records must be 16-byte aligned");
    SMRF_CHECK_ARG(!classification || (class_offset >= 0 && class_offset < record_length), "class offset outside the record");
    const int64_t n_tiles = (n + las::kTile - 1) / las::kTile;
    const size_t smem = 2 * (size_t)las::kTile * record_length;
    SMRF_CHECK_ARG(smem <= 200 * 1024, "record too long for the staging tiles");
    const bool words = (record_length & 3) == 0;
    auto kern = words ? las::decode_kernel<true> : las::decode_kernel<false>;
    if (smem > 48 * 1024) SMRF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((220 * 1024) / smem);
    per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
    int64_t g = (int64_t)num_sms() * per_sm;
    if (g > n_tiles) g = n_tiles;
    kern<<<(unsigned)g, las::kThreads, smem, (cudaStream_t)stream>>>(
        records, n, record_length, scale3_host[0], scale3_host[1], scale3_host[2], offset3_host[0], offset3_host[1],
        offset3_host[2], x, y, z, classification, class_offset, n_tiles);
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int smrf_las_write_class(uint8_t* records, int64_t n, int record_length, int class_offset, int keep_mask,
                         const uint8_t* is_object_point, int ground_code, int object_code, void* stream) {
    SMRF_CHECK_ARG(n >= 0, "negative point count");
    if (n == 0) return 0;
    SMRF_CHECK_ARG(records && is_object_point, "null pointer");
    SMRF_CHECK_ARG(record_length >= 12 && class_offset >= 0 && class_offset < record_length, "class offset outside the record");
    SMRF_CHECK_ARG((keep_mask & ~0xff) == 0 && (ground_code & ~0xff) == 0 && (object_code & ~0xff) == 0, "codes must be bytes");
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (g > cap) g = cap;
    las::write_class_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(records, n, record_length, class_offset, keep_mask,
                                                                          is_object_point, ground_code, object_code);
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

}  // extern "C"
