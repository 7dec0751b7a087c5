// Register-marching fused disk opening for float32 grids (sm_100a).
//
//   this = dilate(erode(last, disk(W)), disk(W));  new = (last - this) > thr;  mask |= new
//
// One CTA owns a strip of XO output columns and a segment of rows and marches down
// it once.  disk(W) is used in chord form: row dy of the disk spans |dx| <= h(dy),
// h(dy) = floor(sqrt(W^2 - dy^2)), so
//     erode(Z)[y][x] = min_dy  rowmin_{h(dy)}(Z[y+dy])[x].
// Warp roles (8 warps, 256 threads so that each thread may hold up to 255 registers):
//   first pass  (4 warps)  rows of `last` arrive in the NZ-stage shared ring Zs by TMA
//                    (cp.async.bulk.tensor, one elected thread issues the row boxes of a
//                    batch and arms the stage's mbarrier with the byte count; cells outside
//                    the image are filled with NaN by the TMA unit, which min/max ignore --
//                    the reference's border rule with no branch); NZ-1 batches are in flight
//                    while one is consumed.  Each thread owns C adjacent columns.  For every
//                    incoming row it grows the horizontal window one cell per side per step
//                    (R_h = min3(R_{h-1}, Z[x-h], Z[x+h]), one FMNMX3 per column) and,
//                    whenever h equals the chord of some dy, folds R_h into the accumulator
//                    of output row (y - dy).  The 2W live output rows per column sit in
//                    registers.  A step reads accumulator file A and writes file B one (two)
//                    slot(s) lower, the next step reads B and writes A: the slot of an output
//                    row moves with the march, indices stay static, and no register is ever
//                    copied (the first versions shifted the file after every group: 18 % of
//                    the instructions).  Finished rows go to the shared ring Es (identity
//                    outside the image: the reference dilates an eroded image that does not
//                    exist there).
//   second pass (4 warps)  the same marching with max over Es; a finished row is compared with
//                    `last` (re-read through L2; float32 screen, float64 decision) and stored
//                    with 16-byte stores; mask / when_dropped bytes are written only where set.
// Es is handed over in batches of U rows through mbarriers (full/empty pairs); Zs is
// private to the first-pass warps (TMA full barriers + one named barrier per batch).
// Surfaces whose rows are not 16-byte aligned cannot be described to the TMA unit; they
// take the cp.async loader of the same kernel (Params::use_tma = 0).
// NEG = true opens -Z instead (the low-outlier pass): open(-Z) = -close(Z), so the roles
// swap min and max and the sign is applied in the epilogue -- no negated copy is made.
// The erosion->dilation intermediate never leaves the SM: HBM traffic per cell-window
// is one read of `last` (+ halo re-reads, L2 hits) and one write of `this`.
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through the runtime, no -lcuda)

#include <type_traits>

#include "opening.cuh"

namespace smrf {
namespace march {

constexpr int kRoleThreads = 128;
constexpr int kThreads = 2 * kRoleThreads;

template <int W_, int C_, bool PAIR_, int MINB_, int NZ_ = 4, int NB_ = 3>
struct CfgT {
    static constexpr int W = W_;
    // PAIR: two incoming rows are folded per step so that every accumulator update is one
    // 3-input min/max (acc, row u's chord, row u+1's chord) instead of two 2-input ones.
    static constexpr bool PAIR = PAIR_;
    static constexpr int C = C_;                                // adjacent columns per thread
    static constexpr int U = 4;                                 // rows per batch (ring stage / hand-over unit)
    static constexpr int S = 2 * W;                             // accumulator slots carried between steps, per column
    static constexpr int EW = kRoleThreads * C;                 // first-pass columns per CTA
    static constexpr int NL = 2 * W + C;                        // elements a thread reads per row
    static constexpr int NQ = (NL + C - 1) / C;                 // ... as C-wide vectors
    static constexpr int NEED = (kRoleThreads - 1) * C + NQ * C;          // floats of a ring row that are read
    static constexpr int NBOX = (NEED + 255) / 256;             // TMA boxes per row (a box is <= 256 elements wide)
    static constexpr int BW = ((NEED + NBOX - 1) / NBOX + 31) / 32 * 32;  // box width: TMA destinations are 128-byte aligned
    static constexpr int COLS = NBOX * BW;                      // ring row length (floats), a multiple of 32
    static constexpr int XO = ((EW - 2 * W) / 4) * 4;           // output columns per CTA (multiple of 4 and of C)
    // The TMA unit wants the first column of a box 16-byte aligned: strips start DX columns left of a multiple
    // of XO so that their ring starts at x0 - 2W = 0 mod 4 (odd radii: DX = 2; 4-wide global accesses then split in two)
    static constexpr int DX = (2 * W) % 4;
    static constexpr int I0 = (2 * W + U - 1) / U * U;          // warm-up rows per pass, whole batches
    static constexpr int NB = NB_;                              // batches in the Es ring
    static constexpr int NZ = NZ_;                              // batches in the Zs ring
    static constexpr int MINB = MINB_;                          // CTAs per SM the register budget is held to
    static constexpr uint32_t kStageBytes = (uint32_t)(U * COLS * sizeof(float));
    static constexpr size_t kSmemBytes = 128 + (size_t)(NZ + NB) * kStageBytes + (2 * NB + NZ) * sizeof(uint64_t);

    __host__ __device__ static constexpr int isqrt(int v) {
        int h = 0;
        while ((h + 1) * (h + 1) <= v) ++h;
        return h;
    }
    __host__ __device__ static constexpr int half(int dy) { return isqrt(W * W - dy * dy); }
};

// The shipped choice per radius comes from B200 sweeps of tools/march_sweep.cu (profiles/r2_march_sweep.log):
//   radii 1..6   this fused kernel (one read + one write of the surface per window; HBM matters here)
//   radii 7..72  two single-role passes (PassCfg below): the FMNMX work dominates, and a pass without the
//                second role's epilogue and hand-over keeps the ALU pipe busier than the fused form does.
template <int W>
struct Cfg : CfgT<W,
                  /*C=*/(W == 4 ? 4 : 2),
                  /*PAIR=*/true,
                  /*MINB=*/(W == 4 ? 1 : ((W == 3 || W == 5) ? 3 : 2)),
                  /*NZ=*/4, /*NB=*/3> {};

struct Params {
    const float* in;
    float* out;
    uint8_t* mask;
    uint8_t* when;
    int64_t ny, nx, pitch, row_lo, row_hi;   // pitch: row stride of in/out in elements (>= nx); mask/when are nx wide
    int seg;
    double thr;
    float thr_screen;                        // float32 differences <= this are certainly <= thr (see the epilogue)
    int widx, vec_ok, use_tma;
};

// f(integral_constant<B>), ..., f(integral_constant<E-1>) in order; split in halves so that the
// instantiation depth stays logarithmic (radius 72 unrolls 144 chords)
template <int B, int E, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (E - B == 1) {
        f(std::integral_constant<int, B>{});
    } else if constexpr (E - B > 1) {
        constexpr int M = B + (E - B) / 2;
        static_for<B, M>(f);
        static_for<M, E>(f);
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(b)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t a = smem_u32(b);
    // try_wait suspends the warp for a hardware-defined interval; the bound turns a lost hand-over (a bug) into a
    // launch failure instead of a hung GPU
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) break;
        if (spins > (1u << 24)) __trap();
    }
}
// one row segment (box = BW x 1 elements) of the 2-D tensor map -> shared memory; completion is
// counted in bytes on `bar`; out-of-image elements arrive as NaN
__device__ __forceinline__ void tma_load_row(float* dst, const CUtensorMap* map, int32_t col, int32_t row, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(col), "r"(row), "r"(smem_u32(bar))
        : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cp_async(float* dst, const float* src) {
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    else if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void role_barrier() {   // the 128 first-pass threads only (barrier 0 is __syncthreads)
    asm volatile("bar.sync 1, 128;" ::: "memory");
}

template <bool IS_MAX>
__device__ __forceinline__ float op2(float a, float b) {
    return IS_MAX ? fmaxf(a, b) : fminf(a, b);
}
template <bool IS_MAX>
__device__ __forceinline__ float op3(float a, float b, float c) {
    float r;
    if (IS_MAX) asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    else asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

template <int C>
__device__ __forceinline__ void load_vec(const float* p, float* dst) {
    if constexpr (C == 4) {
        float4 v = *reinterpret_cast<const float4*>(p);
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    } else if constexpr (C == 2) {
        float2 v = *reinterpret_cast<const float2*>(p);
        dst[0] = v.x; dst[1] = v.y;
    } else {
        dst[0] = *p;
    }
}
template <int C>
__device__ __forceinline__ void store_vec(float* p, const float* src) {
    if constexpr (C == 4) *reinterpret_cast<float4*>(p) = make_float4(src[0], src[1], src[2], src[3]);
    else if constexpr (C == 2) *reinterpret_cast<float2*>(p) = make_float2(src[0], src[1]);
    else *p = src[0];
}

// global rows: C-wide access at an address that is only 8-byte aligned when HALF (odd radii, see CfgT::DX)
template <int C, bool HALF>
__device__ __forceinline__ void load_gvec(const float* p, float* dst) {
    if constexpr (C == 4 && HALF) {
        load_vec<2>(p, dst);
        load_vec<2>(p + 2, dst + 2);
    } else {
        load_vec<C>(p, dst);
    }
}
template <int C, bool HALF>
__device__ __forceinline__ void store_gvec(float* p, const float* src) {
    if constexpr (C == 4 && HALF) {
        store_vec<2>(p, src);
        store_vec<2>(p + 2, src + 2);
    } else {
        store_vec<C>(p, src);
    }
}

// The thread's window of one ring row: local index W + c is the centre of column c.  With several
// columns per thread the row is read once with vector loads; with one column per thread (large
// radii: the registers belong to the accumulators) every sample is read where it is used.
template <typename K>
struct RowWindow {
    float z[K::C == 1 ? 1 : K::NQ * K::C];
    const float* s;
    __device__ __forceinline__ explicit RowWindow(const float* __restrict__ srow) : s(srow) {
        if constexpr (K::C > 1) {
#pragma unroll
            for (int i = 0; i < K::NQ; ++i) load_vec<K::C>(srow + i * K::C, z + i * K::C);
        }
    }
    template <int I>
    __device__ __forceinline__ float at() const {
        if constexpr (K::C == 1) return s[I];
        else return z[I];
    }
};

// One incoming ring row.  Slot a of `in` is the accumulator of the output row (this row - W + a);
// the row contributes the chord of dy = W - a to it.  Slot 0 is completed (returned in `fin`), every
// other slot moves one down into `out`, the newest output row (dy = -W, half-length 0) enters at the top.
template <typename K, bool IS_MAX>
__device__ __forceinline__ void single_step(const float* __restrict__ srow, const float (&in)[K::S][K::C],
                                            float (&out)[K::S][K::C], float (&fin)[K::C]) {
    constexpr int C = K::C, W = K::W;
    const RowWindow<K> z(srow);
    float R[C];
    static_for<0, C>([&](auto CC) {
        constexpr int c = decltype(CC)::value;
        R[c] = z.template at<W + c>();
        fin[c] = op2<IS_MAX>(in[0][c], R[c]);
        out[2 * W - 1][c] = R[c];
    });
    static_for<1, W + 1>([&](auto H) {
        constexpr int h = decltype(H)::value;
        static_for<0, C>([&](auto CC) {
            constexpr int c = decltype(CC)::value;
            R[c] = op3<IS_MAX>(R[c], z.template at<W + c - h>(), z.template at<W + c + h>());
        });
        static_for<0, W>([&](auto DY) {
            constexpr int dy = decltype(DY)::value;
            if constexpr (K::half(dy) == h) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    out[W - dy - 1][c] = op2<IS_MAX>(in[W - dy][c], R[c]);
                    if constexpr (dy != 0) out[W + dy - 1][c] = op2<IS_MAX>(in[W + dy][c], R[c]);
                }
            }
        });
    });
}

// Two incoming ring rows at once.  Slot a of `in` is the accumulator of output row (row0 - W + a):
// row0 contributes the chord of dy = W - a, row1 the chord of dy + 1, so each slot takes one 3-input
// min/max.  The horizontal windows of both rows grow in lockstep; an update is issued as soon as the
// wider of its two chords is available.  Slots 0 and 1 are completed (fin0, fin1), the others move two
// down into `out`, two new output rows enter at the top.
template <typename K, bool IS_MAX>
__device__ __forceinline__ void pair_step(const float* __restrict__ srow0, const float* __restrict__ srow1,
                                          const float (&in)[K::S][K::C], float (&out)[K::S][K::C], float (&fin0)[K::C],
                                          float (&fin1)[K::C]) {
    constexpr int C = K::C, W = K::W;
    const RowWindow<K> z0(srow0), z1(srow1);
    float R0[W + 1][C], R1[W + 1][C];   // statically indexed: only the live windows occupy registers
    static_for<0, W + 1>([&](auto H) {
        constexpr int h = decltype(H)::value;
        static_for<0, C>([&](auto CC) {
            constexpr int c = decltype(CC)::value;
            if constexpr (h == 0) {
                R0[0][c] = z0.template at<W + c>();
                R1[0][c] = z1.template at<W + c>();
                fin0[c] = op2<IS_MAX>(in[0][c], R0[0][c]);         // row0 is dy = +W of slot 0
                out[2 * W - 1][c] = R1[0][c];                      // row1 is dy = -W of the newest output row
            } else {
                R0[h][c] = op3<IS_MAX>(R0[h - 1][c], z0.template at<W + c - h>(), z0.template at<W + c + h>());
                R1[h][c] = op3<IS_MAX>(R1[h - 1][c], z1.template at<W + c - h>(), z1.template at<W + c + h>());
            }
        });
        // slot a = W - dy  (dy of row0 in [-W, W-1]; row1 sees dy + 1)
        static_for<-W, W>([&](auto DY) {
            constexpr int dy = decltype(DY)::value;
            constexpr int h0 = K::half(dy < 0 ? -dy : dy);
            constexpr int h1 = K::half(dy + 1 < 0 ? -(dy + 1) : dy + 1);
            if constexpr ((h0 > h1 ? h0 : h1) == h) {
                constexpr int a = W - dy;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    if constexpr (dy == -W) out[a - 2][c] = op2<IS_MAX>(R0[h0][c], R1[h1][c]);   // first terms of a new output row
                    else if constexpr (a == 1) fin1[c] = op3<IS_MAX>(in[a][c], R0[h0][c], R1[h1][c]);
                    else out[a - 2][c] = op3<IS_MAX>(in[a][c], R0[h0][c], R1[h1][c]);
                }
            }
        });
    });
}

// cp.async loader (rows not describable to the TMA unit): batch `g` of `last` rows [zr0 + g*U, +U) into
// its Zs stage; warp wi copies rows wi, wi+4, ..; VL floats per copy; anything outside the image is
// written as `ident`.
template <typename K, int VL>
__device__ __forceinline__ void issue_group(const Params& p, float* Zs, int g, int64_t zc0, int64_t zr0, float ident,
                                            int wi, int lane) {
    constexpr int NCH = K::COLS / VL;
    float* slot = Zs + (size_t)(g % K::NZ) * K::U * K::COLS;
#pragma unroll
    for (int rr0 = 0; rr0 < K::U; rr0 += 4) {
        const int rr = rr0 + wi;
        if (rr >= K::U) break;
        const int64_t r = zr0 + (int64_t)g * K::U + rr;
        const bool rowok = r >= 0 && r < p.ny;
        const float* src = p.in + r * p.pitch + zc0;
        float* dst = slot + (size_t)rr * K::COLS;
        for (int cc = lane; cc < NCH; cc += 32) {
            const int64_t g0 = zc0 + (int64_t)cc * VL;
            if (rowok && g0 >= 0 && g0 + VL <= p.nx) {
                cp_async<4 * VL>(dst + cc * VL, src + cc * VL);
            } else {
#pragma unroll
                for (int e = 0; e < VL; ++e)
                    dst[cc * VL + e] = (rowok && g0 + e >= 0 && g0 + e < p.nx) ? __ldg(src + cc * VL + e) : ident;
            }
        }
    }
}

template <typename K>
__device__ __forceinline__ void issue_group_any(const Params& p, float* Zs, int g, int64_t zc0, int64_t zr0,
                                                float ident, int wi, int lane) {
    if (p.vec_ok) issue_group<K, 4>(p, Zs, g, zc0, zr0, ident, wi, lane);      // zc0 = 0 mod 4 (CfgT::DX)
    else issue_group<K, 1>(p, Zs, g, zc0, zr0, ident, wi, lane);
}

// TMA loader: one thread arms the stage's barrier with the batch's byte count and issues its U * NBOX row boxes
template <typename K>
__device__ __forceinline__ void tma_issue_batch(const CUtensorMap* map, float* Zs, uint64_t* zfull, int g, int64_t zc0,
                                                int64_t zr0) {
    const int st = g % K::NZ;
    float* slot = Zs + (size_t)st * K::U * K::COLS;
    mbar_expect_tx(&zfull[st], K::kStageBytes);
#pragma unroll
    for (int rr = 0; rr < K::U; ++rr)
#pragma unroll
        for (int b = 0; b < K::NBOX; ++b)
            tma_load_row(slot + rr * K::COLS + b * K::BW, map, (int32_t)(zc0 + b * K::BW),
                         (int32_t)(zr0 + (int64_t)g * K::U + rr), &zfull[st]);
}

template <typename K, bool NEG>
__global__ void __launch_bounds__(kThreads, K::MINB) open_march_kernel(const Params p, const __grid_constant__ CUtensorMap tmap) {
    constexpr int C = K::C, U = K::U, S = K::S, W = K::W;
    constexpr bool E_MAX = NEG;        // erosion of -Z is -(dilation of Z)
    constexpr bool D_MAX = !NEG;
    extern __shared__ unsigned char smem_base[];
    // TMA destinations want 128-byte alignment: align the dynamic segment by hand
    unsigned char* smem_raw = smem_base + ((128u - (smem_u32(smem_base) & 127u)) & 127u);
    float* Zs = reinterpret_cast<float*>(smem_raw);
    float* Es = Zs + (size_t)K::NZ * U * K::COLS;
    uint64_t* efull = reinterpret_cast<uint64_t*>(Es + (size_t)K::NB * U * K::COLS);
    uint64_t* eempty = efull + K::NB;
    uint64_t* zfull = eempty + K::NB;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < K::NB; ++i) {
            mbar_init(&efull[i], kRoleThreads);
            mbar_init(&eempty[i], kRoleThreads);
        }
        for (int i = 0; i < K::NZ; ++i) mbar_init(&zfull[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t x0 = (int64_t)blockIdx.x * K::XO - K::DX;
    const int64_t y0 = p.row_lo + (int64_t)blockIdx.y * p.seg;
    const int64_t y1 = (y0 + p.seg < p.row_hi) ? y0 + p.seg : p.row_hi;
    const int nOut = (int)(y1 - y0);
    // Both passes emit from their batch I0/U on, so that the first second-pass output is row
    // y0: the second pass consumes first-pass rows from e0 = y0 + W - I0, and the first pass
    // consumes rows of `last` from zr0 = e0 + W - I0.
    const int nDg = (nOut + U - 1) / U;          // emitting batches of the second pass
    const int nEg = nDg + K::I0 / U;             // emitting batches of the first pass = batches the second consumes
    const int nZg = nEg + K::I0 / U;             // batches the first pass consumes
    const int64_t e0 = y0 + W - K::I0;
    const int64_t zr0 = e0 + W - K::I0;
    const float e_ident = E_MAX ? -INFINITY : INFINITY;   // "no sample" for the first pass
    const float d_ident = D_MAX ? -INFINITY : INFINITY;   // ... and for the second

    if (tid < kRoleThreads) {
        // ------------------------------------------------------------- first pass (erosion)
        const int te = tid, wi = tid >> 5, lane = tid & 31;
        const int64_t zc0 = x0 - 2 * W;
        const bool tma = p.use_tma != 0;
        if (tma) {
            if (tid == 0)
                for (int g = 0; g < K::NZ - 1 && g < nZg; ++g) tma_issue_batch<K>(&tmap, Zs, zfull, g, zc0, zr0);
        } else {
            issue_group_any<K>(p, Zs, 0, zc0, zr0, e_ident, wi, lane);
        }
        float accA[S][C], accB[S][C];
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int c = 0; c < C; ++c) accA[s][c] = e_ident;
        const int64_t ecol0 = x0 - W + C * te;
        bool colok[C];
        bool cols_in = true;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            colok[c] = (ecol0 + c >= 0) && (ecol0 + c < p.nx);
            cols_in = cols_in && colok[c];
        }
        if (!tma) {
            cp_async_wait_all();
            role_barrier();
        }
        int slot = 0;
        uint32_t phase = 0;
#pragma unroll 1
        for (int g = 0; g < nZg; ++g) {
            const int zst = g % K::NZ;
            if (tma) {
                // every first-pass thread left batch g-1 through the barrier below: its stage takes batch g-1+NZ
                if (tid == 0 && g + K::NZ - 1 < nZg) tma_issue_batch<K>(&tmap, Zs, zfull, g + K::NZ - 1, zc0, zr0);
                mbar_wait(&zfull[zst], (uint32_t)((g / K::NZ) & 1));
            } else if (g + 1 < nZg) {
                issue_group_any<K>(p, Zs, g + 1, zc0, zr0, e_ident, wi, lane);
            }
            const float* zb = Zs + (size_t)zst * U * K::COLS + C * te;
            const int kg = g - K::I0 / U;
            const bool emit = kg >= 0;
            float* eb = Es + (size_t)slot * U * K::COLS + C * te;
            auto put = [&](int u, float (&fin)[C]) {
                const int64_t e = e0 + (int64_t)kg * U + u;
                const bool rowok = e >= 0 && e < p.ny;
                if (!(rowok && cols_in)) {
#pragma unroll
                    for (int c = 0; c < C; ++c) fin[c] = (rowok && colok[c]) ? fin[c] : d_ident;
                }
                store_vec<C>(eb + u * K::COLS, fin);
            };
            if constexpr (K::PAIR) {
                float f0[C], f1[C];
                pair_step<K, E_MAX>(zb, zb + K::COLS, accA, accB, f0, f1);
                if (emit) {
                    mbar_wait(&eempty[slot], phase ^ 1);     // the second pass has left this slot (waited for as late as possible)
                    put(0, f0); put(1, f1);
                }
                pair_step<K, E_MAX>(zb + 2 * K::COLS, zb + 3 * K::COLS, accB, accA, f0, f1);
                if (emit) { put(2, f0); put(3, f1); }
            } else {
                float f0[C];
                single_step<K, E_MAX>(zb, accA, accB, f0);
                if (emit) {
                    mbar_wait(&eempty[slot], phase ^ 1);
                    put(0, f0);
                }
                single_step<K, E_MAX>(zb + K::COLS, accB, accA, f0);
                if (emit) put(1, f0);
                single_step<K, E_MAX>(zb + 2 * K::COLS, accA, accB, f0);
                if (emit) put(2, f0);
                single_step<K, E_MAX>(zb + 3 * K::COLS, accB, accA, f0);
                if (emit) put(3, f0);
            }
            if (emit) {
                mbar_arrive(&efull[slot]);
                if (++slot == K::NB) { slot = 0; phase ^= 1; }
            }
            if (!tma) cp_async_wait_all();
            role_barrier();
        }
    } else {
        // ------------------------------------------------------------- second pass + threshold
        const int td = tid - kRoleThreads;
        float accA[S][C], accB[S][C];
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int c = 0; c < C; ++c) accA[s][c] = d_ident;
        const int64_t gx = x0 + C * td;
        const bool dvalid = (td < K::XO / C) && (gx < p.nx) && (gx + C > 0);
        // fast threads: all C columns inside the image and reachable with vector accesses (everything but the strips at
        // the image's left / right border and unaligned surfaces)
        const bool fast = dvalid && p.vec_ok && (gx >= 0) && (gx + C - 1 < p.nx);
        constexpr bool HALF = K::DX != 0;
        const bool want_mask = p.mask != nullptr;
        int slot = 0;
        uint32_t phase = 0;
#pragma unroll 1
        for (int g = 0; g < nEg; ++g) {
            mbar_wait(&efull[slot], phase);
            const float* eb = Es + (size_t)slot * U * K::COLS + C * td;
            const int dg = g - K::I0 / U;
            // rows [r0, r0 + nv) of this batch are produced; everything per row below is pointer + u * pitch
            const int64_t r0 = y0 + (int64_t)(dg < 0 ? 0 : dg) * U;
            const int nv = dg < 0 ? 0 : (int)((y1 - r0) < U ? (y1 - r0) : U);
            const float* lp = p.in + r0 * p.pitch + gx;
            float* op = p.out + r0 * p.pitch + gx;             // only dereferenced if p.out
            const int64_t m0off = r0 * p.nx + gx;
            // the re-read of `last` is issued before the compute so that its L2 latency hides under it
            auto fetch = [&](int u, float (&l)[C]) {
                if (u >= nv) return;
                if (fast) load_gvec<C, HALF>(lp + u * p.pitch, l);
                else if (dvalid) {
#pragma unroll
                    for (int c = 0; c < C; ++c) l[c] = (gx + c >= 0 && gx + c < p.nx) ? __ldg(lp + u * p.pitch + c) : 0.f;
                }
            };
            auto put = [&](int u, const float (&fin)[C], const float (&l)[C]) {
                if (u >= nv || !dvalid) return;
                float o[C];
#pragma unroll
                for (int c = 0; c < C; ++c) o[c] = NEG ? -fin[c] : fin[c];
                if (p.out) {
                    if (fast) store_gvec<C, HALF>(op + u * p.pitch, o);
                    else {
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            if (gx + c >= 0 && gx + c < p.nx) op[u * p.pitch + c] = o[c];
                    }
                }
                if (want_mask) {
                    // float32 screen: the rounded difference d32 is within 2^-24 relative of the exact one, so
                    // d32 <= thr_screen = thr (1 - 2^-20) rounded down means certainly not (exact difference > thr);
                    // a NaN difference never exceeds a threshold, so dropping it from the max is harmless.
                    // (-Z) - open(-Z) == close(Z) - Z exactly.
                    float dmax = NEG ? fin[0] - l[0] : l[0] - fin[0];
#pragma unroll
                    for (int c = 1; c < C; ++c) dmax = fmaxf(dmax, NEG ? fin[c] - l[c] : l[c] - fin[c]);
                    if (!(dmax <= p.thr_screen)) {
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const double df = NEG ? __dsub_rn((double)fin[c], (double)l[c])
                                                  : __dsub_rn((double)l[c], (double)fin[c]);
                            if ((gx + c >= 0) && (gx + c < p.nx) && (df > p.thr)) {
                                p.mask[m0off + u * p.nx + c] = 1;
                                if (p.when) p.when[m0off + u * p.nx + c] = (uint8_t)p.widx;
                            }
                        }
                    }
                }
            };
            if constexpr (K::PAIR) {
                float l0[C], l1[C], f0[C], f1[C];
                fetch(0, l0); fetch(1, l1);
                pair_step<K, D_MAX>(eb, eb + K::COLS, accA, accB, f0, f1);
                put(0, f0, l0); put(1, f1, l1);
                fetch(2, l0); fetch(3, l1);
                pair_step<K, D_MAX>(eb + 2 * K::COLS, eb + 3 * K::COLS, accB, accA, f0, f1);
                mbar_arrive(&eempty[slot]);      // the batch has been read: hand the slot back before the last stores
                put(2, f0, l0); put(3, f1, l1);
            } else {
                float l0[C], f0[C];
                fetch(0, l0);
                single_step<K, D_MAX>(eb, accA, accB, f0);
                put(0, f0, l0);
                fetch(1, l0);
                single_step<K, D_MAX>(eb + K::COLS, accB, accA, f0);
                put(1, f0, l0);
                fetch(2, l0);
                single_step<K, D_MAX>(eb + 2 * K::COLS, accA, accB, f0);
                put(2, f0, l0);
                fetch(3, l0);
                single_step<K, D_MAX>(eb + 3 * K::COLS, accB, accA, f0);
                mbar_arrive(&eempty[slot]);
                put(3, f0, l0);
            }
            if (++slot == K::NB) { slot = 0; phase ^= 1; }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Radii 41..72 (0.25 m cells, BASELINE.json configs[4]): two single-role passes.
//
// The fused kernel keeps 2W accumulators per column in registers for BOTH passes and each of its
// roles covers only 128 C columns of which 2W are halo; past W ~ 40 neither fits.  Here the erosion
// and the dilation are separate launches of one kernel: all T threads run the same pass, one column
// per thread (2W registers of accumulators), every column of the CTA's strip is an output column
// (only the ring rows carry the 2W-column halo, no thread computes a halo column), and the
// intermediate goes through the `tmp` plane in HBM / L2 -- 8 bytes per cell-window more traffic on a
// kernel that is FMNMX-bound by a factor > 5 at these radii.  Rows arrive by TMA exactly as above.
template <int W_, int T_, int C_, bool PAIR_, int MINB_ = 1, int NZ_ = 4>
struct PassCfg {
    static constexpr int W = W_;
    static constexpr bool PAIR = PAIR_;
    static constexpr int C = C_;                                // adjacent columns per thread
    static constexpr int T = T_;                                // threads per CTA
    static constexpr int XO = T * C;                            // output columns per CTA: every column of the strip
    static constexpr int U = 4;
    static constexpr int S = 2 * W;
    static constexpr int NL = 2 * W + C;
    static constexpr int NQ = (NL + C - 1) / C;
    // Strips start DX columns left of a multiple of XO so that the ring's first column x0 - W is 0 mod 4 (the TMA
    // unit wants 16-byte aligned boxes, and the threads' C-wide shared-memory windows stay aligned); the global
    // accesses of the outputs are then only aligned to GA = gcd(4, W mod 4 ...) elements.
    static constexpr int DX = (4 - W % 4) % 4;
    static constexpr int GA = (W % 4 == 0) ? 4 : ((W % 2 == 0) ? 2 : 1);      // alignment (elements) of x0 + C * t
    static constexpr int NEED = (T - 1) * C + NQ * C;
    static constexpr int NBOX = (NEED + 255) / 256;
    static constexpr int BW = ((NEED + NBOX - 1) / NBOX + 31) / 32 * 32;
    static constexpr int COLS = NBOX * BW;
    static constexpr int I0 = (2 * W + U - 1) / U * U;
    static constexpr int NZ = NZ_;
    static constexpr int MINB = MINB_;
    static constexpr uint32_t kStageBytes = (uint32_t)(U * COLS * sizeof(float));
    static constexpr size_t kSmemBytes = 128 + (size_t)NZ * kStageBytes + 2 * NZ * sizeof(uint64_t);
    __host__ __device__ static constexpr int isqrt(int v) {
        int h = 0;
        while ((h + 1) * (h + 1) <= v) ++h;
        return h;
    }
    __host__ __device__ static constexpr int half(int dy) { return isqrt(W * W - dy * dy); }
};

struct PassParams {
    const float* in;      // the plane this pass streams (last for the erosion, tmp for the dilation)
    float* out;           // tmp for the erosion, `this` for the dilation (may be null there)
    const float* last;    // dilation only: the window's input surface, for the threshold
    uint8_t* mask;
    uint8_t* when;
    int64_t ny, nx, pitch, row_lo, row_hi;   // rows [row_lo, row_hi) are produced
    int seg;
    double thr;
    float thr_screen;
    int widx, use_tma, vec_ok;
};

// C-wide global access at an address aligned to GA elements
template <int C, int GA>
__device__ __forceinline__ void load_galign(const float* p, float* dst) {
    constexpr int V = (GA < C) ? GA : C;
#pragma unroll
    for (int i = 0; i < C; i += V) load_vec<V>(p + i, dst + i);
}
template <int C, int GA>
__device__ __forceinline__ void store_galign(float* p, const float* src) {
    constexpr int V = (GA < C) ? GA : C;
#pragma unroll
    for (int i = 0; i < C; i += V) store_vec<V>(p + i, src + i);
}

template <typename K, bool IS_MAX, bool FINAL>
__global__ void __launch_bounds__(K::T, K::MINB) open_pass_kernel(const PassParams p, const __grid_constant__ CUtensorMap tmap) {
    constexpr int U = K::U, S = K::S, W = K::W, C = K::C;
    extern __shared__ unsigned char smem_base[];
    unsigned char* smem_raw = smem_base + ((128u - (smem_u32(smem_base) & 127u)) & 127u);
    float* Zs = reinterpret_cast<float*>(smem_raw);
    uint64_t* zfull = reinterpret_cast<uint64_t*>(Zs + (size_t)K::NZ * U * K::COLS);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < K::NZ; ++i) mbar_init(&zfull[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t x0 = (int64_t)blockIdx.x * K::XO - K::DX;
    const int64_t y0 = p.row_lo + (int64_t)blockIdx.y * p.seg;
    const int64_t y1 = (y0 + p.seg < p.row_hi) ? y0 + p.seg : p.row_hi;
    const int nOut = (int)(y1 - y0);
    const int nG = (nOut + U - 1) / U + K::I0 / U;       // batches streamed; emission starts at batch I0 / U
    const int64_t zr0 = y0 + W - K::I0;                  // first streamed row
    const int64_t zc0 = x0 - W;                          // first ring column (0 mod 4)
    const float ident = IS_MAX ? -INFINITY : INFINITY;
    const bool tma = p.use_tma != 0;
    auto load_batch = [&](int g) {                       // the generic loader: every thread copies, 4 bytes at a time
        float* slot = Zs + (size_t)(g % K::NZ) * U * K::COLS;
        for (int i = tid; i < U * K::COLS; i += K::T) {
            const int rr = i / K::COLS, cc = i - rr * K::COLS;
            const int64_t r = zr0 + (int64_t)g * U + rr, c = zc0 + cc;
            slot[i] = (r >= 0 && r < p.ny && c >= 0 && c < p.nx) ? __ldg(p.in + r * p.pitch + c) : ident;
        }
    };
    if (tma) {
        if (tid == 0)
            for (int g = 0; g < K::NZ - 1 && g < nG; ++g) tma_issue_batch<K>(&tmap, Zs, zfull, g, zc0, zr0);
    } else {
        load_batch(0);
        __syncthreads();
    }
    float accA[S][C], accB[S][C];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int c = 0; c < C; ++c) accA[s][c] = ident;
    const int64_t gx = x0 + (int64_t)C * tid;
    const bool anycol = (gx < p.nx) && (gx + C > 0);
    const bool fast = anycol && p.vec_ok && (gx >= 0) && (gx + C - 1 < p.nx);    // see the fused kernel
    const bool want_mask = FINAL && p.mask != nullptr;
    const int64_t y_end = y1 < p.ny ? y1 : p.ny;
#pragma unroll 1
    for (int g = 0; g < nG; ++g) {
        const int zst = g % K::NZ;
        if (tma) {
            // every thread left batch g-1 through the barrier at the end of the loop: its stage takes batch g-1+NZ
            if (tid == 0 && g + K::NZ - 1 < nG) tma_issue_batch<K>(&tmap, Zs, zfull, g + K::NZ - 1, zc0, zr0);
            mbar_wait(&zfull[zst], (uint32_t)((g / K::NZ) & 1));
        } else if (g + 1 < nG) {
            load_batch(g + 1);
        }
        const float* zb = Zs + (size_t)zst * U * K::COLS + C * tid;
        const int kg = g - K::I0 / U;
        // rows [r0, r0 + nv) of this batch are produced (r0 >= 0: row_lo >= 0)
        const int64_t r0 = y0 + (int64_t)(kg < 0 ? 0 : kg) * U;
        const int nv = kg < 0 ? 0 : (int)((y_end - r0) < U ? (y_end - r0 < 0 ? 0 : y_end - r0) : U);
        const float* lp = p.last + r0 * p.pitch + gx;          // only dereferenced if FINAL
        float* op = p.out + r0 * p.pitch + gx;                 // only dereferenced if p.out
        const int64_t m0off = r0 * p.nx + gx;
        auto fetch = [&](int u, float (&l)[C]) {
            if (!FINAL || u >= nv) return;
            if (fast) load_galign<C, K::GA>(lp + u * p.pitch, l);
            else if (anycol) {
#pragma unroll
                for (int c = 0; c < C; ++c) l[c] = (gx + c >= 0 && gx + c < p.nx) ? __ldg(lp + u * p.pitch + c) : 0.f;
            }
        };
        auto put = [&](int u, const float (&f)[C], const float (&l)[C]) {
            if (u >= nv || !anycol) return;
            if (p.out) {
                if (fast) store_galign<C, K::GA>(op + u * p.pitch, f);
                else {
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        if (gx + c >= 0 && gx + c < p.nx) op[u * p.pitch + c] = f[c];
                }
            }
            if (want_mask) {
                float dmax = l[0] - f[0];
#pragma unroll
                for (int c = 1; c < C; ++c) dmax = fmaxf(dmax, l[c] - f[c]);
                if (!(dmax <= p.thr_screen)) {           // float32 screen, float64 decision (see the fused kernel)
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        if (gx + c >= 0 && gx + c < p.nx && __dsub_rn((double)l[c], (double)f[c]) > p.thr) {
                            p.mask[m0off + u * p.nx + c] = 1;
                            if (p.when) p.when[m0off + u * p.nx + c] = (uint8_t)p.widx;
                        }
                    }
                }
            }
        };
        if constexpr (K::PAIR) {
            float l0[C], l1[C], f0[C], f1[C];
            fetch(0, l0); fetch(1, l1);
            pair_step<K, IS_MAX>(zb, zb + K::COLS, accA, accB, f0, f1);
            put(0, f0, l0); put(1, f1, l1);
            fetch(2, l0); fetch(3, l1);
            pair_step<K, IS_MAX>(zb + 2 * K::COLS, zb + 3 * K::COLS, accB, accA, f0, f1);
            put(2, f0, l0); put(3, f1, l1);
        } else {
            float l0[C], f0[C];
            fetch(0, l0);
            single_step<K, IS_MAX>(zb, accA, accB, f0);
            put(0, f0, l0);
            fetch(1, l0);
            single_step<K, IS_MAX>(zb + K::COLS, accB, accA, f0);
            put(1, f0, l0);
            fetch(2, l0);
            single_step<K, IS_MAX>(zb + 2 * K::COLS, accA, accB, f0);
            put(2, f0, l0);
            fetch(3, l0);
            single_step<K, IS_MAX>(zb + 3 * K::COLS, accB, accA, f0);
            put(3, f0, l0);
        }
        __syncthreads();      // everyone has left batch g: its stage may be refilled
    }
}

}  // namespace march

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (the library does not link libcuda)
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TensorMapEncodeTiledFn tensor_map_encoder() {
    static TensorMapEncodeTiledFn fn = []() -> TensorMapEncodeTiledFn {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &f, 12000, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (TensorMapEncodeTiledFn)f;
    }();
    return fn;
}

// 2-D map of a float32 surface [ny][pitch] (nx valid columns) with boxes of `box_w` x 1 elements and NaN for
// everything outside [0, ny) x [0, nx).  Returns false when the surface cannot be described (alignment).
inline bool make_row_tensor_map(CUtensorMap* map, const float* base, int64_t ny, int64_t nx, int64_t pitch, int box_w) {
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return false;
    if ((((uintptr_t)base) & 15) || (pitch % 4) || nx >= (int64_t)1 << 31 || ny >= (int64_t)1 << 31) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)nx, (cuuint64_t)ny};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, 1u};
    const cuuint32_t estr[2] = {1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA) == CUDA_SUCCESS;
}

inline float screen_threshold(double thr) {
    // largest float that is certainly below thr * (1 - 2^-20); anything not finite or not positive disables the screen
    if (!(thr > 0.0) || !(thr < 1e30)) return -INFINITY;
    float t = (float)(thr * (1.0 - 9.5367431640625e-7));
    return nextafterf(t, -INFINITY);
}

template <typename K, bool NEG>
int launch_open_march_cfg(const float* in, float* out, uint8_t* mask, uint8_t* when, int64_t ny, int64_t nx,
                          int64_t pitch, double thr, int widx, int64_t row_lo, int64_t row_hi, cudaStream_t st) {
    // per device: the attribute belongs to the function on the current device
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        SMRF_CUDA(cudaFuncSetAttribute(march::open_march_kernel<K, NEG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)K::kSmemBytes));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const int64_t rows = row_hi - row_lo;
    const int nstrips = (int)((nx + K::DX + K::XO - 1) / K::XO);
    // Segment the rows so that (waves of CTAs) x (rows marched per CTA, incl. the 2*I0 warm-up rows)
    // is smallest: few long segments waste SMs, many short ones waste warm-up.
    const int64_t slots = (int64_t)num_sms() * K::MINB;
    const int64_t min_seg = 2 * K::I0 < 32 ? 32 : 2 * K::I0;
    int64_t max_segs = rows / min_seg;
    if (max_segs < 1) max_segs = 1;
    if (max_segs > 4096) max_segs = 4096;
    int64_t best_cost = -1, best_n = 1;
    for (int64_t n = 1; n <= max_segs; ++n) {
        const int64_t ctas = n * nstrips;
        const int64_t waves = (ctas + slots - 1) / slots;
        const int64_t cost = waves * ((rows + n - 1) / n + 2 * K::I0);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_n = n; }
    }
    int seg = (int)((rows + best_n - 1) / best_n);
    seg = (seg + K::U - 1) / K::U * K::U;
    const int nsegs = (int)((rows + seg - 1) / seg);
    march::Params p;
    p.in = in; p.out = out; p.mask = mask; p.when = when;
    p.ny = ny; p.nx = nx; p.pitch = pitch; p.row_lo = row_lo; p.row_hi = row_hi;
    p.seg = seg; p.thr = thr; p.thr_screen = screen_threshold(thr); p.widx = widx;
    p.vec_ok = (pitch % 4 == 0) && (((uintptr_t)in & 15) == 0) && (out == nullptr || ((uintptr_t)out & 15) == 0);
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    p.use_tma = (!open_no_tma() && make_row_tensor_map(&map, in, ny, nx, pitch, K::BW)) ? 1 : 0;
    dim3 grid((unsigned)nstrips, (unsigned)nsegs);
    march::open_march_kernel<K, NEG><<<grid, march::kThreads, K::kSmemBytes, st>>>(p, map);
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

// erosion into `tmp`, then dilation + threshold from `tmp` (radii the fused kernel does not cover)
template <typename K>
int launch_open_passes_cfg(const float* in, float* out, float* tmp, uint8_t* mask, uint8_t* when, int64_t ny, int64_t nx,
                           int64_t pitch, double thr, int widx, int64_t row_lo, int64_t row_hi, cudaStream_t st) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        SMRF_CUDA(cudaFuncSetAttribute(march::open_pass_kernel<K, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)K::kSmemBytes));
        SMRF_CUDA(cudaFuncSetAttribute(march::open_pass_kernel<K, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)K::kSmemBytes));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const int nstrips = (int)((nx + K::DX + K::XO - 1) / K::XO);
    auto plan = [&](int64_t rows, int* seg_out, int* nsegs_out) {
        const int64_t slots = (int64_t)num_sms() * K::MINB;
        int64_t max_segs = rows / (2 * K::I0);
        if (max_segs < 1) max_segs = 1;
        if (max_segs > 4096) max_segs = 4096;
        int64_t best_cost = -1, best_n = 1;
        for (int64_t n = 1; n <= max_segs; ++n) {
            const int64_t waves = (n * nstrips + slots - 1) / slots;
            const int64_t cost = waves * ((rows + n - 1) / n + K::I0);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_n = n; }
        }
        int seg = (int)((rows + best_n - 1) / best_n);
        seg = (seg + K::U - 1) / K::U * K::U;
        *seg_out = seg;
        *nsegs_out = (int)((rows + seg - 1) / seg);
    };
    const int64_t e_lo = row_lo - K::W < 0 ? 0 : row_lo - K::W;
    const int64_t e_hi = row_hi + K::W > ny ? ny : row_hi + K::W;
    march::PassParams p;
    memset(&p, 0, sizeof(p));
    p.ny = ny; p.nx = nx; p.pitch = pitch; p.thr = thr; p.thr_screen = screen_threshold(thr); p.widx = widx;
    p.vec_ok = (pitch % 4 == 0) && (((uintptr_t)in & 15) == 0) && (((uintptr_t)tmp & 15) == 0) &&
               (out == nullptr || ((uintptr_t)out & 15) == 0);
    CUtensorMap map;
    int seg, nsegs;
    // pass 1: tmp = erode(in) on rows [e_lo, e_hi)
    memset(&map, 0, sizeof(map));
    p.in = in; p.out = tmp; p.last = nullptr; p.mask = nullptr; p.when = nullptr;
    p.row_lo = e_lo; p.row_hi = e_hi;
    plan(e_hi - e_lo, &seg, &nsegs);
    p.seg = seg;
    p.use_tma = (!open_no_tma() && make_row_tensor_map(&map, in, ny, nx, pitch, K::BW)) ? 1 : 0;
    march::open_pass_kernel<K, false, false><<<dim3((unsigned)nstrips, (unsigned)nsegs), K::T, K::kSmemBytes, st>>>(p, map);
    SMRF_LAUNCH_CHECK();
    // pass 2: out = dilate(tmp) on rows [row_lo, row_hi), threshold against `in`.  Rows of tmp outside
    // [e_lo, e_hi) are only ever streamed as warm-up rows of output rows that are not produced.
    memset(&map, 0, sizeof(map));
    p.in = tmp; p.out = out; p.last = in; p.mask = mask; p.when = when;
    p.row_lo = row_lo; p.row_hi = row_hi;
    plan(row_hi - row_lo, &seg, &nsegs);
    p.seg = seg;
    p.use_tma = (!open_no_tma() && make_row_tensor_map(&map, tmp, ny, nx, pitch, K::BW)) ? 1 : 0;
    march::open_pass_kernel<K, true, true><<<dim3((unsigned)nstrips, (unsigned)nsegs), K::T, K::kSmemBytes, st>>>(p, map);
    SMRF_LAUNCH_CHECK();
    count_launches(2);
    return 0;
}

template <int W>
int launch_open_passes_f32(const float* in, float* out, float* tmp, uint8_t* mask, uint8_t* when, int64_t ny, int64_t nx,
                           int64_t pitch, double thr, int widx, int64_t row_lo, int64_t row_hi, cudaStream_t st) {
    // per-radius shape from the B200 sweeps: columns per thread as wide as the register file allows (4 W C accumulators
    // + windows), then as many resident CTAs as fit
    using K = std::conditional_t<(W <= 11), march::PassCfg<W, 128, 4, true, 3>,
              std::conditional_t<(W <= 20), march::PassCfg<W, 256, 2, true, 2>,
              std::conditional_t<(W <= 40), march::PassCfg<W, 256, 2, true, 1>,
              std::conditional_t<(W <= 48), march::PassCfg<W, 256, 2, false, 1>,
                                            march::PassCfg<W, 256, 1, true, 1>>>>>;
    return launch_open_passes_cfg<K>(in, out, tmp, mask, when, ny, nx, pitch, thr, widx, row_lo, row_hi, st);
}

template <int W, bool NEG>
int launch_open_march_f32(const float* in, float* out, uint8_t* mask, uint8_t* when, int64_t ny, int64_t nx,
                          int64_t pitch, double thr, int widx, int64_t row_lo, int64_t row_hi, cudaStream_t st) {
    return launch_open_march_cfg<march::Cfg<W>, NEG>(in, out, mask, when, ny, nx, pitch, thr, widx, row_lo, row_hi, st);
}

}  // namespace smrf
