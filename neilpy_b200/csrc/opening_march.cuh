// Register-marching fused disk opening for float32 grids (sm_100a).
//
//   this = dilate(erode(last, disk(W)), disk(W));  new = (last - this) > thr;  mask |= new
//
// One CTA owns a strip of XO output columns and a segment of rows and marches down
// it once.  disk(W) is used in chord form: row dy of the disk spans |dx| <= h(dy),
// h(dy) = floor(sqrt(W^2 - dy^2)), so
//     erode(Z)[y][x] = min_dy  rowmin_{h(dy)}(Z[y+dy])[x].
// Warp roles (8 warps, 256 threads so that each thread may hold up to 255 registers):
//   E      (4 warps) stream rows of `last` into the double-buffered shared ring Zs with
//                    cp.async one batch ahead (identity +inf outside the image), and
//                    each thread owns C=4 adjacent columns.  For every incoming row it
//                    grows the horizontal window one cell per side per step
//                    (R_h = min3(R_{h-1}, Z[x-h], Z[x+h]), one FMNMX3 per column) and,
//                    whenever h equals the chord of some dy, folds R_h into the
//                    accumulator of output row (y - dy).  The 2W+U live output rows
//                    per column sit in registers; rows are processed in groups of U
//                    (unrolled, so accumulator indices are static) and the accumulator
//                    file is shifted by U registers after each group, which keeps the
//                    loop body small enough to live in the instruction cache.
//                    Finished erosion rows go to the shared ring Es (identity -inf
//                    outside the image: the reference dilates an eroded image that does
//                    not exist there).
//   D      (4 warps) the same marching with max over Es; a finished row is compared with
//                    `last` (re-read through L2) in float64 and stored with 16-byte
//                    vector stores; mask / when_dropped bytes are written only where set.
// Es is handed over in batches of RB rows through mbarriers (full/empty pairs); Zs is
// private to the E warps (one named barrier per batch).
// NEG = true opens -Z instead (the low-outlier pass): open(-Z) = -close(Z), so the roles
// swap min and max and the sign is applied in the epilogue -- no negated copy is made.
// The erosion->dilation intermediate never leaves the SM: HBM traffic per cell-window
// is one read of `last` (+ halo re-reads, L2 hits) and one write of `this`.
#pragma once
#include <type_traits>

#include "opening.cuh"

namespace smrf {
namespace march {

constexpr int kRoleThreads = 128;
constexpr int kThreads = 2 * kRoleThreads;

template <int W_, int C_, bool PAIR_, int MINB_, int U_>
struct CfgT {
    static constexpr int W = W_;
    // PAIR: two incoming rows are folded per step so that every accumulator update is one
    // 3-input min/max (acc, row u's chord, row u+1's chord) instead of two 2-input ones.  It
    // needs both rows' neighbourhoods in registers.
    static constexpr bool PAIR = PAIR_;
    static constexpr int C = C_;                                // adjacent columns per thread
    static constexpr int U = U_;                                // rows per group (= hand-over batch)
    static constexpr int A = 2 * W + U;                         // live accumulators per column
    static constexpr int EW = kRoleThreads * C;                 // first-pass columns per CTA
    static constexpr int NL = 2 * W + C;                        // elements a thread reads per row
    static constexpr int NQ = (NL + C - 1) / C;                 // ... as C-wide vectors
    static constexpr int COLS = ((kRoleThreads - 1) * C + NQ * C + 3) / 4 * 4;   // ring row length (floats)
    static constexpr int XO = ((EW - 2 * W) / 4) * 4;           // output columns per CTA (multiple of 4 and of C)
    static constexpr int I0 = (2 * W + U - 1) / U * U;          // warm-up rows per pass, whole groups
    static constexpr int NB = 3;                                // batches in the Es ring
    static constexpr int MINB = MINB_;                          // CTAs per SM the register budget is held to
    static constexpr size_t kSmemBytes = (size_t)(2 + NB) * U * COLS * sizeof(float) + 2 * NB * sizeof(uint64_t);

    __host__ __device__ static constexpr int isqrt(int v) {
        int h = 0;
        while ((h + 1) * (h + 1) <= v) ++h;
        return h;
    }
    __host__ __device__ static constexpr int half(int dy) { return isqrt(W * W - dy * dy); }
};

// The shipped choice per radius, taken from the variant sweep tools/march_sweep.cu on B200
// (profiles/r1_march_sweep.log).  What matters most is how many warps an SM can hold (the
// kernel is bound by FMNMX issue and its latency: 2 CTAs/SM whenever ~128 registers suffice),
// then the number of min/max instructions per cell (PAIR), then shared-memory traffic (C).
template <int W>
struct Cfg : CfgT<W,
                  /*C=*/((W <= 10 || W == 17 || W == 18) ? 4 : 2),
                  /*PAIR=*/(W != 2 && W <= 24),
                  /*MINB=*/(W <= 5 ? 1 : (W <= 16 || W == 19 || W == 20 ? 2 : 1)),
                  /*U=*/4> {};

struct Params {
    const float* in;
    float* out;
    uint8_t* mask;
    uint8_t* when;
    int64_t ny, nx, pitch, row_lo, row_hi;   // pitch: row stride of in/out in elements (>= nx); mask/when are nx wide
    int seg;
    double thr;
    int widx, vec_ok;
};

template <int B, int E, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}
static_assert(true, "static_for bounds may be negative");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t a = smem_u32(b);
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!ok);
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

// One incoming ring row (row u of the current group) for one thread.  acc[a] is the
// accumulator of the output row (group base - W + a); this row contributes the chord of
// dy to acc[u - dy + W]: its first term to acc[u + 2W] (dy = -W, assigned) and the last
// term of acc[u] (dy = +W), which is returned in `fin`.  srow points at the thread's first
// element: local index W + c is the centre of column c.
template <typename K, bool IS_MAX, int u>
__device__ __forceinline__ void chord_step(const float* __restrict__ srow, float (&acc)[K::A][K::C],
                                           float (&fin)[K::C]) {
    constexpr int C = K::C, W = K::W;
    float z[K::NQ * C];
#pragma unroll
    for (int i = 0; i < K::NQ; ++i) load_vec<C>(srow + i * C, z + i * C);
    float R[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        R[c] = z[W + c];
        fin[c] = op2<IS_MAX>(acc[u][c], R[c]);     // chord of dy = +W has half-length 0
        acc[u + 2 * W][c] = R[c];                  // ... and so has dy = -W: the newest output row
    }
    static_for<1, W + 1>([&](auto H) {
        constexpr int h = decltype(H)::value;
#pragma unroll
        for (int c = 0; c < C; ++c) R[c] = op3<IS_MAX>(R[c], z[W + c - h], z[W + c + h]);
        static_for<0, W>([&](auto DY) {
            constexpr int dy = decltype(DY)::value;
            if constexpr (K::half(dy) == h) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    acc[u + W - dy][c] = op2<IS_MAX>(acc[u + W - dy][c], R[c]);
                    if constexpr (dy != 0) acc[u + W + dy][c] = op2<IS_MAX>(acc[u + W + dy][c], R[c]);
                }
            }
        });
    });
}

// Two incoming ring rows (u, u+1 of the current group) at once.  Row u contributes the chord
// of dy to acc[u + W - dy], row u+1 the chord of dy + 1 to the same accumulator, so each
// accumulator takes one 3-input min/max.  The horizontal windows R_h of both rows grow in
// lockstep; an update is issued as soon as the wider of its two chords is available.
// fin0 / fin1 are the output rows completed by row u and row u+1.
template <typename K, bool IS_MAX, int u>
__device__ __forceinline__ void chord_pair(const float* __restrict__ srow0, const float* __restrict__ srow1,
                                           float (&acc)[K::A][K::C], float (&fin0)[K::C], float (&fin1)[K::C]) {
    constexpr int C = K::C, W = K::W;
    float z0[K::NQ * C], z1[K::NQ * C];
#pragma unroll
    for (int i = 0; i < K::NQ; ++i) {
        load_vec<C>(srow0 + i * C, z0 + i * C);
        load_vec<C>(srow1 + i * C, z1 + i * C);
    }
    float R0[W + 1][C], R1[W + 1][C];   // statically indexed: only the live windows occupy registers
    static_for<0, W + 1>([&](auto H) {
        constexpr int h = decltype(H)::value;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if constexpr (h == 0) {
                R0[0][c] = z0[W + c];
                R1[0][c] = z1[W + c];
                fin0[c] = op2<IS_MAX>(acc[u][c], R0[0][c]);        // row u is dy = +W of output u
                acc[u + 2 * W + 1][c] = R1[0][c];                  // row u+1 is dy = -W of the newest output
            } else {
                R0[h][c] = op3<IS_MAX>(R0[h - 1][c], z0[W + c - h], z0[W + c + h]);
                R1[h][c] = op3<IS_MAX>(R1[h - 1][c], z1[W + c - h], z1[W + c + h]);
            }
        }
        // accumulator a = u + W - dy  (dy of row u in [-W, W-1]; row u+1 sees dy + 1)
        static_for<-W, W>([&](auto DY) {
            constexpr int dy = decltype(DY)::value;
            constexpr int h0 = K::half(dy < 0 ? -dy : dy);
            constexpr int h1 = K::half(dy + 1 < 0 ? -(dy + 1) : dy + 1);
            if constexpr ((h0 > h1 ? h0 : h1) == h) {
                constexpr int a = u + W - dy;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    if constexpr (dy == -W) acc[a][c] = op2<IS_MAX>(R0[h0][c], R1[h1][c]);   // first terms of output u+2W
                    else acc[a][c] = op3<IS_MAX>(acc[a][c], R0[h0][c], R1[h1][c]);
                }
            }
        });
    });
#pragma unroll
    for (int c = 0; c < C; ++c) fin1[c] = acc[u + 1][c];   // completed by row u+1 (its dy = +W, folded above at dy = W-1)
}

// The first-pass warps stream group `g` of `last` rows [zr0 + g*U, +U) into its Zs slot:
// warp wi copies rows wi, wi+4, ..; VL floats per cp.async (16 / 8 / 4 bytes); anything
// outside the image is written as `ident`.
template <typename K, int VL>
__device__ __forceinline__ void issue_group(const Params& p, float* Zs, int g, int64_t zc0, int64_t zr0, float ident,
                                            int wi, int lane) {
    constexpr int NCH = K::COLS / VL;
    float* slot = Zs + (size_t)(g & 1) * K::U * K::COLS;
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
    if (p.vec_ok && (K::W % 2 == 0)) issue_group<K, 4>(p, Zs, g, zc0, zr0, ident, wi, lane);
    else if (p.vec_ok) issue_group<K, 2>(p, Zs, g, zc0, zr0, ident, wi, lane);
    else issue_group<K, 1>(p, Zs, g, zc0, zr0, ident, wi, lane);
}

template <typename K, bool NEG>
__global__ void __launch_bounds__(kThreads, K::MINB) open_march_kernel(const Params p) {
    constexpr int C = K::C, U = K::U, A = K::A, W = K::W;
    constexpr bool E_MAX = NEG;        // erosion of -Z is -(dilation of Z)
    constexpr bool D_MAX = !NEG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Zs = reinterpret_cast<float*>(smem_raw);
    float* Es = Zs + (size_t)2 * U * K::COLS;
    uint64_t* efull = reinterpret_cast<uint64_t*>(Es + (size_t)K::NB * U * K::COLS);
    uint64_t* eempty = efull + K::NB;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < K::NB; ++i) {
            mbar_init(&efull[i], kRoleThreads);
            mbar_init(&eempty[i], kRoleThreads);
        }
    }
    __syncthreads();

    const int64_t x0 = (int64_t)blockIdx.x * K::XO;
    const int64_t y0 = p.row_lo + (int64_t)blockIdx.y * p.seg;
    const int64_t y1 = (y0 + p.seg < p.row_hi) ? y0 + p.seg : p.row_hi;
    const int nOut = (int)(y1 - y0);
    // Both passes emit from their group I0/U on, so that the first second-pass output is row
    // y0: the second pass consumes first-pass rows from e0 = y0 + W - I0, and the first pass
    // consumes rows of `last` from zr0 = e0 + W - I0.
    const int nDg = (nOut + U - 1) / U;          // emitting groups of the second pass
    const int nEg = nDg + K::I0 / U;             // emitting groups of the first pass = groups the second consumes
    const int nZg = nEg + K::I0 / U;             // groups the first pass consumes
    const int64_t e0 = y0 + W - K::I0;
    const int64_t zr0 = e0 + W - K::I0;
    const float e_ident = E_MAX ? -INFINITY : INFINITY;   // "no sample" for the first pass
    const float d_ident = D_MAX ? -INFINITY : INFINITY;   // ... and for the second

    if (tid < kRoleThreads) {
        // ------------------------------------------------------------- first pass (erosion)
        const int te = tid, wi = tid >> 5, lane = tid & 31;
        const int64_t zc0 = x0 - 2 * W;
        issue_group_any<K>(p, Zs, 0, zc0, zr0, e_ident, wi, lane);
        float acc[A][C];
#pragma unroll
        for (int s = 0; s < A; ++s)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[s][c] = e_ident;
        const int64_t ecol0 = x0 - W + C * te;
        bool colok[C];
#pragma unroll
        for (int c = 0; c < C; ++c) colok[c] = (ecol0 + c >= 0) && (ecol0 + c < p.nx);
        cp_async_wait_all();
        role_barrier();
        int slot = 0;
        uint32_t phase = 0;
#pragma unroll 1
        for (int g = 0; g < nZg; ++g) {
            // group g+1 streams in while group g is consumed; its slot was last read in group
            // g-1, which every first-pass thread left through the barrier below
            if (g + 1 < nZg) issue_group_any<K>(p, Zs, g + 1, zc0, zr0, e_ident, wi, lane);
            const float* zb = Zs + (size_t)(g & 1) * U * K::COLS + C * te;
            const int kg = g - K::I0 / U;
            const bool emit = kg >= 0;
            if (emit) mbar_wait(&eempty[slot], phase ^ 1);
            float* eb = Es + (size_t)slot * U * K::COLS + C * te;
            auto put = [&](int u, float (&fin)[C]) {
                const int64_t e = e0 + (int64_t)kg * U + u;
                const bool rowok = e >= 0 && e < p.ny;
#pragma unroll
                for (int c = 0; c < C; ++c) fin[c] = (rowok && colok[c]) ? fin[c] : d_ident;
                store_vec<C>(eb + u * K::COLS, fin);
            };
            if constexpr (K::PAIR) {
                static_for<0, U / 2>([&](auto UU) {
                    constexpr int u = 2 * decltype(UU)::value;
                    float fin0[C], fin1[C];
                    chord_pair<K, E_MAX, u>(zb + u * K::COLS, zb + (u + 1) * K::COLS, acc, fin0, fin1);
                    if (emit) { put(u, fin0); put(u + 1, fin1); }
                });
            } else {
                static_for<0, U>([&](auto UU) {
                    constexpr int u = decltype(UU)::value;
                    float fin[C];
                    chord_step<K, E_MAX, u>(zb + u * K::COLS, acc, fin);
                    if (emit) put(u, fin);
                });
            }
            if (emit) {
                mbar_arrive(&efull[slot]);
                if (++slot == K::NB) { slot = 0; phase ^= 1; }
            }
#pragma unroll
            for (int s = 0; s < 2 * W; ++s)
#pragma unroll
                for (int c = 0; c < C; ++c) acc[s][c] = acc[s + U][c];
            cp_async_wait_all();
            role_barrier();
        }
    } else {
        // ------------------------------------------------------------- second pass + threshold
        const int td = tid - kRoleThreads;
        float acc[A][C];
#pragma unroll
        for (int s = 0; s < A; ++s)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[s][c] = d_ident;
        const int64_t gx = x0 + C * td;
        const bool dvalid = (td < K::XO / C) && (gx < p.nx);
        const bool vec = p.vec_ok && (gx + C - 1 < p.nx);
        int slot = 0;
        uint32_t phase = 0;
#pragma unroll 1
        for (int g = 0; g < nEg; ++g) {
            mbar_wait(&efull[slot], phase);
            const float* eb = Es + (size_t)slot * U * K::COLS + C * td;
            const int dg = g - K::I0 / U;
            // the re-read of `last` is issued before the compute so that its L2 latency hides under it
            auto fetch = [&](int u, float (&l)[C]) -> bool {
                const int64_t d = y0 + (int64_t)dg * U + u;
                const bool emit = dvalid && dg >= 0 && d < y1;
#pragma unroll
                for (int c = 0; c < C; ++c) l[c] = 0.f;
                if (emit) {
                    const int64_t off = d * p.pitch + gx;
                    if (vec) load_vec<C>(p.in + off, l);
                    else {
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            if (gx + c < p.nx) l[c] = __ldg(p.in + off + c);
                    }
                }
                return emit;
            };
            auto put = [&](int u, const float (&fin)[C], const float (&l)[C]) {
                const int64_t off = (y0 + (int64_t)dg * U + u) * p.pitch + gx;
                const int64_t moff = (y0 + (int64_t)dg * U + u) * p.nx + gx;
                if (p.out) {
                    float o[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) o[c] = NEG ? -fin[c] : fin[c];
                    if (vec) store_vec<C>(p.out + off, o);
                    else {
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            if (gx + c < p.nx) p.out[off + c] = o[c];
                    }
                }
                if (p.mask) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        // (-Z) - open(-Z) == close(Z) - Z exactly
                        const double df = NEG ? __dsub_rn((double)fin[c], (double)l[c])
                                              : __dsub_rn((double)l[c], (double)fin[c]);
                        if ((gx + c < p.nx) && (df > p.thr)) {
                            p.mask[moff + c] = 1;
                            if (p.when) p.when[moff + c] = (uint8_t)p.widx;
                        }
                    }
                }
            };
            if constexpr (K::PAIR) {
                static_for<0, U / 2>([&](auto UU) {
                    constexpr int u = 2 * decltype(UU)::value;
                    float l0[C], l1[C], fin0[C], fin1[C];
                    const bool emit0 = fetch(u, l0), emit1 = fetch(u + 1, l1);
                    chord_pair<K, D_MAX, u>(eb + u * K::COLS, eb + (u + 1) * K::COLS, acc, fin0, fin1);
                    if (emit0) put(u, fin0, l0);
                    if (emit1) put(u + 1, fin1, l1);
                });
            } else {
                static_for<0, U>([&](auto UU) {
                    constexpr int u = decltype(UU)::value;
                    float l[C], fin[C];
                    const bool emit = fetch(u, l);
                    chord_step<K, D_MAX, u>(eb + u * K::COLS, acc, fin);
                    if (emit) put(u, fin, l);
                });
            }
            mbar_arrive(&eempty[slot]);
            if (++slot == K::NB) { slot = 0; phase ^= 1; }
#pragma unroll
            for (int s = 0; s < 2 * W; ++s)
#pragma unroll
                for (int c = 0; c < C; ++c) acc[s][c] = acc[s + U][c];
        }
    }
}

}  // namespace march

template <typename K, bool NEG>
int launch_open_march_cfg(const float* in, float* out, uint8_t* mask, uint8_t* when, int64_t ny, int64_t nx,
                          int64_t pitch, double thr, int widx, int64_t row_lo, int64_t row_hi, cudaStream_t st) {
    constexpr int W = K::W;
    static bool attr_set = false;
    if (!attr_set) {
        SMRF_CUDA(cudaFuncSetAttribute(march::open_march_kernel<K, NEG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)K::kSmemBytes));
        attr_set = true;
    }
    const int64_t rows = row_hi - row_lo;
    const int nstrips = (int)((nx + K::XO - 1) / K::XO);
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
    p.seg = seg; p.thr = thr; p.widx = widx;
    p.vec_ok = (pitch % 4 == 0) && (((uintptr_t)in & 15) == 0) && (out == nullptr || ((uintptr_t)out & 15) == 0);
    dim3 grid((unsigned)nstrips, (unsigned)nsegs);
    march::open_march_kernel<K, NEG><<<grid, march::kThreads, K::kSmemBytes, st>>>(p);
    SMRF_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

template <int W, bool NEG>
int launch_open_march_f32(const float* in, float* out, uint8_t* mask, uint8_t* when, int64_t ny, int64_t nx,
                          int64_t pitch, double thr, int widx, int64_t row_lo, int64_t row_hi, cudaStream_t st) {
    return launch_open_march_cfg<march::Cfg<W>, NEG>(in, out, mask, when, ny, nx, pitch, thr, widx, row_lo, row_hi, st);
}

}  // namespace smrf
