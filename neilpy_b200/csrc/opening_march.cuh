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
//                    accumulator of output row (y - dy).  The 2W+1 live output rows
//                    per column sit in registers; the row loop is unrolled 2W+1 times
//                    so that the rotating accumulator index is static.
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

template <int W>
struct Cfg {
    static constexpr int N = 2 * W + 1;
    static constexpr int C = 4;
    static constexpr int EW = kRoleThreads * C;                // erosion columns per CTA
    static constexpr int NL = 2 * W + C;                       // elements a thread reads per row
    static constexpr int NQ = (NL + 3) / 4;                    // ... as 16-byte quads
    static constexpr int COLS = (kRoleThreads - 1) * C + NQ * 4;  // ring row length (floats)
    static constexpr int XO = ((EW - 2 * W) / 4) * 4;          // output columns per CTA
    static constexpr int RB = W <= 6 ? 8 : 4;                  // rows per hand-over batch
    static constexpr int NB = 3;                               // batches in the Es ring
    static constexpr int ZRING = 2 * RB;                       // Zs: double buffer
    static constexpr int ERING = NB * RB;
    static constexpr int MINB = W <= 2 ? 3 : (W <= 5 ? 2 : 1);  // CTAs per SM the register budget allows
    static constexpr size_t kSmemBytes = (size_t)(ZRING + ERING) * COLS * sizeof(float) + 2 * NB * sizeof(uint64_t);

    __host__ __device__ static constexpr int isqrt(int v) {
        int h = 0;
        while ((h + 1) * (h + 1) <= v) ++h;
        return h;
    }
    __host__ __device__ static constexpr int half(int dy) { return isqrt(W * W - dy * dy); }
};

struct Params {
    const float* in;
    float* out;
    uint8_t* mask;
    uint8_t* when;
    int64_t ny, nx, row_lo, row_hi;
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
__device__ __forceinline__ void role_barrier() {   // the 128 E threads only (barrier 0 is __syncthreads)
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

// One incoming ring row for one thread: updates the 2W+1 rotating accumulators of its C
// columns and returns in `fin` the output row that this row completes (local row i - W).
// J = i mod (2W+1) is static.  srow points at the thread's first element: local index
// W + c is the centre of column c.
template <int W, bool IS_MAX, int J>
__device__ __forceinline__ void chord_step(const float* __restrict__ srow, float (&acc)[2 * W + 1][4], float (&fin)[4]) {
    using K = Cfg<W>;
    constexpr int N = K::N;
    float z[K::NQ * 4];
    const float4* q = reinterpret_cast<const float4*>(srow);
#pragma unroll
    for (int i = 0; i < K::NQ; ++i) {
        float4 v = q[i];
        z[4 * i + 0] = v.x; z[4 * i + 1] = v.y; z[4 * i + 2] = v.z; z[4 * i + 3] = v.w;
    }
    float R[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) R[c] = z[W + c];
    {   // chord of dy = +-W has half-length 0: the oldest output gets its last term, the newest its first
        constexpr int s_old = (J + N - W) % N;
        constexpr int s_new = (J + W) % N;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            fin[c] = op2<IS_MAX>(acc[s_old][c], R[c]);
            acc[s_new][c] = R[c];
        }
    }
    static_for<1, W + 1>([&](auto H) {
        constexpr int h = decltype(H)::value;
#pragma unroll
        for (int c = 0; c < 4; ++c) R[c] = op3<IS_MAX>(R[c], z[W + c - h], z[W + c + h]);
        static_for<0, W>([&](auto DY) {
            constexpr int dy = decltype(DY)::value;
            if constexpr (K::half(dy) == h) {
                if constexpr (dy == 0) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[J][c] = op2<IS_MAX>(acc[J][c], R[c]);
                } else {
                    constexpr int s1 = (J + N - dy) % N;
                    constexpr int s2 = (J + dy) % N;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        acc[s1][c] = op2<IS_MAX>(acc[s1][c], R[c]);
                        acc[s2][c] = op2<IS_MAX>(acc[s2][c], R[c]);
                    }
                }
            }
        });
    });
}

// E threads: issue the copies of batch `b` of `last` rows [zr0 + b*RB, +RB) into its Zs slot.
// VL floats per copy (16/8/4 bytes); anything outside the image is written as `ident`.
template <int W, int VL>
__device__ __forceinline__ void issue_batch(const Params& p, float* Zs, int b, int nZ, int64_t zc0, int64_t zr0,
                                            float ident, int te) {
    using K = Cfg<W>;
    constexpr int NCH = K::COLS / VL;   // chunks per row (COLS is a multiple of 4)
    constexpr int TOT = NCH * K::RB;
    float* slot = Zs + (size_t)(b & 1) * K::RB * K::COLS;
#pragma unroll 2
    for (int ch = te; ch < TOT; ch += kRoleThreads) {
        const int rr = ch / NCH;
        const int cc = ch - rr * NCH;
        const int i = b * K::RB + rr;
        const int64_t r = zr0 + i;
        const int64_t g = zc0 + (int64_t)cc * VL;
        float* dst = slot + (size_t)rr * K::COLS + cc * VL;
        const bool rowok = i < nZ && r >= 0 && r < p.ny;
        if (rowok && g >= 0 && g + VL <= p.nx) {
            cp_async<4 * VL>(dst, p.in + r * p.nx + g);
        } else {
#pragma unroll
            for (int e = 0; e < VL; ++e) {
                const bool ok = rowok && g + e >= 0 && g + e < p.nx;
                dst[e] = ok ? __ldg(p.in + r * p.nx + g + e) : ident;
            }
        }
    }
}

template <int W>
__device__ __forceinline__ void issue_batch_any(const Params& p, float* Zs, int b, int nZ, int64_t zc0, int64_t zr0,
                                                float ident, int te) {
    if (p.vec_ok && (W % 2 == 0)) issue_batch<W, 4>(p, Zs, b, nZ, zc0, zr0, ident, te);
    else if (p.vec_ok) issue_batch<W, 2>(p, Zs, b, nZ, zc0, zr0, ident, te);
    else issue_batch<W, 1>(p, Zs, b, nZ, zc0, zr0, ident, te);
}

template <int W, bool NEG>
__global__ void __launch_bounds__(kThreads, Cfg<W>::MINB) open_march_kernel(const Params p) {
    using K = Cfg<W>;
    constexpr int N = K::N;
    constexpr bool E_MAX = NEG;        // erosion of -Z is -(dilation of Z)
    constexpr bool D_MAX = !NEG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Zs = reinterpret_cast<float*>(smem_raw);
    float* Es = Zs + (size_t)K::ZRING * K::COLS;
    uint64_t* efull = reinterpret_cast<uint64_t*>(Es + (size_t)K::ERING * K::COLS);
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
    const int nZ = nOut + 4 * W;   // rows of `last` consumed: [y0 - 2W, y1 + 2W)
    const int nE = nOut + 2 * W;   // erosion rows produced:  [y0 - W, y1 + W)
    const float e_ident = E_MAX ? -INFINITY : INFINITY;   // "no sample" for the first pass
    const float d_ident = D_MAX ? -INFINITY : INFINITY;   // ... and for the second

    if (tid < kRoleThreads) {
        // ------------------------------------------------------------- first pass (erosion)
        const int te = tid;
        const int64_t zc0 = x0 - 2 * W, zr0 = y0 - 2 * W;
        issue_batch_any<W>(p, Zs, 0, nZ, zc0, zr0, e_ident, te);
        float acc[N][4];
#pragma unroll
        for (int s = 0; s < N; ++s)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[s][c] = e_ident;
        const int64_t ecol0 = x0 - W + 4 * te;
        bool colok[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) colok[c] = (ecol0 + c >= 0) && (ecol0 + c < p.nx);
        cp_async_wait_all();
        role_barrier();
        for (int ib = 0; ib < nZ; ib += N) {
            static_for<0, N>([&](auto JJ) {
                constexpr int J = decltype(JJ)::value;
                const int i = ib + J;
                if (i < nZ) {
                    const int b = i / K::RB;
                    // batch b+1 streams in while batch b is consumed; its slot was last read in
                    // batch b-1, which every E thread left through the barrier below
                    if ((i % K::RB) == 0 && (b + 1) * K::RB < nZ) issue_batch_any<W>(p, Zs, b + 1, nZ, zc0, zr0, e_ident, te);
                    float fin[4];
                    chord_step<W, E_MAX, J>(Zs + (size_t)(i % K::ZRING) * K::COLS + 4 * te, acc, fin);
                    if (i >= 2 * W) {
                        const int k = i - 2 * W;
                        const int kb = k / K::RB;
                        if ((k % K::RB) == 0) mbar_wait(&eempty[kb % K::NB], ((kb / K::NB) & 1) ^ 1);
                        const int64_t e = y0 - W + k;
                        const bool rowok = e >= 0 && e < p.ny;
                        float4 v;
                        v.x = (rowok && colok[0]) ? fin[0] : d_ident;
                        v.y = (rowok && colok[1]) ? fin[1] : d_ident;
                        v.z = (rowok && colok[2]) ? fin[2] : d_ident;
                        v.w = (rowok && colok[3]) ? fin[3] : d_ident;
                        *reinterpret_cast<float4*>(Es + (size_t)(k % K::ERING) * K::COLS + 4 * te) = v;
                        if ((k % K::RB) == K::RB - 1 || k == nE - 1) mbar_arrive(&efull[kb % K::NB]);
                    }
                    if ((i % K::RB) == K::RB - 1) {
                        cp_async_wait_all();
                        role_barrier();
                    }
                }
            });
        }
    } else {
        // ------------------------------------------------------------- second pass + threshold
        const int td = tid - kRoleThreads;
        float acc[N][4];
#pragma unroll
        for (int s = 0; s < N; ++s)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[s][c] = d_ident;
        const int64_t gx = x0 + 4 * td;
        const bool dvalid = (td < K::XO / 4) && (gx < p.nx);
        const bool vec = p.vec_ok && (gx + 3 < p.nx);
        for (int ib = 0; ib < nE; ib += N) {
            static_for<0, N>([&](auto JJ) {
                constexpr int J = decltype(JJ)::value;
                const int k = ib + J;
                if (k < nE) {
                    const int b = k / K::RB;
                    if ((k % K::RB) == 0) mbar_wait(&efull[b % K::NB], (b / K::NB) & 1);
                    const bool emit = dvalid && (k >= 2 * W);
                    const int64_t off = (y0 + (k - 2 * W)) * p.nx + gx;
                    float l[4] = {0.f, 0.f, 0.f, 0.f};
                    if (emit) {   // issue the re-read of `last` before the compute so L2 latency hides under it
                        if (vec) {
                            float4 t = __ldg(reinterpret_cast<const float4*>(p.in + off));
                            l[0] = t.x; l[1] = t.y; l[2] = t.z; l[3] = t.w;
                        } else {
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                if (gx + c < p.nx) l[c] = __ldg(p.in + off + c);
                        }
                    }
                    float fin[4];
                    chord_step<W, D_MAX, J>(Es + (size_t)(k % K::ERING) * K::COLS + 4 * td, acc, fin);
                    if ((k % K::RB) == K::RB - 1 || k == nE - 1) mbar_arrive(&eempty[b % K::NB]);
                    if (emit) {
                        if (p.out) {
                            float o[4];
#pragma unroll
                            for (int c = 0; c < 4; ++c) o[c] = NEG ? -fin[c] : fin[c];
                            if (vec) *reinterpret_cast<float4*>(p.out + off) = make_float4(o[0], o[1], o[2], o[3]);
                            else {
#pragma unroll
                                for (int c = 0; c < 4; ++c)
                                    if (gx + c < p.nx) p.out[off + c] = o[c];
                            }
                        }
                        if (p.mask) {
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                // (-Z) - open(-Z) == close(Z) - Z exactly
                                const double d = NEG ? __dsub_rn((double)fin[c], (double)l[c])
                                                     : __dsub_rn((double)l[c], (double)fin[c]);
                                if ((gx + c < p.nx) && (d > p.thr)) {
                                    p.mask[off + c] = 1;
                                    if (p.when) p.when[off + c] = (uint8_t)p.widx;
                                }
                            }
                        }
                    }
                }
            });
        }
    }
}

}  // namespace march

template <int W, bool NEG>
int launch_open_march_f32(const float* in, float* out, uint8_t* mask, uint8_t* when, int64_t ny, int64_t nx,
                          double thr, int widx, int64_t row_lo, int64_t row_hi, cudaStream_t st) {
    using K = march::Cfg<W>;
    static bool attr_set = false;
    if (!attr_set) {
        SMRF_CUDA(cudaFuncSetAttribute(march::open_march_kernel<W, NEG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)K::kSmemBytes));
        attr_set = true;
    }
    const int64_t rows = row_hi - row_lo;
    const int nstrips = (int)((nx + K::XO - 1) / K::XO);
    // Segment the rows so that (waves of CTAs) x (rows marched per CTA, incl. the 4W warm-up rows)
    // is smallest: few long segments waste SMs, many short ones waste warm-up.
    const int64_t slots = (int64_t)num_sms() * K::MINB;
    const int64_t min_seg = 4 * W < 32 ? 32 : 4 * W;
    int64_t max_segs = rows / min_seg;
    if (max_segs < 1) max_segs = 1;
    if (max_segs > 4096) max_segs = 4096;
    int64_t best_cost = -1, best_n = 1;
    for (int64_t n = 1; n <= max_segs; ++n) {
        const int64_t ctas = n * nstrips;
        const int64_t waves = (ctas + slots - 1) / slots;
        const int64_t cost = waves * ((rows + n - 1) / n + 4 * W);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_n = n; }
    }
    const int seg = (int)((rows + best_n - 1) / best_n);
    const int nsegs = (int)((rows + seg - 1) / seg);
    march::Params p;
    p.in = in; p.out = out; p.mask = mask; p.when = when;
    p.ny = ny; p.nx = nx; p.row_lo = row_lo; p.row_hi = row_hi;
    p.seg = seg; p.thr = thr; p.widx = widx;
    p.vec_ok = (nx % 4 == 0) && (((uintptr_t)in & 15) == 0) && (out == nullptr || ((uintptr_t)out & 15) == 0);
    dim3 grid((unsigned)nstrips, (unsigned)nsegs);
    march::open_march_kernel<W, NEG><<<grid, march::kThreads, K::kSmemBytes, st>>>(p);
    SMRF_LAUNCH_CHECK();
    return 0;
}

}  // namespace smrf
