// Shared helpers for libsmrf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/smrf_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsmrf_b200 is written for sm_100a (B200) only"
#endif

namespace smrf {

void set_error(const char* fmt, ...);
void count_launches(int n);   // kernels launched by this process (smrf_launch_count)

#define SMRF_CHECK_ARG(cond, msg)                              \
    do {                                                       \
        if (!(cond)) {                                         \
            ::smrf::set_error("%s: %s", __func__, msg);        \
            return SMRF_E_ARG;                                 \
        }                                                      \
    } while (0)

#define SMRF_CUDA(call)                                                                 \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            ::smrf::set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e__)); \
            return (int)e__;                                                            \
        }                                                                               \
    } while (0)

#define SMRF_LAUNCH_CHECK() SMRF_CUDA(cudaGetLastError())

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// ---- order-preserving float <-> signed-int keys (atomicMin/atomicMax on ints) ----
__device__ __forceinline__ int f32_key(float v) {
    int b = __float_as_int(v);
    return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float f32_unkey(int k) {
    return __int_as_float(k ^ ((k >> 31) & 0x7fffffff));
}
__device__ __forceinline__ long long f64_key(double v) {
    long long b = __double_as_longlong(v);
    return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double f64_unkey(long long k) {
    return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL));
}

template <typename T>
struct KeyOf;
template <>
struct KeyOf<float> {
    using type = int;
    static __device__ __forceinline__ int key(float v) { return f32_key(v); }
    static __device__ __forceinline__ float unkey(int k) { return f32_unkey(k); }
    static constexpr int empty_min = 0x7fffffff;            // above key(+inf)
    static constexpr int empty_max = (int)0x80000000;       // below key(-inf)
};
template <>
struct KeyOf<double> {
    using type = long long;
    static __device__ __forceinline__ long long key(double v) { return f64_key(v); }
    static __device__ __forceinline__ double unkey(long long k) { return f64_unkey(k); }
    static constexpr long long empty_min = 0x7fffffffffffffffLL;
    static constexpr long long empty_max = (long long)0x8000000000000000ULL;
};

template <typename T>
__device__ __forceinline__ T quiet_nan();
template <>
__device__ __forceinline__ float quiet_nan<float>() { return __int_as_float(0x7fc00000); }
template <>
__device__ __forceinline__ double quiet_nan<double>() { return __longlong_as_double(0x7ff8000000000000LL); }

// ---- point stream loaders: everything is widened to float64 before any arithmetic ----
template <int FMT>
struct PointLoader;
template <>
struct PointLoader<SMRF_PTS_SOA_F64> {
    const double *x, *y, *z;
    __device__ __forceinline__ void xy(int64_t i, double& px, double& py) const { px = x[i]; py = y[i]; }
    __device__ __forceinline__ void xyz(int64_t i, double& px, double& py, double& pz) const {
        px = x[i]; py = y[i]; pz = z[i];
    }
};
template <>
struct PointLoader<SMRF_PTS_XYZW_F32> {
    const float4* p;
    const void *unused_y, *unused_z;
    __device__ __forceinline__ void xy(int64_t i, double& px, double& py) const {
        float4 v = __ldg(p + i); px = (double)v.x; py = (double)v.y;
    }
    __device__ __forceinline__ void xyz(int64_t i, double& px, double& py, double& pz) const {
        float4 v = __ldg(p + i); px = (double)v.x; py = (double)v.y; pz = (double)v.z;
    }
};
template <>
struct PointLoader<SMRF_PTS_SOA_F32> {
    const float *x, *y, *z;
    __device__ __forceinline__ void xy(int64_t i, double& px, double& py) const { px = (double)x[i]; py = (double)y[i]; }
    __device__ __forceinline__ void xyz(int64_t i, double& px, double& py, double& pz) const {
        px = (double)x[i]; py = (double)y[i]; pz = (double)z[i];
    }
};

// ~t * (x, y) exactly as affine.Affine.__mul__ evaluates it on float64 arrays:
// (vx*sa + vy*sb) + sc, every product and sum rounded separately (no FMA).
struct Inv6 {
    double ra, rb, rc, rd, re, rf;
};
__device__ __forceinline__ void affine_apply(const Inv6& t, double x, double y, double& c, double& r) {
    c = __dadd_rn(__dadd_rn(__dmul_rn(x, t.ra), __dmul_rn(y, t.rb)), t.rc);
    r = __dadd_rn(__dadd_rn(__dmul_rn(x, t.rd), __dmul_rn(y, t.re)), t.rf);
}

}  // namespace smrf
