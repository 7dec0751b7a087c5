// Issue-rate microbenchmark for the min/max instructions the opening kernels are made of (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 tools/fmnmx_bench.cu -o tools/_bin/fmnmx_bench
// Every thread runs ILP independent dependency chains of one instruction kind for ITER iterations; the
// kernel is launched with enough warps to fill every scheduler, and the rate is reported as
// thread-instructions per clock per SM (SM clocks from clock64 on one SM, confirmed by the event time).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

enum Kind { FMNMX2 = 0, FMNMX3 = 1, FFMA = 2, MIX_FMNMX3_FFMA = 3, IMNMX3 = 4, FMNMX3_CHAIN1 = 5, MIX_FMNMX3_LDS = 6 };

template <int KIND, int ILP>
__global__ void __launch_bounds__(256) bench(float* out, const float* in, int iters, long long* clocks) {
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = in[i];
    __syncthreads();
    float a[ILP], b[ILP];
    int ia[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        a[i] = in[threadIdx.x + i];
        b[i] = in[threadIdx.x + 32 + i];
        ia[i] = __float_as_int(a[i]);
    }
    const float c0 = in[threadIdx.x & 31], c1 = in[(threadIdx.x & 31) + 64];
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (KIND == FMNMX2) {
                    asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
                    asm volatile("max.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(c0));
                } else if (KIND == FMNMX3 || KIND == FMNMX3_CHAIN1) {
                    asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c0));
                    asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(b[i]) : "f"(a[i]), "f"(c1));
                } else if (KIND == FFMA) {
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c0), "f"(b[i]));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b[i]) : "f"(c1), "f"(a[i]));
                }
                else if (KIND == MIX_FMNMX3_FFMA) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c0)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b[i]) : "f"(c1), "f"(c0)); }
                else if (KIND == IMNMX3) { ia[i] = __vimin3_s32(ia[i], __float_as_int(b[i]) + r, __float_as_int(c0)); }
                else if (KIND == MIX_FMNMX3_LDS) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c0)); if ((i & 3) == 0) b[i] = sm[(threadIdx.x + r * 32 + i) & 1023]; }
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i] + b[i] + __int_as_float(ia[i]);
    out[blockIdx.x * 256 + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clocks = t1 - t0;
}

template <int KIND, int ILP>
void run(const char* name, int per_iter_instr, int ctas_per_sm, float* out, float* in, long long* dclk, int sms) {
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<KIND, ILP><<<sms * ctas_per_sm, 256>>>(out, in, 10, dclk);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench<KIND, ILP><<<sms * ctas_per_sm, 256>>>(out, in, iters, dclk);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long clk = 0; cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost);
    const double instr_per_sm = (double)iters * 8 * ILP * per_iter_instr * 256.0 * ctas_per_sm;   // thread-instructions per SM
    printf("%-28s ILP %2d  %d x 256 threads/SM  %7.1f thread-instr/clk/SM  (%.3f ms, %lld clk on SM0, %.0f MHz)\n", name, ILP,
           ctas_per_sm, instr_per_sm / (double)clk, ms, clk, clk / (ms * 1e3));
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out, *in; long long* dclk;
    cudaMalloc(&out, sms * 8 * 256 * 4); cudaMalloc(&in, 4096 * 4); cudaMalloc(&dclk, 8);
    float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 1.f + (i * 37 % 101) * 0.01f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int c : {1, 2, 4}) {
        if (c == 1) {
            run<FMNMX2, 8>("FMNMX (2-input)", 2, 1, out, in, dclk, sms);
            run<FMNMX3, 8>("FMNMX3 (3-input)", 2, 1, out, in, dclk, sms);
            run<FMNMX3_CHAIN1, 1>("FMNMX3 one chain", 2, 1, out, in, dclk, sms);
            run<FMNMX3, 2>("FMNMX3 two chains", 2, 1, out, in, dclk, sms);
            run<FMNMX3, 4>("FMNMX3 four chains", 2, 1, out, in, dclk, sms);
            run<FFMA, 8>("FFMA", 2, 1, out, in, dclk, sms);
            run<MIX_FMNMX3_FFMA, 8>("FMNMX3 + FFMA (1:1)", 2, 1, out, in, dclk, sms);
            run<IMNMX3, 8>("VIMNMX3 (s32, 3-input)", 1, 1, out, in, dclk, sms);
            run<MIX_FMNMX3_LDS, 8>("FMNMX3 + LDS (4:1)", 1, 1, out, in, dclk, sms);
        } else if (c == 2) {
            run<FMNMX2, 8>("FMNMX (2-input)", 2, 2, out, in, dclk, sms);
            run<FMNMX3, 8>("FMNMX3 (3-input)", 2, 2, out, in, dclk, sms);
            run<FFMA, 8>("FFMA", 2, 2, out, in, dclk, sms);
            run<MIX_FMNMX3_FFMA, 8>("FMNMX3 + FFMA (1:1)", 2, 2, out, in, dclk, sms);
        } else {
            run<FMNMX2, 8>("FMNMX (2-input)", 2, 4, out, in, dclk, sms);
            run<FMNMX3, 8>("FMNMX3 (3-input)", 2, 4, out, in, dclk, sms);
            run<FFMA, 8>("FFMA", 2, 4, out, in, dclk, sms);
        }
    }
    return 0;
}
