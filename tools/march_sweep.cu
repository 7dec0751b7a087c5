// Variant sweep for the register-marching opening kernel (development tool, not shipped):
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -DSWEEP_W=18 tools/march_sweep.cu -o tools/_bin/sweep18
// times every (C, PAIR, MINB, U) variant of radius SWEEP_W on an n x n float32 surface and
// checks that all variants agree bit for bit.
#include <stdarg.h>
#include <stdlib.h>
#include <vector>

#include "../neilpy_b200/csrc/opening_march.cuh"

namespace smrf {
void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n");
}
void count_launches(int) {}
bool open_no_tma() { const char* e = getenv("SMRF_OPEN_NO_TMA"); return e && e[0] == '1'; }
}
using namespace smrf;

static float *d_in, *d_out, *d_ref;
static uint8_t* d_mask;
static int64_t N;
static std::vector<float> h_ref, h_out;

template <typename K>
void run(const char* name) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, march::open_march_kernel<K, false>);
    if (fa.localSizeBytes > 256) { printf("W=%2d %-22s regs %3d local %4zu  skipped (spills)\n", K::W, name, fa.numRegs, (size_t)fa.localSizeBytes); return; }
    cudaMemset(d_out, 0xff, N * N * 4);
    cudaMemset(d_mask, 0, N * N);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rc = launch_open_march_cfg<K, false>(d_in, d_out, d_mask, nullptr, N, N, N, 0.15 * K::W, 0, 0, N, 0);
    if (rc) { printf("W=%d %s launch failed %d\n", K::W, name, rc); return; }
    cudaDeviceSynchronize();
    const int reps = 5;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) launch_open_march_cfg<K, false>(d_in, d_out, d_mask, nullptr, N, N, N, 0.15 * K::W, 0, 0, N, 0);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, march::open_march_kernel<K, false>, march::kThreads, K::kSmemBytes);
    const char* ok = "";
    if (h_ref.empty()) { h_ref.resize(N * N); cudaMemcpy(h_ref.data(), d_out, N * N * 4, cudaMemcpyDeviceToHost); ok = "ref"; }
    else {
        cudaMemcpy(h_out.data(), d_out, N * N * 4, cudaMemcpyDeviceToHost);
        ok = memcmp(h_out.data(), h_ref.data(), N * N * 4) == 0 ? "same" : "DIFFERENT";
    }
    printf("W=%2d %-22s regs %3d local %3zu occ %d smem %6zu  %8.3f ms  %7.1f Gcw/s  frac %.3f  %s %s\n", K::W, name, fa.numRegs,
           (size_t)fa.localSizeBytes, occ, K::kSmemBytes, ms, N * N / ms / 1e6, N * N * 10.0 / (ms * 1e-3) / 6549.8e9, ok,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
    fflush(stdout);
}

static float* d_tmp;
template <typename K>
void run_pass(const char* name) {
    cudaFuncAttributes fa, fb;
    cudaFuncGetAttributes(&fa, march::open_pass_kernel<K, false, false>);
    cudaFuncGetAttributes(&fb, march::open_pass_kernel<K, true, true>);
    if (fa.localSizeBytes > 256 || fb.localSizeBytes > 256) { printf("W=%2d %-26s regs %3d/%3d local %4zu/%4zu  skipped (spills)\n", K::W, name, fa.numRegs, fb.numRegs, (size_t)fa.localSizeBytes, (size_t)fb.localSizeBytes); return; }
    cudaMemset(d_out, 0xff, N * N * 4);
    cudaMemset(d_mask, 0, N * N);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rc = launch_open_passes_cfg<K>(d_in, d_out, d_tmp, d_mask, nullptr, N, N, N, 0.15 * K::W, 0, 0, N, 0);
    if (rc) { printf("W=%d %s launch failed %d\n", K::W, name, rc); return; }
    cudaDeviceSynchronize();
    const int reps = 5;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) launch_open_passes_cfg<K>(d_in, d_out, d_tmp, d_mask, nullptr, N, N, N, 0.15 * K::W, 0, 0, N, 0);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    const char* ok = "";
    if (h_ref.empty()) { h_ref.resize(N * N); cudaMemcpy(h_ref.data(), d_out, N * N * 4, cudaMemcpyDeviceToHost); ok = "ref"; }
    else {
        cudaMemcpy(h_out.data(), d_out, N * N * 4, cudaMemcpyDeviceToHost);
        ok = memcmp(h_out.data(), h_ref.data(), N * N * 4) == 0 ? "same" : "DIFFERENT";
    }
    printf("W=%2d %-26s regs %3d/%3d local %3zu smem %6zu  %8.3f ms  %7.1f Gcw/s  frac %.3f  %s %s\n", K::W, name, fa.numRegs, fb.numRegs,
           (size_t)fa.localSizeBytes, K::kSmemBytes, ms, N * N / ms / 1e6, N * N * 10.0 / (ms * 1e-3) / 6549.8e9, ok,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
    fflush(stdout);
}

int main(int argc, char** argv) {
    N = argc > 1 ? atoll(argv[1]) : 8192;
    cudaMalloc(&d_in, N * N * 4); cudaMalloc(&d_out, N * N * 4); cudaMalloc(&d_tmp, N * N * 4); cudaMalloc(&d_mask, N * N);
    std::vector<float> h(N * N);
    unsigned s = 12345;
    for (int64_t y = 0; y < N; ++y)
        for (int64_t x = 0; x < N; ++x) {
            s = s * 1664525u + 1013904223u;
            h[y * N + x] = 100.f + 20.f * sinf(x * 0.01f) * cosf(y * 0.013f) + (s >> 8) * (1.0f / 16777216.f) + (((x / 97) + (y / 83)) % 7 == 0 ? 12.f : 0.f);
        }
    cudaMemcpy(d_in, h.data(), N * N * 4, cudaMemcpyHostToDevice);
    h_out.resize(N * N);
    constexpr int W = SWEEP_W;
    using namespace march;
#if SWEEP_W <= 40
    run<Cfg<W>>("shipped fused Cfg<W>");
    run<CfgT<W, 4, true, 1, 4, 3>>("fused C4 pair MINB1");
    run<CfgT<W, 2, true, 2, 4, 3>>("fused C2 pair MINB2");
    run_pass<PassCfg<W, 256, 4, true, 1>>("2pass T256 C4 pair MINB1");
    run_pass<PassCfg<W, 128, 4, true, 2>>("2pass T128 C4 pair MINB2");
    run_pass<PassCfg<W, 256, 2, true, 2>>("2pass T256 C2 pair MINB2");
    run_pass<PassCfg<W, 256, 2, true, 1>>("2pass T256 C2 pair MINB1");
    run_pass<PassCfg<W, 128, 4, true, 3>>("2pass T128 C4 pair MINB3");
#else
    run_pass<PassCfg<W, 256, 1, true, 1>>("2pass T256 C1 pair MINB1");
    run_pass<PassCfg<W, 256, 1, false, 1>>("2pass T256 C1 sngl MINB1");
    run_pass<PassCfg<W, 256, 2, false, 1>>("2pass T256 C2 sngl MINB1");
#endif
    return 0;
}
