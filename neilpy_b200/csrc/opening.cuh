// Internal interface between the opening translation units.
#pragma once
#include "common.cuh"

namespace smrf {

// radii 1..SMRF_FUSED_MAX_W run the fused two-role marching kernel (one pass over HBM), radii up to
// SMRF_MARCH_MAX_W the two single-role passes (opening_march.cuh; the split follows the B200 sweeps);
// the build compiles one translation unit per radius
#define SMRF_FUSED_MAX_W 6
#ifndef SMRF_MARCH_MAX_W
#define SMRF_MARCH_MAX_W 72
#endif

bool open_march_available(int dtype, int window, int negate);
const char* open_march_name(int dtype, int window);
bool open_force_generic();  // env SMRF_OPEN_IMPL=generic
bool open_no_tma();         // env SMRF_OPEN_NO_TMA=1: the cp.async loader even where TMA is possible (diagnostics)

// `pitch` = row stride of in / out / tmp in elements (>= nx); mask and when are nx wide
int open_window_march(const void* in, void* out, void* tmp, uint8_t* mask, uint8_t* when, int64_t ny, int64_t nx,
                      int64_t pitch, int dtype, int w, double thr, int widx, int negate, int64_t row_lo, int64_t row_hi,
                      cudaStream_t st);

int open_window_generic(const void* in, void* out, void* tmp, uint8_t* mask, uint8_t* when, int64_t ny, int64_t nx,
                        int64_t pitch, int dtype, int w, double thr, int widx, int negate, int64_t row_lo,
                        int64_t row_hi, cudaStream_t st);

// per-radius launcher, one explicit instantiation per radius (opening_march_inst.cu, -DSMRF_W=<radius>)
template <int W>
int launch_open_radius_f32(const float* in, float* out, float* tmp, uint8_t* mask, uint8_t* when, int64_t ny, int64_t nx,
                           int64_t pitch, double thr, int widx, int negate, int64_t row_lo, int64_t row_hi,
                           cudaStream_t st);

}  // namespace smrf
