// Internal interface between the opening translation units.
#pragma once
#include "common.cuh"

namespace smrf {

// largest disk radius the register-marching kernels are instantiated for
#ifndef SMRF_MARCH_MAX_W
#define SMRF_MARCH_MAX_W 40
#endif

bool open_march_available(int dtype, int window, int negate);
const char* open_march_name(int dtype, int window);
bool open_force_generic();  // env SMRF_OPEN_IMPL=generic

// `pitch` = row stride of in / out / tmp in elements (>= nx); mask and when are nx wide
int open_window_march(const void* in, void* out, uint8_t* mask, uint8_t* when, int64_t ny, int64_t nx, int64_t pitch,
                      int dtype, int w, double thr, int widx, int negate, int64_t row_lo, int64_t row_hi,
                      cudaStream_t st);

int open_window_generic(const void* in, void* out, void* tmp, uint8_t* mask, uint8_t* when, int64_t ny, int64_t nx,
                        int64_t pitch, int dtype, int w, double thr, int widx, int negate, int64_t row_lo,
                        int64_t row_hi, cudaStream_t st);

// per-radius launchers, one explicit instantiation per (W) spread over several TUs
template <int W, bool NEG>
int launch_open_march_f32(const float* in, float* out, uint8_t* mask, uint8_t* when, int64_t ny, int64_t nx,
                          int64_t pitch, double thr, int widx, int64_t row_lo, int64_t row_hi, cudaStream_t st);

}  // namespace smrf
