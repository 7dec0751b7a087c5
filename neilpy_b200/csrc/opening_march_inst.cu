// The marching opening kernels of ONE disk radius (-DSMRF_W=<radius>): the kernels are fully unrolled per
// radius, so the build compiles this file once per radius and the units build in parallel.
// NEG (open -Z) is only needed for the low-outlier pass, which uses radius 1 (neilpy.py:1744).
#include "opening_march.cuh"

#ifndef SMRF_W
#error "define SMRF_W"
#endif

namespace smrf {

template <>
int launch_open_radius_f32<SMRF_W>(const float* in, float* out, float* tmp, uint8_t* mask, uint8_t* when, int64_t ny,
                                   int64_t nx, int64_t pitch, double thr, int widx, int negate, int64_t row_lo,
                                   int64_t row_hi, cudaStream_t st) {
#if SMRF_W == 1
    if (negate) return launch_open_march_f32<1, true>(in, out, mask, when, ny, nx, pitch, thr, widx, row_lo, row_hi, st);
#endif
    if (negate) {
        set_error("open_window_march: negate is instantiated for radius 1 only");
        return SMRF_E_UNSUPPORTED;
    }
#if SMRF_W <= SMRF_FUSED_MAX_W
    (void)tmp;
    return launch_open_march_f32<SMRF_W, false>(in, out, mask, when, ny, nx, pitch, thr, widx, row_lo, row_hi, st);
#else
    return launch_open_passes_f32<SMRF_W>(in, out, tmp, mask, when, ny, nx, pitch, thr, widx, row_lo, row_hi, st);
#endif
}

}  // namespace smrf
