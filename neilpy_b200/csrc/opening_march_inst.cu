// Explicit instantiations of the register-marching opening kernels for radii
// SMRF_W_LO..SMRF_W_HI (the build compiles this file once per radius group so the
// per-radius kernels build in parallel).  NEG (open -Z) is only needed for the
// low-outlier pass, which uses radius 1 (neilpy.py:1744).
#include "opening_march.cuh"

#ifndef SMRF_W_LO
#error "define SMRF_W_LO / SMRF_W_HI"
#endif

namespace smrf {
#define SMRF_INST(W, NEG)                                                                                     \
    template int launch_open_march_f32<W, NEG>(const float*, float*, uint8_t*, uint8_t*, int64_t, int64_t, int64_t, double, \
                                               int, int64_t, int64_t, cudaStream_t);
#if SMRF_W_LO <= 1 && 1 <= SMRF_W_HI
SMRF_INST(1, false)
SMRF_INST(1, true)
#endif
#if SMRF_W_LO <= 2 && 2 <= SMRF_W_HI
SMRF_INST(2, false)
#endif
#if SMRF_W_LO <= 3 && 3 <= SMRF_W_HI
SMRF_INST(3, false)
#endif
#if SMRF_W_LO <= 4 && 4 <= SMRF_W_HI
SMRF_INST(4, false)
#endif
#if SMRF_W_LO <= 5 && 5 <= SMRF_W_HI
SMRF_INST(5, false)
#endif
#if SMRF_W_LO <= 6 && 6 <= SMRF_W_HI
SMRF_INST(6, false)
#endif
#if SMRF_W_LO <= 7 && 7 <= SMRF_W_HI
SMRF_INST(7, false)
#endif
#if SMRF_W_LO <= 8 && 8 <= SMRF_W_HI
SMRF_INST(8, false)
#endif
#if SMRF_W_LO <= 9 && 9 <= SMRF_W_HI
SMRF_INST(9, false)
#endif
#if SMRF_W_LO <= 10 && 10 <= SMRF_W_HI
SMRF_INST(10, false)
#endif
#if SMRF_W_LO <= 11 && 11 <= SMRF_W_HI
SMRF_INST(11, false)
#endif
#if SMRF_W_LO <= 12 && 12 <= SMRF_W_HI
SMRF_INST(12, false)
#endif
#if SMRF_W_LO <= 13 && 13 <= SMRF_W_HI
SMRF_INST(13, false)
#endif
#if SMRF_W_LO <= 14 && 14 <= SMRF_W_HI
SMRF_INST(14, false)
#endif
#if SMRF_W_LO <= 15 && 15 <= SMRF_W_HI
SMRF_INST(15, false)
#endif
#if SMRF_W_LO <= 16 && 16 <= SMRF_W_HI
SMRF_INST(16, false)
#endif
#if SMRF_W_LO <= 17 && 17 <= SMRF_W_HI
SMRF_INST(17, false)
#endif
#if SMRF_W_LO <= 18 && 18 <= SMRF_W_HI
SMRF_INST(18, false)
#endif
#if SMRF_W_LO <= 19 && 19 <= SMRF_W_HI
SMRF_INST(19, false)
#endif
#if SMRF_W_LO <= 20 && 20 <= SMRF_W_HI
SMRF_INST(20, false)
#endif
#if SMRF_W_LO <= 21 && 21 <= SMRF_W_HI
SMRF_INST(21, false)
#endif
#if SMRF_W_LO <= 22 && 22 <= SMRF_W_HI
SMRF_INST(22, false)
#endif
#if SMRF_W_LO <= 23 && 23 <= SMRF_W_HI
SMRF_INST(23, false)
#endif
#if SMRF_W_LO <= 24 && 24 <= SMRF_W_HI
SMRF_INST(24, false)
#endif
#if SMRF_W_LO <= 25 && 25 <= SMRF_W_HI
SMRF_INST(25, false)
#endif
#if SMRF_W_LO <= 26 && 26 <= SMRF_W_HI
SMRF_INST(26, false)
#endif
#if SMRF_W_LO <= 27 && 27 <= SMRF_W_HI
SMRF_INST(27, false)
#endif
#if SMRF_W_LO <= 28 && 28 <= SMRF_W_HI
SMRF_INST(28, false)
#endif
#if SMRF_W_LO <= 29 && 29 <= SMRF_W_HI
SMRF_INST(29, false)
#endif
#if SMRF_W_LO <= 30 && 30 <= SMRF_W_HI
SMRF_INST(30, false)
#endif
#if SMRF_W_LO <= 31 && 31 <= SMRF_W_HI
SMRF_INST(31, false)
#endif
#if SMRF_W_LO <= 32 && 32 <= SMRF_W_HI
SMRF_INST(32, false)
#endif
#if SMRF_W_LO <= 33 && 33 <= SMRF_W_HI
SMRF_INST(33, false)
#endif
#if SMRF_W_LO <= 34 && 34 <= SMRF_W_HI
SMRF_INST(34, false)
#endif
#if SMRF_W_LO <= 35 && 35 <= SMRF_W_HI
SMRF_INST(35, false)
#endif
#if SMRF_W_LO <= 36 && 36 <= SMRF_W_HI
SMRF_INST(36, false)
#endif
#if SMRF_W_LO <= 37 && 37 <= SMRF_W_HI
SMRF_INST(37, false)
#endif
#if SMRF_W_LO <= 38 && 38 <= SMRF_W_HI
SMRF_INST(38, false)
#endif
#if SMRF_W_LO <= 39 && 39 <= SMRF_W_HI
SMRF_INST(39, false)
#endif
#if SMRF_W_LO <= 40 && 40 <= SMRF_W_HI
SMRF_INST(40, false)
#endif
}  // namespace smrf
