// Dispatch from a run-time radius to the per-radius register-marching kernels.
#include <stdlib.h>
#include <string.h>

#include "opening.cuh"

namespace smrf {

#define SMRF_DECL(W)                                                                                              \
    extern template int launch_open_march_f32<W, false>(const float*, float*, uint8_t*, uint8_t*, int64_t, int64_t, \
                                                        int64_t, double, int, int64_t, int64_t, cudaStream_t);
SMRF_DECL(1) SMRF_DECL(2) SMRF_DECL(3) SMRF_DECL(4) SMRF_DECL(5) SMRF_DECL(6) SMRF_DECL(7) SMRF_DECL(8) SMRF_DECL(9) SMRF_DECL(10) SMRF_DECL(11) SMRF_DECL(12) SMRF_DECL(13) SMRF_DECL(14) SMRF_DECL(15) SMRF_DECL(16) SMRF_DECL(17) SMRF_DECL(18) SMRF_DECL(19) SMRF_DECL(20) SMRF_DECL(21) SMRF_DECL(22) SMRF_DECL(23) SMRF_DECL(24) SMRF_DECL(25) SMRF_DECL(26) SMRF_DECL(27) SMRF_DECL(28) SMRF_DECL(29) SMRF_DECL(30) SMRF_DECL(31) SMRF_DECL(32) SMRF_DECL(33) SMRF_DECL(34) SMRF_DECL(35) SMRF_DECL(36) SMRF_DECL(37) SMRF_DECL(38) SMRF_DECL(39) SMRF_DECL(40)
#undef SMRF_DECL
extern template int launch_open_march_f32<1, true>(const float*, float*, uint8_t*, uint8_t*, int64_t, int64_t, int64_t,
                                                   double, int, int64_t, int64_t, cudaStream_t);

bool open_march_available(int dtype, int window, int negate) {
    if (dtype != SMRF_F32) return false;
    if (negate) return window == 1;
    return window >= 1 && window <= SMRF_MARCH_MAX_W;
}

const char* open_march_name(int dtype, int window) {
    (void)dtype;
    if (window == 2 || window > 24) return "march_f32_fused";
    return "march_f32_fused_rowpair";
}

bool open_force_generic() {
    const char* e = getenv("SMRF_OPEN_IMPL");
    return e && strcmp(e, "generic") == 0;
}

int open_window_march(const void* in, void* out, uint8_t* mask, uint8_t* when, int64_t ny, int64_t nx, int64_t pitch,
                      int dtype, int w, double thr, int widx, int negate, int64_t row_lo, int64_t row_hi,
                      cudaStream_t st) {
    if (dtype != SMRF_F32) {
        set_error("open_window_march: float32 only");
        return SMRF_E_UNSUPPORTED;
    }
    const float* i = (const float*)in;
    float* o = (float*)out;
    if (negate) {
        if (w == 1) return launch_open_march_f32<1, true>(i, o, mask, when, ny, nx, pitch, thr, widx, row_lo, row_hi, st);
        set_error("open_window_march: negate is instantiated for radius 1 only");
        return SMRF_E_UNSUPPORTED;
    }
#define SMRF_CASE(W) \
    case W: return launch_open_march_f32<W, false>(i, o, mask, when, ny, nx, pitch, thr, widx, row_lo, row_hi, st);
    switch (w) {
        SMRF_CASE(1) SMRF_CASE(2) SMRF_CASE(3) SMRF_CASE(4) SMRF_CASE(5) SMRF_CASE(6) SMRF_CASE(7) SMRF_CASE(8) SMRF_CASE(9) SMRF_CASE(10) SMRF_CASE(11) SMRF_CASE(12) SMRF_CASE(13) SMRF_CASE(14) SMRF_CASE(15) SMRF_CASE(16) SMRF_CASE(17) SMRF_CASE(18) SMRF_CASE(19) SMRF_CASE(20) SMRF_CASE(21) SMRF_CASE(22) SMRF_CASE(23) SMRF_CASE(24) SMRF_CASE(25) SMRF_CASE(26) SMRF_CASE(27) SMRF_CASE(28) SMRF_CASE(29) SMRF_CASE(30) SMRF_CASE(31) SMRF_CASE(32) SMRF_CASE(33) SMRF_CASE(34) SMRF_CASE(35) SMRF_CASE(36) SMRF_CASE(37) SMRF_CASE(38) SMRF_CASE(39) SMRF_CASE(40)
        default: break;
    }
#undef SMRF_CASE
    set_error("open_window_march: radius %d not instantiated", w);
    return SMRF_E_UNSUPPORTED;
}

}  // namespace smrf
