// Dispatch from a run-time radius to the per-radius marching kernels (one translation unit per radius).
#include <stdlib.h>
#include <string.h>

#include "opening.cuh"

namespace smrf {

typedef int (*RadiusFn)(const float*, float*, float*, uint8_t*, uint8_t*, int64_t, int64_t, int64_t, double, int, int,
                        int64_t, int64_t, cudaStream_t);

template <int W>
struct Table {
    static void fill(RadiusFn* t) {
        t[W] = &launch_open_radius_f32<W>;
        Table<W - 1>::fill(t);
    }
};
template <>
struct Table<0> {
    static void fill(RadiusFn*) {}
};

static RadiusFn radius_fn(int w) {
    static RadiusFn table[SMRF_MARCH_MAX_W + 1] = {};
    static bool ready = false;
    if (!ready) {
        Table<SMRF_MARCH_MAX_W>::fill(table);
        ready = true;
    }
    return (w >= 1 && w <= SMRF_MARCH_MAX_W) ? table[w] : nullptr;
}

bool open_march_available(int dtype, int window, int negate) {
    if (dtype != SMRF_F32) return false;
    if (negate) return window == 1;
    return window >= 1 && window <= SMRF_MARCH_MAX_W;
}

const char* open_march_name(int dtype, int window) {
    (void)dtype;
    if (window > SMRF_FUSED_MAX_W) return "march_f32_two_pass_rowpair_tma";
    return "march_f32_fused_rowpair_tma";
}

bool open_force_generic() {
    const char* e = getenv("SMRF_OPEN_IMPL");
    return e && strcmp(e, "generic") == 0;
}

bool open_no_tma() {
    const char* e = getenv("SMRF_OPEN_NO_TMA");
    return e && e[0] == '1';
}

int open_window_march(const void* in, void* out, void* tmp, uint8_t* mask, uint8_t* when, int64_t ny, int64_t nx,
                      int64_t pitch, int dtype, int w, double thr, int widx, int negate, int64_t row_lo, int64_t row_hi,
                      cudaStream_t st) {
    if (dtype != SMRF_F32) {
        set_error("open_window_march: float32 only");
        return SMRF_E_UNSUPPORTED;
    }
    RadiusFn fn = radius_fn(w);
    if (!fn) {
        set_error("open_window_march: radius %d not instantiated", w);
        return SMRF_E_UNSUPPORTED;
    }
    return fn((const float*)in, (float*)out, (float*)tmp, mask, when, ny, nx, pitch, thr, widx, negate, row_lo, row_hi, st);
}

}  // namespace smrf
