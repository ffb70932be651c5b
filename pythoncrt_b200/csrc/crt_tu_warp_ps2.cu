// crt_tu_warp_ps2.cu — translation unit of the single-pass warp block kernel, crt_fused_warp_ps2.cuh
#define CRT_TU_WARP_PS2
#include "crt_fused_warp_ps2.cuh"

namespace crt {
int launch_warp_ps2(LaunchEnv& env, const WarpPs2Plan& pl, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state,
                    int has_prev, cudaStream_t st, int* launches, bool pdl) {
    return run_warp_ps2(env, pl, d, f, in, out, state, has_prev, st, launches, pdl);
}
}  // namespace crt
