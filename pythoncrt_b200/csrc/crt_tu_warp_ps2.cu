// crt_tu_warp_ps2.cu — translation unit of the single-pass warp block kernels (crt_fused_warp_src.cuh: source-driven;
// crt_fused_warp_ps2.cuh: output-driven, opt-in)
#define CRT_TU_WARP_PS2
#include "crt_fused_warp_ps2.cuh"
#include "crt_fused_warp_src.cuh"

namespace crt {
int launch_warp_ps2(LaunchEnv& env, const WarpPs2Plan& pl, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state,
                    int has_prev, cudaStream_t st, int* launches, bool pdl) {
    return run_warp_ps2(env, pl, d, f, in, out, state, has_prev, st, launches, pdl);
}
int launch_warp_src(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, int has_prev, cudaStream_t st,
                    int* launches, bool pdl, const WsTile* d_tiles, int ntiles, int ntx, int nty, const Ps2Maps* maps) {
    return run_warp_src(env, d, f, in, out, state, has_prev, st, launches, pdl, d_tiles, ntiles, ntx, nty, maps);
}
}  // namespace crt
