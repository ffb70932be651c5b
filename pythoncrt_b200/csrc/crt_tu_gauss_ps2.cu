// crt_tu_gauss_ps2.cu — translation unit of the block-resolution gaussian-bloom kernels, crt_fused_gauss_ps2.cuh
#define CRT_TU_GAUSS_PS2
#include "crt_fused_gauss_ps2.cuh"

namespace crt {
int launch_fused_gauss_ps2(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, float* q_out,
                           int has_prev, cudaStream_t st, int* launches, bool pdl, const Ps2Maps* maps, const CUtensorMap* gmap_in) {
    return run_fused_gauss_ps2(env, d, f, in, out, state, q_out, has_prev, st, launches, pdl, maps, gmap_in);
}
}  // namespace crt
