// crt_tu_gauss.cu — translation unit of the general packed-FP32 gaussian tile kernels, crt_fused_gauss.cuh
#define CRT_TU_GAUSS
#include "crt_fused_gauss.cuh"

namespace crt {
int launch_fused_gauss(LaunchEnv& env, int th, int nt, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state,
                       float* q_out, int has_prev, cudaStream_t st, int* launches) {
    return run_fused_gauss(env, th, nt, d, f, in, out, state, q_out, has_prev, st, launches);
}
}  // namespace crt
