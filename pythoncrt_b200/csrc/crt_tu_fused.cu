// crt_tu_fused.cu — translation unit of the general fused tile kernel and of the second pass of the two-pass
// path (crt_fused.cuh, crt_gather_tile.cuh)
#define CRT_TU_FUSED
#include "crt_gather_tile.cuh"
#include "crt_gather_box.cuh"

namespace crt {
int launch_fused(LaunchEnv& env, const FusedPlan& pl, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state,
                 float* q_out, int has_prev, cudaStream_t st, int* launches) {
    return run_fused(env, pl, d, f, in, out, state, q_out, has_prev, st, launches);
}
int launch_gather(LaunchEnv&, const Dev& d, const FrameDev& f, const float* qimg, uint8_t* out, float* state, int has_prev, cudaStream_t st,
                  int* launches, const CUtensorMap* map_st) {
    return run_gather_any(d, f, qimg, out, state, has_prev, st, launches, map_st);
}
int launch_gather_box(LaunchEnv& env, const Dev& d, const FrameDev& f, const float* qimg, uint8_t* out, int has_prev, cudaStream_t st,
                      int* launches, const int* origin, int bw, int bh, size_t smem, const CUtensorMap* map_q, const CUtensorMap* map_st) {
    return run_gather_box(env, d, f, qimg, out, has_prev, st, launches, origin, bw, bh, smem, map_q, map_st);
}
}  // namespace crt
