// crt_tu_ps2.cu — translation unit of the pixel_size-2 block kernels (fast bloom / no bloom), crt_fused_ps2.cuh
#define CRT_TU_PS2
#include "crt_fused_ps2.cuh"

namespace crt {
int launch_fused_ps2(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, float* q_out, int has_prev,
                     cudaStream_t st, int* launches, bool pdl, const Ps2Maps* maps) {
    return run_fused_ps2(env, d, f, in, out, state, q_out, has_prev, st, launches, pdl, maps);
}
int clip_max_frames() { return CLIP_MAX_FRAMES; }
int launch_fused_ps2_clip(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, cudaStream_t st,
                          int* launches, const Ps2Maps* maps, int nf, const FrameVar* fv, int* sync) {
    if (nf < 1 || nf > CLIP_MAX_FRAMES) return 4;
    ClipArgs ca;
    ca.nf = nf; ca.sync = sync; ca.release = env_int("CRT_CLIP_RELEASE", 1);
    if (ca.release == 0) ca.release = 1; else if (ca.release < 0) ca.release = 0;      // 0 would be the racy mode: only as -1, for the race hunt in tests/_probe
    for (int i = 0; i < nf; ++i) ca.fv[i] = fv[i];
    return run_fused_ps2_clip(env, d, f, in, out, state, st, launches, maps, ca);
}
}  // namespace crt
