// crt_launch.h — host-callable launchers of the kernel families.  Each family is compiled in its own
// translation unit (crt_tu_*.cu) so that the library builds in parallel; crt_abi.cu sees only these
// prototypes and never instantiates a kernel template itself.
//
// LaunchEnv carries what a launcher needs to remember between calls — the SM count of the context's
// device, which kernels already had their opt-in shared-memory size set, how many CTAs of a persistent
// kernel the device holds.  It lives in the crt_ctx (one context per device and host thread), so the
// launchers keep no process-wide mutable state: the reference's two-worker export pool
// (crt_filter.py:1015-1017), one context per worker, is safe, and so are GPUs of different sizes.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>

#include "crt_math.cuh"

namespace crt {

struct LaunchEnv {
    int device = 0;
    int sms = 148;
    std::map<const void*, int> memo;      // kernel entry point -> remembered value (configured size / resident CTAs)
    // true the first time `value` exceeds what was remembered for `key` in this context
    bool raise(const void* key, int value) {
        auto it = memo.find(key);
        if (it != memo.end() && it->second >= value) return false;
        memo[key] = value;
        return true;
    }
};

struct FusedPlan;
struct WarpPs2Plan;
struct WsTile;
struct Ps2Maps;
struct Scratch;

int launch_fused_ps2(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, float* q_out, int has_prev,
                     cudaStream_t st, int* launches, bool pdl, const Ps2Maps* maps);
// clip mode: frames [maps->frame, maps->frame + nf) in one launch (nf <= clip_max_frames()); in / out = the first of them;
// fv = their scalars (host); sync = [1 + tiles] ints, zeroed on st before the launch
int clip_max_frames();
int launch_fused_ps2_clip(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, cudaStream_t st,
                          int* launches, const Ps2Maps* maps, int nf, const FrameVar* fv, int* sync);
bool fused_gauss_ps2_clip_supported(const Dev& d, int K);
int launch_fused_gauss_ps2_clip(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, cudaStream_t st,
                                int* launches, const Ps2Maps* maps, const CUtensorMap* gmap_in, int nf, const FrameVar* fv, int* sync);
int launch_warp_ps2(LaunchEnv& env, const WarpPs2Plan& pl, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state,
                    int has_prev, cudaStream_t st, int* launches, bool pdl);
int launch_warp_src(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, int has_prev, cudaStream_t st,
                    int* launches, bool pdl, const WsTile* d_tiles, int ntiles, int ntx, int nty, const Ps2Maps* maps);
int launch_fused_gauss_ps2(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, float* q_out,
                           int has_prev, cudaStream_t st, int* launches, bool pdl, const Ps2Maps* maps, const CUtensorMap* gmap_in);
int launch_fused_gauss(LaunchEnv& env, int th, int nt, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state,
                       float* q_out, int has_prev, cudaStream_t st, int* launches);
int launch_fused(LaunchEnv& env, const FusedPlan& pl, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state,
                 float* q_out, int has_prev, cudaStream_t st, int* launches);
int launch_gather(LaunchEnv& env, const Dev& d, const FrameDev& f, const float* qimg, uint8_t* out, float* state, int has_prev,
                  cudaStream_t st, int* launches, const CUtensorMap* map_st);
int launch_gather_box(LaunchEnv& env, const Dev& d, const FrameDev& f, const float* qimg, uint8_t* out, int has_prev, cudaStream_t st,
                      int* launches, const int* origin, int bw, int bh, size_t smem, const CUtensorMap* map_q, const CUtensorMap* map_st);
// staged kernels: bloom plane(s), pre-warp image, output; *dominant_first = launches before the output kernel
int launch_staged(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, int has_prev, float* img,
                  const Scratch& s, cudaStream_t st, int* launches);
int launch_resize_state(const float* src, int sw, float* dst, int dw, int dh, const Lerp1* cx, const Lerp1* cy, cudaStream_t st);
int launch_noise_gen(float* plane, int n_cells, uint64_t seed, uint64_t frame_index, cudaStream_t st);
int launch_glitch_gen(int32_t* offs, int rows, int nseg, int variant, float amp_px, uint64_t seed, uint64_t key, cudaStream_t st);

}  // namespace crt
