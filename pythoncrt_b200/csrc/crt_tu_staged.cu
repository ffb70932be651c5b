// crt_tu_staged.cu — translation unit of the staged kernels and the counter-based generators (crt_kernels.cuh)
#include "crt_kernels.cuh"
#include "crt_launch.h"

namespace crt {

// One frame through the staged kernels: bloom plane, pre-warp image, output.
int launch_staged(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, int has_prev, float* img,
                  const Scratch& s, cudaStream_t st, int* launches) {
    dim3 blk(32, 8);
    if (d.bloom_mode == 1) {
        dim3 grd((d.hw + 31) / 32, (d.hh + 7) / 8);
        k_bloom_down<<<grd, blk, 0, st>>>(d, in, s.ds); ++*launches;
    } else if (d.bloom_mode == 2) {
        const int r = d.ksize / 2;
        const size_t smem = ((size_t)(GAUSS_TH + 2 * r) * (GAUSS_TW + 2 * r) + (size_t)(GAUSS_TH + 2 * r) * GAUSS_TW) * 3 * sizeof(float);
        if (smem > 48 * 1024 && env.raise((const void*)k_bloom_gauss, (int)smem) &&
            cudaFuncSetAttribute(k_bloom_gauss, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
        dim3 grd((d.W + GAUSS_TW - 1) / GAUSS_TW, (d.H + GAUSS_TH - 1) / GAUSS_TH);
        k_bloom_gauss<<<grd, blk, smem, st>>>(d, in, s.bl); ++*launches;
    }
    dim3 grd((d.W + 31) / 32, (d.H + 7) / 8);
    if (d.warp_on) { k_pre_warp<<<grd, blk, 0, st>>>(d, f, in, s); ++*launches; }
    k_output<<<grd, blk, 0, st>>>(d, f, in, s, has_prev, state, out, img); ++*launches;
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

int launch_resize_state(const float* src, int sw, float* dst, int dw, int dh, const Lerp1* cx, const Lerp1* cy, cudaStream_t st) {
    k_resize_state<<<dim3((dw + 31) / 32, (dh + 7) / 8), dim3(32, 8), 0, st>>>(src, sw, dst, dw, dh, cx, cy);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

int launch_noise_gen(float* plane, int n_cells, uint64_t seed, uint64_t frame_index, cudaStream_t st) {
    k_noise_gen<<<((n_cells + 3) / 4 + 255) / 256, 256, 0, st>>>(plane, n_cells, seed, frame_index);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

int launch_glitch_gen(int32_t* offs, int rows, int nseg, int variant, float amp_px, uint64_t seed, uint64_t key, cudaStream_t st) {
    const int grid = variant == 0 ? (rows + GLITCH_THREADS - 1) / GLITCH_THREADS : (rows + GLITCH_ROWS - 1) / GLITCH_ROWS;
    k_glitch_gen<<<grid, GLITCH_THREADS, (size_t)rows * sizeof(float), st>>>(offs, rows, nseg, variant, amp_px, seed, key);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

}  // namespace crt
