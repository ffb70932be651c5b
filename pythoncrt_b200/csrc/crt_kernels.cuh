// crt_kernels.cuh — STAGED kernels: the general path of the chain.  Every
// parameter combination runs here through global-memory intermediates
// (half-size / blurred bloom plane, pre-warp image).  The fused tile kernel in
// crt_fused.cuh replaces it whenever the parameters allow (crt_abi.cu decides).
#pragma once
#include "crt_stages.cuh"

namespace crt {

// ---- fast bloom, pass 1: threshold + 2x down-scale into the half-size plane -----------
__global__ void __launch_bounds__(256) k_bloom_down(Dev d, const uint8_t* __restrict__ in, float* __restrict__ ds) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < d.hw && j < d.hh) bloom_down_cell(d, in, ds, j, i);
}

// ---- gaussian bloom: threshold + separable blur, tile staged through shared memory ----
constexpr int GAUSS_TW = 32, GAUSS_TH = 16;
__global__ void __launch_bounds__(256) k_bloom_gauss(Dev d, const uint8_t* __restrict__ in, float* __restrict__ bl) {
    extern __shared__ float smem[];
    const int K = d.ksize, r = K >> 1;
    const int SW = GAUSS_TW + 2 * r, SH = GAUSS_TH + 2 * r;
    float* S = smem;                       // [SH][SW][3] thresholded source, REPLICATE border
    float* R = smem + (size_t)SH * SW * 3; // [SH][TW][3] row-pass result
    const int x0 = blockIdx.x * GAUSS_TW, y0 = blockIdx.y * GAUSS_TH;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nth = blockDim.x * blockDim.y;
    for (int e = tid; e < SH * SW; e += nth) {
        int sy = e / SW, sx = e - sy * SW;
        int yy = imin(imax(y0 - r + sy, 0), d.H - 1), xx = imin(imax(x0 - r + sx, 0), d.W - 1);
        F3 v = bloom_src(d, graded_input(d, in, yy, xx));
        S[e * 3 + 0] = v.x; S[e * 3 + 1] = v.y; S[e * 3 + 2] = v.z;
    }
    __syncthreads();
    for (int e = tid; e < SH * GAUSS_TW * 3; e += nth) {
        int sy = e / (GAUSS_TW * 3), rem = e - sy * (GAUSS_TW * 3);
        R[e] = gauss_row(S + (size_t)sy * SW * 3 + rem, 3, d.taps, K);
    }
    __syncthreads();
    for (int e = tid; e < GAUSS_TH * GAUSS_TW * 3; e += nth) {
        int gy = e / (GAUSS_TW * 3), rem = e - gy * (GAUSS_TW * 3);
        int gx = rem / 3, y = y0 + gy, x = x0 + gx;
        if (y < d.H && x < d.W)
            bl[((size_t)y * d.W + x) * 3 + (rem - gx * 3)] = gauss_col(R + (size_t)(gy + r) * GAUSS_TW * 3 + rem, GAUSS_TW * 3, d.taps, K);
    }
}

// ---- stages 0-10 into the pre-warp image (only when the warp gather follows) ------------
__global__ void __launch_bounds__(256) k_pre_warp(Dev d, FrameDev f, const uint8_t* __restrict__ in, Scratch s) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= d.W || y >= d.H) return;
    F3 v = pre_warp_pixel(d, f, in, s, y, x, d.lut_fwd, d.lut_inv);
    float* q = s.q + ((size_t)y * d.W + x) * 3;
    q[0] = v.x; q[1] = v.y; q[2] = v.z;
}

// ---- output: glitch shift, warp gather, text-after, persistence, quantise ----------------
__global__ void __launch_bounds__(256) k_output(Dev d, FrameDev f, const uint8_t* __restrict__ in, Scratch s, int has_prev,
                                                float* __restrict__ state, uint8_t* __restrict__ out, float* __restrict__ img_out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= d.W || y >= d.H) return;
    F3 v = post_pixel(d, f, in, s, y, x, d.lut_fwd, d.lut_inv);
    if (img_out) {   // apply_static_effects: the float image before persistence (:861)
        float* o = img_out + ((size_t)y * d.W + x) * 3;
        o[0] = v.x; o[1] = v.y; o[2] = v.z;
        return;
    }
    finish_pixel(d, v, has_prev, state, out, y, x);
}

// ---- persistence state of another frame size -> this frame size (GUI: cv2.resize INTER_LINEAR, crt_filter.py:689-690) ----
// cv2's arithmetic (oracle/cv_restated.py resize_linear): coordinates in double, float32 weights, rows first, every lerp
// fma(q - p, w, p).  cx / cy are the per-axis coordinate tables (crt_derive.h linear_coords).
__global__ void __launch_bounds__(256) k_resize_state(const float* __restrict__ src, int sw, float* __restrict__ dst, int dw, int dh,
                                                      const Lerp1* __restrict__ cx, const Lerp1* __restrict__ cy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const Lerp1 ax = cx[x], ay = cy[y];
    const float* r0 = src + (size_t)ay.s0 * sw * 3;
    const float* r1 = src + (size_t)ay.s1 * sw * 3;
    float* o = dst + ((size_t)y * dw + x) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float a = lerp_cv(r0[ax.s0 * 3 + ch], r0[ax.s1 * 3 + ch], ax.w), b = lerp_cv(r1[ax.s0 * 3 + ch], r1[ax.s1 * 3 + ch], ax.w);
        o[ch] = lerp_cv(a, b, ay.w);
    }
}

// ---- counter-based generators ---------------------------------------------------------------
// N(0,1) plane [gh][gw] keyed (seed, frame_index, cell): four cells per Philox call.
__global__ void __launch_bounds__(256) k_noise_gen(float* __restrict__ plane, int n_cells, uint64_t seed, uint64_t frame_index) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g * 4 >= n_cells) return;
    U4 c; c.x = (uint32_t)g; c.y = 0x6e6f6973u /* "nois" */; c.z = (uint32_t)frame_index; c.w = (uint32_t)(frame_index >> 32);
    U4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    float n[4];
    box_muller(r.x, r.y, n[0], n[1]);
    box_muller(r.z, r.w, n[2], n[3]);
    for (int k = 0; k < 4; ++k) if (g * 4 + k < n_cells) plane[g * 4 + k] = n[k];
}

__device__ __forceinline__ float normal_at(uint64_t seed, uint64_t frame_index, uint32_t row, uint32_t col, uint32_t stream) {
    U4 c; c.x = row; c.y = col ^ (stream << 28) ^ 0x676c6974u /* "glit" */; c.z = (uint32_t)frame_index; c.w = (uint32_t)(frame_index >> 32);
    U4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    float a, b; box_muller(r.x, r.y, a, b);
    return a;
}
__device__ __forceinline__ float uniform_at(uint64_t seed, uint64_t frame_index, uint32_t row, uint32_t stream) {
    U4 c; c.x = row; c.y = (stream << 28) ^ 0x756e6966u; c.z = (uint32_t)frame_index; c.w = (uint32_t)(frame_index >> 32);
    U4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

// Glitch offsets with the distribution of the reference's generators
// (gui: crt_filter.py:672-679, export: :845-853), drawn from Philox instead of
// numpy's PCG64, keyed by the reference's own seed (crt_abi.cu glitch_key).
// gui: one offset per row, rows spread over the grid.  export: offset = rint(walk[row] + N(row, segment) * 0.7 amp) with
// walk = clip(cumsum(N(0,1)) * 0.1, +-0.4 amp) a random walk down the band.  Every CTA owns GLITCH_ROWS rows; it first
// rebuilds the walk's prefix for rows [0, its last row] (one N(0,1) per row, a block scan: at most ~2000 draws) and then
// fills its rows' segments.  Round 1 ran this on ONE CTA: 63 us per 4K frame, 241 us at 8K (ncu launch list, run 20) —
// a quarter of the full chain's frame; spread over the grid it is a few microseconds.
constexpr int GLITCH_THREADS = 256;
constexpr int GLITCH_ROWS = 8;
__global__ void __launch_bounds__(GLITCH_THREADS) k_glitch_gen(int32_t* __restrict__ offs, int rows, int nseg, int variant, float amp_px,
                                                               uint64_t seed, uint64_t key) {
    extern __shared__ float walk[];          // [rows up to this CTA's last row]
    __shared__ float warp_tot[GLITCH_THREADS / 32];
    const int t = threadIdx.x;
    const float frows = fmaxf(1.0f, (float)rows);
    if (variant == 0) {                      // gui: one offset per row
        const int rr = blockIdx.x * GLITCH_THREADS + t;
        if (rr < rows) {
            float amp = amp_px * expf(-3.0f * ((float)rr / frows));
            float base = clampf(0.5f * normal_at(seed, key, rr, 0, 1), -1.0f, 1.0f);
            if (uniform_at(seed, key, rr, 2) < 0.03f) base += uniform_at(seed, key, rr, 3) < 0.5f ? -1.0f : 1.0f;
            offs[rr] = (int)rintf(clampf(base * amp, -amp, amp));
        }
        return;
    }
    // export: prefix of the walk for rows [0, r1), each thread a contiguous run of rows
    const int r0 = blockIdx.x * GLITCH_ROWS, r1 = imin(rows, r0 + GLITCH_ROWS);
    const int per = (r1 + GLITCH_THREADS - 1) / GLITCH_THREADS;
    const int a = t * per, b = imin(r1, a + per);
    float local = 0.f;
    for (int rr = a; rr < b; ++rr) { local += normal_at(seed, key, rr, 0, 1); walk[rr] = local; }
    float incl = local;
    for (int o = 1; o < 32; o <<= 1) { float n = __shfl_up_sync(0xffffffffu, incl, o); if ((t & 31) >= o) incl += n; }
    if ((t & 31) == 31) warp_tot[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        float w = t < GLITCH_THREADS / 32 ? warp_tot[t] : 0.f, wi = w;
        for (int o = 1; o < 32; o <<= 1) { float n = __shfl_up_sync(0xffffffffu, wi, o); if (t >= o) wi += n; }
        if (t < GLITCH_THREADS / 32) warp_tot[t] = wi - w;       // exclusive prefix over warps
    }
    __syncthreads();
    const float base = warp_tot[t >> 5] + (incl - local);
    for (int rr = imax(a, r0); rr < b; ++rr) {                   // only this CTA's rows are needed below
        float amp = amp_px * (1.0f - (float)rr / frows);
        walk[rr] = clampf((walk[rr] + base) * 0.1f, -amp * 0.4f, amp * 0.4f);
    }
    __syncthreads();
    for (int e = t; e < (r1 - r0) * nseg; e += GLITCH_THREADS) {
        const int rr = r0 + e / nseg, sg = e - (e / nseg) * nseg;
        float amp = amp_px * (1.0f - (float)rr / frows);
        offs[(size_t)rr * nseg + sg] = (int)rintf(walk[rr] + normal_at(seed, key, rr, sg + 1, 4) * (amp * 0.7f));
    }
}

}  // namespace crt
