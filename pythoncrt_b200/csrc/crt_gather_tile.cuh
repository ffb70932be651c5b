// crt_gather_tile.cuh — second pass of the two-pass path with the persistence state moved by TMA.
//
// Same arithmetic as k_gather (crt_fused.cuh): glitch shift, cv2.remap 4-tap gather from the pre-warp image,
// text layer, persistence, quantise.  What changes is how the state crosses HBM: the tile's previous state
// (16 rows x 768 bytes) arrives by one tensor-map copy issued before the taps are computed, the blend reads it
// from shared memory and writes the new state back into the same tile, and the tile leaves with one TMA store —
// instead of three 16-byte loads and stores per thread at a 48-byte stride (see k_fused_ps2_pipe and
// tests/_probe/store_probe.cu).  Needs W % 4 == 0 and a 16-byte aligned state buffer; run_gather_any falls back
// to k_gather otherwise.
#pragma once
#include "crt_fused.cuh"
#include "crt_tma.cuh"

namespace crt {

#if defined(__CUDACC__)

constexpr int GT_BYTES = GATHER_TH * FTW * 3 * 4;            // state tile: 16 x 192 float32

template <bool WARP>
__global__ void __launch_bounds__(256) k_gather_tile(Dev d, FrameDev f, const float* __restrict__ qimg, uint8_t* __restrict__ out, int has_prev,
                                                     const __grid_constant__ CUtensorMap map_st) {
    __shared__ __align__(128) float s_st[GATHER_TH * FTW * 3];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * FTW, y0 = blockIdx.y * GATHER_TH;
    const int trow = tid / FROW_THREADS, y = y0 + trow, xb = x0 + (tid % FROW_THREADS) * 4;
    griddep_launch_dependents();        // the next frame's first pass may fill this kernel's last wave (it waits before it writes)
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        if (has_prev) { mbar_expect_tx(&bar, GT_BYTES); tma_load_2d(s_st, &map_st, x0 * 3, y0, &bar); }
    }
    __syncthreads();                    // barrier initialised before anyone waits on it
    const bool active = y < d.H && xb < d.W;
    if (active) {
        const float yn = WARP ? warp_norm((float)y, d.warp_cy, d.warp_dy) : 0.f;
        F3 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {                                // W % 4 == 0: whole quads
            const int gx = glitch_src_x(d, f, y, xb + k);
            if (WARP) {
                const Taps t = warp_taps_n(d, warp_norm((float)gx, d.warp_cx, d.warp_dx), yn);
                F3 a[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ty = t.iy + (j >> 1), tx = t.ix + (j & 1);
                    const bool ok = ty >= 0 && ty < d.H && tx >= 0 && tx < d.W;
                    a[j] = ok ? load_f3(qimg + ((size_t)ty * d.W + tx) * 3) : mk3(0.f, 0.f, 0.f);
                }
                v[k] = mk3(gather4_fast(a[0].x, a[1].x, a[2].x, a[3].x, t), gather4_fast(a[0].y, a[1].y, a[2].y, a[3].y, t),
                           gather4_fast(a[0].z, a[1].z, a[2].z, a[3].z, t));
            } else {
                v[k] = load_f3(qimg + ((size_t)y * d.W + gx) * 3);
            }
            if (d.text_mode == 2) v[k] = text_blend(d, v[k], y, gx);
        }
        float4* sp = reinterpret_cast<float4*>(s_st + (trow * FTW + (xb - x0)) * 3);
        float res[12];
        if (has_prev) {
            mbar_wait(&bar, 0);                                      // the tile's previous state has landed
            const float4 pa = sp[0], pb = sp[1], pc = sp[2];
            const float prev[12] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w, pc.x, pc.y, pc.z, pc.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                res[k * 3] = blend_fast(prev[k * 3], v[k].x, d.persist, d.persist_q);
                res[k * 3 + 1] = blend_fast(prev[k * 3 + 1], v[k].y, d.persist, d.persist_q);
                res[k * 3 + 2] = blend_fast(prev[k * 3 + 2], v[k].z, d.persist, d.persist_q);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) { res[k * 3] = v[k].x; res[k * 3 + 1] = v[k].y; res[k * 3 + 2] = v[k].z; }
        }
        sp[0] = make_float4(res[0], res[1], res[2], res[3]);
        sp[1] = make_float4(res[4], res[5], res[6], res[7]);
        sp[2] = make_float4(res[8], res[9], res[10], res[11]);
        if (out) {
            uint32_t* op = reinterpret_cast<uint32_t*>(out + ((size_t)y * d.W + xb) * 3);
            op[0] = pack4(res[0], res[1], res[2], res[3]);
            op[1] = pack4(res[4], res[5], res[6], res[7]);
            op[2] = pack4(res[8], res[9], res[10], res[11]);
        }
    }
    fence_proxy_async();                // the new state in shared memory -> visible to the TMA engine
    __syncthreads();
    if (tid == 0) {                     // one coalesced store per tile; rows / columns outside the frame are clipped
        tma_store_2d(&map_st, s_st, x0 * 3, y0);
        bulk_commit();
        bulk_wait_read();               // the tile has been read before the CTA (and its shared memory) goes away
    }
}

#if defined(CRT_TU_FUSED)
// map_st: float32 [H][W*3] tensor map of the state buffer with a 192 x 16 box, or null (plain k_gather)
inline int run_gather_any(const Dev& d, const FrameDev& f, const float* qimg, uint8_t* out, float* state, int has_prev, cudaStream_t st,
                          int* launches, const CUtensorMap* map_st) {
    if (!map_st || !state || (d.W & 3)) return run_gather(d, f, qimg, out, state, has_prev, st, launches);
    dim3 grid((d.W + FTW - 1) / FTW, (d.H + GATHER_TH - 1) / GATHER_TH);
    if (d.warp_on) k_gather_tile<true><<<grid, 256, 0, st>>>(d, f, qimg, out, has_prev, *map_st);
    else k_gather_tile<false><<<grid, 256, 0, st>>>(d, f, qimg, out, has_prev, *map_st);
    ++*launches;
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

#endif  // CRT_TU_FUSED

#endif  // __CUDACC__

}  // namespace crt
