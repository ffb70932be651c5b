// crt_gather_box.cuh — second pass of the two-pass path for warped chains: cv2.remap gather from a
// shared-memory copy of the tile's source footprint.
//
// k_gather / k_gather_tile read the four bilinear taps of every output pixel straight from the pre-warp image in
// global memory: twelve scalar loads per pixel, four bounds checks, 64-bit address arithmetic — 187 thread-instructions
// per pixel, 66 us per 4K frame (ncu, round 1), more than the whole first pass.  Here
//   * the host planner (plan_gather_box) evaluates, with the kernel's own float32 map, the source footprint of every
//     32 x 32 output tile, and uploads one box origin per tile; all boxes have the same (worst-case) size;
//   * the CTA's first instruction issues ONE tiled tensor-map copy (TMA) of that box of the pre-warp image
//     (float32 [H][W*3], box bw*3 x bh) and one of the tile's previous state; rows and columns outside the frame are
//     zero-filled by the copy engine, which is exactly cv2.remap's BORDER_CONSTANT rule (a tap outside contributes 0,
//     crt_filter.py:347), so the gather needs no bounds checks;
//   * while the copies are in flight every thread computes the taps of its four pixels (numpy-order float32 map, the
//     two divisions by per-clip constants through div_const: bit-identical, 3 instructions);
//   * taps are read with 32-bit shared-memory addressing; the new state goes back into the state tile and leaves with
//     one TMA store; only the packed uint8 pixels are stored by the threads.
// Glitch rows (horizontal shifts of up to a few glitch_amp, wrapping around the frame, crt_filter.py:852-857) move taps
// out of the box: those pixels (and every pixel of a clip whose footprint does not fit) take the global-memory path
// of k_gather, per pixel.  ~39 KB shared memory per CTA -> five 256-thread CTAs per SM.
// Measured and NOT adopted (round 2, run 30): launching this kernel with programmatic stream serialisation and computing the
// taps before griddepcontrol.wait (so that they overlap the first pass's last wave): 10 158 vs 10 596 frames/s on configs[2].
#pragma once
#include "crt_fused.cuh"
#include "crt_tma.cuh"

namespace crt {

constexpr int GB_TW = 32, GB_TH = 32;          // output tile; thread = 4 pixels of one row
constexpr int GB_NT = 256;
constexpr int GB_MAX_BW = 84;                  // box width limit: 252 floats (a tensor-map box dimension is at most 256 elements)
constexpr int GB_MAX_BH = 64;

struct GatherBoxPlan {
    bool ok = false;
    const char* why = "not planned";
    int bw = 0, bh = 0;                          // box size in pixels / rows (bw % 4 == 0: the float row pitch bw * 3 is a multiple of 16 bytes)
    std::vector<int> origin;                     // [tiles][2]: x0 (multiple of 4 pixels, may be negative), y0
    size_t smem = 0;
};

// Footprint boxes of all output tiles.  slack_px widens the boxes of tiles that intersect the glitch band (rows >= gy0).
inline GatherBoxPlan plan_gather_box(const Dev& d, int gy0, int slack_px) {
    GatherBoxPlan pl;
    if (!d.warp_on) { pl.why = "no warp"; return pl; }
    if ((d.W & 3) != 0) { pl.why = "frame width is not a multiple of 4"; return pl; }
    if ((size_t)d.W * d.H * 3 >= ((size_t)1 << 31)) { pl.why = "frame too large for 32-bit indexing"; return pl; }
    const int tiles_x = (d.W + GB_TW - 1) / GB_TW, tiles_y = (d.H + GB_TH - 1) / GB_TH;
    std::vector<float> xn(d.W), yn(d.H);
    for (int x = 0; x < d.W; ++x) xn[x] = warp_norm((float)x, d.warp_cx, d.warp_dx);
    for (int y = 0; y < d.H; ++y) yn[y] = warp_norm((float)y, d.warp_cy, d.warp_dy);
    pl.origin.resize((size_t)tiles_x * tiles_y * 2);
    int bw = 4, bh = 2;
    for (int ty = 0; ty < tiles_y; ++ty)
        for (int tx = 0; tx < tiles_x; ++tx) {
            const int ox0 = tx * GB_TW, oy0 = ty * GB_TH, ox1 = imin(ox0 + GB_TW, d.W) - 1, oy1 = imin(oy0 + GB_TH, d.H) - 1;
            int x0 = 0x7fffffff, y0 = 0x7fffffff, x1 = -0x7fffffff, y1 = -0x7fffffff;
            auto tap = [&](int x, int y) {
                const Taps t = warp_taps_n(d, xn[x], yn[y]);
                x0 = imin(x0, t.ix); x1 = imax(x1, t.ix + 1); y0 = imin(y0, t.iy); y1 = imax(y1, t.iy + 1);
            };
            if (d.warp_mono) {           // extremes on the perimeter
                for (int x = ox0; x <= ox1; ++x) { tap(x, oy0); tap(x, oy1); }
                for (int y = oy0; y <= oy1; ++y) { tap(ox0, y); tap(ox1, y); }
            } else {
                for (int y = oy0; y <= oy1; ++y) for (int x = ox0; x <= ox1; ++x) tap(x, y);
            }
            // taps far outside the frame read zeros wherever the box lies: clamp the box to one pixel around the frame
            x0 = imax(x0, -4); y0 = imax(y0, -1); x1 = imin(x1, d.W + 3); y1 = imin(y1, d.H);
            if (x1 < x0) { x0 = -4; x1 = -1; }
            if (y1 < y0) { y0 = -1; y1 = 0; }
            if (oy1 >= gy0) { x0 -= slack_px; x1 += slack_px; }
            x0 = (x0 >= 0 ? x0 : x0 - 3) / 4 * 4;              // floor to a multiple of 4 pixels (16-byte aligned float offset)
            pl.origin[((size_t)ty * tiles_x + tx) * 2] = x0;
            pl.origin[((size_t)ty * tiles_x + tx) * 2 + 1] = y0;
            bw = imax(bw, x1 - x0 + 1); bh = imax(bh, y1 - y0 + 1);
        }
    bw = (bw + 3) & ~3;
    if (bw > GB_MAX_BW || bh > GB_MAX_BH) { pl.why = "warp footprint larger than one tensor-map box"; return pl; }
    // the kernel divides by max(1, cx) / max(1, cy) through div_const: every operand it can meet is checked here against
    // the IEEE quotient the reference's numpy expression produces (crt_filter.py:340-341)
    const float rx = rcp_rn(d.warp_dx), ry = rcp_rn(d.warp_dy);
    for (int x = 0; x < d.W; ++x) if (div_const(fsub((float)x, d.warp_cx), d.warp_dx, rx) != xn[x]) { pl.why = "div_const differs from the IEEE quotient"; return pl; }
    for (int y = 0; y < d.H; ++y) if (div_const(fsub((float)y, d.warp_cy), d.warp_dy, ry) != yn[y]) { pl.why = "div_const differs from the IEEE quotient"; return pl; }
    pl.bw = bw; pl.bh = bh;
    pl.smem = (size_t)bw * 3 * bh * sizeof(float) + (size_t)GB_TW * 3 * GB_TH * sizeof(float) + 128;
    pl.ok = true; pl.why = "";
    return pl;
}

#if defined(__CUDACC__)

template <bool GLITCH>
__global__ void __launch_bounds__(GB_NT, 5) k_gather_box(Dev d, FrameDev f, const float* __restrict__ qimg, uint8_t* __restrict__ out, int has_prev,
                                                         const int2* __restrict__ origin, int bw, int bh, float rcp_dx, float rcp_dy,
                                                         const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_st) {
    extern __shared__ __align__(128) unsigned char gsm[];
    float* const s_st = reinterpret_cast<float*>(gsm);                           // [32][96]: state tile
    float* const s_q = s_st + GB_TH * GB_TW * 3;                                 // [bh][bw * 3]: footprint of the pre-warp image
    __shared__ __align__(8) uint64_t bar_q, bar_st;
    const int tid = threadIdx.x;
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const int ox0 = blockIdx.x * GB_TW, oy0 = blockIdx.y * GB_TH;
    const int trow = tid >> 3, y = oy0 + trow, xb = ox0 + (tid & 7) * 4;
    griddep_launch_dependents();        // the next frame's first pass may start its state-independent phases
    const int2 org = origin[tile];
    const int pitch = bw * 3;
    if (tid == 0) {
        mbar_init(&bar_q, 1); mbar_init(&bar_st, 1);
        fence_mbar_init();
    }
    griddep_wait();                     // the first pass of this frame (previous kernel in the stream) has written the pre-warp image
    if (tid == 0) {
        mbar_expect_tx(&bar_q, (uint32_t)(pitch * bh * 4));
        tma_load_2d_hint(s_q, &map_q, org.x * 3, org.y, &bar_q, L2_EVICT_NORMAL);       // neighbouring tiles' boxes overlap
        if (has_prev) { mbar_expect_tx(&bar_st, GB_TH * GB_TW * 3 * 4); tma_load_2d_hint(s_st, &map_st, ox0 * 3, oy0, &bar_st, L2_EVICT_FIRST); }
    }
    __syncthreads();                    // barriers initialised before anyone waits on them
    const bool active = y < d.H && xb < d.W;
    float res[12];
    if (active) {
        // taps of the four pixels while the copies are in flight: apply_barrel_warp's float32 map in numpy's order
        // (crt_math.cuh warp_taps), the divisions by max(1, cx) / max(1, cy) through div_const
        const float yn = div_const(fsub((float)y, d.warp_cy), d.warp_dy, rcp_dy);
        Taps tp[4];
        int gxs[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            gxs[k] = GLITCH ? glitch_src_x(d, f, y, xb + k) : xb + k;
            tp[k] = warp_taps_n(d, div_const(fsub((float)gxs[k], d.warp_cx), d.warp_dx, rcp_dx), yn);
        }
        mbar_wait(&bar_q, 0);                                            // the footprint has landed
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const Taps& t = tp[k];
            const int lx = t.ix - org.x, ly = t.iy - org.y;
            F3 a[4];
            if ((unsigned)lx < (unsigned)(bw - 1) && (unsigned)ly < (unsigned)(bh - 1)) {      // both columns and both rows inside the box
                const float* base = s_q + ly * pitch + lx * 3;
                a[0] = load_f3(base); a[1] = load_f3(base + 3); a[2] = load_f3(base + pitch); a[3] = load_f3(base + pitch + 3);
            } else {                                                         // glitch shift / far outside the frame: global memory, per tap
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ty = t.iy + (j >> 1), tx = t.ix + (j & 1);
                    const bool ok = ty >= 0 && ty < d.H && tx >= 0 && tx < d.W;
                    a[j] = ok ? load_f3(qimg + ((size_t)ty * d.W + tx) * 3) : mk3(0.f, 0.f, 0.f);
                }
            }
            F3 v = mk3(gather4_fast(a[0].x, a[1].x, a[2].x, a[3].x, t), gather4_fast(a[0].y, a[1].y, a[2].y, a[3].y, t),
                       gather4_fast(a[0].z, a[1].z, a[2].z, a[3].z, t));
            if (d.text_mode == 2) v = text_blend(d, v, y, gxs[k]);
            res[k * 3] = v.x; res[k * 3 + 1] = v.y; res[k * 3 + 2] = v.z;
        }
        float4* sp = reinterpret_cast<float4*>(s_st + (trow * GB_TW + (xb - ox0)) * 3);
        if (has_prev) {
            mbar_wait(&bar_st, 0);                                       // the tile's previous state has landed
            const float4 pa = sp[0], pb = sp[1], pc = sp[2];
            const float prev[12] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w, pc.x, pc.y, pc.z, pc.w};
#pragma unroll
            for (int k = 0; k < 12; ++k) res[k] = blend_fast(prev[k], res[k], d.persist, d.persist_q);
        }
        sp[0] = make_float4(res[0], res[1], res[2], res[3]);
        sp[1] = make_float4(res[4], res[5], res[6], res[7]);
        sp[2] = make_float4(res[8], res[9], res[10], res[11]);
        if (out) {
            uint32_t* op = reinterpret_cast<uint32_t*>(out + ((size_t)y * d.W + xb) * 3);
            __stcs(op, pack4(res[0], res[1], res[2], res[3]));                   // streaming: written once, never read by this library
            __stcs(op + 1, pack4(res[4], res[5], res[6], res[7]));
            __stcs(op + 2, pack4(res[8], res[9], res[10], res[11]));
        }
    }
    fence_proxy_async();                // the new state in shared memory -> visible to the TMA engine
    __syncthreads();
    if (tid == 0) {                     // one coalesced store per tile; rows / columns outside the frame are clipped
        tma_store_2d_hint(&map_st, s_st, ox0 * 3, oy0, L2_EVICT_FIRST);          // read again only a frame later
        bulk_commit();
        bulk_wait_read();               // the tile has been read before the CTA (and its shared memory) goes away
    }
}

#if defined(CRT_TU_FUSED)
// origin: device copy of GatherBoxPlan::origin; map_q / map_st: tensor maps of the pre-warp image (box bw*3 x bh) and of the
// state buffer (box 96 x 32), both float32 [H][W*3]
inline int run_gather_box(LaunchEnv& env, const Dev& d, const FrameDev& f, const float* qimg, uint8_t* out, int has_prev, cudaStream_t st,
                          int* launches, const int* origin, int bw, int bh, size_t smem, const CUtensorMap* map_q, const CUtensorMap* map_st) {
    auto kern = f.goffs ? k_gather_box<true> : k_gather_box<false>;
    if (env.raise((const void*)kern, (int)smem) &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
    const dim3 grid((d.W + GB_TW - 1) / GB_TW, (d.H + GB_TH - 1) / GB_TH);
    kern<<<grid, GB_NT, smem, st>>>(d, f, qimg, out, has_prev, reinterpret_cast<const int2*>(origin), bw, bh, rcp_rn(d.warp_dx), rcp_rn(d.warp_dy),
                                    *map_q, *map_st);
    ++*launches;
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}
#endif  // CRT_TU_FUSED

#endif  // __CUDACC__

}  // namespace crt
