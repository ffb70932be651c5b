// crt_fused_warp_ps2.cuh — single-pass barrel warp for the pixel_size-2 chains (BASELINE.json configs[2]).
//
// apply_barrel_warp (crt_filter.py:331-348, called at :649) samples the PROCESSED image — after bloom, triad, scanlines,
// vignette, flicker — so a warped output tile needs stages 0-10 over the tile's source footprint.  The two-pass path
// materialises that image for the whole frame in HBM (12 B/px written by the block kernel, read back by k_gather:
// ~54 B/px moved, two kernels, 125 us per 4K frame); the general single-pass kernel (k_fused<.., WARP>) keeps the
// footprint in shared memory but evaluates every stage per pixel (207 us).  This kernel combines the two ideas:
//   phase 0  cv2.remap taps of the tile's perimeter (the map is monotone, crt_derive.h warp_mono) -> footprint box Q,
//            clipped to the frame and aligned to 4 x 2 pixel patches
//   phase 1  ONE graded value per 2x2 block of Q (+ one halo block): the block kernels' arithmetic (crt_fused_ps2.cuh)
//   phase 2  stages 5-10 over Q, one 4 x 2 patch per thread and step, through ps2_patch_tail (bloom from a 4 x 3 block
//            neighbourhood in registers, composite triad LUT, row / column mask tables) into a float32 tile in shared memory
//   phase 3  per output pixel: cv2.remap's 1/32-pixel bilinear taps gathered from that tile (taps outside the frame
//            contribute 0), persistence blend against the state in HBM, packed uint8 store.
// Every byte of frame, state and output crosses HBM once (30 B/px); the pre-warp image never leaves the SM.
// 512 threads per CTA, output tile 64 x 32, two CTAs per SM; shared memory is sized by the host planner
// (plan_warp_ps2) from the exact worst-case footprint over all tiles, evaluated with the kernel's own float32 map.
// Falls back to the two-pass path when the footprint does not fit (strong warps), for pincushion maps that are not
// monotone, and for glitch (whose row shifts wrap around the frame) or a text layer after the warp.
#pragma once
#include "crt_fused_ps2.cuh"

namespace crt {

constexpr int WP_NT = 512;
constexpr int WP_TW = 64, WP_TH = 32;               // output tile
constexpr int WP_MAX_QW = 128, WP_MAX_QH = 64;      // static bounds of the per-tile mask tables

struct WarpPs2Plan {
    bool ok = false;
    const char* why = "not planned";
    int qw = 0, qh = 0;          // worst-case aligned footprint (pixels)
    int tpitch = 0;              // floats per footprint row in the shared-memory tile
    int bpitch = 0, bh = 0;      // block arrays: pitch (floats) and rows
    size_t smem = 0;
};

CRT_HD bool warp_ps2_supported(const Dev& d, bool glitch_on) {
    return d.warp_on && d.warp_mono && d.pix_uniform == 2 && d.even_dims && (d.W & 3) == 0 && !glitch_on && d.text_mode == 0 && d.bloom_mode != 2;
}

// footprint box of one output tile, aligned to whole 4 x 2 patches (W % 4 == 0, H even: stays inside the frame)
CRT_HD Box warp_ps2_align(const Box& q) {
    Box a;
    a.x0 = q.x0 & ~3; a.x1 = q.x1 | 3; a.y0 = q.y0 & ~1; a.y1 = q.y1 | 1;
    return a;
}

inline WarpPs2Plan plan_warp_ps2(const Dev& d, bool glitch_on) {
    WarpPs2Plan pl;
    if (!warp_ps2_supported(d, glitch_on)) { pl.why = "not a pixel_size-2 chain with a monotone warp"; return pl; }
    if ((size_t)d.W * d.H * 3 >= ((size_t)1 << 31)) { pl.why = "frame too large for 32-bit indexing"; return pl; }
    std::vector<float> xn(d.W), yn(d.H);
    for (int x = 0; x < d.W; ++x) xn[x] = warp_norm((float)x, d.warp_cx, d.warp_dx);
    for (int y = 0; y < d.H; ++y) yn[y] = warp_norm((float)y, d.warp_cy, d.warp_dy);
    int qw = 0, qh = 0;
    for (int oy0 = 0; oy0 < d.H; oy0 += WP_TH)
        for (int ox0 = 0; ox0 < d.W; ox0 += WP_TW) {
            const int ox1 = imin(ox0 + WP_TW, d.W) - 1, oy1 = imin(oy0 + WP_TH, d.H) - 1;
            int x0 = 0x7fffffff, y0 = 0x7fffffff, x1 = -0x7fffffff, y1 = -0x7fffffff;
            auto tap = [&](int x, int y) {
                const Taps t = warp_taps_n(d, xn[x], yn[y]);
                x0 = imin(x0, t.ix); x1 = imax(x1, t.ix + 1); y0 = imin(y0, t.iy); y1 = imax(y1, t.iy + 1);
            };
            for (int x = ox0; x <= ox1; ++x) { tap(x, oy0); tap(x, oy1); }
            for (int y = oy0; y <= oy1; ++y) { tap(ox0, y); tap(ox1, y); }
            bool empty = false;
            const Box q = footprint_box(d, x0, y0, x1, y1, &empty);
            if (empty) continue;
            const Box a = warp_ps2_align(q);
            qw = imax(qw, box_w(a)); qh = imax(qh, box_h(a));
        }
    if (qw == 0) { qw = 4; qh = 2; }
    if (qw > WP_MAX_QW || qh > WP_MAX_QH) { pl.why = "warp footprint larger than the tile tables"; return pl; }
    pl.qw = qw; pl.qh = qh;
    pl.tpitch = qw * 3 + 4;                          // +4 floats: rows start 16 bytes apart modulo 128 (fewer bank conflicts on the gather)
    pl.bh = qh / 2 + 2;
    pl.bpitch = (qw / 2 + 2 + 3) & ~1;               // even, >= blocks + 2
    const int nblk = d.bloom_mode == 1 && d.thr_on ? 2 : 1;
    pl.smem = sizeof(float) * ((size_t)pl.tpitch * qh + (size_t)nblk * 3 * pl.bh * pl.bpitch);
    if (pl.smem + 14 * 1024 > 113 * 1024) { pl.why = "warp footprint does not leave room for two CTAs per SM"; return pl; }
    pl.ok = true; pl.why = "";
    return pl;
}

#if defined(__CUDACC__)

struct WarpPs2Geo { int tpitch, bpitch, bh, cap_qw, cap_qh; };

template <bool BLOOM, bool FAST, bool THR>
__global__ void __launch_bounds__(WP_NT, 2) k_warp_ps2(Dev d, FrameDev f, const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                       float* __restrict__ state, int has_prev, WarpPs2Geo g) {
    extern __shared__ __align__(16) float wsm[];
    __shared__ __align__(16) float s_lut[2 * 1028];
    __shared__ __align__(16) float s_sel[3][12];
    float* const s_fwd = s_lut;
    float* const s_inv = s_lut + 1028;
    __shared__ float s_unit[256];
    __shared__ __align__(16) float s_pow[POW_TAB_FLOATS];
    __shared__ float s_rows[2 * WP_MAX_QH], s_cols[2 * WP_MAX_QW];
    __shared__ int s_box[4], s_geo[12];
    float* const T = wsm;                                        // [qh][tpitch]: pre-warp image over the footprint, interleaved RGB
    float* const Ub = T + g.tpitch * g.cap_qh;                   // [3][bh][bpitch]: graded block values
    float* const Sb = Ub + 3 * g.bh * g.bpitch;                  // [3][bh][bpitch]: thresholded bloom source (THR only)
    const int tid = threadIdx.x, lane = tid & 31;
    const int ox0 = blockIdx.x * WP_TW, oy0 = blockIdx.y * WP_TH;
    const int ox1 = imin(ox0 + WP_TW, d.W) - 1, oy1 = imin(oy0 + WP_TH, d.H) - 1;
    griddep_launch_dependents();        // the next frame's kernel may begin its state-independent phases (see launch_pdl)

    const float* lut_a = d.triad_comp ? d.triad_comp : d.lut_fwd;
    const float* lut_b = d.triad_comp ? d.triad_comp + 1028 : d.lut_inv;
    if (d.triad_mode >= 2 && tid < 256) {
        reinterpret_cast<float4*>(s_fwd)[tid] = reinterpret_cast<const float4*>(lut_a)[tid];
        reinterpret_cast<float4*>(s_inv)[tid] = reinterpret_cast<const float4*>(lut_b)[tid];
        if (tid == 0) { s_fwd[1024] = lut_a[1024]; s_inv[1024] = lut_b[1024]; }
    }
    if (tid < 256) s_unit[tid] = __fdiv_rn((float)tid, 255.0f);
    ps2_fill_sel(s_sel, tid, d.bgr);
    if (d.col_gamma) for (int i = tid; i < POW_TAB_FLOATS; i += WP_NT) s_pow[i] = d.pow_tab[i];
    if (tid < 4) s_box[tid] = (tid < 2) ? 0x7fffffff : -0x7fffffff;
    __syncthreads();

    // ---- phase 0: footprint of the tile in the pre-warp image (monotone map: extremes on the perimeter) ----
    {
        int bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -0x7fffffff, by1 = -0x7fffffff;
        const int tw = ox1 - ox0 + 1, thh = oy1 - oy0 + 1, nper = 2 * tw + 2 * thh;
        if (tid < nper) {
            int x, y;
            if (tid < 2 * tw) { x = ox0 + (tid < tw ? tid : tid - tw); y = tid < tw ? oy0 : oy1; }
            else { const int j = tid - 2 * tw; y = oy0 + (j < thh ? j : j - thh); x = j < thh ? ox0 : ox1; }
            const Taps t = warp_taps_n(d, warp_norm((float)x, d.warp_cx, d.warp_dx), warp_norm((float)y, d.warp_cy, d.warp_dy));
            bx0 = t.ix; bx1 = t.ix + 1; by0 = t.iy; by1 = t.iy + 1;
        }
        if (tid < ((nper + 31) & ~31)) {          // warps that hold perimeter samples
            bx0 = __reduce_min_sync(0xffffffffu, bx0); by0 = __reduce_min_sync(0xffffffffu, by0);
            bx1 = __reduce_max_sync(0xffffffffu, bx1); by1 = __reduce_max_sync(0xffffffffu, by1);
            if (lane == 0) { atomicMin(&s_box[0], bx0); atomicMin(&s_box[1], by0); atomicMax(&s_box[2], bx1); atomicMax(&s_box[3], by1); }
        }
    }
    __syncthreads();
    if (tid == 0) {
        bool empty = false;
        const Box q0 = footprint_box(d, s_box[0], s_box[1], s_box[2], s_box[3], &empty);
        const Box a = warp_ps2_align(q0);
        const int qw = empty ? 0 : box_w(a), qh = empty ? 0 : box_h(a);
        s_geo[0] = a.x0; s_geo[1] = a.y0; s_geo[2] = qw; s_geo[3] = qh; s_geo[4] = empty;
        s_geo[5] = (int)make_magic(qw / 2 + 2);           // / blocks per row
        s_geo[6] = (int)make_magic(qw / 4);               // / patches per row
    }
    __syncthreads();
    const int qx0 = s_geo[0], qy0 = s_geo[1], qw = s_geo[2], qh = s_geo[3];
    const bool q_empty = s_geo[4] != 0;
    // planning guarantees the footprint fits; a violated bound must never corrupt memory
    if (qw > g.cap_qw || qh > g.cap_qh) {
        if (tid == 0) printf("crt_b200: warp footprint %dx%d exceeds capacity %dx%d\n", qw, qh, g.cap_qw, g.cap_qh);
        return;
    }
    MaskTabs mt{s_rows, s_cols, s_rows + WP_MAX_QH, s_cols + WP_MAX_QW};
    const int BP = g.bpitch, BH = g.bh;

    if (!q_empty) {
        // per-footprint mask tables (same expressions as the block kernels' per-tile tables)
        if (tid < qh) {
            const int y = qy0 + tid;
            if (d.scan_mode == 1) mt.row_scan[tid] = scan_row(d, f, y);
            else if (d.scan_mode == 2) { double t = ((double)y + f.phase) * d.scan_inv_period; mt.row_scan[tid] = (float)(t - floor(t)); }
            if (d.vig_mode == 1) { const float ny = ((float)y - d.vig_cy) * d.vig_iry; mt.row_vig[tid] = ny * ny; }
        } else if (tid >= 128 && tid < 128 + qw) {
            const int c = tid - 128, x = qx0 + c;
            if (d.scan_mode == 2) { double t = (d.scan_tan * (double)x) * d.scan_inv_period; mt.col_scan[c] = (float)(t - floor(t)); }
            if (d.vig_mode == 1) { const float nx = ((float)x - d.vig_cx) * d.vig_irx; mt.col_vig[c] = nx * nx; }
        }
        // ---- phase 1: one graded value per 2x2 block of Q + one halo block (clamped = cv2's edge rule) ----
        const int qbw = qw / 2 + 2, qbh = qh / 2 + 2;
        const int gbx0 = (qx0 >> 1) - 1, gby0 = (qy0 >> 1) - 1;
        const unsigned magic_b = (unsigned)s_geo[5];
        const int a0 = d.aberr != 0 ? d.aberr_mod : 0;
        for (int u = tid; u < qbw * qbh; u += WP_NT) {
            const int bj = fastdiv(u, magic_b), bi = u - bj * qbw;
            const int sx = 2 * imin(imax(gbx0 + bi, 0), d.hw - 1), sy = 2 * imin(imax(gby0 + bj, 0), d.hh - 1);
            const uint8_t* row = in + (size_t)sy * d.W * 3;
            const uint32_t r0 = row[wrap(sx - a0, d.W) * 3 + 0], r1 = row[sx * 3 + 1], r2 = row[wrap(sx + a0, d.W) * 3 + 2];
            const F3 v1 = colour(d, mk3(s_unit[r0], s_unit[r1], s_unit[r2]), s_pow);
            float* ub = Ub + bj * BP + bi;
            ub[0] = v1.x; ub[BH * BP] = v1.y; ub[2 * BH * BP] = v1.z;
            if (BLOOM && THR) {
                const F3 sv = bloom_src(d, v1);
                float* sb = Sb + bj * BP + bi;
                sb[0] = sv.x; sb[BH * BP] = sv.y; sb[2 * BH * BP] = sv.z;
            }
        }
    }
    __syncthreads();

    // ---- phase 2: stages 5-10 over Q, one 4 x 2 patch per thread and step, into the shared-memory tile ----
    if (!q_empty) {
        const int qpw = qw >> 2, npatch = qpw * (qh >> 1);
        const unsigned magic_p = (unsigned)s_geo[6];
        for (int u = tid; u < npatch; u += WP_NT) {
            const int pj = fastdiv(u, magic_p), pi = u - pj * qpw;
            const int xb = qx0 + 4 * pi, y0 = qy0 + 2 * pj;
            const int bi = 2 * pi + 1, bj = pj + 1;                  // first of the patch's two blocks, in halo coordinates
            float blr[4][3], t1[2][3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) { t1[0][ch] = Ub[(ch * BH + bj) * BP + bi]; t1[1][ch] = Ub[(ch * BH + bj) * BP + bi + 1]; }
            auto row_begin = [&](int r) {        // cv2's 2x up-scale of the bloom for patch row r (see k_fused_ps2)
#pragma unroll
                for (int ch = 0; ch < (BLOOM ? 3 : 0); ++ch) {
                    const float* src = (THR ? Sb : Ub) + ch * BH * BP;
                    float h[2][4];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float2 ca = *reinterpret_cast<const float2*>(src + (bj - 1 + r + q) * BP + bi - 1);
                        const float2 cb = *reinterpret_cast<const float2*>(src + (bj - 1 + r + q) * BP + bi + 1);
                        const float d01 = fsub(ca.y, ca.x), d12 = fsub(cb.x, ca.y), d23 = fsub(cb.y, cb.x);
                        h[q][0] = ffma(d01, 0.75f, ca.x); h[q][1] = ffma(d12, 0.25f, ca.y);
                        h[q][2] = ffma(d12, 0.75f, ca.y); h[q][3] = ffma(d23, 0.25f, cb.x);
                    }
                    const float w = r == 0 ? 0.75f : 0.25f;
#pragma unroll
                    for (int k = 0; k < 4; ++k) blr[k][ch] = ffma(fsub(h[1][k], h[0][k]), w, h[0][k]);
                }
            };
            // q_out != nullptr + state_in_smem: the patch's float values go to the tile (no blend, no uint8 output)
            ps2_patch_tail<BLOOM, FAST>(d, f, mt, s_fwd, s_inv, s_sel, nullptr, nullptr, T /* non-null marker */, 0, qx0, qy0, qx0 + qw - 1, qy0 + qh - 1,
                                        xb, y0, t1, [&](int, int k) { return mk3(blr[k][0], blr[k][1], blr[k][2]); },
                                        T + (2 * pj) * g.tpitch + 12 * pi, true, row_begin, g.tpitch);
        }
    }
    __syncthreads();
    griddep_wait();         // previous kernel of the stream complete: the state may be touched from here on

    // ---- phase 3: cv2.remap gather from the tile, persistence, quantise; 4 pixels of one row per thread ----
    const int y = oy0 + (tid >> 4), xb = ox0 + 4 * (tid & 15);
    if (y > oy1 || xb > ox1) return;
    const float yn = warp_norm((float)y, d.warp_cy, d.warp_dy);
    const int o = (y * d.W + xb) * 3;
    float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa, pc = pa;
    if (has_prev) {
        const float4* sp = reinterpret_cast<const float4*>(state + o);
        pa = sp[0]; pb = sp[1]; pc = sp[2];
    }
    const float prev[12] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w, pc.x, pc.y, pc.z, pc.w};
    float res[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        F3 v = mk3(0.f, 0.f, 0.f);
        if (!q_empty) {
            const Taps t = warp_taps_n(d, warp_norm((float)(xb + k), d.warp_cx, d.warp_dx), yn);
            const int lx = t.ix - qx0, ly = t.iy - qy0;
            const float* base = T + ly * g.tpitch + lx * 3;
            F3 a[4];
            if (lx >= 0 && lx + 1 < qw && ly >= 0 && ly + 1 < qh) {         // all four taps inside Q (hence inside the frame)
                a[0] = load_f3(base); a[1] = load_f3(base + 3); a[2] = load_f3(base + g.tpitch); a[3] = load_f3(base + g.tpitch + 3);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ty = t.iy + (j >> 1), tx = t.ix + (j & 1);
                    const bool ok = ty >= 0 && ty < d.H && tx >= 0 && tx < d.W;          // inside the frame == inside Q (Q is the clipped box)
                    a[j] = ok ? load_f3(T + (ty - qy0) * g.tpitch + (tx - qx0) * 3) : mk3(0.f, 0.f, 0.f);
                }
            }
            v = mk3(gather4_fast(a[0].x, a[1].x, a[2].x, a[3].x, t), gather4_fast(a[0].y, a[1].y, a[2].y, a[3].y, t),
                    gather4_fast(a[0].z, a[1].z, a[2].z, a[3].z, t));
        }
        if (has_prev) {
            v.x = blend_fast(prev[k * 3], v.x, d.persist, d.persist_q);
            v.y = blend_fast(prev[k * 3 + 1], v.y, d.persist, d.persist_q);
            v.z = blend_fast(prev[k * 3 + 2], v.z, d.persist, d.persist_q);
        }
        res[k * 3] = v.x; res[k * 3 + 1] = v.y; res[k * 3 + 2] = v.z;
    }
    if (state) {
        float4* sp = reinterpret_cast<float4*>(state + o);
        sp[0] = make_float4(res[0], res[1], res[2], res[3]);
        sp[1] = make_float4(res[4], res[5], res[6], res[7]);
        sp[2] = make_float4(res[8], res[9], res[10], res[11]);
    }
    if (out) {
        uint32_t* op = reinterpret_cast<uint32_t*>(out + o);
        op[0] = pack4(res[0], res[1], res[2], res[3]);
        op[1] = pack4(res[4], res[5], res[6], res[7]);
        op[2] = pack4(res[8], res[9], res[10], res[11]);
    }
}

#if defined(CRT_TU_WARP_PS2)      // launcher: compiled only in crt_tu_warp_ps2.cu
inline int run_warp_ps2(LaunchEnv& env, const WarpPs2Plan& pl, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state,
                        int has_prev, cudaStream_t st, int* launches, bool pdl) {
    const bool fast = d.triad_mode == 2 && d.triad_comp && d.vig_mode <= 1;
    const bool thr = d.bloom_mode == 1 && d.thr_on;
    auto kern = d.bloom_mode == 1 ? (thr ? (fast ? k_warp_ps2<true, true, true> : k_warp_ps2<true, false, true>)
                                         : (fast ? k_warp_ps2<true, true, false> : k_warp_ps2<true, false, false>))
                                  : (fast ? k_warp_ps2<false, true, false> : k_warp_ps2<false, false, false>);
    if (env.raise((const void*)kern, (int)pl.smem) &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess) return 2;
    const dim3 grid((d.W + WP_TW - 1) / WP_TW, (d.H + WP_TH - 1) / WP_TH);
    const WarpPs2Geo g{pl.tpitch, pl.bpitch, pl.bh, pl.qw, pl.qh};
    const cudaError_t e = launch_pdl(kern, grid, dim3(WP_NT), pl.smem, st, pdl, d, f, in, out, state, has_prev, g);
    ++*launches;
    return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : 2;
}
#endif  // CRT_TU_WARP_PS2

#endif  // __CUDACC__

}  // namespace crt
