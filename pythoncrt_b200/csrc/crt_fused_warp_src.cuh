// crt_fused_warp_src.cuh — single-pass barrel warp, SOURCE-driven: 30 B/px, the pre-warp image never leaves the SM.
//
// apply_barrel_warp (crt_filter.py:331-348, called at :649) samples the processed image, so a warped OUTPUT tile needs
// stages 0-10 over its source footprint — 1.6-1.8x the tile's pixels once aligned (k_warp_ps2, crt_fused_warp_ps2.cuh: 354
// instructions per pixel, slower than two passes).  This kernel turns the assignment round: the frame of SOURCE pixels is cut
// into fixed 64 x 32 tiles (stride 62 x 30: a bilinear tap pair needs one more column and row), each evaluated exactly like
// a tile of the block kernel k_fused_ps2_pipe (input bytes by TMA, one graded value per 2x2 block, bloom from block values,
// composite triad LUT, mask tables) into a float32 tile in shared memory — the expensive phase runs over 1.10x the frame.
// Every OUTPUT pixel belongs to exactly one work item: the source tile that holds its top-left tap (ix, iy), or — when all four
// taps fall outside the frame (the black border of a barrel warp: value 0, only the persistence blend is left) — the
// 64 x 32 OUTPUT tile it lies in, a light "border item" without a source tile (round 2, run 36: handing those pixels to the
// nearest edge tile made the four corner tiles 24x heavier than the rest and the kernel 2x slower).  The host planner
// (plan_warp_src) evaluates the map for every output pixel once per clip — with the kernel's own float32 arithmetic — and
// gives every tile the bounding box of its output pixels; the kernel walks that box in 4-pixel quads, keeps the pixels it
// owns, gathers their cv2.remap taps from the tile (taps outside the frame contribute 0, :347), blends with the persistence
// state in HBM and writes state + packed uint8.  Quads owned entirely take 16-byte state accesses, quads cut by a tile
// boundary scalar ones (a pixel is written by exactly one CTA).
// Requirements: a pixel_size-2 chain with fast / no bloom (the block kernel's), a monotone map, no glitch (its row shifts
// wrap around the frame) and no text layer after the warp; the clip and state 16-byte aligned, W % 8 == 0.
#pragma once
#include "crt_fused_ps2.cuh"

namespace crt {

constexpr int WS_SX = P2_TW - 2, WS_SY = P2_TH - 2;      // tile stride in source pixels (62 x 30)
constexpr int WS_ORG = -2;                               // first tile origin: tap index -1 (left / top tap outside the frame) is owned too

struct WsTile {              // one source tile and the output pixels it owns (32 bytes)
    int x0, y0;              // tile origin in the source frame (even; may be -2)
    int bx0, by0;            // bounding box of the owned output pixels: x start (multiple of 4), y start
    int bw4, bh;             // quads per box row, box rows
    unsigned magic;          // make_magic(bw4)
    int kind;                // WS_EDGE / WS_INTERIOR source tile, or WS_BORDER item (no source tile: x0, y0 unused)
};
enum : int { WS_EDGE = 0, WS_INTERIOR = 1, WS_BORDER = 2 };

// all four taps of (ix, iy) outside the frame: the pixel's value is 0 and it belongs to a border item
CRT_HD bool ws_outside(int ix, int iy, int W, int H) { return ix < -1 || ix >= W || iy < -1 || iy >= H; }
// column / row of the source tile that owns tap index i, for i in [-1, n_px - 1] (identical on host and device)
CRT_HD int ws_owner(int i, int stride) { return (i - WS_ORG) / stride; }

struct WarpSrcPlan {
    bool ok = false;
    const char* why = "not planned";
    int ntx = 0, nty = 0, n_source = 0;
    std::vector<WsTile> tiles;           // non-empty source tiles (row-major), then the border items
};

inline WarpSrcPlan plan_warp_src(const Dev& d, bool glitch_on) {
    WarpSrcPlan pl;
    if (!(d.warp_on && d.warp_mono && d.pix_uniform == 2 && d.even_dims && (d.W & 7) == 0 && !glitch_on && d.text_mode == 0 && d.bloom_mode != 2)) {
        pl.why = "not a pixel_size-2 chain with a monotone warp"; return pl;
    }
    if (!fused_ps2_pipe_supported(d)) { pl.why = "the TMA input tile needs |aberration| <= 6"; return pl; }
    if ((size_t)d.W * d.H * 3 >= ((size_t)1 << 31)) { pl.why = "frame too large for 32-bit indexing"; return pl; }
    std::vector<float> xn(d.W), yn(d.H);
    for (int x = 0; x < d.W; ++x) xn[x] = warp_norm((float)x, d.warp_cx, d.warp_dx);
    for (int y = 0; y < d.H; ++y) yn[y] = warp_norm((float)y, d.warp_cy, d.warp_dy);
    // the kernel divides by max(1, cx) / max(1, cy) through div_const: every operand is checked against the IEEE quotient
    const float rx = rcp_rn(d.warp_dx), ry = rcp_rn(d.warp_dy);
    for (int x = 0; x < d.W; ++x) if (div_const(fsub((float)x, d.warp_cx), d.warp_dx, rx) != xn[x]) { pl.why = "div_const differs from the IEEE quotient"; return pl; }
    for (int y = 0; y < d.H; ++y) if (div_const(fsub((float)y, d.warp_cy), d.warp_dy, ry) != yn[y]) { pl.why = "div_const differs from the IEEE quotient"; return pl; }
    const int ntx = (d.W - 1 - WS_ORG) / WS_SX + 1, nty = (d.H - 1 - WS_ORG) / WS_SY + 1;       // tap indices -1 .. W-1 / H-1
    struct BB { int x0, y0, x1, y1; };
    const int otx = (d.W + P2_TW - 1) / P2_TW, oty = (d.H + P2_TH - 1) / P2_TH;     // output tiles (border items)
    std::vector<BB> bb((size_t)ntx * nty, BB{0x7fffffff, 0x7fffffff, -1, -1}), ob((size_t)otx * oty, BB{0x7fffffff, 0x7fffffff, -1, -1});
    for (int y = 0; y < d.H; ++y)
        for (int x = 0; x < d.W; ++x) {
            const Taps t = warp_taps_n(d, xn[x], yn[y]);
            BB& b = ws_outside(t.ix, t.iy, d.W, d.H) ? ob[(size_t)(y / P2_TH) * otx + x / P2_TW]
                                                     : bb[(size_t)ws_owner(t.iy, WS_SY) * ntx + ws_owner(t.ix, WS_SX)];
            if (x < b.x0) b.x0 = x;
            if (x > b.x1) b.x1 = x;
            if (y < b.y0) b.y0 = y;
            if (y > b.y1) b.y1 = y;
        }
    auto push = [&](const BB& b, int x0, int y0, int kind) {
        WsTile t;
        t.x0 = x0; t.y0 = y0;
        t.bx0 = b.x0 & ~3; t.by0 = b.y0;
        t.bw4 = (b.x1 - t.bx0) / 4 + 1; t.bh = b.y1 - b.y0 + 1;
        t.magic = t.bw4 == 1 ? 0u : 0xffffffffu / (unsigned)t.bw4 + 1u;       // fastdiv: exact for q * bw4 < 2^32
        t.kind = kind;
        pl.tiles.push_back(t);
    };
    // source tiles first (row-major), border items last: they are light, and a CTA never meets a source tile after one
    for (int tj = 0; tj < nty; ++tj)
        for (int ti = 0; ti < ntx; ++ti) {
            const BB& b = bb[(size_t)tj * ntx + ti];
            if (b.x1 < 0) continue;
            const int x0 = WS_ORG + ti * WS_SX, y0 = WS_ORG + tj * WS_SY;
            push(b, x0, y0, (x0 >= 0 && y0 >= 0 && x0 + P2_TW <= d.W && y0 + P2_TH <= d.H) ? WS_INTERIOR : WS_EDGE);
        }
    pl.n_source = (int)pl.tiles.size();
    for (size_t i = 0; i < ob.size(); ++i) if (ob[i].x1 >= 0) push(ob[i], 0, 0, WS_BORDER);
    pl.ntx = ntx; pl.nty = nty;
    pl.ok = !pl.tiles.empty(); pl.why = pl.ok ? "" : "no tiles";
    return pl;
}

#if defined(__CUDACC__)

constexpr int WS_SMEM = P2_ST_BYTES + 2 * P2_RAW_BYTES;      // the float32 tile + the double-buffered input bytes

template <bool BLOOM, bool FAST, bool THR, int SPEC = 0>
__global__ void __launch_bounds__(P2_NT, THR ? 3 : 4) k_warp_src(Dev d_arg, FrameDev f_arg, const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                                   float* __restrict__ state, int has_prev, const WsTile* __restrict__ tiles, int ntiles,
                                                                   int ntx, int nty, float rcp_dx, float rcp_dy,
                                                                   const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_st,
                                                                   int frame) {
    Dev d = d_arg;
    FrameDev f = f_arg;
    specialise<SPEC>(d, f);
    extern __shared__ __align__(128) unsigned char dsm[];
    float* const T = reinterpret_cast<float*>(dsm);                                  // [TH][TW*3]: pre-warp image of the source tile
    uint8_t* const s_raw = dsm + P2_ST_BYTES;                                        // [2][18][256]
    __shared__ __align__(16) float s_lut[2 * 1028];
    __shared__ __align__(16) float s_sel[3][12];
    float* const s_fwd = s_lut;
    float* const s_inv = s_lut + 1028;
    __shared__ float s_unit[256];
    __shared__ __align__(16) float s_pow[POW_TAB_FLOATS];
    __shared__ float s_rows[2 * P2_TH], s_cols[2 * P2_TW];
    __shared__ __align__(16) float Us[3][P2_BH][P2_BW + 2];
    __shared__ __align__(16) float Ss[THR ? 3 : 1][THR ? P2_BH : 1][P2_BW + 2];
    __shared__ __align__(8) uint64_t bar_in[2];
    const int tid = threadIdx.x;
    griddep_launch_dependents();
    const int a0 = d.aberr != 0 ? d.aberr_mod : 0;
    const int as = a0 > (d.W >> 1) ? a0 - d.W : a0;                         // signed shift (aberr_mod is taken modulo W)
    const int aa = as < 0 ? -as : as;
    if (tid == 0) {
        mbar_init(&bar_in[0], 1); mbar_init(&bar_in[1], 1);
        fence_mbar_init();
        if ((int)blockIdx.x < ntiles && tiles[blockIdx.x].kind != WS_BORDER) {          // first tile's input: independent of the previous kernel
            const WsTile t0 = tiles[blockIdx.x];
            mbar_expect_tx(&bar_in[0], P2_RAW_BYTES);
            tma_load_2d_hint(s_raw, &map_in, (6 * ((t0.x0 >> 1) - 1) - 3 * aa) & ~15, frame * d.hh + (t0.y0 >> 1) - 1, &bar_in[0], L2_EVICT_FIRST);
        }
    }
    const float* lut_a = d.triad_comp ? d.triad_comp : d.lut_fwd;
    const float* lut_b = d.triad_comp ? d.triad_comp + 1028 : d.lut_inv;
    if (d.triad_mode >= 2) {
        reinterpret_cast<float4*>(s_fwd)[tid] = reinterpret_cast<const float4*>(lut_a)[tid];
        reinterpret_cast<float4*>(s_inv)[tid] = reinterpret_cast<const float4*>(lut_b)[tid];
        if (tid == 0) { s_fwd[1024] = lut_a[1024]; s_inv[1024] = lut_b[1024]; }
    }
    s_unit[tid] = __fdiv_rn((float)tid, 255.0f);
    ps2_fill_sel(s_sel, tid, d.bgr);
    if (d.col_gamma) for (int i = tid; i < POW_TAB_FLOATS; i += blockDim.x) s_pow[i] = d.pow_tab[i];
    MaskTabs mt{s_rows, s_cols, s_rows + P2_TH, s_cols + P2_TW};
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const WsTile tl = tiles[tile];
        const int ox0 = tl.x0, oy0 = tl.y0;
        const int ox1 = imin(ox0 + P2_TW, d.W) - 1, oy1 = imin(oy0 + P2_TH, d.H) - 1;
        const int gbx0 = (ox0 >> 1) - 1, gby0 = (oy0 >> 1) - 1;
        const int buf = it & 1;
        const bool source = tl.kind != WS_BORDER;               // CTA-uniform; border items come last in the list
        if (has_prev && tid == 0) {
            // The previous state of this item's output box is read at the very end of the iteration, from wherever the map put it:
            // pull it into L2 now with tiled TMA prefetches (192-float x 32-row boxes), a whole tile's worth of work ahead.
            // (ncu, run 37: 39 % of the stall samples sat on the first use of those loads.)  Safe before griddepcontrol.wait:
            // a prefetch only moves lines into L2, where the previous kernel's stores land as well.
            for (int yy = tl.by0; yy < tl.by0 + tl.bh; yy += P2_TH)
                for (int xx = tl.bx0; xx < tl.bx0 + 4 * tl.bw4; xx += P2_TW) tma_prefetch_2d(&map_st, xx * 3, yy);
        }
        if (source) {
        if (tid == 0 && tile + (int)gridDim.x < ntiles) {      // next tile's input into the other buffer (last read two barriers ago)
            const WsTile tn = tiles[tile + gridDim.x];
            if (tn.kind != WS_BORDER) {
                mbar_expect_tx(&bar_in[buf ^ 1], P2_RAW_BYTES);
                tma_load_2d_hint(s_raw + (buf ^ 1) * P2_RAW_BYTES, &map_in, (6 * ((tn.x0 >> 1) - 1) - 3 * aa) & ~15, frame * d.hh + (tn.y0 >> 1) - 1,
                                 &bar_in[buf ^ 1], L2_EVICT_FIRST);
            }
        }
        if (tid < P2_TH) {
            const int y = oy0 + tid;
            if (d.scan_mode == 1) mt.row_scan[tid] = scan_row(d, f, y);
            else if (d.scan_mode == 2) { double t = ((double)y + f.phase) * d.scan_inv_period; mt.row_scan[tid] = (float)(t - floor(t)); }
            if (d.vig_mode == 1) { const float ny = ((float)y - d.vig_cy) * d.vig_iry; mt.row_vig[tid] = ny * ny; }
        } else if (tid >= 64 && tid < 64 + P2_TW) {
            const int c = tid - 64, x = ox0 + c;
            if (d.scan_mode == 2) { double t = (d.scan_tan * (double)x) * d.scan_inv_period; mt.col_scan[c] = (float)(t - floor(t)); }
            if (d.vig_mode == 1) { const float nx = ((float)x - d.vig_cx) * d.vig_irx; mt.col_vig[c] = nx * nx; }
        }
        // Tables staged before the loop.  Later tiles: the mask tables above are written while slower warps may still be in the
        // previous tile's gather, which reads neither them nor Us; the barrier after phase 1 orders them before the tail.
        if (it == 0) __syncthreads();

        // ---- phase 1: one graded value per 2x2 block of the source tile + one halo block (clamped = cv2's edge rule) ----
        mbar_wait(&bar_in[buf], (it >> 1) & 1);                 // this tile's input bytes have landed
        {
            constexpr int NIT = (P2_BW * P2_BH + P2_NT - 1) / P2_NT;
            const bool x_inside = 2 * gbx0 - aa >= 0 && 2 * (gbx0 + P2_BW - 1) + aa < d.W;       // tile-uniform
            const uint8_t* rawb = s_raw + buf * P2_RAW_BYTES;
            const int xoff = 6 * gbx0 - ((6 * gbx0 - 3 * aa) & ~15);
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
                const int u = tid + i * P2_NT;
                if (u < P2_BW * P2_BH) {
                    const int bj = u / P2_BW, bi = u - bj * P2_BW;
                    uint32_t r0, r1, r2;
                    if (x_inside) {
                        const uint8_t* p = rawb + (imin(imax(gby0 + bj, 0), d.hh - 1) - gby0) * P2_RAW_W + 6 * bi + xoff;
                        r0 = p[-3 * as]; r1 = p[1]; r2 = p[3 * as + 2];
                    } else {
                        const int sx = 2 * imin(imax(gbx0 + bi, 0), d.hw - 1), sy = 2 * imin(imax(gby0 + bj, 0), d.hh - 1);
                        const uint8_t* row = in + (size_t)sy * d.W * 3;
                        r0 = row[wrap(sx - a0, d.W) * 3 + 0]; r1 = row[sx * 3 + 1]; r2 = row[wrap(sx + a0, d.W) * 3 + 2];
                    }
                    const F3 v1 = colour(d, mk3(s_unit[r0], s_unit[r1], s_unit[r2]), s_pow);
                    Us[0][bj][bi] = v1.x; Us[1][bj][bi] = v1.y; Us[2][bj][bi] = v1.z;
                    if (BLOOM && THR) {
                        const F3 sv = bloom_src(d, v1);
                        Ss[0][bj][bi] = sv.x; Ss[1][bj][bi] = sv.y; Ss[2][bj][bi] = sv.z;
                    }
                }
            }
        }
        __syncthreads();        // (also: every warp has left the previous tile's gather, T may be overwritten)

        // ---- phase 2: stages 5-10 of the source tile into T (the block kernel's tail, no blend, no output) ----
        {
            const int tx = tid & 15, ty = tid >> 4;
            const int xb = ox0 + 4 * tx, y0 = oy0 + 2 * ty;
            if (xb <= ox1 && y0 <= oy1) {
                const int bi = 2 * tx + 1, bj = ty + 1;
                float blr[4][3], t1[2][3];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) { t1[0][ch] = Us[ch][bj][bi]; t1[1][ch] = Us[ch][bj][bi + 1]; }
                auto row_begin = [&](int r) {        // cv2's 2x up-scale of the bloom for patch row r (see k_fused_ps2)
#pragma unroll
                    for (int ch = 0; ch < (BLOOM ? 3 : 0); ++ch) {
                        const float (*src)[P2_BW + 2] = THR ? Ss[THR ? ch : 0] : Us[ch];
                        float h[2][4];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const float2 ca = *reinterpret_cast<const float2*>(&src[bj - 1 + r + q][bi - 1]);
                            const float2 cb = *reinterpret_cast<const float2*>(&src[bj - 1 + r + q][bi + 1]);
                            const float d01 = fsub(ca.y, ca.x), d12 = fsub(cb.x, ca.y), d23 = fsub(cb.y, cb.x);
                            h[q][0] = ffma(d01, 0.75f, ca.x); h[q][1] = ffma(d12, 0.25f, ca.y);
                            h[q][2] = ffma(d12, 0.75f, ca.y); h[q][3] = ffma(d23, 0.25f, cb.x);
                        }
                        const float w = r == 0 ? 0.75f : 0.25f;
#pragma unroll
                        for (int k = 0; k < 4; ++k) blr[k][ch] = ffma(fsub(h[1][k], h[0][k]), w, h[0][k]);
                    }
                };
                // q_out != nullptr + state_in_smem: the patch's float values go to T; the specialised tail only on tiles inside the frame
                ps2_patch_tail<BLOOM, FAST>(d, f, mt, s_fwd, s_inv, s_sel, nullptr, nullptr, T /* marker */, 0, ox0, oy0, ox1, oy1, xb, y0, t1,
                                            [&](int, int k) { return mk3(blr[k][0], blr[k][1], blr[k][2]); },
                                            T + (y0 - oy0) * (P2_TW * 3) + 12 * tx, true, row_begin, P2_TW * 3, ox0 >= 0 && oy0 >= 0);
            }
        }
        }                       // source tile
        __syncthreads();
        griddep_wait();         // previous kernel of the stream complete: the state may be touched from here on

        // ---- phase 3: the output pixels this item owns: taps, gather from T, persistence, quantise ----
        const int my_ti = (ox0 - WS_ORG) / WS_SX, my_tj = (oy0 - WS_ORG) / WS_SY;
        const int nq = tl.bw4 * tl.bh;
        for (int q = tid; q < nq; q += P2_NT) {
            const int row = fastdiv(q, tl.magic), c4 = q - row * tl.bw4;
            const int y = tl.by0 + row, xb = tl.bx0 + 4 * c4;
            const int o = (y * d.W + xb) * 3;
            float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa, pc = pa;      // previous state of the quad: loaded first, used last
            if (has_prev) {
                const float4* sp = reinterpret_cast<const float4*>(state + o);
                pa = sp[0]; pb = sp[1]; pc = sp[2];
            }
            const float yn = div_const(fsub((float)y, d.warp_cy), d.warp_dy, rcp_dy);
            Taps tp[4];
            unsigned own = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                tp[k] = warp_taps_n(d, div_const(fsub((float)(xb + k), d.warp_cx), d.warp_dx, rcp_dx), yn);
                // (comparing the tap against the tile's core range instead of dividing measured 5 % slower: run 39)
                const bool outside = ws_outside(tp[k].ix, tp[k].iy, d.W, d.H);
                const bool mine = source ? (!outside && ws_owner(tp[k].ix, WS_SX) == my_ti && ws_owner(tp[k].iy, WS_SY) == my_tj)
                                         : (outside && xb + k < d.W);
                if (mine) own |= 1u << k;
            }
            if (!own) continue;
            float res[12];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const Taps& t = tp[k];
                const int lx = t.ix - ox0, ly = t.iy - oy0;
                F3 a[4];
                if (!source) {
                    a[0] = a[1] = a[2] = a[3] = mk3(0.f, 0.f, 0.f);            // border item: every tap outside the frame
                } else if (tl.kind == WS_INTERIOR) {
                    const float* base = T + (ly * P2_TW + lx) * 3;              // an owned pixel's four taps lie inside the tile
                    if (own >> k & 1) { a[0] = load_f3(base); a[1] = load_f3(base + 3); a[2] = load_f3(base + P2_TW * 3); a[3] = load_f3(base + P2_TW * 3 + 3); }
                    else a[0] = a[1] = a[2] = a[3] = mk3(0.f, 0.f, 0.f);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int ty = t.iy + (j >> 1), tx = t.ix + (j & 1);
                        const bool ok = (own >> k & 1) && (unsigned)ty < (unsigned)d.H && (unsigned)tx < (unsigned)d.W;      // inside the frame (hence inside T)
                        a[j] = ok ? load_f3(T + ((ty - oy0) * P2_TW + (tx - ox0)) * 3) : mk3(0.f, 0.f, 0.f);
                    }
                }
                res[k * 3] = gather4_fast(a[0].x, a[1].x, a[2].x, a[3].x, t);
                res[k * 3 + 1] = gather4_fast(a[0].y, a[1].y, a[2].y, a[3].y, t);
                res[k * 3 + 2] = gather4_fast(a[0].z, a[1].z, a[2].z, a[3].z, t);
            }
            if (has_prev) {
                const float prev[12] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w, pc.x, pc.y, pc.z, pc.w};
#pragma unroll
                for (int e = 0; e < 12; ++e) res[e] = blend_fast(prev[e], res[e], d.persist, d.persist_q);
            }
            if (own == 0xfu) {                                   // the whole quad: 16-byte state stores, packed uint8 stores
                if (state) {
                    float4* sp = reinterpret_cast<float4*>(state + o);
                    sp[0] = make_float4(res[0], res[1], res[2], res[3]);
                    sp[1] = make_float4(res[4], res[5], res[6], res[7]);
                    sp[2] = make_float4(res[8], res[9], res[10], res[11]);
                }
                uint32_t* op = reinterpret_cast<uint32_t*>(out + o);
                __stcs(op, pack4(res[0], res[1], res[2], res[3]));
                __stcs(op + 1, pack4(res[4], res[5], res[6], res[7]));
                __stcs(op + 2, pack4(res[8], res[9], res[10], res[11]));
            } else {                                             // a quad cut by a tile boundary: only the owned pixels, scalar stores
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!(own >> k & 1)) continue;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        if (state) state[o + k * 3 + ch] = res[k * 3 + ch];
                        out[o + k * 3 + ch] = (uint8_t)quantise_fast(res[k * 3 + ch]);
                    }
                }
            }
        }
    }
}

#if defined(CRT_TU_WARP_PS2)      // launcher: compiled only in crt_tu_warp_ps2.cu
inline int run_warp_src(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, int has_prev,
                        cudaStream_t st, int* launches, bool pdl, const WsTile* d_tiles, int ntiles, int ntx, int nty, const Ps2Maps* maps) {
    const bool fast = d.triad_mode == 2 && d.triad_comp && d.vig_mode <= 1;
    const bool thr = d.bloom_mode == 1 && d.thr_on;
    auto kern = !thr ? (d.bloom_mode == 1 ? (fast ? k_warp_src<true, true, false> : k_warp_src<true, false, false>)
                                          : (fast ? k_warp_src<false, true, false> : k_warp_src<false, false, false>))
                     : (fast ? k_warp_src<true, true, true> : k_warp_src<true, false, true>);
    static const bool use_spec = env_int("CRT_SPEC", 1) != 0;
    if (use_spec && !thr && d.bloom_mode == 1 && fast) {
        if (spec_matches(SPEC_DEFAULT, d, f.flicker_on != 0, fast)) kern = k_warp_src<true, true, false, SPEC_DEFAULT>;
        else if (spec_matches(SPEC_SLANTED, d, f.flicker_on != 0, fast)) kern = k_warp_src<true, true, false, SPEC_SLANTED>;
    }
    if (env.raise((const void*)kern, WS_SMEM) &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM) != cudaSuccess) return 2;
    const int resident = env.sms * (thr ? 3 : 4);
    const cudaError_t e = launch_pdl(kern, dim3(ntiles < resident ? ntiles : resident), dim3(P2_NT), (size_t)WS_SMEM, st, pdl,
                                     d, f, in, out, state, has_prev, d_tiles, ntiles, ntx, nty, rcp_rn(d.warp_dx), rcp_rn(d.warp_dy), maps->in, maps->st, maps->frame);
    ++*launches;
    return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : 2;
}
#endif  // CRT_TU_WARP_PS2

#endif  // __CUDACC__

}  // namespace crt
