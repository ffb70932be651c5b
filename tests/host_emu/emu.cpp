// tests/host_emu/emu.cpp — TEST INFRASTRUCTURE, never shipped, never imported by
// the product.  Compiles the kernels' own per-pixel arithmetic
// (pythoncrt_b200/csrc/crt_math.cuh, crt_stages.cuh, crt_derive.h) with g++ and
// runs the staged pipeline as plain loops, so the arithmetic can be checked
// against the oracle on a machine without a GPU (-ffp-contract=off; the exact
// float32 wrappers keep numpy's operation order).  The CUDA thread mapping,
// shared-memory tiling and the fused kernel are NOT covered here: those are
// checked on the B200 by tests marked `gpu`.
#include <cstring>
#include <string>
#include <vector>

#include "../../pythoncrt_b200/csrc/crt_derive.h"
#include "../../pythoncrt_b200/csrc/crt_stages.cuh"
#include "../../pythoncrt_b200/csrc/crt_fused.cuh"   // host-side tile planner only
#include "../../pythoncrt_b200/csrc/crt_fused_warp_src.cuh"   // host-side planner of the source-driven warp kernel
#include "../../pythoncrt_b200/csrc/crt_policy.h"             // scheduling policy of the C ABI (pure functions)

using namespace crt;

extern "C" int emu_sizeof_params() { return (int)sizeof(crt_params); }
extern "C" int emu_sizeof_frame() { return (int)sizeof(crt_frame); }

extern "C" int emu_process(const crt_params* p, int W, int H, const void* const* tabs, const size_t* tab_bytes,
                           const uint8_t* in, uint8_t* out, float* state, int state_valid, float* img_out,
                           const crt_frame* frames, int n_frames, char* err, int errlen) {
    TablePtrs t{};
    for (int i = 0; i < CRT_TABLE_COUNT; ++i) { t.tab[i] = tabs[i]; t.bytes[i] = tab_bytes[i]; }
    const int hw = W / 2 > 1 ? W / 2 : 1, hh = H / 2 > 1 ? H / 2 : 1;
    std::vector<Lerp1> dn_x = linear_coords(hw, W), dn_y = linear_coords(hh, H), up_x = linear_coords(W, hw), up_y = linear_coords(H, hh);
    std::vector<Lerp1> nz_x, nz_y;
    if (p->noise_strength > 0.0 && p->grain_size > 1) {
        nz_x = linear_coords(W, W / p->grain_size > 1 ? W / p->grain_size : 1);
        nz_y = linear_coords(H, H / p->grain_size > 1 ? H / p->grain_size : 1);
    }
    t.dn_x = dn_x.data(); t.dn_y = dn_y.data(); t.up_x = up_x.data(); t.up_y = up_y.data();
    t.nz_x = nz_x.empty() ? nullptr : nz_x.data(); t.nz_y = nz_y.empty() ? nullptr : nz_y.data();
    alignas(16) static float pow_tab[POW_TAB_FLOATS];
    fill_pow_table(pow_tab);
    t.pow_tab = pow_tab;
    Dev d{};
    std::string e;
    int rc = derive_dev(*p, W, H, t, &d, &e);
    if (rc) { if (err) { strncpy(err, e.c_str(), errlen - 1); err[errlen - 1] = 0; } return rc; }
    const size_t px = (size_t)W * H;
    std::vector<float> ds((size_t)hw * hh * 3), bl, q, S, R;
    Scratch s{ds.data(), nullptr, nullptr};
    if (d.bloom_mode == 2) { bl.resize(px * 3); S.resize(px * 3); R.resize(px * 3); s.bl = bl.data(); }
    if (d.warp_on) { q.resize(px * 3); s.q = q.data(); }
    const GlitchGeom gg = glitch_geom(*p, W, H);
    const bool persist = p->persistence > 0.0 && !img_out;
    for (int n = 0; n < n_frames; ++n) {
        const uint8_t* fin = in + (size_t)n * px * 3;
        FrameDev f = derive_frame(*p, frames[n]);
        f.noise = frames[n].d_noise;
        if (gg.rows > 0 && frames[n].d_glitch_offs) {
            f.goffs = frames[n].d_glitch_offs; f.gy0 = frames[n].glitch_y0; f.gseg = frames[n].glitch_seg_len; f.gnseg = frames[n].glitch_segments;
        }
        if (d.bloom_mode == 1) {
            for (int j = 0; j < d.hh; ++j) for (int i = 0; i < d.hw; ++i) bloom_down_cell(d, fin, ds.data(), j, i);
        } else if (d.bloom_mode == 2) {
            const int K = d.ksize, r = K / 2;
            for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
                F3 v = bloom_src(d, graded_input(d, fin, y, x));
                float* o = &S[((size_t)y * W + x) * 3]; o[0] = v.x; o[1] = v.y; o[2] = v.z;
            }
            std::vector<float> line((size_t)(W + 2 * r) * 3), col((size_t)(H + 2 * r));
            for (int y = 0; y < H; ++y) {
                for (int x = -r; x < W + r; ++x) { int xx = x < 0 ? 0 : (x >= W ? W - 1 : x); memcpy(&line[(size_t)(x + r) * 3], &S[((size_t)y * W + xx) * 3], 12); }
                for (int x = 0; x < W; ++x) for (int c = 0; c < 3; ++c) R[((size_t)y * W + x) * 3 + c] = gauss_row(&line[(size_t)x * 3 + c], 3, d.taps, K);
            }
            for (int x = 0; x < W; ++x) for (int c = 0; c < 3; ++c) {
                for (int y = -r; y < H + r; ++y) { int yy = y < 0 ? 0 : (y >= H ? H - 1 : y); col[y + r] = R[((size_t)yy * W + x) * 3 + c]; }
                for (int y = 0; y < H; ++y) bl[((size_t)y * W + x) * 3 + c] = gauss_col(&col[y + r], 1, d.taps, K);
            }
        }
        if (d.warp_on)
            for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
                F3 v = pre_warp_pixel(d, f, fin, s, y, x, d.lut_fwd, d.lut_inv);
                float* o = &q[((size_t)y * W + x) * 3]; o[0] = v.x; o[1] = v.y; o[2] = v.z;
            }
        const int has_prev = persist && (state_valid || n > 0);
        for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
            F3 v = post_pixel(d, f, fin, s, y, x, d.lut_fwd, d.lut_inv);
            if (img_out) { float* o = img_out + ((size_t)n * px + (size_t)y * W + x) * 3; o[0] = v.x; o[1] = v.y; o[2] = v.z; }
            else finish_pixel(d, v, has_prev, state, out + (size_t)n * px * 3, y, x);
        }
    }
    return 0;
}

// Host-side tile planner of the fused kernel (plan_fused): out = {ok, th, cap_px, cap_aux, smem bytes}
extern "C" int emu_plan(const crt_params* p, int W, int H, const void* const* tabs, const size_t* tab_bytes, int pix_uniform,
                        long long* out, char* why, int whylen) {
    TablePtrs t{};
    for (int i = 0; i < CRT_TABLE_COUNT; ++i) { t.tab[i] = tabs[i]; t.bytes[i] = tab_bytes[i]; }
    const int hw = W / 2 > 1 ? W / 2 : 1, hh = H / 2 > 1 ? H / 2 : 1;
    std::vector<Lerp1> dn_x = linear_coords(hw, W), dn_y = linear_coords(hh, H), up_x = linear_coords(W, hw), up_y = linear_coords(H, hh);
    std::vector<Lerp1> nz(1);
    t.dn_x = dn_x.data(); t.dn_y = dn_y.data(); t.up_x = up_x.data(); t.up_y = up_y.data(); t.nz_x = nz.data(); t.nz_y = nz.data();
    t.pix_uniform = pix_uniform;
    Dev d{};
    std::string e;
    int rc = derive_dev(*p, W, H, t, &d, &e);
    if (rc) { strncpy(why, e.c_str(), whylen - 1); why[whylen - 1] = 0; return rc; }
    FusedPlan pl = plan_fused(d, glitch_active(*p));
    out[0] = pl.ok; out[1] = pl.th; out[2] = pl.cap_px; out[3] = pl.cap_aux; out[4] = (long long)pl.smem;
    strncpy(why, pl.why, whylen - 1); why[whylen - 1] = 0;
    return 0;
}

// Planner of the source-driven warp kernel (csrc/crt_fused_warp_src.cuh): build the tile list, then check the partition the
// kernel relies on for EVERY output pixel — its owner tile exists, the tile's box of output quads contains it, and every tap
// that lies inside the frame lies inside the owner's 64 x 32 source tile; pixels whose taps all fall outside the frame belong
// to exactly one border item.  out = {ok, tiles, violations, largest box in quads, pixels covered by boxes}.
extern "C" int emu_check_warp_src(const crt_params* p, int W, int H, const void* const* tabs, const size_t* tab_bytes, int pix_uniform,
                                  long long* out, char* why, int whylen) {
    TablePtrs t{};
    for (int i = 0; i < CRT_TABLE_COUNT; ++i) { t.tab[i] = tabs[i]; t.bytes[i] = tab_bytes[i]; }
    const int hw = W / 2 > 1 ? W / 2 : 1, hh = H / 2 > 1 ? H / 2 : 1;
    std::vector<Lerp1> dn_x = linear_coords(hw, W), dn_y = linear_coords(hh, H), up_x = linear_coords(W, hw), up_y = linear_coords(H, hh);
    std::vector<Lerp1> nz(1);
    t.dn_x = dn_x.data(); t.dn_y = dn_y.data(); t.up_x = up_x.data(); t.up_y = up_y.data(); t.nz_x = nz.data(); t.nz_y = nz.data();
    t.pix_uniform = pix_uniform;
    Dev d{};
    std::string e;
    int rc = derive_dev(*p, W, H, t, &d, &e);
    if (rc) { strncpy(why, e.c_str(), whylen - 1); why[whylen - 1] = 0; return rc; }
    const WarpSrcPlan pl = plan_warp_src(d, glitch_active(*p));
    strncpy(why, pl.why, whylen - 1); why[whylen - 1] = 0;
    out[0] = pl.ok; out[1] = (long long)pl.tiles.size(); out[2] = 0; out[3] = 0; out[4] = 0;
    if (!pl.ok) return 0;
    const int otx = (W + P2_TW - 1) / P2_TW;
    std::vector<int> index((size_t)pl.ntx * pl.nty, -1);
    std::vector<std::vector<int>> border((size_t)otx * ((H + P2_TH - 1) / P2_TH));
    for (size_t i = 0; i < pl.tiles.size(); ++i) {
        const WsTile& tl = pl.tiles[i];
        if (tl.kind == WS_BORDER) border[(size_t)(tl.by0 / P2_TH) * otx + tl.bx0 / P2_TW].push_back((int)i);
        else index[(size_t)((tl.y0 - WS_ORG) / WS_SY) * pl.ntx + (tl.x0 - WS_ORG) / WS_SX] = (int)i;
        if ((long long)tl.bw4 * tl.bh > out[3]) out[3] = (long long)tl.bw4 * tl.bh;
        out[4] += (long long)tl.bw4 * 4 * tl.bh;
        if (tl.kind != WS_BORDER && (tl.x0 % 2 || tl.y0 % 2)) ++out[2];
        if (tl.bx0 % 4) ++out[2];
        if ((tl.kind == WS_BORDER) != ((int)i >= pl.n_source)) ++out[2];          // border items come last
    }
    long long bad = 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const Taps tp = warp_taps_n(d, warp_norm((float)x, d.warp_cx, d.warp_dx), warp_norm((float)y, d.warp_cy, d.warp_dy));
            if (ws_outside(tp.ix, tp.iy, W, H)) {                // exactly one border item: the pixel's output tile, box containing it
                const std::vector<int>& v = border[(size_t)(y / P2_TH) * otx + x / P2_TW];
                if (v.size() != 1) { ++bad; continue; }
                const WsTile& tl = pl.tiles[v[0]];
                if (x < tl.bx0 || x >= tl.bx0 + 4 * tl.bw4 || y < tl.by0 || y >= tl.by0 + tl.bh) ++bad;
                for (int j = 0; j < 4; ++j) {
                    const int ty = tp.iy + (j >> 1), tx = tp.ix + (j & 1);
                    if (ty >= 0 && ty < H && tx >= 0 && tx < W) ++bad;          // "outside" means all four taps
                }
                continue;
            }
            const int k = index[(size_t)ws_owner(tp.iy, WS_SY) * pl.ntx + ws_owner(tp.ix, WS_SX)];
            if (k < 0) { ++bad; continue; }
            const WsTile& tl = pl.tiles[k];
            if (x < tl.bx0 || x >= tl.bx0 + 4 * tl.bw4 || y < tl.by0 || y >= tl.by0 + tl.bh) ++bad;
            for (int j = 0; j < 4; ++j) {
                const int ty = tp.iy + (j >> 1), tx = tp.ix + (j & 1);
                if (ty < 0 || ty >= H || tx < 0 || tx >= W) { if (tl.kind == WS_INTERIOR) ++bad; continue; }
                if (tx < tl.x0 || tx >= tl.x0 + P2_TW || ty < tl.y0 || ty >= tl.y0 + P2_TH) ++bad;
            }
        }
    out[2] += bad;
    return 0;
}

// pow_unit on arrays, for the accuracy test
// n / d through div_const (csrc/crt_math.cuh) next to the compiler's IEEE division
extern "C" void emu_div_const(const float* n, float* out, float* want, int count, float d) {
    const float y = rcp_rn(d);
    for (int i = 0; i < count; ++i) { out[i] = div_const(n[i], d, y); volatile float q = n[i] / d; want[i] = q; }
}

extern "C" void emu_pow_unit(const float* x, float* out, int n, double y) {
    alignas(16) static float pow_tab[POW_TAB_FLOATS];
    fill_pow_table(pow_tab);
    for (int i = 0; i < n; ++i) out[i] = pow_unit(x[i], (float)y, pow_tab);
}

// ---- scheduling policy (crt_policy.h): out = {tile_h per frame, clip tile_h, clip size ok, shards, auto prefers clip, halo} ----
extern "C" void emu_policy(int W, int H, int sms, int per_sm, int halo_blocks, int gaussian, int shards_wanted, int n_frames, double persistence,
                           int* out) {
    const int resident = sms * per_sm;
    out[0] = policy_tile_h(W, H, sms, per_sm, halo_blocks, 0);
    out[1] = policy_clip_tile_h(W, H, resident, 0);
    out[2] = policy_clip_size_ok(W, H, resident, 0, 0) ? 1 : 0;
    out[3] = policy_shards(shards_wanted, -1, n_frames, persistence, W, H);
    out[4] = policy_auto_prefers_clip(gaussian != 0, W, H, resident) ? 1 : 0;
    out[5] = policy_halo_frames(persistence);
}
