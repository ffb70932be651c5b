// crt_policy.h — host-side scheduling policy of the C ABI as pure functions (no CUDA, no context): tile heights, temporal shards,
// the choice between clip mode and shards.  crt_abi.cu calls them with the environment overrides it has read; tests/host_emu
// compiles them for the CPU suite (tests/test_host_logic.py pins the decisions DESIGN.md quotes).
#pragma once
#include <cmath>
#include <cstddef>

namespace crt {

constexpr int POLICY_TW = 64, POLICY_TH = 32;      // tile of the pixel_size-2 block kernels (P2_TW x P2_TH, crt_fused_ps2.cuh)

inline int policy_tiles(int W, int H, int th) { return ((W + POLICY_TW - 1) / POLICY_TW) * ((H + th - 1) / th); }

// Tile height of the single-pass block kernels for a frame processed ALONE on the GPU.  Their CTAs are persistent and walk the
// tiles with a fixed stride, so a frame costs ceil(tiles / resident CTAs) tile times: 1080p in 64 x 32 tiles is 1020 tiles for
// 444 (gaussian bloom, 3 CTAs per SM) or 592 (4 per SM) CTAs = 3 resp. 2 rounds of which the last is 30 % resp. 72 % full.
// A lower tile (28 rows: 1170 tiles, 1.98 rounds) fills the rounds; the cost model is rounds x (rows + a), a = the rows' worth
// of halo and per-tile overhead, and 32 rows stay unless the model gains 6 %.  4K: 6.9 rounds, nothing to gain.  Concurrent
// temporal shards fill each other's partial rounds and keep 32 rows (process_sharded).  forced = CRT_TILE_H (0: none, < 0: 32).
inline int policy_tile_h(int W, int H, int sms, int per_sm, int halo_blocks, int forced) {
    if (forced >= 8 && forced <= POLICY_TH && !(forced & 1)) return forced;
    if (forced < 0 || halo_blocks < 0) return POLICY_TH;       // halo_blocks < 0: only on request
    const int slots = sms * per_sm;
    // frames that do not even half fill the GPU in 32-row tiles (VGA: 150 tiles for 592 CTAs): lower tiles until three fifths of
    // the CTAs have one — measured (run 66, VGA, TMA-pipelined kernel): 128 k frames/s in 32-row tiles of the plain kernel,
    // 137 k with 16 rows, 148 k with 12 (400 tiles), 130 k with 8
    if (2 * policy_tiles(W, H, POLICY_TH) < slots) {
        for (int th = POLICY_TH - 2; th > 12; th -= 2)
            if (5 * policy_tiles(W, H, th) >= 3 * slots) return th;
        return 12;
    }
    const double a = 2.0 + 1.2 * halo_blocks;
    auto cost = [&](int th) { return (double)((policy_tiles(W, H, th) + slots - 1) / slots) * (th + a); };
    int best = POLICY_TH;
    for (int th = POLICY_TH - 2; th >= 20; th -= 2)
        if (cost(th) < cost(best)) best = th;
    return cost(best) <= 0.94 * cost(POLICY_TH) ? best : POLICY_TH;
}

// Tile height of clip-mode launches: the tallest tile that gives the resident CTAs one and a half tiles each (1080p and up: 32
// rows; 720p: 16), not below 16 rows.  Measured (run 63): a tile costs ~8 us of mostly fixed latency whatever its height (TMA round
// trips, barriers, the flag protocol), so low tiles only pay as far as they are needed to keep every CTA busy — 720p 9.45 us per
// frame with 24 rows, 8.63 with 16; VGA in 8-row tiles 8.1 us per frame, slower than one launch per frame with programmatic
// dependent launch (119 k against 130 k frames/s): frames that small stay out of clip mode.  forced = CRT_CLIP_TH.
inline int policy_clip_tile_h(int W, int H, int resident, int forced) {
    if (forced >= 8 && forced <= POLICY_TH && !(forced & 1)) return forced;
    for (int th = POLICY_TH; th > 16; th -= 2)
        if (2 * policy_tiles(W, H, th) >= 3 * resident) return th;
    return 16;
}
// ... and whether the frame is large enough for clip mode at that height (a tile's frames are a serial chain: with fewer tiles
// than resident CTAs the chain is the bound).  min_tiles = CRT_CLIP_MIN_TILES (<= 0: one tile per resident CTA).
inline bool policy_clip_size_ok(int W, int H, int resident, int forced_th, int min_tiles) {
    return policy_tiles(W, H, policy_clip_tile_h(W, H, resident, forced_th)) >= (min_tiles > 0 ? min_tiles : resident);
}

// warm-up frames of a temporal shard: persistence^k <= 1/2040 (an eighth of an LSB)
inline int policy_halo_frames(double persistence) {
    if (!(persistence > 0.0)) return 0;
    return (int)std::ceil(std::log(1.0 / 2040.0) / std::log(persistence));
}

// Number of intra-GPU temporal shards: wanted = crt_set_shards (0 automatic, 1 off, k at most k), forced = CRT_SHARDS (< 0: none)
inline int policy_shards(int wanted, int forced, int n_frames, double persistence, int W, int H) {
    if (forced >= 0) wanted = forced;
    if (wanted == 1 || n_frames < 2) return 1;
    // automatic mode: an 8K frame fills the GPU on its own (16 000+ tiles per kernel); concurrent shards only thrash L2 there
    // (measured, run 33: BASELINE configs[4] 1 245 frames/s on one stream, 1 203 with three shards)
    if (wanted == 0 && (size_t)W * H >= (size_t)24 << 20) return 1;
    const int halo = policy_halo_frames(persistence);
    // a shard must be worth its warm-up: at least 8 halos (<= 12.5 % extra frames) and 48 frames long
    const int min_chunk = halo * 8 > 48 ? halo * 8 : 48;
    int k = n_frames / min_chunk;
    const int cap = wanted == 0 ? 4 : wanted;
    if (k > cap) k = cap;
    return k < 1 ? 1 : k;
}

// Automatic mode, a clip that could be sharded AND could run in clip mode: clip mode (one stream, the serial recurrence exactly)
// where it measures faster than concurrent shards of per-frame launches — the fast-bloom / no-bloom kernel between one and four
// rounds of tiles per frame (1080p: 80-84 k against 79 k frames/s); at 4K (6.9 rounds) the shards keep a 2-3 % edge, at 720p
// (0.8 rounds) a large one (176 k against 115 k), and so they do for the gaussian kernel (BASELINE configs[1]: 56-58 k against 54 k).
inline bool policy_auto_prefers_clip(bool gaussian, int W, int H, int resident) {
    const int ntiles = policy_tiles(W, H, POLICY_TH);
    return !gaussian && ntiles >= resident && ntiles < 4 * resident;
}

}  // namespace crt
