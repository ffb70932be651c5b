// crt_abi.cu — implementation of include/crt_b200.h: context, tables, launch logic.
// Pure CUDA runtime; no torch, no CPU compute path.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/crt_b200.h"
#include "crt_derive.h"
#include "crt_fused_ps2.cuh"      // types and constants only: the kernels are instantiated in the crt_tu_*.cu units
#include "crt_fused_warp_ps2.cuh"
#include "crt_fused_warp_src.cuh"
#include "crt_fused_gauss_ps2.cuh"
#include "crt_gather_box.cuh"
#include "crt_launch.h"
#include "crt_policy.h"

using namespace crt;

namespace {

thread_local std::string g_create_error;

struct HostRing {                       // crt_process_host staging
    static constexpr int SLOTS = 3;
    uint8_t* d_in[SLOTS] = {nullptr, nullptr, nullptr};
    uint8_t* d_out[SLOTS] = {nullptr, nullptr, nullptr};
    cudaEvent_t in_ready[SLOTS] = {}, done[SLOTS] = {}, out_ready[SLOTS] = {};
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
    int chunk_frames = 0;
    bool ready = false;
};

}  // namespace

struct crt_ctx;
struct Shard {                          // one extra temporal shard of a clip: its own context, stream and buffers
    crt_ctx* kid = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    float* state = nullptr;             // [H][W][3] persistence state of the shard
    uint8_t* halo_out = nullptr;        // outputs of the warm-up frames (discarded)
    size_t halo_cap = 0;
    unsigned long long version = 0;     // configuration of the parent this kid was last synchronised with
};

struct crt_ctx {
    int device = 0, W = 0, H = 0;
    crt_params p{};
    bool have_params = false;
    void* tab[CRT_TABLE_COUNT] = {};
    size_t tab_bytes[CRT_TABLE_COUNT] = {};
    std::vector<int32_t> h_pix_x, h_pix_y;                 // host copies (tile planning, pix_uniform)
    std::vector<float> h_cols, h_fwd, h_inv;               // host copies of the triad tables (composite LUT)
    float* pow_tab = nullptr;                              // device tables of pow_unit
    float* comp_lut = nullptr;                             // device [2][1028] (16-byte aligned tables of 1025)
    std::vector<Lerp1> h_dn_x, h_dn_y, h_up_x, h_up_y;     // host copies of the fast-bloom coordinate tables
    Lerp1 *dn_x = nullptr, *dn_y = nullptr, *up_x = nullptr, *up_y = nullptr, *nz_x = nullptr, *nz_y = nullptr;
    int nz_grain = 0;
    Scratch scratch{nullptr, nullptr, nullptr};
    float* noise_buf = nullptr;
    int32_t* glitch_buf = nullptr;
    size_t glitch_cap = 0;
    float* state = nullptr;             // host-API persistence state
    int state_valid = 0;
    HostRing ring;
    Dev dev{};
    bool dev_ok = false;
    int policy = 0;
    LaunchEnv env;                      // per-context launch state (SM count, configured kernels): no process-wide statics
    CUtensorMap map_gather{};           // float32 [H][W*3] map of the state buffer with the gather kernel's 192 x 16 box
    const void* map_gather_ptr = nullptr;
    int* d_clip_sync = nullptr; int clip_sync_cap = 0;      // clip mode: item counter + per-tile frame counters
    std::vector<int> prof_frames;       // frames covered by each profile sample (clip mode: a launch covers many)
    int tile_h = 32;                    // tile height of the single-pass block kernels when a frame runs alone (choose_tile_h)
    Ps2Maps maps{};                     // tensor maps of the TMA-pipelined block kernel, valid for (maps_in, maps_frames, maps_st)
    const void* maps_in = nullptr; const void* maps_st = nullptr; int maps_frames = 0;
    CUtensorMap map_gin{};              // block gaussian kernel: the clip's even rows with a 256 x NBY(K) box
    const void* map_gin_ptr = nullptr; int map_gin_frames = 0, map_gin_k = 0;
    FusedPlan plan{};                   // single-pass fused kernel
    FusedPlan plan_q{};                 // two-pass: fused first pass without warp/glitch/text-after, then k_gather
    WarpPs2Plan plan_w{};               // single-pass warp on the block kernels (crt_fused_warp_ps2.cuh)
    WarpSrcPlan plan_s{};               // source-driven single-pass warp (crt_fused_warp_src.cuh)
    WsTile* d_ws_tiles = nullptr;       // device copy of plan_s.tiles
    GatherBoxPlan plan_g{};             // second pass with the footprint staged by TMA (crt_gather_box.cuh)
    int* d_origin = nullptr;            // device copy of plan_g.origin
    CUtensorMap map_gq{}, map_gst{};    // tensor maps of the pre-warp image (box of plan_g) and of the state (96 x 32 box)
    const void* map_gq_ptr = nullptr; const void* map_gst_ptr = nullptr; int map_gq_bw = 0, map_gq_bh = 0;
    Dev dev_q{};                        // parameter block of that first pass
    std::vector<uint8_t> h_tab[CRT_TABLE_COUNT];    // host copy of every table (replayed into the shard contexts)
    unsigned long long version = 1;     // bumped by crt_set_params / crt_set_table / crt_set_policy
    int shards_wanted = 1;              // crt_set_shards: 1 = off, 0 = auto, k > 1 = at most k
    std::vector<Shard> shards;          // shards 1 .. K-1 (shard 0 is this context on the caller's stream)
    cudaEvent_t ev_fork = nullptr;
    std::vector<cudaEvent_t> prof_ev;   // start/stop pairs (crt_profile_begin/end)
    int prof_cap = 0, prof_n = 0;
    int prof_every = 1, prof_tick = 0;  // one launch in prof_every is timed (crt_profile_sample_every)
    bool prof_on = false, prof_open = false;
    std::string err;
};

namespace {

int fail(crt_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, CRT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)

int upload(crt_ctx* ctx, Lerp1** dst, const std::vector<Lerp1>& v) {
    if (*dst) { cudaFree(*dst); *dst = nullptr; }
    CU(cudaMalloc((void**)dst, v.size() * sizeof(Lerp1)));
    CU(cudaMemcpy(*dst, v.data(), v.size() * sizeof(Lerp1), cudaMemcpyHostToDevice));
    return CRT_OK;
}

// Derive the device parameter block (crt_derive.h) from the context's parameters and tables.
int choose_tile_h(int W, int H, int sms, int per_sm, int halo_blocks);

int build_dev(crt_ctx* ctx) {
    const crt_params& p = ctx->p;
    if (p.noise_strength > 0.0 && p.grain_size > 1 && ctx->nz_grain != p.grain_size) {
        const int gh = ctx->H / p.grain_size > 1 ? ctx->H / p.grain_size : 1, gw = ctx->W / p.grain_size > 1 ? ctx->W / p.grain_size : 1;
        int rc = upload(ctx, &ctx->nz_x, linear_coords(ctx->W, gw)); if (rc) return rc;
        rc = upload(ctx, &ctx->nz_y, linear_coords(ctx->H, gh)); if (rc) return rc;
        ctx->nz_grain = p.grain_size;
    }
    TablePtrs t{};
    for (int i = 0; i < CRT_TABLE_COUNT; ++i) { t.tab[i] = ctx->tab[i]; t.bytes[i] = ctx->tab_bytes[i]; }
    // pixelate tables of the regular form ps * (i / ps) let the fused kernel grade each source pixel once
    t.pix_uniform = 0;
    if (p.pixel_size > 1 && (int)ctx->h_pix_x.size() == ctx->W && (int)ctx->h_pix_y.size() == ctx->H) {
        bool uni = true;
        for (int x = 0; x < ctx->W && uni; ++x) uni = ctx->h_pix_x[x] == p.pixel_size * (x / p.pixel_size);
        for (int y = 0; y < ctx->H && uni; ++y) uni = ctx->h_pix_y[y] == p.pixel_size * (y / p.pixel_size);
        t.pix_uniform = uni ? p.pixel_size : 0;
    }
    t.pow_tab = ctx->pow_tab;
    t.dn_x = ctx->dn_x; t.dn_y = ctx->dn_y; t.up_x = ctx->up_x; t.up_y = ctx->up_y; t.nz_x = ctx->nz_x; t.nz_y = ctx->nz_y;
    std::string err;
    int rc = derive_dev(p, ctx->W, ctx->H, t, &ctx->dev, &err);
    if (rc) return fail(ctx, rc, err);
    // composite triad tables for the fused kernels
    ctx->dev.triad_comp = nullptr;
    if (ctx->dev.triad_mode == 2 && (int)ctx->h_cols.size() == ctx->W * 3 && ctx->h_fwd.size() == 1025 && ctx->h_inv.size() == 1025) {
        std::vector<float> comp;
        int x0 = 0, x1 = -1;
        if (build_triad_comp(ctx->h_cols.data(), ctx->W, ctx->h_fwd.data(), ctx->h_inv.data(), &comp, &x0, &x1)) {
            // device layout [2][1028] so that both tables start on a 16-byte boundary
            if (!ctx->comp_lut) CU(cudaMalloc((void**)&ctx->comp_lut, 2 * 1028 * sizeof(float)));
            CU(cudaMemcpy(ctx->comp_lut, comp.data(), 1025 * sizeof(float), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(ctx->comp_lut + 1028, comp.data() + 1025, 1025 * sizeof(float), cudaMemcpyHostToDevice));
            ctx->dev.triad_comp = ctx->comp_lut; ctx->dev.comp_x0 = x0; ctx->dev.comp_x1 = x1;
        }
    }
    ctx->dev_ok = true;
    Dev hd = ctx->dev;                                   // same block with HOST coordinate tables, for planning
    hd.dn_x = ctx->h_dn_x.data(); hd.dn_y = ctx->h_dn_y.data(); hd.up_x = ctx->h_up_x.data(); hd.up_y = ctx->h_up_y.data();
    ctx->plan = plan_fused(hd, glitch_active(p));
    ctx->plan_q = FusedPlan{};
    ctx->tile_h = P2_TH;
    if (ctx->plan.ok && ctx->plan.ps2 && ctx->plan.gauss_k <= 13) {
        // measured (run 41, 1080p): default chain 21.5 -> 20.5 us and 58 954 -> 63 215 frames/s on one stream with 28 rows; the
        // gaussian kernel LOSES with any lower tile (28.4 us at 32 rows, 30.6 at 26, 30.3 at 28, 28.9 at 30: its per-tile fixed
        // cost — four barriers, the state tile's load behind the previous store — outweighs the partial round) -> 32 unless forced
        const bool thr = ctx->dev.bloom_mode == 1 && ctx->dev.thr_on;
        ctx->tile_h = ctx->plan.gauss_k ? choose_tile_h(ctx->W, ctx->H, ctx->env.sms, 3, -1)
                                        : choose_tile_h(ctx->W, ctx->H, ctx->env.sms, thr ? 3 : 4, 1);
    }
    // measured (round 2, run 3): the single-pass warp block kernel is slower than the two-pass path on BASELINE configs[2]
    // (152 vs 125 us at 4K: the aligned footprint of a 64 x 32 tile holds 1.6-1.8x its pixels) -> opt-in only
    ctx->plan_w = env_int("CRT_WARP_PS2", 0) ? plan_warp_ps2(hd, glitch_active(p)) : WarpPs2Plan{};
    ctx->plan_g = GatherBoxPlan{};
    ctx->plan_s = WarpSrcPlan{};
    // measured (round 2, runs 36-39): the source-driven single pass (crt_fused_warp_src.cuh) halves the DRAM traffic of the
    // two-pass path but runs 120.5 us against 106.8 us at 4K (ownership tests + cut quads + phase barriers) -> opt-in only
    if (ctx->dev.warp_on && env_int("CRT_WARP_SRC", 0) && !ctx->plan_w.ok) {
        ctx->plan_s = plan_warp_src(hd, glitch_active(p));
        if (ctx->plan_s.ok) {
            if (ctx->d_ws_tiles) { cudaFree(ctx->d_ws_tiles); ctx->d_ws_tiles = nullptr; }
            CU(cudaMalloc((void**)&ctx->d_ws_tiles, ctx->plan_s.tiles.size() * sizeof(WsTile)));
            CU(cudaMemcpy(ctx->d_ws_tiles, ctx->plan_s.tiles.data(), ctx->plan_s.tiles.size() * sizeof(WsTile), cudaMemcpyHostToDevice));
        }
    }
    const bool gather_needed = ctx->dev.warp_on || glitch_active(p) || ctx->dev.text_mode == 2;
    if (!ctx->plan.ok || gather_needed || env_int("CRT_TWO_PASS", 0)) {
        Dev hq = hd;
        hq.warp_on = 0; hq.text_mode = hd.text_mode == 1 ? 1 : 0;
        ctx->plan_q = plan_fused(hq, false);
        ctx->dev_q = ctx->dev;
        ctx->dev_q.warp_on = 0; ctx->dev_q.text_mode = hq.text_mode;
        // Prefer the two-pass path when its first pass is the pixel_size-2 block kernel: measured faster
        // than the single-pass footprint kernel on BASELINE configs[2] (6 090 vs 4 851 frames/s at 4K).
        const int pref = env_int("CRT_TWO_PASS", -1);
        if (ctx->plan.ok && ctx->plan_q.ok && ctx->dev.warp_on && (pref == 1 || (pref != 0 && ctx->plan_q.ps2))) ctx->plan.ok = false;
        if (ctx->plan.ok && ctx->plan_q.ok && !ctx->dev.warp_on && pref == 1) ctx->plan.ok = false;
        if (ctx->dev.warp_on && env_int("CRT_GATHER_BOX", 1)) {
            const GlitchGeom gg = glitch_geom(p, ctx->W, ctx->H);
            ctx->plan_g = plan_gather_box(hd, gg.rows > 0 ? gg.y0 : ctx->H + 1, gg.rows > 0 ? env_int("CRT_GATHER_SLACK", 12) : 0);
            if (ctx->plan_g.ok) {
                if (ctx->d_origin) { cudaFree(ctx->d_origin); ctx->d_origin = nullptr; }
                CU(cudaMalloc((void**)&ctx->d_origin, ctx->plan_g.origin.size() * sizeof(int)));
                CU(cudaMemcpy(ctx->d_origin, ctx->plan_g.origin.data(), ctx->plan_g.origin.size() * sizeof(int), cudaMemcpyHostToDevice));
            }
        }
    }
    return CRT_OK;
}

int ensure_scratch(crt_ctx* ctx) {
    const Dev& d = ctx->dev;
    const size_t px = (size_t)d.W * d.H;
    if (d.bloom_mode == 1 && !ctx->scratch.ds) CU(cudaMalloc((void**)&ctx->scratch.ds, (size_t)d.hw * d.hh * 3 * sizeof(float)));
    if (d.bloom_mode == 2 && !ctx->scratch.bl) CU(cudaMalloc((void**)&ctx->scratch.bl, px * 3 * sizeof(float)));
    if (d.warp_on && !ctx->scratch.q) CU(cudaMalloc((void**)&ctx->scratch.q, px * 3 * sizeof(float)));
    return CRT_OK;
}

// Key of the generated glitch pattern: the reference's own seed formula (gui crt_filter.py:670, export :841), so that the
// pattern changes exactly when the reference's does (the GUI holds a pattern for 20 phase-px, the export for 0.5).
uint64_t glitch_key(const crt_params& p, int W, int H, const crt_frame& fr) {
    const double k = p.variant == CRT_VARIANT_EXPORT ? 2.0 : 0.05;
    return (uint64_t)(((long long)(fabs(fr.phase_px) * k) + ((long long)W << 10) + ((long long)H << 1)) & 0xFFFFFFFFll);
}

int gen_glitch(crt_ctx* ctx, const crt_frame& fr, const GlitchGeom& g, int32_t* d_offs, cudaStream_t st) {
    if (g.rows <= 0) return CRT_OK;
    if (g.rows > 12000) return fail(ctx, CRT_ERR_UNSUPPORTED, "glitch band taller than 12000 rows");
    if (launch_glitch_gen(d_offs, g.rows, g.nseg, ctx->p.variant == CRT_VARIANT_EXPORT ? 1 : 0, (float)ctx->p.glitch_amp_px,
                          ctx->p.noise_seed ^ 0x9E3779B97F4A7C15ull, glitch_key(ctx->p, ctx->W, ctx->H, fr), st))
        return fail(ctx, CRT_ERR_CUDA, std::string("glitch generator launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    return CRT_OK;
}

// bench.py timing hook: events around ALL kernels of a frame (generators, first pass, gather), one frame in prof_every
void prof_mark(crt_ctx* ctx, cudaStream_t st, bool stop) {
    if (!ctx->prof_on || ctx->prof_n >= ctx->prof_cap) return;
    if (!stop) ctx->prof_open = (ctx->prof_tick++ % ctx->prof_every) == 0;
    if (!ctx->prof_open) return;
    cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n + (stop ? 1 : 0)], st);
    if (stop) { if ((int)ctx->prof_frames.size() <= ctx->prof_n) ctx->prof_frames.resize(ctx->prof_n + 1, 1); ctx->prof_frames[ctx->prof_n] = 1; ++ctx->prof_n; ctx->prof_open = false; }
}

// ... and around a clip-mode launch, which covers `frames` frames
void prof_mark_clip(crt_ctx* ctx, cudaStream_t st, bool stop, int frames) {
    if (!ctx->prof_on || ctx->prof_n >= ctx->prof_cap) return;
    cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n + (stop ? 1 : 0)], st);
    if (stop) { if ((int)ctx->prof_frames.size() <= ctx->prof_n) ctx->prof_frames.resize(ctx->prof_n + 1, 1); ctx->prof_frames[ctx->prof_n++] = frames; }
}

// One frame through the staged kernels.
int run_staged(crt_ctx* ctx, const FrameDev& f, const uint8_t* d_in, uint8_t* d_out, float* d_state, int has_prev, float* d_img,
               cudaStream_t st, int* launches) {
    int rc = ensure_scratch(ctx); if (rc) return rc;
    return launch_staged(ctx->env, ctx->dev, f, d_in, d_out, d_state, has_prev, d_img, ctx->scratch, st, launches);
}

// Tile height of the per-frame single-pass block kernels: policy_tile_h (crt_policy.h); CRT_TILE_H overrides.
int choose_tile_h(int W, int H, int sms, int per_sm, int halo_blocks) {
    return policy_tile_h(W, H, sms, per_sm, halo_blocks, env_int("CRT_TILE_H", 0));
}

// Tensor maps for k_fused_ps2_pipe: the clip as uint8 [frames * H/2 even rows][W*3] (box 256 x 18) and the
// persistence state as float32 [H][W*3] (box 192 x 32).  Returns false when the variant cannot be used.
bool prepare_ps2_maps(crt_ctx* ctx, const Dev& d, const uint8_t* d_in, int n_frames, const float* d_state, int th = P2_TH) {
    if (!d_state || !fused_ps2_pipe_supported(d)) return false;
    if (((uintptr_t)d_in & 15) || ((uintptr_t)d_state & 15)) return false;
    if (ctx->maps_in == d_in && ctx->maps_frames == n_frames && ctx->maps_st == d_state && ctx->maps.th == th) return true;
    ctx->maps_in = nullptr;
    const uint64_t W3 = (uint64_t)d.W * 3;
    // even rows of all frames as one 2-D tensor: the frame pitch W3 * H is (H / 2) row pitches of 2 * W3
    const uint64_t in_dims[2] = {W3, (uint64_t)(d.H / 2) * n_frames}, in_strides[1] = {2 * W3};
    const uint32_t in_box[2] = {(uint32_t)P2_RAW_W, (uint32_t)P2_BH};
    if (!tma_encode(&ctx->maps.in, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_in, in_dims, in_strides, in_box)) return false;
    const uint64_t st_dims[2] = {W3, (uint64_t)d.H}, st_strides[1] = {W3 * 4};
    const uint32_t st_box[2] = {(uint32_t)P2_TW * 3, (uint32_t)th};
    if (!tma_encode(&ctx->maps.st, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_state, st_dims, st_strides, st_box)) return false;
    ctx->maps_in = d_in; ctx->maps_frames = n_frames; ctx->maps_st = d_state; ctx->maps.th = th;
    return true;
}

// ---- one call = prepare (validation, path selection, tensor maps) + one launch_frame per frame ----------------------
struct Call {
    bool want_warp = false, want_warp_src = false, want_fused = false, want_two_pass = false, pipe = false, gather_box = false, persist = false;
    const CUtensorMap* gather_map = nullptr;
    const CUtensorMap* gauss_in = nullptr;      // input-tile map of the block gaussian kernel, when usable
    GlitchGeom gg{};
    size_t frame_px = 0;
    int fused_used = 0;
};

// d_in / n_frames describe the whole clip the frames of this call are taken from (the TMA-pipelined kernels address the
// clip as ONE tensor: a frame is selected by its index), d_state the state buffer THIS context blends against.
int prepare_call(crt_ctx* ctx, const uint8_t* d_in, int n_frames, const uint8_t* d_out, float* d_state, const float* d_img, Call* c,
                 bool concurrent = false, int tile_h = 0) {
    if (!ctx->have_params) return fail(ctx, CRT_ERR_INVALID, "crt_set_params has not been called");
    if (!ctx->dev_ok) { int rc = build_dev(ctx); if (rc) return rc; }
    if (n_frames < 0 || (n_frames > 0 && (!d_in || (!d_out && !d_img)))) return fail(ctx, CRT_ERR_INVALID, "null buffer");
    const crt_params& p = ctx->p;
    const Dev& d = ctx->dev;
    c->persist = p.persistence > 0.0 && !d_img;
    if (c->persist && !d_state) return fail(ctx, CRT_ERR_INVALID, "persistence > 0 needs a state buffer");
    // the kernels move the state with 16-byte and the pixels with 4-byte accesses (device allocations are aligned far beyond that)
    // (frames whose width is not a multiple of 4 take scalar accesses and have no such requirement)
    if ((d.W & 3) == 0 && (((uintptr_t)d_state & 15) || ((uintptr_t)d_out & 3) || ((uintptr_t)d_img & 15)))
        return fail(ctx, CRT_ERR_INVALID, "the state / float image buffer must be 16-byte aligned, the output clip 4-byte aligned");
    CU(cudaSetDevice(ctx->device));
    c->frame_px = (size_t)d.W * d.H;
    c->gg = glitch_geom(p, d.W, d.H);
    // crt_process_static (d_img) takes the same kernels: the float image leaves through the path the pre-warp image of the
    // two-pass path takes (q_out), or is what the gather writes as "state" when no previous state is blended in
    c->want_warp = ctx->policy != 1 && ctx->plan_w.ok && !d_img;          // single-pass warp block kernel (output-driven, opt-in)
    // source-driven single-pass warp: needs the clip's tensor map (16-byte aligned clip) and, for its 16-byte state accesses, an aligned state
    c->want_warp_src = ctx->policy != 1 && !c->want_warp && ctx->plan_s.ok && !d_img && d_out && d_state && !((uintptr_t)d_in & 15) &&
                       prepare_ps2_maps(ctx, d, d_in, n_frames, d_state);
    c->want_fused = ctx->policy != 1 && !c->want_warp && !c->want_warp_src && ctx->plan.ok;
    c->want_two_pass = ctx->policy != 1 && !c->want_warp && !c->want_warp_src && !c->want_fused && ctx->plan_q.ok;
    if (ctx->policy == 2 && !c->want_warp && !c->want_warp_src && !c->want_fused && !c->want_two_pass)
        return fail(ctx, CRT_ERR_UNSUPPORTED, std::string("fused kernel not available: ") + ctx->plan.why);
    c->fused_used = (c->want_warp || c->want_warp_src || c->want_fused) ? 1 : c->want_two_pass ? 2 : 0;
    if (c->want_two_pass && !ctx->scratch.q) CU(cudaMalloc((void**)&ctx->scratch.q, c->frame_px * 3 * sizeof(float)));
    // second pass of the two-pass path with the state moved by TMA: needs a tensor map of the state buffer (192 x 16 box)
    static const bool use_gather_tile = env_int("CRT_GATHER_TILE", 1) != 0;
    c->gather_map = nullptr;
    if (c->want_two_pass && use_gather_tile && !d_img && d_state && !((uintptr_t)d_state & 15) && !(d.W & 3)) {
        if (ctx->map_gather_ptr != d_state) {
            const uint64_t W3 = (uint64_t)d.W * 3;
            const uint64_t dims[2] = {W3, (uint64_t)d.H}, strides[1] = {W3 * 4};
            const uint32_t box[2] = {(uint32_t)FTW * 3, (uint32_t)GATHER_TH};
            if (tma_encode(&ctx->map_gather, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_state, dims, strides, box)) ctx->map_gather_ptr = d_state;
        }
        if (ctx->map_gather_ptr == d_state) c->gather_map = &ctx->map_gather;
    }
    // second pass with the footprint staged by TMA: tensor maps of the pre-warp image (box from the plan) and of the state
    c->gather_box = false;
    if (c->want_two_pass && ctx->plan_g.ok && !d_img && d_state && !((uintptr_t)d_state & 15)) {
        const uint64_t W3 = (uint64_t)d.W * 3;
        const uint64_t dims[2] = {W3, (uint64_t)d.H}, strides[1] = {W3 * 4};
        bool ok = true;
        if (ctx->map_gq_ptr != ctx->scratch.q || ctx->map_gq_bw != ctx->plan_g.bw || ctx->map_gq_bh != ctx->plan_g.bh) {
            const uint32_t box[2] = {(uint32_t)ctx->plan_g.bw * 3, (uint32_t)ctx->plan_g.bh};
            ok = tma_encode(&ctx->map_gq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ctx->scratch.q, dims, strides, box);
            if (ok) { ctx->map_gq_ptr = ctx->scratch.q; ctx->map_gq_bw = ctx->plan_g.bw; ctx->map_gq_bh = ctx->plan_g.bh; }
        }
        if (ok && ctx->map_gst_ptr != d_state) {
            const uint32_t box[2] = {(uint32_t)GB_TW * 3, (uint32_t)GB_TH};
            ok = tma_encode(&ctx->map_gst, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_state, dims, strides, box);
            if (ok) ctx->map_gst_ptr = d_state;
        }
        c->gather_box = ok;
    }
    // TMA-pipelined block kernels: tensor maps of the clip and of this context's state buffer (or pre-warp image)
    // (single pass, frames one after the other on the GPU: tile height matched to the number of resident CTAs)
    const int th = (c->want_fused && tile_h) ? tile_h : (c->want_fused && !concurrent) ? ctx->tile_h : P2_TH;
    c->pipe = (c->want_fused && ctx->plan.ps2 && c->persist && prepare_ps2_maps(ctx, d, d_in, n_frames, d_state, th)) ||
              (c->want_two_pass && ctx->plan_q.ps2 && prepare_ps2_maps(ctx, ctx->dev_q, d_in, n_frames, ctx->scratch.q));
    // block gaussian kernel: input tile by TMA (a K-specific box over the clip's even rows)
    c->gauss_in = nullptr;
    if (c->pipe && c->want_fused && ctx->plan.gauss_k && fused_gauss_ps2_tin_ok(d, ctx->plan.gauss_k) && !((uintptr_t)d_in & 15)) {
        const int K = ctx->plan.gauss_k;
        if (ctx->map_gin_ptr != d_in || ctx->map_gin_frames != n_frames || ctx->map_gin_k != K) {
            const uint64_t W3 = (uint64_t)d.W * 3;
            const uint64_t dims[2] = {W3, (uint64_t)(d.H / 2) * n_frames}, strides[1] = {2 * W3};
            const uint32_t box[2] = {(uint32_t)P2_RAW_W, (uint32_t)gauss_ps2_geo(K).nby};
            if (tma_encode(&ctx->map_gin, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_in, dims, strides, box)) {
                ctx->map_gin_ptr = d_in; ctx->map_gin_frames = n_frames; ctx->map_gin_k = K;
            } else ctx->map_gin_ptr = nullptr;
        }
        if (ctx->map_gin_ptr == d_in) c->gauss_in = &ctx->map_gin;
    }
    return CRT_OK;
}

// Frame `index` of the clip (d_in + index frames) -> out_i (uint8) / img_i (float image); blends against d_state when has_prev.
int launch_frame(crt_ctx* ctx, const Call& c, int index, const uint8_t* in_i, uint8_t* out_i, float* img_i, float* d_state, int has_prev,
                 bool pdl, const crt_frame& fr, cudaStream_t st, int* launches_out) {
    const crt_params& p = ctx->p;
    const Dev& d = ctx->dev;
    const GlitchGeom& gg = c.gg;
    int launches = 0;
    FrameDev f = derive_frame(p, fr);
    prof_mark(ctx, st, false);          // whole frame: generators + every pass
    if (d.noise_on) {
        f.noise = fr.d_noise;
        if (!f.noise) {
            if (p.noise_mode != 1) return fail(ctx, CRT_ERR_INVALID, "noise_strength > 0 with noise_mode 0 needs crt_frame.d_noise");
            if (!ctx->noise_buf) CU(cudaMalloc((void**)&ctx->noise_buf, c.frame_px * sizeof(float)));
            if (launch_noise_gen(ctx->noise_buf, d.gh * d.gw, p.noise_seed, fr.frame_index, st)) return fail(ctx, CRT_ERR_CUDA, "noise generator launch failed");
            ++launches;
            f.noise = ctx->noise_buf;
        }
    }
    if (gg.rows > 0) {
        if (fr.d_glitch_offs) {
            if (fr.glitch_rows != gg.rows || fr.glitch_y0 != gg.y0 || fr.glitch_seg_len <= 0 ||
                fr.glitch_segments != (d.W + fr.glitch_seg_len - 1) / fr.glitch_seg_len)
                return fail(ctx, CRT_ERR_INVALID, "injected glitch table geometry does not match the parameters");
            f.goffs = fr.d_glitch_offs; f.gy0 = fr.glitch_y0; f.gseg = fr.glitch_seg_len; f.gnseg = fr.glitch_segments;
        } else {
            if (p.glitch_mode != 1) return fail(ctx, CRT_ERR_INVALID, "glitch on with glitch_mode 0 needs crt_frame.d_glitch_offs");
            size_t need = (size_t)gg.rows * gg.nseg;
            if (need > ctx->glitch_cap) {
                if (ctx->glitch_buf) cudaFree(ctx->glitch_buf);
                ctx->glitch_buf = nullptr; ctx->glitch_cap = 0;
                CU(cudaMalloc((void**)&ctx->glitch_buf, need * sizeof(int32_t)));
                ctx->glitch_cap = need;
            }
            int rc = gen_glitch(ctx, fr, gg, ctx->glitch_buf, st); if (rc) return rc;
            ++launches;
            f.goffs = ctx->glitch_buf; f.gy0 = gg.y0; f.gseg = gg.seg_len; f.gnseg = gg.nseg;
        }
    }
    float* state_i = (c.persist || (d_state && !img_i)) ? d_state : nullptr;
    float* q_i = nullptr;               // single-pass kernels: where the float image goes instead of out / state
    if (img_i) { if (c.want_two_pass) state_i = img_i; else q_i = img_i; }
    ctx->maps.frame = index;
    int rc;
    if (c.want_warp) {
        rc = launch_warp_ps2(ctx->env, ctx->plan_w, d, f, in_i, out_i, state_i, has_prev, st, &launches, pdl);
    } else if (c.want_warp_src) {
        rc = launch_warp_src(ctx->env, d, f, in_i, out_i, state_i, has_prev, st, &launches, pdl, ctx->d_ws_tiles, (int)ctx->plan_s.tiles.size(),
                             ctx->plan_s.ntx, ctx->plan_s.nty, &ctx->maps);
    } else if (c.want_fused) {
        const Ps2Maps* maps = c.pipe ? &ctx->maps : nullptr;
        rc = (ctx->plan.ps2 && ctx->plan.gauss_k) ? launch_fused_gauss_ps2(ctx->env, d, f, in_i, out_i, state_i, q_i, has_prev, st, &launches, pdl, maps, c.gauss_in)
           : ctx->plan.ps2 ? launch_fused_ps2(ctx->env, d, f, in_i, out_i, state_i, q_i, has_prev, st, &launches, pdl, maps)
           : ctx->plan.gauss_k ? launch_fused_gauss(ctx->env, ctx->plan.th, ctx->plan.nt, d, f, in_i, out_i, state_i, q_i, has_prev, st, &launches)
                               : launch_fused(ctx->env, ctx->plan, d, f, in_i, out_i, state_i, q_i, has_prev, st, &launches);
    } else if (c.want_two_pass) {
        const FusedPlan& pq = ctx->plan_q;
        const Ps2Maps* maps = c.pipe ? &ctx->maps : nullptr;
        rc = (pq.ps2 && pq.gauss_k) ? launch_fused_gauss_ps2(ctx->env, ctx->dev_q, f, in_i, nullptr, nullptr, ctx->scratch.q, 0, st, &launches, pdl, maps, nullptr)
           : pq.ps2 ? launch_fused_ps2(ctx->env, ctx->dev_q, f, in_i, nullptr, nullptr, ctx->scratch.q, 0, st, &launches, pdl, maps)
           : pq.gauss_k ? launch_fused_gauss(ctx->env, pq.th, pq.nt, ctx->dev_q, f, in_i, nullptr, nullptr, ctx->scratch.q, 0, st, &launches)
                        : launch_fused(ctx->env, pq, ctx->dev_q, f, in_i, nullptr, nullptr, ctx->scratch.q, 0, st, &launches);
        if (!rc) rc = c.gather_box ? launch_gather_box(ctx->env, d, f, ctx->scratch.q, out_i, has_prev, st, &launches, ctx->d_origin, ctx->plan_g.bw,
                                                       ctx->plan_g.bh, ctx->plan_g.smem, &ctx->map_gq, &ctx->map_gst)
                                   : launch_gather(ctx->env, d, f, ctx->scratch.q, out_i, state_i, has_prev, st, &launches, c.gather_map);
    }
    else rc = run_staged(ctx, f, in_i, out_i, state_i, has_prev, img_i, st, &launches);
    prof_mark(ctx, st, true);
    *launches_out += launches;
    if (rc == CRT_ERR_CUDA) return fail(ctx, rc, std::string("kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    return rc;
}

// Clip mode applies to the single-pass block kernel with fast bloom / no bloom whose frames need nothing generated per frame
// (no noise plane, no glitch table) — the CLI's default chain among them; the caller's state must be blended (persistence > 0).
int clip_resident(const crt_ctx* ctx) {
    return ctx->env.sms * ((ctx->plan.gauss_k || (ctx->dev.bloom_mode == 1 && ctx->dev.thr_on)) ? 3 : 4);
}
// Tile height of clip-mode launches: policy_clip_tile_h (crt_policy.h); CRT_CLIP_TH overrides.
int clip_tile_h(const crt_ctx* ctx) { return policy_clip_tile_h(ctx->W, ctx->H, clip_resident(ctx), env_int("CRT_CLIP_TH", 0)); }

bool clip_wanted(crt_ctx* ctx, const uint8_t* d_out, const float* d_state, const float* d_img, int n_frames) {
    const bool use_clip = env_int("CRT_CLIP", 1) != 0;      // read per call: tests switch it inside one process
    if (!use_clip || !ctx || !ctx->have_params || n_frames < 2 || !d_out || !d_state || d_img || ctx->policy == 1) return false;
    if (!ctx->dev_ok && build_dev(ctx)) return false;
    const crt_params& p = ctx->p;
    if (!(ctx->plan.ok && ctx->plan.ps2 && p.persistence > 0.0 && !ctx->dev.noise_on && !glitch_active(p) && !ctx->plan_w.ok && !ctx->plan_s.ok))
        return false;
    if (ctx->plan.gauss_k && !fused_gauss_ps2_clip_supported(ctx->dev, ctx->plan.gauss_k)) return false;
    // (a tile's frames are a serial chain: only frames with a tile per resident CTA — policy_clip_size_ok, crt_policy.h)
    return policy_clip_size_ok(ctx->W, ctx->H, clip_resident(ctx), env_int("CRT_CLIP_TH", 0), env_int("CRT_CLIP_MIN_TILES", 0));
}

int process_impl(crt_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, float* d_state, int state_valid, float* d_img,
                 const crt_frame* frames, int n_frames, cudaStream_t st, crt_launch_info* info) {
    if (!ctx) return CRT_ERR_INVALID;
    if (n_frames > 0 && !frames) return fail(ctx, CRT_ERR_INVALID, "null buffer");
    Call c;
    const bool try_clip = clip_wanted(ctx, d_out, d_state, d_img, n_frames);
    int rc = prepare_call(ctx, d_in, n_frames, d_out, d_state, d_img, &c, try_clip, try_clip ? clip_tile_h(ctx) : 0); if (rc) return rc;
    static const bool use_pdl = env_int("CRT_PDL", 1) != 0;
    const size_t fb = c.frame_px * 3;
    int launches = 0, i = 0, clip_frames = 0;
    if (try_clip && c.pipe && c.want_fused && c.persist && (!ctx->plan.gauss_k || c.gauss_in)) {
        // Clip mode (crt_fused_ps2.cuh, ClipArgs): runs of up to clip_max_frames() frames in ONE launch each, the frames of a
        // run chained tile by tile through the persistence state.  A first frame without a valid state goes alone (no blend).
        if (!state_valid) {
            rc = launch_frame(ctx, c, 0, d_in, d_out, nullptr, d_state, 0, false, frames[0], st, &launches);
            if (rc) return rc;
            i = 1;
        }
        const Dev& d = ctx->dev;
        const int ntiles = ((d.W + P2_TW - 1) / P2_TW) * ((d.H + ctx->maps.th - 1) / ctx->maps.th);
        if (ctx->clip_sync_cap < 1 + ntiles) {
            if (ctx->d_clip_sync) cudaFree(ctx->d_clip_sync);
            ctx->d_clip_sync = nullptr; ctx->clip_sync_cap = 0;
            CU(cudaMalloc((void**)&ctx->d_clip_sync, (size_t)(1 + ntiles) * sizeof(int)));
            ctx->clip_sync_cap = 1 + ntiles;
        }
        const int max_run = clip_max_frames();
        std::vector<FrameVar> fv;
        while (n_frames - i >= 2) {
            const int nf = n_frames - i < max_run ? n_frames - i : max_run;
            fv.resize(nf);
            FrameDev f0{};
            for (int j = 0; j < nf; ++j) {
                const FrameDev f = derive_frame(ctx->p, frames[i + j]);
                if (j == 0) f0 = f;
                fv[j].phase32 = f.phase32; fv[j].phase = f.phase; fv[j].flicker = f.flicker;
            }
            prof_mark_clip(ctx, st, false, nf);
            CU(cudaMemsetAsync(ctx->d_clip_sync, 0, (size_t)(1 + ntiles) * sizeof(int), st));
            ctx->maps.frame = i;
            rc = ctx->plan.gauss_k ? launch_fused_gauss_ps2_clip(ctx->env, d, f0, d_in + (size_t)i * fb, d_out + (size_t)i * fb, d_state, st, &launches,
                                                                 &ctx->maps, c.gauss_in, nf, fv.data(), ctx->d_clip_sync)
                                   : launch_fused_ps2_clip(ctx->env, d, f0, d_in + (size_t)i * fb, d_out + (size_t)i * fb, d_state, st, &launches, &ctx->maps,
                                                           nf, fv.data(), ctx->d_clip_sync);
            prof_mark_clip(ctx, st, true, rc ? 0 : nf);
            if (rc == 4) break;          // this parameter set has no clip-mode kernel after all: the frames go one launch each (below)
            if (rc) return fail(ctx, CRT_ERR_CUDA, std::string("clip-mode launch failed: ") + cudaGetErrorString(cudaGetLastError()));
            i += nf; clip_frames += nf;
        }
    }
    for (; i < n_frames; ++i) {
        // Frames after the first may overlap the previous frame's kernel tail (launch_pdl, crt_fused.cuh).  Never frame 0:
        // its input may come from the caller's immediately preceding kernel.
        rc = launch_frame(ctx, c, i, d_in + (size_t)i * fb, d_out ? d_out + (size_t)i * fb : nullptr, d_img ? d_img + (size_t)i * fb : nullptr, d_state,
                          c.persist && (state_valid || i > 0), i > 0 && use_pdl && clip_frames == 0, frames[i], st, &launches);
        if (rc) return rc;
    }
    if (info) info->reserved[2] = clip_frames;
    if (info) { info->kernels_launched = launches; info->fused = c.fused_used; }
    return CRT_OK;
}

// ---- intra-GPU temporal shards ------------------------------------------------------------------------------------------
// One stream of frame kernels does not fill a B200: the kernels are latency bound and every frame ends in a
// partial wave.  The chain's only cross-frame dependency is the persistence recurrence s_t = clip(p s_{t-1} + (1-p) x_t)
// (crt_filter.py:1092), whose memory of the past decays as p^k, so a clip can be cut into contiguous temporal shards exactly
// as it is across GPUs (pythoncrt_b200/clip.py): shard k > 0 starts `halo` frames early from an empty state,
// discards those outputs and then matches the serial run to within p^halo <= 1/2040 (an eighth of an LSB).  Here the shards
// run CONCURRENTLY on one GPU, each on its own stream with its own context (scratch buffers, tensor maps), launches
// interleaved frame by frame.  Measured (round 2, profiles/tools/concurrency_probe.py): default chain at 4K 18.7k -> 21.9k
// frames/s with 3 shards, 1080p gaussian chain 40.3k -> 55.7k, VGA 113k -> 185k with 4.
int halo_frames(double persistence) { return policy_halo_frames(persistence); }

int choose_shards(const crt_ctx* ctx, int n_frames) {
    static const int forced = env_int("CRT_SHARDS", -1);
    return policy_shards(ctx->shards_wanted, forced, n_frames, ctx->p.persistence, ctx->W, ctx->H);
}

int sync_shards(crt_ctx* ctx, int K) {
    if (!ctx->ev_fork) CU(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    while ((int)ctx->shards.size() < K - 1) {
        Shard s;
        int rc = crt_create(ctx->device, ctx->W, ctx->H, &s.kid);
        if (rc) return fail(ctx, rc, std::string("shard context: ") + crt_last_error(nullptr));
        CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        ctx->shards.push_back(s);
    }
    for (int k = 0; k < K - 1; ++k) {
        Shard& s = ctx->shards[k];
        if (s.version == ctx->version) continue;
        for (int t = 0; t < CRT_TABLE_COUNT; ++t) {
            int rc = crt_set_table(s.kid, t, ctx->h_tab[t].empty() ? nullptr : ctx->h_tab[t].data(), ctx->h_tab[t].size());
            if (rc) return fail(ctx, rc, std::string("shard table: ") + crt_last_error(s.kid));
        }
        int rc = crt_set_params(s.kid, &ctx->p);
        if (!rc) rc = crt_set_policy(s.kid, ctx->policy);
        if (rc) return fail(ctx, rc, std::string("shard parameters: ") + crt_last_error(s.kid));
        s.version = ctx->version;
    }
    return CRT_OK;
}

int process_sharded(crt_ctx* ctx, int K, const uint8_t* d_in, uint8_t* d_out, float* d_state, int state_valid, const crt_frame* frames,
                    int n_frames, cudaStream_t st, crt_launch_info* info) {
    int rc = sync_shards(ctx, K); if (rc) return rc;
    const int halo = halo_frames(ctx->p.persistence);
    const size_t fb = (size_t)ctx->W * ctx->H * 3;
    struct Part { crt_ctx* c; cudaStream_t st; float* state; uint8_t* halo_out; int warm, a, b; Call call; };
    std::vector<Part> parts(K);
    const int base = n_frames / K, extra = n_frames % K;
    int longest = 0;
    for (int k = 0; k < K; ++k) {
        Part& pt = parts[k];
        pt.a = k * base + (k < extra ? k : extra);
        pt.b = pt.a + base + (k < extra ? 1 : 0);
        pt.warm = k == 0 ? 0 : pt.a - halo;                   // >= 0: a shard is at least 8 halos long (choose_shards)
        if (k == 0) { pt.c = ctx; pt.st = st; pt.state = d_state; pt.halo_out = nullptr; }
        else {
            Shard& s = ctx->shards[k - 1];
            if (!s.state) CU(cudaMalloc((void**)&s.state, fb * sizeof(float)));
            if (halo > 0 && s.halo_cap < (size_t)halo * fb) {
                if (s.halo_out) cudaFree(s.halo_out);
                s.halo_out = nullptr; s.halo_cap = 0;
                CU(cudaMalloc((void**)&s.halo_out, (size_t)halo * fb));
                s.halo_cap = (size_t)halo * fb;
            }
            pt.c = s.kid; pt.st = s.stream; pt.state = s.state; pt.halo_out = s.halo_out;
        }
        rc = prepare_call(pt.c, d_in, n_frames, d_out, pt.state, nullptr, &pt.call, true);
        if (rc) return k == 0 ? rc : fail(ctx, rc, std::string("shard: ") + crt_last_error(pt.c));
        if (pt.b - pt.warm > longest) longest = pt.b - pt.warm;
    }
    // fork: the shards' streams start after whatever the caller's stream has queued so far (the clip may come from a kernel there)
    CU(cudaEventRecord(ctx->ev_fork, st));
    for (int k = 1; k < K; ++k) CU(cudaStreamWaitEvent(parts[k].st, ctx->ev_fork, 0));
    static const bool use_pdl = env_int("CRT_PDL", 1) != 0;
    int launches = 0;
    for (int j = 0; j < longest; ++j)                       // launches interleaved frame by frame so that every stream stays fed
        for (int k = 0; k < K; ++k) {
            Part& pt = parts[k];
            const int i = pt.warm + j;
            if (i >= pt.b) continue;
            uint8_t* out_i = i < pt.a ? pt.halo_out + (size_t)(i - pt.warm) * fb : d_out + (size_t)i * fb;
            const int has_prev = pt.call.persist && ((k == 0 && state_valid) || j > 0);
            rc = launch_frame(pt.c, pt.call, i, d_in + (size_t)i * fb, out_i, nullptr, pt.state, has_prev, j > 0 && use_pdl, frames[i], pt.st, &launches);
            if (rc) return k == 0 ? rc : fail(ctx, rc, std::string("shard: ") + crt_last_error(pt.c));
        }
    // join, then hand the caller the state after the LAST frame of the clip
    for (int k = 1; k < K; ++k) {
        CU(cudaEventRecord(ctx->shards[k - 1].done, parts[k].st));
        CU(cudaStreamWaitEvent(st, ctx->shards[k - 1].done, 0));
    }
    if (d_state) CU(cudaMemcpyAsync(d_state, parts[K - 1].state, fb * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (info) { info->kernels_launched = launches; info->fused = parts[0].call.fused_used; info->reserved[0] = K; info->reserved[1] = halo; }
    return CRT_OK;
}

}  // namespace

extern "C" {

int crt_abi_version(void) { return CRT_B200_ABI_VERSION; }

const char* crt_last_error(const crt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int crt_create(int device, int width, int height, crt_ctx** out_ctx) {
    crt_ctx* ctx = nullptr;   // for the CU macro
    if (!out_ctx || width < 2 || height < 2 || width > 32768 || height > 32768) return fail(nullptr, CRT_ERR_INVALID, "bad size");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return fail(nullptr, CRT_ERR_NO_DEVICE, "no CUDA device: the CRT chain has no CPU path");
    if (device < 0 || device >= n) return fail(nullptr, CRT_ERR_INVALID, "device index out of range");
    CU(cudaSetDevice(device));
    crt_ctx* c = new (std::nothrow) crt_ctx();
    if (!c) return fail(nullptr, CRT_ERR_INVALID, "out of host memory");
    c->device = device; c->W = width; c->H = height;
    c->env.device = device;
    if (cudaDeviceGetAttribute(&c->env.sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || c->env.sms < 1) c->env.sms = 148;
    ctx = c;
    const int hw = width / 2 > 1 ? width / 2 : 1, hh = height / 2 > 1 ? height / 2 : 1;
    c->h_dn_x = linear_coords(hw, width); c->h_dn_y = linear_coords(hh, height);
    c->h_up_x = linear_coords(width, hw); c->h_up_y = linear_coords(height, hh);
    int rc = upload(c, &c->dn_x, c->h_dn_x);
    if (!rc) rc = upload(c, &c->dn_y, c->h_dn_y);
    if (!rc) rc = upload(c, &c->up_x, c->h_up_x);
    if (!rc) rc = upload(c, &c->up_y, c->h_up_y);
    if (!rc) {
        float tab[POW_TAB_FLOATS];
        fill_pow_table(tab);
        if (cudaMalloc((void**)&c->pow_tab, sizeof(tab)) != cudaSuccess || cudaMemcpy(c->pow_tab, tab, sizeof(tab), cudaMemcpyHostToDevice) != cudaSuccess)
            rc = fail(c, CRT_ERR_CUDA, "pow table upload failed");
    }
    if (rc) { g_create_error = c->err; crt_destroy(c); return rc; }
    *out_ctx = c;
    return CRT_OK;
}

int crt_destroy(crt_ctx* ctx) {
    if (!ctx) return CRT_OK;
    cudaSetDevice(ctx->device);
    for (auto& t : ctx->tab) if (t) cudaFree(t);
    for (Lerp1* t : {ctx->dn_x, ctx->dn_y, ctx->up_x, ctx->up_y, ctx->nz_x, ctx->nz_y}) if (t) cudaFree(t);
    if (ctx->scratch.ds) cudaFree(ctx->scratch.ds);
    if (ctx->scratch.bl) cudaFree(ctx->scratch.bl);
    if (ctx->scratch.q) cudaFree(ctx->scratch.q);
    if (ctx->noise_buf) cudaFree(ctx->noise_buf);
    if (ctx->glitch_buf) cudaFree(ctx->glitch_buf);
    if (ctx->state) cudaFree(ctx->state);
    if (ctx->comp_lut) cudaFree(ctx->comp_lut);
    if (ctx->d_origin) cudaFree(ctx->d_origin);
    if (ctx->d_ws_tiles) cudaFree(ctx->d_ws_tiles);
    if (ctx->d_clip_sync) cudaFree(ctx->d_clip_sync);
    for (Shard& sh : ctx->shards) {
        if (sh.stream) { cudaStreamSynchronize(sh.stream); cudaStreamDestroy(sh.stream); }
        if (sh.done) cudaEventDestroy(sh.done);
        if (sh.state) cudaFree(sh.state);
        if (sh.halo_out) cudaFree(sh.halo_out);
        crt_destroy(sh.kid);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->pow_tab) cudaFree(ctx->pow_tab);
    for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
    HostRing& r = ctx->ring;
    for (int s = 0; s < HostRing::SLOTS; ++s) {
        if (r.d_in[s]) cudaFree(r.d_in[s]);
        if (r.d_out[s]) cudaFree(r.d_out[s]);
        if (r.in_ready[s]) cudaEventDestroy(r.in_ready[s]);
        if (r.done[s]) cudaEventDestroy(r.done[s]);
        if (r.out_ready[s]) cudaEventDestroy(r.out_ready[s]);
    }
    if (r.s_in) cudaStreamDestroy(r.s_in);
    if (r.s_comp) cudaStreamDestroy(r.s_comp);
    if (r.s_out) cudaStreamDestroy(r.s_out);
    delete ctx;
    return CRT_OK;
}

int crt_set_params(crt_ctx* ctx, const crt_params* params) {
    if (!ctx || !params) return CRT_ERR_INVALID;
    if (params->pixel_size < 1 || params->grain_size < 0 || params->text_mode < 0 || params->text_mode > 2 || params->vignette_on < 0 ||
        params->vignette_on > 2 || !(params->persistence >= 0.0 && params->persistence < 1.0) || params->channel_order < 0 ||
        params->channel_order > 1)
        return fail(ctx, CRT_ERR_INVALID, "parameter out of range");
    ctx->p = *params;
    ctx->have_params = true;
    ctx->dev_ok = false;
    ++ctx->version;
    return CRT_OK;
}

int crt_set_table(crt_ctx* ctx, int table, const void* h_data, size_t bytes) {
    if (!ctx || table < 0 || table >= CRT_TABLE_COUNT) return CRT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (ctx->tab[table] && ctx->tab_bytes[table] != bytes) { CU(cudaDeviceSynchronize()); cudaFree(ctx->tab[table]); ctx->tab[table] = nullptr; ctx->tab_bytes[table] = 0; }
    ++ctx->version;
    if (!h_data || bytes == 0) {
        if (ctx->tab[table]) { CU(cudaDeviceSynchronize()); cudaFree(ctx->tab[table]); }
        ctx->tab[table] = nullptr; ctx->tab_bytes[table] = 0; ctx->dev_ok = false;
        ctx->h_tab[table].clear();
        return CRT_OK;
    }
    ctx->h_tab[table].assign((const uint8_t*)h_data, (const uint8_t*)h_data + bytes);
    if (!ctx->tab[table]) CU(cudaMalloc(&ctx->tab[table], bytes));
    CU(cudaMemcpy(ctx->tab[table], h_data, bytes, cudaMemcpyHostToDevice));
    ctx->tab_bytes[table] = bytes;
    if (table == CRT_TABLE_TRIAD_COLS) ctx->h_cols.assign((const float*)h_data, (const float*)h_data + bytes / 4);
    if (table == CRT_TABLE_LUT_FWD) ctx->h_fwd.assign((const float*)h_data, (const float*)h_data + bytes / 4);
    if (table == CRT_TABLE_LUT_INV) ctx->h_inv.assign((const float*)h_data, (const float*)h_data + bytes / 4);
    if (table == CRT_TABLE_PIXELATE_X) ctx->h_pix_x.assign((const int32_t*)h_data, (const int32_t*)h_data + bytes / 4);
    if (table == CRT_TABLE_PIXELATE_Y) ctx->h_pix_y.assign((const int32_t*)h_data, (const int32_t*)h_data + bytes / 4);
    ctx->dev_ok = false;
    return CRT_OK;
}

int crt_set_policy(crt_ctx* ctx, int policy) {
    if (!ctx || policy < 0 || policy > 2) return CRT_ERR_INVALID;
    ctx->policy = policy;
    ++ctx->version;
    return CRT_OK;
}

int crt_process(crt_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, float* d_state, int state_valid, const crt_frame* frames, int n_frames,
                void* stream, crt_launch_info* info) {
    if (!ctx) return CRT_ERR_INVALID;
    int K = (ctx->have_params && d_in && d_out && frames) ? choose_shards(ctx, n_frames) : 1;
    // automatic mode: clip mode instead of shards where it measures faster (policy_auto_prefers_clip, crt_policy.h);
    // crt_set_shards(1) always gives clip mode where it applies
    if (K > 1 && ctx->shards_wanted == 0 && env_int("CRT_SHARDS", -1) < 0 && clip_wanted(ctx, d_out, d_state, nullptr, n_frames) &&
        (policy_auto_prefers_clip(ctx->plan.gauss_k != 0, ctx->W, ctx->H, clip_resident(ctx)) || env_int("CRT_CLIP_AUTO", 0))) K = 1;
    if (K > 1 && (d_state || !(ctx->p.persistence > 0.0)))
        return process_sharded(ctx, K, d_in, d_out, d_state, state_valid, frames, n_frames, (cudaStream_t)stream, info);
    if (info) { info->reserved[0] = 1; info->reserved[1] = 0; }
    return process_impl(ctx, d_in, d_out, d_state, state_valid, nullptr, frames, n_frames, (cudaStream_t)stream, info);
}

int crt_set_shards(crt_ctx* ctx, int shards) {
    if (!ctx || shards < 0 || shards > 16) return CRT_ERR_INVALID;
    ctx->shards_wanted = shards;
    return CRT_OK;
}

int crt_process_static(crt_ctx* ctx, const uint8_t* d_in, float* d_img, const crt_frame* frames, int n_frames, void* stream, crt_launch_info* info) {
    if (!d_img) return fail(ctx, CRT_ERR_INVALID, "null image buffer");
    return process_impl(ctx, d_in, nullptr, nullptr, 0, d_img, frames, n_frames, (cudaStream_t)stream, info);
}

int crt_reset_state(crt_ctx* ctx) {
    if (!ctx) return CRT_ERR_INVALID;
    ctx->state_valid = 0;
    return CRT_OK;
}

int crt_process_host(crt_ctx* ctx, const uint8_t* h_in, uint8_t* h_out, const crt_frame* frames, int n_frames, crt_launch_info* info) {
    if (!ctx || !h_in || !h_out || !frames || n_frames < 0) return CRT_ERR_INVALID;
    if (!ctx->have_params) return fail(ctx, CRT_ERR_INVALID, "crt_set_params has not been called");
    CU(cudaSetDevice(ctx->device));
    const size_t fbytes = (size_t)ctx->W * ctx->H * 3;
    HostRing& r = ctx->ring;
    if (!r.ready) {
        size_t target = 48u << 20;                                     // ~48 MB per chunk
        r.chunk_frames = (int)(target / fbytes); if (r.chunk_frames < 1) r.chunk_frames = 1; if (r.chunk_frames > 256) r.chunk_frames = 256;
        CU(cudaStreamCreateWithFlags(&r.s_in, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&r.s_comp, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&r.s_out, cudaStreamNonBlocking));
        for (int s = 0; s < HostRing::SLOTS; ++s) {
            CU(cudaMalloc((void**)&r.d_in[s], fbytes * r.chunk_frames));
            CU(cudaMalloc((void**)&r.d_out[s], fbytes * r.chunk_frames));
            CU(cudaEventCreateWithFlags(&r.in_ready[s], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&r.done[s], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&r.out_ready[s], cudaEventDisableTiming));
        }
        r.ready = true;
    }
    if (!ctx->state) CU(cudaMalloc((void**)&ctx->state, fbytes * sizeof(float)));
    int total_launches = 0, fused = 0, chunk_idx = 0;
    for (int f0 = 0; f0 < n_frames; f0 += r.chunk_frames, ++chunk_idx) {
        const int s = chunk_idx % HostRing::SLOTS;
        const int nf = n_frames - f0 < r.chunk_frames ? n_frames - f0 : r.chunk_frames;
        // slot reuse: the copy-out that last used this slot must have drained, and its compute too
        if (chunk_idx >= HostRing::SLOTS) { CU(cudaStreamWaitEvent(r.s_in, r.done[s], 0)); CU(cudaStreamWaitEvent(r.s_comp, r.out_ready[s], 0)); }
        CU(cudaMemcpyAsync(r.d_in[s], h_in + (size_t)f0 * fbytes, fbytes * nf, cudaMemcpyHostToDevice, r.s_in));
        CU(cudaEventRecord(r.in_ready[s], r.s_in));
        CU(cudaStreamWaitEvent(r.s_comp, r.in_ready[s], 0));
        crt_launch_info li{};
        int rc = process_impl(ctx, r.d_in[s], r.d_out[s], ctx->state, ctx->state_valid, nullptr, frames + f0, nf, r.s_comp, &li);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        total_launches += li.kernels_launched; fused = li.fused;
        if (ctx->p.persistence > 0.0) ctx->state_valid = 1;
        CU(cudaEventRecord(r.done[s], r.s_comp));
        CU(cudaStreamWaitEvent(r.s_out, r.done[s], 0));
        CU(cudaMemcpyAsync(h_out + (size_t)f0 * fbytes, r.d_out[s], fbytes * nf, cudaMemcpyDeviceToHost, r.s_out));
        CU(cudaEventRecord(r.out_ready[s], r.s_out));
    }
    CU(cudaStreamSynchronize(r.s_out));
    CU(cudaStreamSynchronize(r.s_comp));
    if (info) { info->kernels_launched = total_launches; info->fused = fused; }
    return CRT_OK;
}

int crt_resize_state(crt_ctx* ctx, const float* d_src, int src_width, int src_height, float* d_dst, void* stream) {
    if (!ctx || !d_src || !d_dst || src_width < 1 || src_height < 1) return CRT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const std::vector<Lerp1> hx = linear_coords(ctx->W, src_width), hy = linear_coords(ctx->H, src_height);
    Lerp1* d_c = nullptr;                       // [W + H] coordinate tables; rare call (the preview window was resized): allocate and free
    CU(cudaMallocAsync((void**)&d_c, (hx.size() + hy.size()) * sizeof(Lerp1), st));
    CU(cudaMemcpyAsync(d_c, hx.data(), hx.size() * sizeof(Lerp1), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_c + hx.size(), hy.data(), hy.size() * sizeof(Lerp1), cudaMemcpyHostToDevice, st));
    const int rc = launch_resize_state(d_src, src_width, d_dst, ctx->W, ctx->H, d_c, d_c + hx.size(), st);
    CU(cudaStreamSynchronize(st));              // the host tables are locals: the copies must have left them
    CU(cudaFreeAsync(d_c, st));
    return rc ? fail(ctx, CRT_ERR_CUDA, std::string("state resize launch failed: ") + cudaGetErrorString(cudaGetLastError())) : CRT_OK;
}

int crt_profile_begin(crt_ctx* ctx, int max_samples) {
    if (!ctx || max_samples < 1) return CRT_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    while ((int)ctx->prof_ev.size() < 2 * max_samples) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        ctx->prof_ev.push_back(e);
    }
    ctx->prof_cap = max_samples; ctx->prof_n = 0; ctx->prof_tick = 0; ctx->prof_open = false; ctx->prof_on = true;
    return CRT_OK;
}

int crt_profile_sample_every(crt_ctx* ctx, int every) {
    if (!ctx || every < 1) return CRT_ERR_INVALID;
    ctx->prof_every = every;
    return CRT_OK;
}

int crt_profile_end(crt_ctx* ctx, double* total_ms, int* samples) {
    if (!ctx || !total_ms || !samples) return CRT_ERR_INVALID;
    ctx->prof_on = false;
    double tot = 0.0;
    for (int i = 0; i < ctx->prof_n; ++i) {
        CU(cudaEventSynchronize(ctx->prof_ev[2 * i + 1]));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
        tot += ms;
    }
    int frames = 0;
    for (int i = 0; i < ctx->prof_n; ++i) frames += ctx->prof_frames[i];
    *total_ms = tot; *samples = frames;
    return CRT_OK;
}

int crt_generate_noise(crt_ctx* ctx, uint64_t frame_index, float* d_plane, void* stream) {
    if (!ctx || !d_plane) return CRT_ERR_INVALID;
    if (!ctx->have_params) return fail(ctx, CRT_ERR_INVALID, "crt_set_params has not been called");
    if (!ctx->dev_ok) { int rc = build_dev(ctx); if (rc) return rc; }
    CU(cudaSetDevice(ctx->device));
    if (launch_noise_gen(d_plane, ctx->dev.gh * ctx->dev.gw, ctx->p.noise_seed, frame_index, (cudaStream_t)stream))
        return fail(ctx, CRT_ERR_CUDA, std::string("noise generator launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    return CRT_OK;
}

int crt_generate_glitch(crt_ctx* ctx, const crt_frame* frame, int32_t* d_offs, void* stream) {
    if (!ctx || !frame || !d_offs) return CRT_ERR_INVALID;
    if (!ctx->have_params) return fail(ctx, CRT_ERR_INVALID, "crt_set_params has not been called");
    CU(cudaSetDevice(ctx->device));
    return gen_glitch(ctx, *frame, glitch_geom(ctx->p, ctx->W, ctx->H), d_offs, (cudaStream_t)stream);
}

}  // extern "C"
