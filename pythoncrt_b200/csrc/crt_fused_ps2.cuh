// crt_fused_ps2.cuh — fused tile kernel specialised for the reference's DEFAULT chain
// shape: pixel_size 2 (regular pixelate tables, even frame size), fast bloom or none, no
// warp / glitch / text layer.  This is the chain BASELINE.json's north_star names.
//
// With pixel_size 2 every aligned 2x2 block of the frame shows ONE source pixel, and the
// fast bloom's 2x down-scale cell of that block is exactly that pixel's value
// (fma(s - s, 0.5, s) == s in cv2.resize's arithmetic).  The tile therefore needs only one
// graded value per block (34 x 18 blocks for a 64 x 32 tile instead of 68 x 36 pixels), no
// down-scale pass, and each thread produces a 4 x 2 pixel patch from a 4 x 3 neighbourhood
// of block values held in registers: cv2's 2x up-scale (rows first, then columns, every
// lerp fma(q - p, w, p) with w = 0.25 / 0.75) costs ~16 instructions per pixel.  Clamped
// block indices reproduce cv2.resize's edge rule exactly (a lerp between equal values
// returns the value).  Everything else (triad LUT, masks, persistence, stores) is shared
// with the general fused kernel (crt_fused.cuh).  ~25 KB static shared memory, no dynamic.
// Two variants: k_fused_ps2 (plain loads) and k_fused_ps2_pipe (tile input and state by TMA, below).
#pragma once
#include "crt_fused.cuh"
#include "crt_launch.h"
#include "crt_tma.cuh"

namespace crt {

constexpr int P2_TW = 64, P2_TH = 32;                 // output tile
constexpr int P2_BW = P2_TW / 2 + 2, P2_BH = P2_TH / 2 + 2;   // blocks incl. one halo block each side (34 x 18)
constexpr int P2_NT = 256;                            // 16 x 16 threads, 4 x 2 pixels each

// ---- compile-time specialisation of the block kernels ---------------------------------------------------------------
// The parameter block is generic: every stage is guarded by a run-time flag (colour grading on?, which scanline mode?,
// analytic vignette or plane?, flicker?).  ncu (round 2, run 13) showed a fifth of the default-chain kernel's instructions
// to be that generality — constant-bank loads of the flags, uniform compares, branches.  A kernel instantiated with a
// non-zero SPEC overwrites those flags in its LOCAL copy of the parameter block with the constants of one common feature
// set; constant propagation then deletes the tests (and the code of stages that are off).  The host launcher picks a SPEC
// only when the clip's parameters match it exactly (ps2_spec below), so results are bit-identical to the generic kernel.
enum : int {
    SP_NOCOLOUR = 1,      // brightness / contrast / gamma / saturation / temperature all identity (:279-305 skipped)
    SP_ALLCOLOUR = 2,     // all four colour stages on (BASELINE configs[1]'s grade)
    SP_SCAN1 = 4,         // scanlines: per-row mask (angle 0, thickness 1)
    SP_SCAN2 = 8,         // scanlines: slanted / shaped plane
    SP_VIG1 = 16,         // analytic vignette
    SP_NOFLICKER = 32,
    SP_RGB = 64,          // channel order RGB
    SP_FASTTAIL = 128,    // the specialised tail's feature set: triad through the composite LUT, no text layer
    SP_NONOISE = 256,
    SP_GRAIN = 512,       // noise on, grain_size > 1 (bilinear up-scale of the draw plane)
    SP_FLICKER = 1024,
    SP_THR = 2048,        // bloom threshold on
    SP_NOTHR = 4096,      // ... off
};
constexpr int SPEC_DEFAULT = SP_NOCOLOUR | SP_SCAN1 | SP_VIG1 | SP_NOFLICKER | SP_RGB | SP_FASTTAIL | SP_NONOISE | SP_NOTHR;      // the CLI's default chain
constexpr int SPEC_SLANTED = SP_NOCOLOUR | SP_SCAN2 | SP_VIG1 | SP_NOFLICKER | SP_RGB | SP_FASTTAIL | SP_NONOISE | SP_NOTHR;      // ... with scanline angle / thickness
constexpr int SPEC_GRADED = SP_ALLCOLOUR | SP_SCAN1 | SP_VIG1 | SP_NOFLICKER | SP_RGB | SP_FASTTAIL | SP_NONOISE | SP_THR;     // BASELINE configs[1]
constexpr int SPEC_FULL = SP_ALLCOLOUR | SP_SCAN2 | SP_VIG1 | SP_FLICKER | SP_RGB | SP_FASTTAIL | SP_GRAIN;             // BASELINE configs[3], [4]: everything on

template <int SPEC>
CRT_HD void specialise(Dev& d, FrameDev& f) {
    if (SPEC & SP_NOCOLOUR) { d.col_sat = 0; d.col_temp = 0; d.col_bc = 0; d.col_gamma = 0; }
    if (SPEC & SP_ALLCOLOUR) { d.col_sat = 1; d.col_temp = 1; d.col_bc = 1; d.col_gamma = 1; }
    if (SPEC & SP_SCAN1) d.scan_mode = 1;
    if (SPEC & SP_SCAN2) d.scan_mode = 2;
    if (SPEC & SP_VIG1) d.vig_mode = 1;
    if (SPEC & SP_NOFLICKER) f.flicker_on = 0;
    if (SPEC & SP_RGB) d.bgr = 0;
    if (SPEC & SP_FASTTAIL) { d.triad_mode = 2; d.text_mode = 0; }
    if (SPEC & SP_NONOISE) d.noise_on = 0;
    if (SPEC & SP_GRAIN) d.noise_on = 1;
    if (SPEC & SP_FLICKER) f.flicker_on = 1;
    if (SPEC & SP_THR) d.thr_on = 1;
    if (SPEC & SP_NOTHR) d.thr_on = 0;
}
// does the clip's parameter set match SPEC exactly?  (host; `flicker_on` is a per-clip property: strength > 0 and hz > 0)
inline bool spec_matches(int spec, const Dev& d, bool flicker_on, bool fast_tail) {
    const bool nocol = !d.col_sat && !d.col_temp && !d.col_bc && !d.col_gamma, allcol = d.col_sat && d.col_temp && d.col_bc && d.col_gamma;
    if ((spec & SP_NOCOLOUR) && !nocol) return false;
    if ((spec & SP_ALLCOLOUR) && !allcol) return false;
    if ((spec & SP_SCAN1) && d.scan_mode != 1) return false;
    if ((spec & SP_SCAN2) && d.scan_mode != 2) return false;
    if ((spec & SP_VIG1) && d.vig_mode != 1) return false;
    if ((spec & SP_NOFLICKER) && flicker_on) return false;
    if ((spec & SP_RGB) && d.bgr) return false;
    if ((spec & SP_FASTTAIL) && !(fast_tail && d.text_mode == 0)) return false;
    if ((spec & SP_NONOISE) && d.noise_on) return false;
    if ((spec & SP_GRAIN) && !(d.noise_on && d.nz_x)) return false;
    if ((spec & SP_FLICKER) && !flicker_on) return false;
    if ((spec & SP_THR) && !d.thr_on) return false;
    if ((spec & SP_NOTHR) && d.thr_on) return false;
    return true;
}

// the TMA-pipelined variants' input box: 256 bytes must cover the tile's blocks, the aberration shift and the alignment slack
inline bool fused_ps2_pipe_supported(const Dev& d) {
    const int a0 = d.aberr != 0 ? d.aberr_mod : 0, as = a0 > (d.W >> 1) ? a0 - d.W : a0;
    return (d.W & 7) == 0 && as >= -6 && as <= 6 && env_int("CRT_PIPE", 1) != 0;
}

CRT_HD bool fused_ps2_supported(const Dev& d, bool glitch_on) {
    return d.pix_uniform == 2 && d.even_dims && (d.W & 3) == 0 && !d.warp_on && !glitch_on && d.text_mode == 0 && d.bloom_mode != 2;
}

#if defined(__CUDACC__)

// Stages 5-10, persistence and stores for a thread's 4 x 2 pixel patch (two 2x2 blocks side by side).
// t1 = graded value of the two blocks, bloom(r, k) = blurred bloom source of pixel k of row r.
struct NoRowHook { __device__ __forceinline__ void operator()(int) const {} };
template <bool BLOOM, bool FAST, typename BloomFn, typename RowFn = NoRowHook>
__device__ __forceinline__ void ps2_patch_tail(const Dev& d, const FrameDev& f, const MaskTabs& mt, const float* s_fwd, const float* s_inv,
                                               const float (*s_sel)[12], float* __restrict__ state, uint8_t* __restrict__ out, float* __restrict__ q_out, int has_prev,
                                               int ox0, int oy0, int ox1, int oy1, int xb, int y0, const float (&t1)[2][3], BloomFn&& bloom,
                                               float* s_prev = nullptr, bool state_in_smem = false, RowFn&& row_begin = RowFn(),
                                               int s_pitch = P2_TW * 3, bool allow_fast = true) {
    // row_begin(r) runs before row r of the patch is evaluated (e.g. to compute that row's bloom values only then)
    // Both rows of the patch are inside the frame (even frame height, even y0), so the two rows' arithmetic is one
    // straight-line block the scheduler can interleave.
    // s_prev: the patch's previous state in shared memory (row pitch s_pitch floats) when a TMA copy fetched it
    auto finish = [&](int r, int y, auto&& pixel) {
        if (s_prev && state_in_smem) {
            // previous state read from (single-pass), and the new state / pre-warp image written back to, the tile buffer
            // in shared memory (it leaves with one TMA store per tile); only the packed uint8 pixels are stored from here
            float4* sp = reinterpret_cast<float4*>(s_prev + r * s_pitch);
            const bool blend = has_prev && !q_out;
            float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa, pc = pa;
            if (blend) { pa = sp[0]; pb = sp[1]; pc = sp[2]; }
            const float prev[12] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w, pc.x, pc.y, pc.z, pc.w};
            float res[12];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                F3 v = pixel(y, xb + k, k);
                if (blend) {
                    v.x = blend_fast(prev[k * 3], v.x, d.persist, d.persist_q);
                    v.y = blend_fast(prev[k * 3 + 1], v.y, d.persist, d.persist_q);
                    v.z = blend_fast(prev[k * 3 + 2], v.z, d.persist, d.persist_q);
                }
                res[k * 3] = v.x; res[k * 3 + 1] = v.y; res[k * 3 + 2] = v.z;
            }
            sp[0] = make_float4(res[0], res[1], res[2], res[3]);
            sp[1] = make_float4(res[4], res[5], res[6], res[7]);
            sp[2] = make_float4(res[8], res[9], res[10], res[11]);
            if (!q_out) {
                uint32_t* op = reinterpret_cast<uint32_t*>(out + (y * d.W + xb) * 3);
                op[0] = pack4(res[0], res[1], res[2], res[3]);
                op[1] = pack4(res[4], res[5], res[6], res[7]);
                op[2] = pack4(res[8], res[9], res[10], res[11]);
            }
        } else if (s_prev) {
            const float4* sp = reinterpret_cast<const float4*>(s_prev + r * s_pitch);
            finish_quad(d, state, out, q_out, has_prev, y, xb, 4, pixel, true, sp[0], sp[1], sp[2]);
        } else {
            finish_quad(d, state, out, q_out, has_prev, y, xb, 4, pixel);
        }
    };
    const bool fast = FAST && allow_fast && ox0 >= d.comp_x0 && ox1 <= d.comp_x1;        // block-uniform
    if (fast) {
        float cvig[4], cscan[4];
        // composite table (bright: s_fwd, dim: s_inv == s_fwd + 1028) per column and channel: the table offset rides on the magic
        // number of the index computation (2^23 + offset: the mantissa of fma.rz(v, 1024, magic) is offset + floor(1024 v)),
        // looked up by the quad's mask phase (s_sel, filled once per CTA by ps2_fill_sel)
        // (measured and NOT adopted, runs 16-17: folding the blend factor (1 - p) into the mask multiplier and the unclipped
        // vignette into one fma per pixel — five instructions per pixel fewer, but 50-200 bytes of spills at 64 registers: slower)
        const int ph0 = xb - 3 * (int)__umulhi((unsigned)xb, 0x55555556u);     // xb % 3
        const float4* selp = reinterpret_cast<const float4*>(s_sel[ph0]);
        const float4 sa = selp[0], sb = selp[1], sc = selp[2];
        const float mg[4][3] = {{sa.x, sa.y, sa.z}, {sa.w, sb.x, sb.y}, {sb.z, sb.w, sc.x}, {sc.y, sc.z, sc.w}};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            cvig[k] = d.vig_mode ? mt.col_vig[xb - ox0 + k] : 0.f;
            cscan[k] = d.scan_mode == 2 ? mt.col_scan[xb - ox0 + k] : 0.f;
        }
        const float vs = d.vig_mode ? d.vig_strength : 0.f;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int y = y0 + r;
            row_begin(r);
            const float rscan = d.scan_mode ? mt.row_scan[y - oy0] : 1.0f;         // mask (mode 1) or phase fraction (mode 2)
            const float flick = f.flicker_on ? f.flicker : 1.0f;
            const float rfac = (d.scan_mode == 2 ? 1.0f : rscan) * flick;
            const float rvig = d.vig_mode ? mt.row_vig[y - oy0] : 0.f;
            auto pixel = [&](int, int, int k) -> F3 {
                F3 v = mk3(t1[k >> 1][0], t1[k >> 1][1], t1[k >> 1][2]);
                if (BLOOM) v = add_bloom(d, v, bloom(r, k));
                float m = rfac * __fmaf_rn(-vs, __saturatef(rvig + cvig[k]), 1.0f);
                if (d.scan_mode == 2) {             // slanted / shaped scanlines: same formula as mask_at
                    float t = rscan + cscan[k];
                    t = t >= 1.0f ? t - 1.0f : t;
                    const float sv = fmaxf(__fmaf_rn(-0.5f, __sinf(__fmaf_rn(6.283185307f, t, -3.14159265f)), 0.5f), 0.0f);
                    const float shaped = (d.scan_inv_sharp == 1.0f) ? sv : __powf(sv, d.scan_inv_sharp);
                    m *= __fmaf_rn(-d.scan_strength, shaped, 1.0f);
                }
                v.x = __saturatef(s_fwd[lut_index_magic(__saturatef(v.x), mg[k][0])] * m);
                v.y = __saturatef(s_fwd[lut_index_magic(__saturatef(v.y), mg[k][1])] * m);
                v.z = __saturatef(s_fwd[lut_index_magic(__saturatef(v.z), mg[k][2])] * m);
                if (d.noise_on) {                   // the same draw on all three channels (:646-648)
                    const float n = noise_at(d, f, y, xb + k);
                    v.x = __saturatef(v.x + n); v.y = __saturatef(v.y + n); v.z = __saturatef(v.z + n);
                }
                return v;
            };
            finish(r, y, pixel);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int y = y0 + r;
            row_begin(r);
            auto pixel = [&](int yy, int x, int k) -> F3 {
                // (tiles of the source-driven warp kernel reach over the frame's edge: nothing to evaluate out there)
                if ((unsigned)x >= (unsigned)d.W || (unsigned)yy >= (unsigned)d.H) return mk3(0.f, 0.f, 0.f);
                F3 v = mk3(t1[k >> 1][0], t1[k >> 1][1], t1[k >> 1][2]);
                if (BLOOM) v = add_bloom(d, v, bloom(r, k));
                return after_bloom_fast(d, f, v, yy, x, s_fwd, s_inv, mt, yy - oy0, x - ox0);
            };
            finish(r, y, pixel);
        }
    }
}

// s_sel[p][3 k + ch]: index magic (2^23 + table offset) column k (0..3) of a quad with xb % 3 == p uses for channel ch
__device__ __forceinline__ void ps2_fill_sel(float (*s_sel)[12], int tid, int bgr) {
    if (tid < 36) {
        const int p = tid / 12, e = tid - 12 * p, k = e / 3, ch = e - 3 * k;
        s_sel[p][e] = ((p + k) % 3 == (bgr ? 2 - ch : ch)) ? 8388608.0f : 8388608.0f + 1028.0f;      // 2^23 (+ offset of the dim table)
    }
}

// FAST: the per-pixel tail is specialised for the default chain's feature set — regular triad mask
// through the composite tables, per-row scanlines (or none), analytic vignette (or none), optional
// flicker folded into the row factor, no noise.  Tiles that touch the mask's irregular edge columns
// and every other feature set take the general tail (after_bloom_fast).
template <bool BLOOM, bool FAST, int MINB, int SPEC = 0>
__global__ void __launch_bounds__(P2_NT, MINB) k_fused_ps2(Dev d_arg, FrameDev f_arg, const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                     float* __restrict__ state, float* __restrict__ q_out, int has_prev) {
    Dev d = d_arg;
    FrameDev f = f_arg;
    specialise<SPEC>(d, f);             // SPEC != 0: feature flags become compile-time constants (see above)
    __shared__ __align__(16) float s_lut[2 * 1028];
    __shared__ __align__(16) float s_sel[3][12];
    float* const s_fwd = s_lut;
    float* const s_inv = s_lut + 1028;
    __shared__ float s_unit[256];
    __shared__ __align__(16) float s_pow[POW_TAB_FLOATS];
    __shared__ float s_rows[2 * P2_TH], s_cols[2 * P2_TW];
    __shared__ __align__(16) float Us[3][P2_BH][P2_BW + 2];      // graded block values, planar (row pitch 36 floats)
    __shared__ __align__(16) float Ss[3][P2_BH][P2_BW + 2];      // thresholded bloom source (only when the threshold is on)
    const int tid = threadIdx.x;
    griddep_launch_dependents();        // the next frame's kernel may begin its state-independent phases (see launch_pdl)
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
    // Persistent CTAs: the tables above are staged once, then the CTA walks over tiles
    // (tile = blockIdx.x, blockIdx.x + gridDim.x, ...; consecutive CTAs work on neighbouring tiles).
    const int tiles_x = (d.W + P2_TW - 1) / P2_TW, ntiles = tiles_x * ((d.H + P2_TH - 1) / P2_TH);
    const int step_y = gridDim.x / tiles_x, step_x = gridDim.x - step_y * tiles_x;      // tile += gridDim.x without a division per tile
    int tby = blockIdx.x / tiles_x, tbx = blockIdx.x - tby * tiles_x;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int ox0 = tbx * P2_TW, oy0 = tby * P2_TH;
    const int ox1 = imin(ox0 + P2_TW, d.W) - 1, oy1 = imin(oy0 + P2_TH, d.H) - 1;
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
    // The LUT / u8->float tables staged before the loop must be visible before the first grading phase.  Later tiles
    // need no barrier here: the mask tables written above are read after the next barrier, and the previous tile's
    // block values died at the barrier that ended its iteration.
    if (tile == (int)blockIdx.x) __syncthreads();

    // ---- phase 1: one graded value per 2x2 block (tile + one halo block, clamped = cv2's edge rule) ----
    const int gbx0 = (ox0 >> 1) - 1, gby0 = (oy0 >> 1) - 1;
    {
        constexpr int NIT = (P2_BW * P2_BH + P2_NT - 1) / P2_NT;          // 3 blocks per thread at most
        uint32_t raw[NIT][3];                                 // 32-bit: a byte array would live in local memory
        const int a0 = d.aberr != 0 ? d.aberr_mod : 0;
        // all loads first: their latencies overlap.  Tiles away from the left / right frame edge need neither the
        // block clamp nor the aberration wrap in x: one 32-bit offset per block, byte offsets per channel.
        const int as = a0 > (d.W >> 1) ? a0 - d.W : a0;                     // signed shift (aberr_mod is taken modulo W)
        const int aa = as < 0 ? -as : as;
        if (2 * gbx0 - aa >= 0 && 2 * (gbx0 + P2_BW - 1) + aa < d.W) {      // tile-uniform
            const int W3 = d.W * 3;
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int u = tid + it * P2_NT;
                if (u < P2_BW * P2_BH) {
                    const int bj = u / P2_BW, bi = u - bj * P2_BW;
                    const int sy = 2 * imin(imax(gby0 + bj, 0), d.hh - 1);
                    const uint8_t* p = in + (unsigned)(sy * W3 + 6 * (gbx0 + bi));      // < 2^31 (checked by plan_fused)
                    raw[it][0] = p[-3 * as];
                    raw[it][1] = p[1];
                    raw[it][2] = p[3 * as + 2];
                }
            }
        } else {
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int u = tid + it * P2_NT;
                if (u < P2_BW * P2_BH) {
                    const int bj = u / P2_BW, bi = u - bj * P2_BW;
                    const int sx = 2 * imin(imax(gbx0 + bi, 0), d.hw - 1), sy = 2 * imin(imax(gby0 + bj, 0), d.hh - 1);
                    const uint8_t* row = in + (size_t)sy * d.W * 3;
                    raw[it][0] = row[wrap(sx - a0, d.W) * 3 + 0];
                    raw[it][1] = row[sx * 3 + 1];
                    raw[it][2] = row[wrap(sx + a0, d.W) * 3 + 2];
                }
            }
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int u = tid + it * P2_NT;
            if (u < P2_BW * P2_BH) {
                const int bj = u / P2_BW, bi = u - bj * P2_BW;
                const F3 v1 = colour(d, mk3(s_unit[raw[it][0]], s_unit[raw[it][1]], s_unit[raw[it][2]]), s_pow);
                Us[0][bj][bi] = v1.x; Us[1][bj][bi] = v1.y; Us[2][bj][bi] = v1.z;
                if (BLOOM && d.thr_on) {
                    const F3 sv = bloom_src(d, v1);
                    Ss[0][bj][bi] = sv.x; Ss[1][bj][bi] = sv.y; Ss[2][bj][bi] = sv.z;
                }
            }
        }
    }
    __syncthreads();
    griddep_wait();         // previous kernel of the stream complete: state / pre-warp image / noise may be touched from here on

    // ---- phase 4: every thread owns 4 x 2 output pixels = two blocks side by side ------------------------
    const int tx = tid & 15, ty = tid >> 4;
    const int xb = ox0 + 4 * tx, y0 = oy0 + 2 * ty;
    if (xb <= ox1 && y0 <= oy1) {
    const int bi = 2 * tx + 1, bj = ty + 1;                 // first of the two blocks, in halo coordinates
    float blr[4][3];                                        // bloom of the patch row being evaluated
    float t1[2][3];                                         // graded value of the two blocks
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) { t1[0][ch] = Us[ch][bj][bi]; t1[1][ch] = Us[ch][bj][bi + 1]; }
    // The bloom of a patch row is computed when that row is evaluated: cells bi-1 .. bi+2 of two block rows
    // (bi - 1 is even: 8-byte aligned pairs), cv2's x-lerps, one y-lerp.  Holding both rows' 24 values across the
    // first row's arithmetic does not fit in 64 registers (they were spilled and reloaded).
    auto row_begin = [&](int r) {
#pragma unroll
        for (int ch = 0; ch < (BLOOM ? 3 : 0); ++ch) {
            const float (*src)[P2_BW + 2] = d.thr_on ? Ss[ch] : Us[ch];
            float h[2][4];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float2 ca = *reinterpret_cast<const float2*>(&src[bj - 1 + r + q][bi - 1]);
                const float2 cb = *reinterpret_cast<const float2*>(&src[bj - 1 + r + q][bi + 1]);
                const float d01 = fsub(ca.y, ca.x), d12 = fsub(cb.x, ca.y), d23 = fsub(cb.y, cb.x);
                h[q][0] = ffma(d01, 0.75f, ca.x);           // x = 2i     : lerp(c[i-1], c[i], 0.75)
                h[q][1] = ffma(d12, 0.25f, ca.y);           // x = 2i + 1 : lerp(c[i], c[i+1], 0.25)
                h[q][2] = ffma(d12, 0.75f, ca.y);           // x = 2i + 2 : lerp(c[i], c[i+1], 0.75)
                h[q][3] = ffma(d23, 0.25f, cb.x);           // x = 2i + 3 : lerp(c[i+1], c[i+2], 0.25)
            }
            const float w = r == 0 ? 0.75f : 0.25f;         // y = 2j: lerp(row j-1, row j, 0.75); y = 2j + 1: lerp(row j, row j+1, 0.25)
#pragma unroll
            for (int k = 0; k < 4; ++k) blr[k][ch] = ffma(fsub(h[1][k], h[0][k]), w, h[0][k]);
        }
    };
    ps2_patch_tail<BLOOM, FAST>(d, f, mt, s_fwd, s_inv, s_sel, state, out, q_out, has_prev, ox0, oy0, ox1, oy1, xb, y0, t1,
                                [&](int, int k) { return mk3(blr[k][0], blr[k][1], blr[k][2]); }, nullptr, false, row_begin);
    }                       // this thread's patch
    __syncthreads();        // everyone is done with this tile's tables / block values
    tbx += step_x; tby += step_y;
    if (tbx >= tiles_x) { tbx -= tiles_x; ++tby; }
    }                       // tile loop
}


// ---- TMA-pipelined variant ----------------------------------------------------------------------
// Same arithmetic as k_fused_ps2; what changes is how a tile's data reaches the SM.  The plain kernel is
// latency bound (ncu, run 28/30: issue slots 46 % busy, long-scoreboard stalls on the first use of the
// input bytes and of the persistence state; removing instructions did not shorten it).  Here
//   * the tile's input bytes — 18 even rows x 256 bytes around what the tile's blocks read — arrive by ONE
//     tiled tensor-map copy (TMA) per tile into a double buffer, issued one tile ahead (for the first tile:
//     at kernel entry, before the previous frame's kernel has finished — see launch_pdl);
//   * the tile's previous state (32 rows x 768 bytes) arrives by one TMA copy issued at the top of the
//     tile's iteration and is consumed after the grading phase; the blend reads it from shared memory;
// so no thread waits on a global load: the copies are in flight while the CTA (and the other CTAs of the
// SM) compute.  The new state is written back into the same shared-memory tile and leaves with ONE TMA
// store per tile (fully coalesced; the 16-byte per-thread stores of the plain kernel reach 3.3 TB/s on
// their own, coalesced stores 5.4 TB/s: tests/_probe/store_probe.cu); the packed uint8 pixels (3 B/px) are
// still stored by the threads.  Tiles on the left / right frame edge,
// where chromatic aberration wraps around (np.roll), read their bytes with the plain loads.
// Measured (round 1, default chain at 4K): 66.0 us per frame against 70.7 us for the plain kernel with
// 4 CTAs per SM (run 35).  Each half alone does not pay: state by TMA with the input on batched byte loads
// 70.3 us (run 36); input by TMA without a state to fetch 62.5 vs 59.7 us (first pass of the two-pass path,
// run 35); with 3 CTAs per SM 72.5 us (run 33).  A first variant with one 1-D bulk copy per tile ROW
// (run 18) was slower still — 96 copies per tile serialise in the copy engine.
// TMA facts measured with tests/_probe/tma_probe.cu: the innermost start coordinate must be 16-byte
// aligned (else: illegal-instruction exception), out-of-range and negative coordinates are zero-filled
// and count towards the barrier's byte count, a [frames][rows][bytes] map with a box of 1 frame never
// completed — hence one 2-D map over the even rows of all frames (the frame pitch is H/2 row pitches).
// Requires W % 8 == 0, |aberration| <= 6, 16-byte aligned clip / state pointers; used when there is a
// state to fetch and the frame has at least ~250 tiles (run 68: 1080p 51 700 vs 44 900 frames/s, VGA neutral).
constexpr int P2_RAW_W = 256, P2_RAW_BYTES = P2_RAW_W * P2_BH;   // input buffer: 18 rows x 256 bytes (the box starts 16-byte aligned)
constexpr int P2_ST_BYTES = P2_TH * P2_TW * 3 * 4;            // state tile: 32 x 192 float32
constexpr int P2_PIPE_SMEM = P2_ST_BYTES + 2 * P2_RAW_BYTES;

struct Ps2Maps {                 // host-encoded tensor maps (crt_abi.cu)
    CUtensorMap in;              // uint8 [frames * H/2 even rows][W*3] (row pitch 2 W*3), box 256 x 18
    CUtensorMap st;              // float32 [H][W*3], box 192 x 32: the state buffer, or the pre-warp image in the two-pass path
    int frame;                   // index of this launch's frame inside `in`
    int th = 32;                 // tile height of the single-pass block kernels (even, <= P2_TH) = rows of st's box: chosen per
                                 // call so that the tiles of a frame fill whole waves of resident CTAs (choose_tile_h, crt_abi.cu)
};

// ---- clip mode: ONE launch for a run of frames ------------------------------------------------------------------------
// One launch per frame leaves a B200 half idle at 1080p: launch latency, table staging and — mostly — the last round of every
// frame, in which a few CTAs run alone on their SMs (kernel alone 26.8 us per 1080p frame against 17.2 us when concurrent
// temporal shards fill the gaps, round 2 run 33).  The only cross-frame dependency of these kernels is the persistence state of
// the SAME tile (crt_filter.py:1092 is per pixel), so a run of frames is one queue of (frame, tile) items, frame-major, worked off
// by persistent CTAs; item (f, t) may fetch its state once done[t] >= f, published by whoever finished (f - 1, t) after its state
// tile's TMA store has completed.  Items are taken with a fixed stride under a cooperative launch (every CTA resident), or from
// an atomic counter — handed out in order and only to running CTAs, so the owner of every item an item waits for is resident:
// no deadlock whatever else shares the GPU.  The result is the serial recurrence exactly — no warm-up halo as with temporal
// shards.  Per-frame scalars (scanline phase, flicker gain) travel as kernel parameters.  DESIGN.md 4.8 has the measurements.
constexpr int P1_ROT = 160;              // phase 1 of the TMA-pipelined kernel: rotation of the thread -> block assignment (see there)
constexpr int CLIP_B = 224;              // the "next item" thread (warp 7: the lightest in phase 1); thread 0 keeps the state traffic
constexpr int CLIP_MAX_FRAMES = 64;      // frames per launch: their scalars travel as kernel parameters (no copy to wait for)
struct ClipArgs {
    int nf = 1;                          // frames in this launch (1: the per-frame launch, everything below unused)
    int* sync = nullptr;                 // [0]: item counter, [1 + t]: frames of tile t completed — zeroed before the launch
    unsigned long long frame_bytes = 0;  // W * H * 3
    int static_items = 0;                // 2: owned tiles — CTA i keeps tiles i, i + G, ... through all frames of the run (no flags at all);
                                         // 1: item = blockIdx.x + k * gridDim.x (the launch is cooperative: every CTA is resident); 0: atomic counter
    int per_sm = 0;                      // owned tiles: > 0 = CTAs per SM when consecutive CTAs share an SM -> rank = (i % per_sm) * SMs + i / per_sm
    int w0_idle = 1;                     // fast-bloom kernel: warp 0 does no phase-1 blocks (CRT_CLIP_W0)
    int release = 1;                     // publication / acquisition mode bits (clip_publish; CRT_CLIP_RELEASE)
    FrameVar fv[CLIP_MAX_FRAMES];
};
// mode bits (ClipArgs.release, CRT_CLIP_RELEASE): 1 = st.release.gpu, 2 = unqualified fence.proxy.async, 4 = fence.acq_rel.gpu on the
// consumer's side, 16 = never trust the early look (always ld.acquire in clip_wait).  Default 1: wait_group, proxy fence, release
// store.  MEASURED (round 2, runs 53-56): with neither 1 nor 2 — cp.async.bulk.wait_group 0, fence.proxy.async.global,
// st.relaxed.gpu — another SM's TMA load of the tile returns stale 32-byte sectors now and then (10-20 thousand values per
// 60-frame 1080p clip): the bulk group's completion makes the store visible to the waiting thread, not to the GPU; a
// MEMBAR.ALL.GPU (either bit) must follow before the flag is raised.  Cost at 4K: 44.6 us per frame without, 47.8 with.
__device__ __forceinline__ void clip_publish(int* flag, int value, int mode) {
    // the TMA store of the state tile has completed (bulk_wait_all: its writes are visible to this thread); release them gpu-wide
    // (SASS: fence.proxy.async without a space is MEMBAR.ALL.GPU + FENCE.VIEW.ASYNC.S, the .global form FENCE.VIEW.ASYNC.G alone;
    // st.release.gpu is MEMBAR.ALL.GPU + STG.STRONG.GPU)
    if (mode & 2) asm volatile("fence.proxy.async;" ::: "memory");
    else asm volatile("fence.proxy.async.global;" ::: "memory");
    if (mode & 1) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
    else asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
// a look at a flag whose result is not needed yet (the load is in flight until its first use) ...
__device__ __forceinline__ int clip_peek(const int* flag) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    return v;
}
// ... and what turns a successful look into an acquire, also for the TMA load (async proxy) that follows
__device__ __forceinline__ void clip_acquire(int mode) {
    if (mode & 4) asm volatile("fence.acq_rel.gpu;" ::: "memory");
    else asm volatile("fence.acquire.gpu;" ::: "memory");            // SASS: CCTL.IVALL
    if (mode & 2) asm volatile("fence.proxy.async;" ::: "memory");
    else asm volatile("fence.proxy.async.global;" ::: "memory");
}
// owned-tiles mode: this thread's completed TMA stores (cp.async.bulk.wait_group 0) become visible to its later TMA loads —
// the gpu-scope membar is what the race above showed to be necessary after the bulk group's completion
__device__ __forceinline__ void clip_settle(int mode) {
    if (mode & 8) asm volatile("fence.proxy.async.global;" ::: "memory");      // (light form: for the race hunt only)
    else asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void clip_wait(const int* flag, int value) {
    int v;
    const long long t0 = clock64();
    for (;;) {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v >= value) break;
        __nanosleep(100);
        if (clock64() - t0 > 4000000000LL) __trap();
    }
    asm volatile("fence.proxy.async.global;" ::: "memory");      // the TMA load that follows reads through the async proxy
}

// THR: the bloom threshold is on (a second block array for the thresholded source; 3 CTAs per SM instead of 4)
// CLIP: clip mode (above); `in` / `out` / `frame` describe the first frame of the run
template <bool BLOOM, bool FAST, bool THR, int SPEC = 0, bool CLIP = false>
__global__ void __launch_bounds__(P2_NT, (THR || (CLIP && SPEC == 0)) ? 3 : 4) k_fused_ps2_pipe(Dev d_arg, FrameDev f_arg, const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                          float* __restrict__ state, float* __restrict__ q_out, int has_prev,
                                                          const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_st,
                                                          int frame, int th, const __grid_constant__ ClipArgs ca) {
    Dev d = d_arg;
    FrameDev f = f_arg;
    specialise<SPEC>(d, f);             // SPEC != 0: feature flags become compile-time constants (see above)
    extern __shared__ __align__(128) unsigned char dsm[];
    float* s_state = reinterpret_cast<float*>(dsm);                                 // [TH][TW*3]
    uint8_t* s_raw = dsm + P2_ST_BYTES;                                               // [2][18][256]
    __shared__ __align__(16) float s_lut[2 * 1028];
    __shared__ __align__(16) float s_sel[3][12];
    float* const s_fwd = s_lut;
    float* const s_inv = s_lut + 1028;
    __shared__ float s_unit[256];
    __shared__ __align__(16) float s_pow[POW_TAB_FLOATS];
    __shared__ float s_rows[2 * P2_TH], s_cols[2 * P2_TW];
    __shared__ __align__(16) float Us[3][P2_BH][P2_BW + 2];
    __shared__ __align__(16) float Ss[THR ? 3 : 1][THR ? P2_BH : 1][P2_BW + 2];
    __shared__ __align__(8) uint64_t bar_in[2], bar_st;
    const int tid = threadIdx.x;
    griddep_launch_dependents();
    const bool use_state = has_prev && !q_out;      // a previous state to fetch
    const bool tile_out = use_state || q_out;       // the result (new state, or the pre-warp image of the two-pass path) leaves through the tile
    const int a0 = d.aberr != 0 ? d.aberr_mod : 0;
    const int as = a0 > (d.W >> 1) ? a0 - d.W : a0;                         // signed shift (aberr_mod is taken modulo W)
    const int aa = as < 0 ? -as : as;
    const int tiles_x = (d.W + P2_TW - 1) / P2_TW, ntiles = tiles_x * ((d.H + th - 1) / th);
    const uint32_t st_bytes = (uint32_t)th * P2_TW * 3 * 4;      // the state box: th rows
    const int step_y = gridDim.x / tiles_x, step_x = gridDim.x - step_y * tiles_x;
    int tby = blockIdx.x / tiles_x, tbx = blockIdx.x - tby * tiles_x;
    // clip mode: items = (frame, tile) pairs, frame-major, from the counter; s_item[] hands the item of the next iteration to the CTA
    __shared__ int s_item[2];
    const int nitems = CLIP ? ntiles * ca.nf : ntiles;
    // owned tiles: which tiles are this CTA's — spread so that the CTAs of one SM own n, n, n, n - 1 tiles whichever way the CTAs were placed
    const int rank = (CLIP && ca.per_sm > 0) ? ((int)blockIdx.x % ca.per_sm) * ((int)gridDim.x / ca.per_sm) + (int)blockIdx.x / ca.per_sm : (int)blockIdx.x;
    int fr = 0;                                     // frame of the current item within the run
    const unsigned magic_tx = make_magic(tiles_x);
    auto split = [&](int item, int& ifr, int& iby, int& ibx) {      // item -> frame, tile row, tile column (item < 2^31, tile < 2^16)
        // ifr on entry: a frame not after the item's (items only grow) -> a compare or two instead of a division
        int t = item - ifr * ntiles;
        while (t >= ntiles) { t -= ntiles; ++ifr; }
        iby = fastdiv(t, magic_tx); ibx = t - iby * tiles_x;
    };
    if (tid == 0) {
        mbar_init(&bar_in[0], 1); mbar_init(&bar_in[1], 1); mbar_init(&bar_st, 1);
        fence_mbar_init();
        int first = blockIdx.x;
        if (CLIP) { first = ca.static_items ? rank : atomicAdd(ca.sync, 1); s_item[0] = first; split(first < nitems ? first : 0, fr, tby, tbx); }
        // first tile's input: independent of the previous kernel
        if (first < nitems) {
            mbar_expect_tx(&bar_in[0], P2_RAW_BYTES);
            tma_load_2d_hint(s_raw, &map_in, (6 * ((tbx * P2_TW >> 1) - 1) - 3 * aa) & ~15, (frame + fr) * d.hh + (tby * th >> 1) - 1, &bar_in[0], L2_EVICT_FIRST);
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
    if (CLIP) __syncthreads();          // s_item[0] and the tables staged above (the per-frame kernel takes this barrier inside the loop)
    int it = 0;
    __shared__ int s_prev[2], s_hint;   // clip mode: tile and frame of the item whose state store is in flight; the current item's flag
    int hint = 0;                       // as thread CLIP_B saw it one iteration ago
    if (CLIP && tid == 0) s_hint = 0;
    for (int tile = CLIP ? s_item[0] : (int)blockIdx.x; tile < nitems; ++it) {
        if (CLIP) {
            split(tile, fr, tby, tbx);
            const FrameVar fv = ca.fv[fr];
            f.phase32 = fv.phase32; f.phase = fv.phase; f.flicker = fv.flicker;
        }
        const int ox0 = tbx * P2_TW, oy0 = tby * th;
        const int ox1 = imin(ox0 + P2_TW, d.W) - 1, oy1 = imin(oy0 + th, d.H) - 1;
        const int gbx0 = (ox0 >> 1) - 1, gby0 = (oy0 >> 1) - 1;
        const uint8_t* const in_f = CLIP ? in + fr * ca.frame_bytes : in;
        uint8_t* const out_f = CLIP ? out + fr * ca.frame_bytes : out;
        // next tile of this CTA
        int nbx = tbx + step_x, nby = tby + step_y, nfr = fr, next = tile + (int)gridDim.x;
        if (nbx >= tiles_x) { nbx -= tiles_x; ++nby; }
        const int buf = it & 1;
        // Clip mode's bookkeeping is split between two threads of different warps and spread over the iteration so that no round
        // trip to L2 sits on the tile's critical path.  Thread CLIP_B ("next item"): the counter's atomic is issued here and
        // consumed after phase 1, where it also fetches the next item's input and takes a look at the next item's flag, which it
        // hands to thread 0 (s_hint) at the end of the iteration.  Thread 0 ("state"): fetches this item's state at once when the
        // flag was already set at that look (the normal case: the tile's previous frame is a whole frame of items behind), awaits
        // the previous store's completion after phase 1, when it has long happened, and publishes it.  Blocking on another CTA's
        // flag comes only after our own publication — no CTA waits while it owes one: no deadlock.
        // OWNED TILES (ca.static_items == 2, opt-in: measured slower, see run_fused_ps2_clip): CTA i keeps tiles i, i + G, ... through every frame of the run, so a tile's
        // previous frame was stored by this very CTA — no flags, no counter, no waiting for anyone.  Thread 0 settles each store
        // (completion + membar) in its slack before the phase-1 barrier of the NEXT item, which covers every later fetch of that
        // tile when the CTA owns two tiles or more; a CTA with a single tile settles at the top (it is the one with time to spare).
        bool owed = false;                   // thread 0: the previous item's completion is still to be published
        const bool own = CLIP && ca.static_items == 2, own_one = own && rank + (int)gridDim.x >= ntiles;
        if (CLIP && tid == CLIP_B) {
            if (own) {
                const int nt = tby * tiles_x + tbx + (int)gridDim.x;
                next = nt < ntiles ? fr * ntiles + nt : (fr + 1 < ca.nf ? (fr + 1) * ntiles + rank : nitems);
            } else next = ca.static_items ? tile + (int)gridDim.x : atomicAdd(ca.sync, 1);
        }
        if (own && tid == 0 && it > 0) {
            bulk_wait_read();                // the previous tile's TMA store has drained the buffer
            if (own_one && fr > 0) { bulk_wait_all(); clip_settle(ca.release); }
            mbar_expect_tx(&bar_st, st_bytes); tma_load_2d(s_state, &map_st, ox0 * 3, oy0, &bar_st);
        }
        if (CLIP && !own && tid == 0 && it > 0) {
            bulk_wait_read();                // the previous tile's TMA store has drained the buffer
            owed = true;
            if (fr > 0 && (s_hint < fr || (ca.release & 16))) {
                bulk_wait_all(); clip_publish(ca.sync + 1 + s_prev[0], s_prev[1] + 1, ca.release); owed = false;
                clip_wait(ca.sync + 1 + tby * tiles_x + tbx, fr);
            } else if (fr > 0) clip_acquire(ca.release);
            mbar_expect_tx(&bar_st, st_bytes); tma_load_2d(s_state, &map_st, ox0 * 3, oy0, &bar_st);
        }
        if (!CLIP && tid == 0) {
            if (it > 0 && tile_out) {            // the previous tile's TMA store must have drained the buffer; then fetch this tile's state
                bulk_wait_read();
                if (use_state) { mbar_expect_tx(&bar_st, st_bytes); tma_load_2d(s_state, &map_st, ox0 * 3, oy0, &bar_st); }
            }
            if (next < nitems) {      // next tile's input into the other buffer (last read two barriers ago)
                mbar_expect_tx(&bar_in[buf ^ 1], P2_RAW_BYTES);
                tma_load_2d_hint(s_raw + (buf ^ 1) * P2_RAW_BYTES, &map_in, (6 * ((nbx * P2_TW >> 1) - 1) - 3 * aa) & ~15,
                                 frame * d.hh + (nby * th >> 1) - 1, &bar_in[buf ^ 1], L2_EVICT_FIRST);
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
        if (!CLIP && it == 0) __syncthreads();      // tables staged before the loop; later tiles need no barrier here (see k_fused_ps2)

        // ---- phase 1: one graded value per 2x2 block ----
        mbar_wait(&bar_in[buf], (it >> 1) & 1);                 // this tile's input bytes have landed
        {
            constexpr int NIT = (P2_BW * P2_BH + P2_NT - 1) / P2_NT;
            const bool x_inside = 2 * gbx0 - aa >= 0 && 2 * (gbx0 + P2_BW - 1) + aa < d.W;       // tile-uniform
            const uint8_t* rawb = s_raw + buf * P2_RAW_BYTES;
            const int xoff = 6 * gbx0 - ((6 * gbx0 - 3 * aa) & ~15);
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
                // blocks are dealt to the threads rotated by P1_ROT: the last, partial round (100 of 612 blocks) then falls to warps 3-6
                // and the two bookkeeping warps (thread 0, thread CLIP_B) reach the barrier early enough to absorb their waits
                // Clip mode: warp 0 — whose thread 0 waits for the previous store, fences and publishes — keeps out of phase 1
                // altogether (612 blocks are three rounds for 224 threads as for 256) and the partial round skips warp 7 (thread
                // CLIP_B).  Measured (run 68): default chain 4K 47.6 -> 46.5 us per frame, 1080p 12.27 -> 12.05.
                int u = ((tid + P1_ROT) & (P2_NT - 1)) + i * P2_NT;
                if (CLIP && ca.w0_idle) {
                    int v = tid - 32 + 196;
                    if (v >= 224) v -= 224;
                    u = tid < 32 ? (1 << 30) : v + i * 224;
                }
                if (u < P2_BW * ((th >> 1) + 2)) {
                    const int bj = u / P2_BW, bi = u - bj * P2_BW;
                    uint32_t r0, r1, r2;
                    if (x_inside) {
                        // buffer row of the (clamped) block row; byte 0 of the buffer is byte (6 gbx0 - 3 aa) & ~15 of the frame row
                        const uint8_t* p = rawb + (imin(imax(gby0 + bj, 0), d.hh - 1) - gby0) * P2_RAW_W + 6 * bi + xoff;
                        r0 = p[-3 * as]; r1 = p[1]; r2 = p[3 * as + 2];
                    } else {
                        const int sx = 2 * imin(imax(gbx0 + bi, 0), d.hw - 1), sy = 2 * imin(imax(gby0 + bj, 0), d.hh - 1);
                        const uint8_t* row = in_f + (size_t)sy * d.W * 3;
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
        if (CLIP && tid == 0) {      // (in the slack the rotation above leaves warp 0; the store was issued a phase ago)
            if (own) { if (it > 0 && !own_one) { bulk_wait_all(); clip_settle(ca.release); } }
            else {
                if (owed) { bulk_wait_all(); clip_publish(ca.sync + 1 + s_prev[0], s_prev[1] + 1, ca.release); }
                s_prev[0] = tby * tiles_x + tbx; s_prev[1] = fr;
            }
        }
        __syncthreads();
        if (CLIP && tid == CLIP_B) {
            s_item[buf ^ 1] = next; split(next < nitems ? next : 0, nfr, nby, nbx);
            if (next < nitems) {      // next item's input into the other buffer (last read in the previous iteration), and a look at its flag
                mbar_expect_tx(&bar_in[buf ^ 1], P2_RAW_BYTES);
                tma_load_2d_hint(s_raw + (buf ^ 1) * P2_RAW_BYTES, &map_in, (6 * ((nbx * P2_TW >> 1) - 1) - 3 * aa) & ~15,
                                 (frame + nfr) * d.hh + (nby * th >> 1) - 1, &bar_in[buf ^ 1], L2_EVICT_FIRST);
                if (nfr > 0 && !own) hint = clip_peek(ca.sync + 1 + nby * tiles_x + nbx);
            }
        }
        griddep_wait();         // previous kernel of the stream complete: state / pre-warp image / noise may be touched from here on
        if (use_state) {
            if (it == 0 && tid == 0) {           // first tile: the state may only be fetched now
                if (CLIP && fr > 0 && ca.static_items != 2) clip_wait(ca.sync + 1 + tby * tiles_x + tbx, fr);
                mbar_expect_tx(&bar_st, st_bytes);
                tma_load_2d(s_state, &map_st, ox0 * 3, oy0, &bar_st);
            }
            mbar_wait(&bar_st, it & 1);
        }

        // ---- phase 4 ----
        const int tx = tid & 15, ty = tid >> 4;
        const int xb = ox0 + 4 * tx, y0 = oy0 + 2 * ty;
        if (xb <= ox1 && y0 <= oy1) {
            const int bi = 2 * tx + 1, bj = ty + 1;
            float blr[4][3];                                     // bloom of the row being evaluated
            float t1[2][3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) { t1[0][ch] = Us[ch][bj][bi]; t1[1][ch] = Us[ch][bj][bi + 1]; }
            // The bloom of a patch row is computed when that row is evaluated (two block rows, x-lerps, one y-lerp):
            // holding both rows' 24 values across the first row's arithmetic does not fit in 64 registers (ncu run 65:
            // 12 spill stores + 12 reloads per thread and tile, reloads stalling on the long scoreboard).
            auto row_begin = [&](int r) {
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
                    const float w = r == 0 ? 0.75f : 0.25f;      // y = 2j: lerp(row j-1, row j, 0.75); y = 2j + 1: lerp(row j, row j+1, 0.25)
#pragma unroll
                    for (int k = 0; k < 4; ++k) blr[k][ch] = ffma(fsub(h[1][k], h[0][k]), w, h[0][k]);
                }
            };
            ps2_patch_tail<BLOOM, FAST>(d, f, mt, s_fwd, s_inv, s_sel, state, out_f, q_out, has_prev, ox0, oy0, ox1, oy1, xb, y0, t1,
                                        [&](int, int k) { return mk3(blr[k][0], blr[k][1], blr[k][2]); },
                                        tile_out ? s_state + (y0 - oy0) * (P2_TW * 3) + 12 * tx : nullptr, tile_out, row_begin);
        }
        if (CLIP && tid == CLIP_B) s_hint = hint;      // (the look's latency has passed behind phase 4)
        if (tile_out) fence_proxy_async();      // the new state in shared memory -> visible to the TMA engine
        __syncthreads();        // everyone is done with this tile's tables, block values and state tile
        if (tile_out && tid == 0) {            // the tile's new state leaves with one coalesced TMA store (rows outside the frame are clipped)
            // the pre-warp image is read back by the next kernel (keep it in L2); the state only a frame later
            // (measured, run 7: marking the state EVICT_FIRST costs the default chain 2 % — part of it is still in L2 a frame later)
            tma_store_2d_hint(&map_st, s_state, ox0 * 3, oy0, q_out ? L2_EVICT_LAST : L2_EVICT_NORMAL);
            bulk_commit();
        }
        if (CLIP) tile = s_item[buf ^ 1];      // written by thread 0 before this iteration's barriers
        else { tile += (int)gridDim.x; tbx = nbx; tby = nby; }
    }
    if (tile_out && tid == 0) {
        bulk_wait_all();      // the last tile's store has completed before the CTA exits
        if (CLIP && it > 0 && ca.static_items != 2) clip_publish(ca.sync + 1 + s_prev[0], s_prev[1] + 1, ca.release);
    }
}


#if defined(CRT_TU_PS2)      // launcher: compiled only in the translation unit that owns these kernels (build.py)
using Ps2PipeKernel = void (*)(Dev, FrameDev, const uint8_t*, uint8_t*, float*, float*, int, const CUtensorMap, const CUtensorMap, int, int, const ClipArgs);
template <bool CLIP>
inline Ps2PipeKernel pick_ps2_pipe(const Dev& d, const FrameDev& f, bool fast, bool thr) {
    Ps2PipeKernel kern = !thr ? (d.bloom_mode == 1 ? (fast ? k_fused_ps2_pipe<true, true, false, 0, CLIP> : k_fused_ps2_pipe<true, false, false, 0, CLIP>)
                                                   : (fast ? k_fused_ps2_pipe<false, true, false, 0, CLIP> : k_fused_ps2_pipe<false, false, false, 0, CLIP>))
                              : (fast ? k_fused_ps2_pipe<true, true, true, 0, CLIP> : k_fused_ps2_pipe<true, false, true, 0, CLIP>);
    static const bool use_spec = env_int("CRT_SPEC", 1) != 0;
    if (use_spec && !thr && d.bloom_mode == 1 && fast) {            // feature sets with a compile-time specialisation
        if (spec_matches(SPEC_DEFAULT, d, f.flicker_on != 0, fast)) kern = k_fused_ps2_pipe<true, true, false, SPEC_DEFAULT, CLIP>;
        else if (spec_matches(SPEC_SLANTED, d, f.flicker_on != 0, fast)) kern = k_fused_ps2_pipe<true, true, false, SPEC_SLANTED, CLIP>;
    }
    return kern;
}

// Clip mode (see ClipArgs): frames [maps->frame, maps->frame + ca.nf) of the clip in one launch; `in` / `out` point at the first of
// them, every frame blends against the state its predecessor left (has_prev = 1).  ca.sync must be zeroed on the stream first.
inline int run_fused_ps2_clip(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, cudaStream_t st,
                              int* launches, const Ps2Maps* maps, ClipArgs& ca) {
    const bool fast = d.triad_mode == 2 && d.triad_comp && d.vig_mode <= 1;
    const bool thr = d.bloom_mode == 1 && d.thr_on;
    auto kern = pick_ps2_pipe<true>(d, f, fast, thr);
    if (env.raise((const void*)kern, P2_PIPE_SMEM) &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_PIPE_SMEM) != cudaSuccess) return 2;
    const int tiles_x = (d.W + P2_TW - 1) / P2_TW;
    const long long nitems = (long long)tiles_x * ((d.H + maps->th - 1) / maps->th) * ca.nf;
    if (ca.nf < 1 || ca.nf > CLIP_MAX_FRAMES) return 4;
    // CTAs the device holds: every CTA of the grid must be able to become resident next to its peers (items wait for items)
    int resident = 0;
    {
        const void* key = (const void*)((const char*)kern + 1);      // (the entry point itself keys the shared-memory opt-in)
        auto it = env.memo.find(key);
        if (it == env.memo.end()) {
            int per_sm = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, P2_NT, (size_t)P2_PIPE_SMEM);
            it = env.memo.emplace(key, env.sms * (per_sm > 0 ? per_sm : 1)).first;
        }
        resident = it->second;
    }
    ca.frame_bytes = (unsigned long long)d.W * d.H * 3;
    // items: CRT_CLIP_ITEMS = 1 fixed stride with per-tile flags (default; cooperative launch: every CTA resident), 0 atomic counter
    // with per-tile flags (also the fallback of 1), 2 owned tiles (no inter-CTA dependency at all, plain launch; 3: ranks permuted).
    // Owned tiles measure SLOWER (run 58-59: default chain 4K 50.2 against 48.0 us per frame, 1080p 15.5 against 12.2, configs[1]
    // 22.8 against 18.7): a CTA works its items off at a latency-bound pace whatever else runs on the SM, so the frame takes
    // ceil(tiles / CTAs) tile latencies — 2 at 1080p where the flag protocols spread 1.72 per CTA evenly.
    const int items = env_int("CRT_CLIP_ITEMS", 1);
    ca.w0_idle = env_int("CRT_CLIP_W0", 1) != 0;
    const dim3 grid((unsigned)(nitems < resident ? nitems : resident));
    cudaError_t e = cudaErrorNotSupported;
    if (items == 2 || items == 3) {      // 3: ranks permuted for a placement that puts consecutive CTAs on one SM
        ca.static_items = 2;
        ca.per_sm = 0;
        const long long ntl = nitems / ca.nf;      // tiles of a frame: every CTA owns at least one
        if (items == 3 && ntl >= resident && resident % env.sms == 0) ca.per_sm = resident / env.sms;
        e = launch_pdl(kern, dim3((unsigned)(ntl < resident ? ntl : resident)), dim3(P2_NT), (size_t)P2_PIPE_SMEM, st, false, d, f, in, out, state,
                       (float*)nullptr, 1, maps->in, maps->st, maps->frame, maps->th, ca);
    } else if (items == 1 && env_int("CRT_CLIP_COOP", 1)) {
        ca.static_items = 1;
        e = launch_coop(kern, grid, dim3(P2_NT), (size_t)P2_PIPE_SMEM, st, d, f, in, out, state, (float*)nullptr, 1, maps->in, maps->st, maps->frame, maps->th, ca);
        if (e != cudaSuccess) cudaGetLastError();
    }
    if (e != cudaSuccess && ca.static_items != 2) {
        ca.static_items = 0;
        e = launch_pdl(kern, grid, dim3(P2_NT), (size_t)P2_PIPE_SMEM, st, false, d, f, in, out, state, (float*)nullptr, 1, maps->in, maps->st, maps->frame,
                       maps->th, ca);
    }
    ++*launches;
    return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : 2;
}

inline int run_fused_ps2(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, float* q_out, int has_prev,
                         cudaStream_t st, int* launches, bool pdl = false, const Ps2Maps* maps = nullptr) {
    dim3 grid((d.W + P2_TW - 1) / P2_TW, (d.H + P2_TH - 1) / P2_TH);
    const int ntiles = (int)(grid.x * grid.y);
    const bool fast = d.triad_mode == 2 && d.triad_comp && d.vig_mode <= 1;
    // TMA-pipelined variant (4 CTAs per SM, 3 with the bloom threshold on): pays off when there is a state to fetch and
    // every CTA walks over several tiles
    static const int pipe_min_tiles = env_int("CRT_PIPE_MIN_TILES", 256);      // measured: wins at 720p (460 tiles, +1.5 %), 1080p (+15 %) and 4K, neutral at VGA (150)
    if (maps && (q_out || has_prev) && (int)grid.x * ((d.H + maps->th - 1) / maps->th) >= pipe_min_tiles) {
        const bool thr = d.bloom_mode == 1 && d.thr_on;
        auto kern = pick_ps2_pipe<false>(d, f, fast, thr);
        // the opt-in shared-memory size is a per-device, per-kernel attribute: set once per context and kernel
        if (env.raise((const void*)kern, P2_PIPE_SMEM) &&
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_PIPE_SMEM) != cudaSuccess) return 2;
        const int resident_pipe = env.sms * (thr ? 3 : 4);
        const int ntiles_th = (int)grid.x * ((d.H + maps->th - 1) / maps->th);      // tiles of maps->th rows
        const cudaError_t e = launch_pdl(kern, dim3(ntiles_th < resident_pipe ? ntiles_th : resident_pipe), dim3(P2_NT), (size_t)P2_PIPE_SMEM, st, pdl,
                                         d, f, in, out, state, q_out, has_prev, maps->in, maps->st, maps->frame, maps->th, ClipArgs{});
        ++*launches;
        return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : 2;
    }
    static const int minb = env_int("CRT_PS2_MINB", 4);
    static const int waves = env_int("CRT_PS2_WAVES", 1);
    static const bool persist = env_int("CRT_PS2_PERSIST", 1) != 0;
    const int resident = persist ? env.sms * (minb == 3 ? 3 : 4) * waves : (1 << 30);      // CTAs the GPU holds at once: SMs x MINB (launch bounds)
    const dim3 pgrid(ntiles < resident ? ntiles : resident);       // persistent 1-D grid
    auto kern = d.bloom_mode == 1 ? (fast ? (minb == 3 ? k_fused_ps2<true, true, 3> : k_fused_ps2<true, true, 4>) : k_fused_ps2<true, false, 4>)
                                  : (fast ? k_fused_ps2<false, true, 4> : k_fused_ps2<false, false, 4>);
    static const bool use_spec_plain = env_int("CRT_SPEC", 1) != 0;
    if (use_spec_plain && minb != 3 && d.bloom_mode == 1 && !d.thr_on && fast) {      // small frames (VGA: BASELINE configs[0]) and first frames
        if (spec_matches(SPEC_DEFAULT, d, f.flicker_on != 0, fast)) kern = k_fused_ps2<true, true, 4, SPEC_DEFAULT>;
        else if (spec_matches(SPEC_SLANTED, d, f.flicker_on != 0, fast)) kern = k_fused_ps2<true, true, 4, SPEC_SLANTED>;
    }
    const cudaError_t e = launch_pdl(kern, pgrid, dim3(P2_NT), 0, st, pdl, d, f, in, out, state, q_out, has_prev);
    ++*launches;
    return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : 2;
}

#endif  // CRT_TU_PS2

#endif  // __CUDACC__

}  // namespace crt
