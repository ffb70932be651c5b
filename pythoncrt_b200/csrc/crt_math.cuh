// crt_math.cuh — per-pixel arithmetic of the CRT effect chain, written once and
// used by every kernel (staged and fused).
//
// Parity contract (oracle/cv_restated.py): everything that feeds
// the triad LUT's floor indexing (crt_filter.py:250, :261) is evaluated in
// float32 with numpy's operation order and NO fused multiply-add, except where
// OpenCV itself uses an fma (resize lerps, gaussian taps).  All such operations
// go through the crt_* wrappers below (__fmul_rn / __fadd_rn never contract);
// the translation unit is additionally compiled with -fmad=false.
//
// The functions are __host__ __device__ so that tests/host_emu can compile the
// very same arithmetic with g++ and compare it with the oracle on a machine
// without a GPU.  That build is test infrastructure; the product never runs it.
#pragma once

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define CRT_HD __host__ __device__ __forceinline__
#else
#define CRT_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define CRT_DEVICE_CODE 1
#else
#define CRT_DEVICE_CODE 0
#endif

#if !defined(__CUDACC__)
// plain g++ build of tests/host_emu: CUDA's sinpif/cospif do not exist in libm
static inline float sinpif(float x) { return (float)sin(3.14159265358979323846 * (double)x); }
static inline float cospif(float x) { return (float)cos(3.14159265358979323846 * (double)x); }
#endif

namespace crt {

#define CRT_POW_TABLE static const
#include "crt_pow_tables.h"     // host copies; kernels read the same values through Dev::pow_tab / shared memory
#undef CRT_POW_TABLE
constexpr int POW_TAB_EXP = 128 * 4;                 // offset of the exp2 table
constexpr int POW_TAB_FLOATS = 128 * 4 + 64 * 2;     // [log: 128 x (invc, logc_hi, logc_lo, 0) | exp2: 64 x (hi, lo)]
inline void fill_pow_table(float* t) {
    for (int i = 0; i < 128; ++i) for (int c = 0; c < 4; ++c) t[4 * i + c] = kPowLog[i][c];
    for (int j = 0; j < 64; ++j) { t[POW_TAB_EXP + 2 * j] = kPowExp2[j][0]; t[POW_TAB_EXP + 2 * j + 1] = kPowExp2[j][1]; }
}

// ---- exact float32 primitives ------------------------------------------------
CRT_HD float fmul(float a, float b) {
#if CRT_DEVICE_CODE
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
CRT_HD float fadd(float a, float b) {
#if CRT_DEVICE_CODE
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
CRT_HD float fsub(float a, float b) {
#if CRT_DEVICE_CODE
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}
CRT_HD float fdiv(float a, float b) {
#if CRT_DEVICE_CODE
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b; return r;
#endif
}
CRT_HD float ffma(float a, float b, float c) {
#if CRT_DEVICE_CODE
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
// np.clip(x, 0, 1)
CRT_HD float sat(float x) {
#if CRT_DEVICE_CODE
    return __saturatef(x);
#else
    return x < 0.f ? 0.f : (x > 1.f ? 1.f : x);
#endif
}
CRT_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
CRT_HD int imin(int a, int b) { return a < b ? a : b; }
CRT_HD int imax(int a, int b) { return a > b ? a : b; }
// a mod n for a in [-n, 2n)
CRT_HD int wrap(int a, int n) { a = a < 0 ? a + n : a; return a >= n ? a - n : a; }
// python-style modulo for arbitrary a
CRT_HD int pymod(int a, int n) { int r = a % n; return r < 0 ? r + n : r; }

struct F3 { float x, y, z; };
CRT_HD F3 mk3(float a, float b, float c) { F3 r; r.x = a; r.y = b; r.z = c; return r; }

// One axis of a cv2.resize(INTER_LINEAR) coordinate table (oracle/cv_restated.py linear_coords)
struct Lerp1 { int32_t s0, s1; float w; };

// ---- device-side parameter block ------------------------------------------------
struct Dev {
    int W, H;
    int bgr;                     // crt_params.channel_order == CRT_ORDER_BGR: channel index 0 is B, index 2 is R
    // stage 1-2: aberration + pixelate (crt_filter.py:571-584)
    int aberr;                   // aberration_px
    int aberr_mod;               // aberration_px mod W, in [0, W): lets the kernels wrap with one compare
    int pix_uniform;             // pixel_size when pix_x[x] == ps*(x/ps) and pix_y likewise (lets tiles de-duplicate), else 0
    const int32_t* pix_x;        // [W] or null
    const int32_t* pix_y;        // [H] or null
    // stage 3: colour (:279-305)
    int col_sat, col_temp, col_bc, col_gamma;
    float sat_f, gain0, gain2, contrast, brightness, inv_gamma;
    const float* pow_tab;        // [POW_TAB_FLOATS] tables of pow_unit (16-byte aligned)
    // text layer (:588-598 / :653-663)
    int text_mode;               // 0 none, 1 before, 2 after
    const uint8_t* text;         // [H][W][4]
    // stage 5: bloom (:599-612)
    int bloom_mode;              // 0 off, 1 fast, 2 gaussian
    float bloom_strength;
    int thr_on; float thr, thr_den;
    float thr_rcp;               // correctly rounded 1 / thr_den (div_const)
    int ksize; const float* taps;
    int hw, hh;                  // half-size plane of the fast path
    int even_dims;               // W and H even: closed-form 2x coordinates
    const Lerp1* dn_x; const Lerp1* dn_y;   // [hw], [hh]   (general sizes)
    const Lerp1* up_x; const Lerp1* up_y;   // [W], [H]
    // stage 6: triad (:238-263)
    int triad_mode;              // 0 off, 1 plain multiply, 2 LUT, 3 LUT + preserve luma
    const float* triad_cols;     // [W][3]
    const float* lut_fwd; const float* lut_inv;   // [1025]
    // Composite tables for the interior columns of a regular triad mask (triad_mode 2 only):
    // comp[0][i] = inv[idx(fwd[i] * bright)], comp[1][i] = inv[idx(fwd[i] * dim)] — the same
    // float32 operations, evaluated once per table entry instead of once per pixel.
    const float* triad_comp;     // [2][1025] or null
    int comp_x0, comp_x1;        // columns [comp_x0, comp_x1] follow the (x % 3 == channel ? bright : dim) pattern
    // stage 7: scanlines (:213-217, :308-328)
    int scan_mode;               // 0 off, 1 per-row float32, 2 slanted/shaped plane
    float scan_strength, scan_c32;
    double scan_tan, scan_inv_period;
    float scan_inv_sharp;
    // stage 8: vignette (:266-276)
    int vig_mode;                // 0 off, 1 analytic, 2 plane
    float vig_strength, vig_cx, vig_cy, vig_irx, vig_iry;
    const float* vig_plane;
    // stage 10: noise (:635-648)
    int noise_on; float noise_scale; int gh, gw;
    const Lerp1* nz_x; const Lerp1* nz_y;   // [W], [H] when grain > 1, else null
    // stage 11: warp (:331-348)
    int warp_on; float warp_cx, warp_cy, warp_dx, warp_dy, warp_k;
    int warp_mono;               // map is monotone in x and in y over the frame: a tile's footprint is spanned by its perimeter
    // stage 14: persistence (:687-694 / :1086-1096)
    float persist, persist_q;    // p and float32(1 - p)
};

struct FrameDev {
    float phase32;               // float32(scanline_phase_px)
    double phase;                // scanline_phase_px
    int flicker_on; float flicker;
    const float* noise;          // [gh][gw] N(0,1) plane or null
    const int32_t* goffs;        // [rows][segments] or null
    int gy0, gseg, gnseg;
};

// the per-frame scalars of FrameDev for a run of frames processed by ONE launch (clip mode, crt_fused_ps2.cuh)
struct FrameVar { float phase32, flicker; double phase; };

// ---- stage 0-4: graded input ----------------------------------------------------
// u8 -> float32 by true division (crt_filter.py:569)
CRT_HD float unit(uint8_t v) { return fdiv((float)v, 255.0f); }

// Raw source bytes of pixel (y, x) after aberration (:571-577: channel 0 rolled by
// +a, channel 2 by -a, circular) and pixelate (:578-584).
CRT_HD void source_bytes(const Dev& d, const uint8_t* __restrict__ in, int y, int x, uint8_t& b0, uint8_t& b1, uint8_t& b2) {
    int sy = d.pix_y ? d.pix_y[y] : y;
    int sx = d.pix_x ? d.pix_x[x] : x;
    const uint8_t* row = in + (size_t)sy * d.W * 3;
    int x0 = sx, x2 = sx;
    if (d.aberr != 0) { x0 = wrap(sx - d.aberr_mod, d.W); x2 = wrap(sx + d.aberr_mod, d.W); }
    b0 = row[x0 * 3 + 0];
    b1 = row[sx * 3 + 1];
    b2 = row[x2 * 3 + 2];
}

// ---- x^y for the colour gamma (:304) ------------------------------------------------------------
// numpy evaluates np.power(img, 1/gamma, dtype=float32) with SVML: within 1 ulp, correctly rounded
// for only ~80 % of inputs, not reproducible elsewhere.  pow_unit is a table-driven float32 scheme that
// keeps double-float (hi, lo) precision only where it is needed:
//   log2 x = (k + logc_hi) + lo          128 mantissa sub-intervals: r = m * invc_i - 1 is ONE fma (|r| <= 2^-8, so
//                                        its rounding error is 2^-33), lo = logc_lo + c1 r - c1 r^2/2 + c1 r^3/3 in plain
//                                        float32 (|lo| < 2^-7: errors ~2^-32), k + logc_hi exact (16 fraction bits);
//   E = y * log2 x                       E0 = y * hi rounded, its exact residual by fma, El = y * lo + residual;
//   x^y = 2^n * T_j * 2^g                n + j / 64 split off E0 with a magic-number add (exact), g = f + El,
//                                        T_j = 2^(j/64) as (hi, lo), 2^g - 1 by a degree-3 polynomial, one final rounding.
// 37 float32 / integer instructions (the 32-entry-table version this replaces: ~65, with Fast2Sum chains for a
// 2^-6 residual), no FP64.  Error < 0.52 ulp over the GUI's gamma range; tests/test_host_emu.py compares with float64.
// Domain: x in [0, 1] (the stage input is clipped), y > 0.  Results below 2^-125 flush to 0.
CRT_HD uint32_t f2u(float f) {
#if CRT_DEVICE_CODE
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
CRT_HD float u2f(uint32_t u) {
#if CRT_DEVICE_CODE
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
CRT_HD float pow_unit(float x, float y, const float* __restrict__ T) {
    // Straight-line code (the two special cases are one select at the end), so that the three channels'
    // evaluations interleave in the instruction stream.
    const uint32_t ix = f2u(x), tmp = ix - 0x3f328000u;
    const uint32_t i = (tmp >> 16) & 127u;
    const uint32_t top = tmp & 0xff800000u;
    const float m = u2f(ix - top);                                         // x = 2^k * m, m in [0.697, 1.395)
    const float kf = (float)((int32_t)top >> 23);
#if CRT_DEVICE_CODE
    const float4 lt = *reinterpret_cast<const float4*>(T + 4 * i);
    const float invc = lt.x, lch = lt.y, lcl = lt.z;
#else
    const float invc = T[4 * i], lch = T[4 * i + 1], lcl = T[4 * i + 2];
#endif
    const float r = ffma(m, invc, -1.0f);                                  // |r| <= 2^-8: one rounding of 2^-33
    const float r2 = fmul(r, r);
    float lo = ffma(r2, ffma((float)(1.4426950408889634 / 3.0), r, (float)(-1.4426950408889634 / 2.0)), lcl);
    lo = ffma(0x1.715476p+0f, r, lo);                                      // + r / ln 2
    const float hi = fadd(kf, lch);                                        // exact
    const float E0 = fmul(y, hi);
    float El = ffma(y, lo, ffma(y, hi, -E0));                              // y * lo + (y * hi - E0): up to y * 2^-7
    const float Eh = fadd(E0, El);                                         // renormalise (Fast2Sum): |El| <= ulp(Eh) / 2 for any y
    El = fsub(El, fsub(Eh, E0));
    const float t = fadd(Eh, 196608.0f);                                   // 1.5 * 2^17: ulp 2^-6, the mantissa holds round(64 Eh)
    const uint32_t ki = f2u(t);
    const float g = fadd(fsub(Eh, fsub(t, 196608.0f)), El);                // exact difference (|.| <= 2^-7) + El
    const uint32_t j = ki & 63u;
#if CRT_DEVICE_CODE
    const float2 tj = *reinterpret_cast<const float2*>(T + POW_TAB_EXP + 2 * j);
    const float th = tj.x, tl = tj.y;
#else
    const float th = T[POW_TAB_EXP + 2 * j], tl = T[POW_TAB_EXP + 2 * j + 1];
#endif
    float e = ffma((float)0.05550410866482158, g, (float)0.2402265069591007);      // ln2^3 / 6, ln2^2 / 2
    e = ffma(e, g, (float)0.6931471805599453);
    e = fmul(e, g);                                                        // 2^g - 1
    const float res = fadd(th, ffma(th, e, tl));
    const float out = u2f(f2u(res) + ((ki & ~63u) << 17));                // * 2^n, n = (round(64 E0) - j) / 64
    // x = 0 (and sub-normals, which the chain never produces) -> 0; results below 2^-125 -> 0
    return (x >= 1.17549435e-38f && Eh >= -125.0f) ? out : 0.0f;
}

// apply_color_adjustments (:279-305), float32, numpy operation order.
CRT_HD F3 colour(const Dev& d, F3 v, const float* __restrict__ pow_tab) {
    if (d.col_sat) {
        // Rec.709 luma in the reference's association ((wR R + wG G) + wB B), whichever index holds R
        const float cr = d.bgr ? v.z : v.x, cb = d.bgr ? v.x : v.z;
        float l = fadd(fadd(fmul(0.2126f, cr), fmul(0.7152f, v.y)), fmul(0.0722f, cb));
        v.x = sat(fadd(l, fmul(fsub(v.x, l), d.sat_f)));
        v.y = sat(fadd(l, fmul(fsub(v.y, l), d.sat_f)));
        v.z = sat(fadd(l, fmul(fsub(v.z, l), d.sat_f)));
    }
    if (d.col_temp) {
        v.x = sat(fmul(v.x, d.gain0));
        v.z = sat(fmul(v.z, d.gain2));
    }
    if (d.col_bc) {
        v.x = sat(fadd(fadd(fmul(fsub(v.x, 0.5f), d.contrast), 0.5f), d.brightness));
        v.y = sat(fadd(fadd(fmul(fsub(v.y, 0.5f), d.contrast), 0.5f), d.brightness));
        v.z = sat(fadd(fadd(fmul(fsub(v.z, 0.5f), d.contrast), 0.5f), d.brightness));
    }
    if (d.col_gamma) {
        // cannot be matched bit for bit with numpy by any implementation (see pow_unit; DESIGN.md, parity)
        v.x = sat(pow_unit(v.x, d.inv_gamma, pow_tab));
        v.y = sat(pow_unit(v.y, d.inv_gamma, pow_tab));
        v.z = sat(pow_unit(v.z, d.inv_gamma, pow_tab));
    }
    return v;
}

// Alpha blend of the rasterised text layer (:595-597)
CRT_HD F3 text_blend(const Dev& d, F3 v, int y, int x) {
    const uint8_t* t = d.text + ((size_t)y * d.W + x) * 4;
    float a = unit(t[3]);
    float ia = fsub(1.0f, a);
    v.x = sat(fadd(fmul(v.x, ia), fmul(unit(t[d.bgr ? 2 : 0]), a)));      // the layer is RGBA
    v.y = sat(fadd(fmul(v.y, ia), fmul(unit(t[1]), a)));
    v.z = sat(fadd(fmul(v.z, ia), fmul(unit(t[d.bgr ? 0 : 2]), a)));
    return v;
}

// Stages 0-4 at pixel (y, x): the image bloom sees.
CRT_HD F3 graded_input(const Dev& d, const uint8_t* __restrict__ in, int y, int x) {
    uint8_t b0, b1, b2;
    source_bytes(d, in, y, x, b0, b1, b2);
    F3 v = colour(d, mk3(unit(b0), unit(b1), unit(b2)), d.pow_tab);
    if (d.text_mode == 1) v = text_blend(d, v, y, x);
    return v;
}

// Same, with the u8 -> float32 conversion read from a 256-entry table of unit() values
// (identical results; saves three IEEE divisions per pixel in the fused kernel).
CRT_HD F3 graded_input_lut(const Dev& d, const uint8_t* __restrict__ in, int y, int x, const float* __restrict__ unit_lut,
                           const float* __restrict__ pow_tab) {
    uint8_t b0, b1, b2;
    source_bytes(d, in, y, x, b0, b1, b2);
    F3 v = colour(d, mk3(unit_lut[b0], unit_lut[b1], unit_lut[b2]), pow_tab);
    if (d.text_mode == 1) v = text_blend(d, v, y, x);
    return v;
}

// Same for a pixel whose pixelate source (sy, sx) is already known (regular pixelate tables:
// the block origin), skipping the index-table loads.
CRT_HD F3 graded_source_lut(const Dev& d, const uint8_t* __restrict__ in, int sy, int sx, int y, int x, const float* __restrict__ unit_lut,
                            const float* __restrict__ pow_tab) {
    const uint8_t* row = in + (size_t)sy * d.W * 3;
    int x0 = sx, x2 = sx;
    if (d.aberr != 0) { x0 = wrap(sx - d.aberr_mod, d.W); x2 = wrap(sx + d.aberr_mod, d.W); }
    F3 v = colour(d, mk3(unit_lut[row[x0 * 3 + 0]], unit_lut[row[sx * 3 + 1]], unit_lut[row[x2 * 3 + 2]]), pow_tab);
    if (d.text_mode == 1) v = text_blend(d, v, y, x);
    return v;
}

// n / d for a divisor fixed per clip, bit-identical to the IEEE quotient: q = RN(n * y) with y = RN(1 / d), the exact
// residual r = n - q d (one fma), q' = RN(q + r y).  Correctly rounded whenever y is (Markstein); checked exhaustively
// against the hardware division for every float n in (1e-9, d] and fifteen divisors (3.6e9 quotients, no mismatch) and
// sampled in tests/test_host_emu.py.  3 instructions instead of the ~9 of a general IEEE division.
CRT_HD float div_const(float n, float d, float y) {
    const float q = fmul(n, y);
    return ffma(ffma(-q, d, n), y, q);
}
// Correctly rounded float32 reciprocal of d (host): the double quotient narrowed, checked against both neighbours.
inline float rcp_rn(float d) {
    float best = (float)(1.0 / (double)d);
    double err = fabs(1.0 - (double)best * (double)d);           // products of two float32 are exact in double
    const float cand[2] = {nextafterf(best, 0.0f), nextafterf(best, INFINITY)};
    for (float c : cand) {
        const double e = fabs(1.0 - (double)c * (double)d);
        if (e < err) { err = e; best = c; }
    }
    return best;
}
// Bloom source: clip((img - thr) / max(1e-6, 1 - thr)) (:602-604)
CRT_HD float bloom_src1(const Dev& d, float v) { return d.thr_on ? sat(div_const(fsub(v, d.thr), d.thr_den, d.thr_rcp)) : v; }
CRT_HD F3 bloom_src(const Dev& d, F3 v) { return mk3(bloom_src1(d, v.x), bloom_src1(d, v.y), bloom_src1(d, v.z)); }

// cv2.resize lerp: fma(q - p, w, p)
CRT_HD float lerp_cv(float p, float q, float w) { return ffma(fsub(q, p), w, p); }

// Coordinates of the 2x fast-bloom resizes (closed form for even sizes, else tables).
CRT_HD Lerp1 down_coord(const Dev& d, const Lerp1* tab, int i) {
    if (d.even_dims) { Lerp1 c; c.s0 = 2 * i; c.s1 = 2 * i + 1; c.w = 0.5f; return c; }
    return tab[i];
}
CRT_HD Lerp1 up_coord(const Dev& d, const Lerp1* tab, int i, int n_half) {
    if (d.even_dims) {
        Lerp1 c;
        int j = i >> 1;
        if (i & 1) { c.s0 = j; c.w = 0.25f; } else { c.s0 = j - 1; c.w = 0.75f; }
        if (c.s0 < 0) { c.s0 = 0; c.w = 0.f; }
        if (c.s0 >= n_half - 1) { c.s0 = n_half - 1; c.w = 0.f; }
        c.s1 = imin(c.s0 + 1, n_half - 1);
        return c;
    }
    return tab[i];
}

// addition of bloom: clip(img + strength * blur) (:611)
CRT_HD F3 add_bloom(const Dev& d, F3 v, F3 bl) {
    v.x = sat(fadd(v.x, fmul(d.bloom_strength, bl.x)));
    v.y = sat(fadd(v.y, fmul(d.bloom_strength, bl.y)));
    v.z = sat(fadd(v.z, fmul(d.bloom_strength, bl.z)));
    return v;
}

// ---- stage 6: triad -------------------------------------------------------------
CRT_HD int lut_index(float v) {           // (np.clip(v,0,1) * 1024).astype(int32); the reference's outer clip to
    return (int)fmul(sat(v), 1024.0f);    // [0, 1024] is a no-op: the product of a value in [0, 1] never exceeds 1024
}
CRT_HD F3 triad(const Dev& d, F3 v, int x, const float* __restrict__ fwd, const float* __restrict__ inv) {
    const float* m = d.triad_cols + (size_t)x * 3;          // the table is in the reference's RGB order
    float m0 = m[d.bgr ? 2 : 0], m1 = m[1], m2 = m[d.bgr ? 0 : 2];
    if (d.triad_mode == 1) return mk3(sat(fmul(v.x, m0)), sat(fmul(v.y, m1)), sat(fmul(v.z, m2)));
    float l0 = fwd[lut_index(v.x)], l1 = fwd[lut_index(v.y)], l2 = fwd[lut_index(v.z)];
    float o0 = fmul(l0, m0), o1 = fmul(l1, m1), o2 = fmul(l2, m2);
    if (d.triad_mode == 3) {
        const float lr = d.bgr ? l2 : l0, lb = d.bgr ? l0 : l2, orr = d.bgr ? o2 : o0, ob = d.bgr ? o0 : o2;
        float before = fadd(fadd(fmul(0.2126f, lr), fmul(0.7152f, l1)), fmul(0.0722f, lb));
        float after = fadd(fadd(fmul(0.2126f, orr), fmul(0.7152f, o1)), fmul(0.0722f, ob));
        float ratio = clampf(fdiv(before, fmaxf(after, 1e-6f)), 0.5f, 2.0f);
        o0 = fmul(o0, ratio); o1 = fmul(o1, ratio); o2 = fmul(o2, ratio);
    }
    return mk3(sat(inv[lut_index(o0)]), sat(inv[lut_index(o1)]), sat(inv[lut_index(o2)]));
}

// ---- stage 7: scanlines ---------------------------------------------------------
// Per-row mask (:213-217): float32 throughout, as numpy evaluates it.
CRT_HD float scan_row(const Dev& d, const FrameDev& f, int y) {
    float arg = fmul(d.scan_c32, fadd((float)y, f.phase32));
    float s = fmul(0.5f, fadd(1.0f, sinf(arg)));
    return fsub(1.0f, fmul(d.scan_strength, s));
}
// Slanted / shaped mask (:308-328): the reference works in float64 and casts to
// float32; the phase is reduced in double (exact), sin/pow run in float32.
CRT_HD float scan_plane(const Dev& d, const FrameDev& f, int y, int x) {
    double u = ((double)y + d.scan_tan * (double)x) + f.phase;
    double t = u * d.scan_inv_period;
    t -= floor(t);                                  // turns in [0, 1)
    float s = 0.5f * (1.0f + sinpif(2.0f * (float)t));
    s = fmaxf(s, 0.0f);
    float shaped = (d.scan_inv_sharp == 1.0f) ? s : powf(s, d.scan_inv_sharp);
    return 1.0f - d.scan_strength * shaped;
}

// ---- stage 8: vignette ----------------------------------------------------------
CRT_HD float vignette(const Dev& d, int y, int x) {
    if (d.vig_mode == 2) return d.vig_plane[(size_t)y * d.W + x];
    float nx = ((float)x - d.vig_cx) * d.vig_irx;
    float ny = ((float)y - d.vig_cy) * d.vig_iry;
    return 1.0f - d.vig_strength * sat(nx * nx + ny * ny);
}

// ---- stage 10: noise ------------------------------------------------------------
CRT_HD float noise_at(const Dev& d, const FrameDev& f, int y, int x) {
    float n;
    if (d.nz_x) {      // grain > 1: cv2.resize(small, (W, H), INTER_LINEAR)
        Lerp1 cx = d.nz_x[x], cy = d.nz_y[y];
        const float* r0 = f.noise + (size_t)cy.s0 * d.gw;
        const float* r1 = f.noise + (size_t)cy.s1 * d.gw;
        float a = lerp_cv(r0[cx.s0], r0[cx.s1], cx.w);
        float b = lerp_cv(r1[cx.s0], r1[cx.s1], cx.w);
        n = lerp_cv(a, b, cy.w);
    } else {
        n = f.noise[(size_t)y * d.W + x];
    }
    return fmul(n, d.noise_scale);
}

// Stages 6-10 at pixel (y, x) given the image after bloom.
CRT_HD F3 after_bloom(const Dev& d, const FrameDev& f, F3 v, int y, int x,
                      const float* __restrict__ fwd, const float* __restrict__ inv, float row_mask) {
    if (d.triad_mode) v = triad(d, v, x, fwd, inv);
    if (d.scan_mode) {
        float m = d.scan_mode == 1 ? row_mask : scan_plane(d, f, y, x);
        v.x = sat(v.x * m); v.y = sat(v.y * m); v.z = sat(v.z * m);
    }
    if (d.vig_mode) {
        float m = vignette(d, y, x);
        v.x = sat(v.x * m); v.y = sat(v.y * m); v.z = sat(v.z * m);
    }
    if (f.flicker_on) { v.x = sat(v.x * f.flicker); v.y = sat(v.y * f.flicker); v.z = sat(v.z * f.flicker); }
    if (d.noise_on) {
        float n = noise_at(d, f, y, x);
        v.x = sat(v.x + n); v.y = sat(v.y + n); v.z = sat(v.z + n);
    }
    return v;
}

// ---- stage 11: warp -------------------------------------------------------------
struct Taps { int ix, iy; float w00, w01, w10, w11; };
// Source coordinates of apply_barrel_warp (:338-346) in numpy's float32 order,
// then cv2.remap's 1/32-pixel quantisation (oracle/cv_restated.py remap_bilinear_const0).
CRT_HD Taps warp_taps(const Dev& d, int y, int x) {
    float xn = fdiv(fsub((float)x, d.warp_cx), d.warp_dx);
    float yn = fdiv(fsub((float)y, d.warp_cy), d.warp_dy);
    float r2 = fadd(fmul(xn, xn), fmul(yn, yn));
    float fac = fadd(1.0f, fmul(d.warp_k, r2));
    float mx = fadd(fmul(fmul(xn, fac), d.warp_cx), d.warp_cx);
    float my = fadd(fmul(fmul(yn, fac), d.warp_cy), d.warp_cy);
    // cv2 saturates the fixed-point coordinate to the int range first
    float qx = clampf(fmul(mx, 32.0f), -1.0e9f, 1.0e9f), qy = clampf(fmul(my, 32.0f), -1.0e9f, 1.0e9f);
    int sx = (int)rintf(qx), sy = (int)rintf(qy);
    Taps t;
    t.ix = sx >> 5; t.iy = sy >> 5;
    float fx = fmul((float)(sx & 31), 0.03125f), fy = fmul((float)(sy & 31), 0.03125f);
    float gx = fsub(1.0f, fx), gy = fsub(1.0f, fy);
    t.w00 = fmul(gy, gx); t.w01 = fmul(gy, fx); t.w10 = fmul(fy, gx); t.w11 = fmul(fy, fx);
    return t;
}
CRT_HD float gather4(float a, float b, float c, float e, const Taps& t) {
    return fadd(fadd(fadd(fmul(a, t.w00), fmul(b, t.w01)), fmul(c, t.w10)), fmul(e, t.w11));
}

// ---- stage 13: glitch -----------------------------------------------------------
// Source column of output (y, x) (:680-684 / :852-857); identity above the band.
CRT_HD int glitch_src_x(const Dev& d, const FrameDev& f, int y, int x) {
    if (!f.goffs || y < f.gy0) return x;
    int off = f.goffs[(size_t)(y - f.gy0) * f.gnseg + x / f.gseg];
    return pymod(x + off, d.W);
}

// ---- stage 14-15: persistence + quantise ------------------------------------------
// Export: clip(p*prev + (1-p)*img) (:1092); GUI: addWeighted without clip (:693) —
// identical here because both operands are already in [0, 1].
CRT_HD float blend(float prev, float v, float p, float q) { return sat(p * prev + q * v); }
// cv2.convertScaleAbs(alpha=255): round-half-even of |255 x|, saturated (:696, :1098)
CRT_HD uint8_t quantise(float v) {
    float s = fabsf(fmul(v, 255.0f));
    int r = (int)rintf(s);
    return (uint8_t)(r > 255 ? 255 : r);
}

// ---- counter-based RNG (Philox4x32-10) ----------------------------------------------
struct U4 { uint32_t x, y, z, w; };
CRT_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if CRT_DEVICE_CODE
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
CRT_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        U4 n; n.x = hi1 ^ c.y ^ k0; n.y = lo1; n.z = hi0 ^ c.w ^ k1; n.w = lo0;
        c = n; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c;
}
// Two independent N(0,1) draws from two 32-bit words (Box-Muller).
CRT_HD void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0, 1)
    float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float r = sqrtf(-2.0f * logf(u1));
    n0 = r * cospif(2.0f * u2); n1 = r * sinpif(2.0f * u2);
}
}  // namespace crt
