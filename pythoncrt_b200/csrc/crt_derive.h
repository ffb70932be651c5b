// crt_derive.h — host-only derivation of the kernel parameter blocks from the C-ABI
// structs.  No CUDA calls: shared by crt_abi.cu (device pointers) and by
// tests/host_emu (host pointers).
#pragma once
#include <cmath>
#include <string>
#include <vector>

#include "../../include/crt_b200.h"
#include "crt_math.cuh"

namespace crt {

struct TablePtrs {
    const void* tab[CRT_TABLE_COUNT];
    size_t bytes[CRT_TABLE_COUNT];
    const Lerp1 *dn_x, *dn_y, *up_x, *up_y, *nz_x, *nz_y;
    const float* pow_tab;    // [POW_TAB_FLOATS] from fill_pow_table, in the caller's address space
    int pix_uniform;     // see Dev::pix_uniform (the caller inspects its host copy of the pixelate tables)
};

// cv2.resize INTER_LINEAR coordinates along one axis (oracle/cv_restated.py linear_coords)
inline std::vector<Lerp1> linear_coords(int n_dst, int n_src) {
    std::vector<Lerp1> t(n_dst);
    const double scale = (double)n_src / (double)n_dst;
    for (int i = 0; i < n_dst; ++i) {
        double fx = ((double)i + 0.5) * scale - 0.5;
        int s0 = (int)std::floor(fx);
        double w = fx - (double)s0;
        if (s0 < 0) { s0 = 0; w = 0.0; }
        if (s0 >= n_src - 1) { s0 = n_src - 1; w = 0.0; }
        t[i].s0 = s0; t[i].s1 = s0 + 1 < n_src ? s0 + 1 : n_src - 1; t[i].w = (float)w;
    }
    return t;
}

inline double py_round(double x) { return std::nearbyint(x); }   // round-half-even, like Python's round()

inline bool bloom_active(const crt_params& p) { return p.bloom_strength > 0.0 && (p.bloom_sigma > 0.0 || p.fast_bloom); }
inline bool glitch_active(const crt_params& p) { return p.glitch_amp_px > 0 && p.glitch_height_frac > 0.0; }

// Derive the device parameter block from crt_params + tables, narrowing doubles to
// float32 exactly where numpy does (weak Python scalars take the array dtype).
// Table pointers are whatever address space the caller computes in.
inline int derive_dev(const crt_params& p, int W, int H, const TablePtrs& t, Dev* out, std::string* err) {
    Dev d{};
    d.W = W; d.H = H;
    d.bgr = p.channel_order == CRT_ORDER_BGR;
    // aberration rolls the R plane by +a and the B plane by -a (:573-575): with B at index 0 the signs swap
    d.aberr = d.bgr ? -p.aberration_px : p.aberration_px;
    d.aberr_mod = ((d.aberr % W) + W) % W;
    if (p.pixel_size > 1) {
        if (t.bytes[CRT_TABLE_PIXELATE_X] != (size_t)W * 4 || t.bytes[CRT_TABLE_PIXELATE_Y] != (size_t)H * 4)
            { *err = "pixel_size > 1 needs CRT_TABLE_PIXELATE_X [W] and _Y [H]"; return CRT_ERR_INVALID; }
        d.pix_x = (const int32_t*)t.tab[CRT_TABLE_PIXELATE_X];
        d.pix_y = (const int32_t*)t.tab[CRT_TABLE_PIXELATE_Y];
        d.pix_uniform = t.pix_uniform;
    }
    d.col_sat = p.saturation != 1.0; d.sat_f = (float)p.saturation;
    d.col_temp = p.temperature != 0.0;
    d.gain0 = (float)fmin(fmax(1.0 + 0.5 * p.temperature, 0.5), 1.5);          // R gain (:296), applied to the index that holds R
    d.gain2 = (float)fmin(fmax(1.0 - 0.5 * p.temperature, 0.5), 1.5);          // B gain (:297)
    if (d.bgr) { const float g = d.gain0; d.gain0 = d.gain2; d.gain2 = g; }
    d.col_bc = p.brightness != 0.0 || p.contrast != 1.0;
    d.contrast = (float)p.contrast; d.brightness = (float)p.brightness;
    d.col_gamma = p.gamma != 1.0 && p.gamma > 0.0;
    d.inv_gamma = d.col_gamma ? (float)(1.0 / p.gamma) : 1.0f;
    d.pow_tab = t.pow_tab;
    d.text_mode = p.text_mode;
    if (p.text_mode) {
        if (t.bytes[CRT_TABLE_TEXT_RGBA] != (size_t)W * H * 4) { *err = "text_mode needs CRT_TABLE_TEXT_RGBA [H][W][4]"; return CRT_ERR_INVALID; }
        d.text = (const uint8_t*)t.tab[CRT_TABLE_TEXT_RGBA];
    }
    // bloom
    d.hw = W / 2 > 1 ? W / 2 : 1; d.hh = H / 2 > 1 ? H / 2 : 1;
    d.even_dims = (W % 2 == 0) && (H % 2 == 0) && W >= 4 && H >= 4;
    if (bloom_active(p)) {
        d.bloom_mode = p.fast_bloom ? 1 : 2;
        d.bloom_strength = (float)p.bloom_strength;
        d.thr_on = p.bloom_threshold > 0.0;
        double thr = fmin(0.99, fmax(0.0, p.bloom_threshold));
        d.thr = (float)thr; d.thr_den = (float)fmax(1e-6, 1.0 - thr);
        d.thr_rcp = rcp_rn(d.thr_den);
        if (d.bloom_mode == 2) {
            int k = (int)py_round(p.bloom_sigma * 3.0) * 2 + 1;
            d.ksize = k > 1 ? k : 1;
            if (t.bytes[CRT_TABLE_GAUSS_TAPS] != (size_t)d.ksize * 4)
                { *err = "gaussian bloom needs CRT_TABLE_GAUSS_TAPS with " + std::to_string(d.ksize) + " float32 taps"; return CRT_ERR_INVALID; }
            d.taps = (const float*)t.tab[CRT_TABLE_GAUSS_TAPS];
            if (d.ksize / 2 > 46) { *err = "gaussian bloom kernel larger than 93 taps (sigma > ~15) is not supported"; return CRT_ERR_UNSUPPORTED; }
        } else {
            d.dn_x = t.dn_x; d.dn_y = t.dn_y; d.up_x = t.up_x; d.up_y = t.up_y;
        }
    }
    // triad
    if (p.triad_on) {
        if (t.bytes[CRT_TABLE_TRIAD_COLS] != (size_t)W * 3 * 4) { *err = "triad_on needs CRT_TABLE_TRIAD_COLS [W][3] float32"; return CRT_ERR_INVALID; }
        d.triad_cols = (const float*)t.tab[CRT_TABLE_TRIAD_COLS];
        const double g = p.triad_gamma;
        if ((!p.triad_preserve_luma && fabs(g - 1.0) < 1e-3) || g <= 0.0) d.triad_mode = 1;
        else {
            d.triad_mode = p.triad_preserve_luma ? 3 : 2;
            if (t.bytes[CRT_TABLE_LUT_FWD] != 1025 * 4 || t.bytes[CRT_TABLE_LUT_INV] != 1025 * 4)
                { *err = "triad needs CRT_TABLE_LUT_FWD / CRT_TABLE_LUT_INV [1025] float32"; return CRT_ERR_INVALID; }
            d.lut_fwd = (const float*)t.tab[CRT_TABLE_LUT_FWD];
            d.lut_inv = (const float*)t.tab[CRT_TABLE_LUT_INV];
        }
    }
    // scanlines
    if (p.scanline_strength > 0.0) {
        d.scan_mode = (p.scanline_angle == 0.0 && p.scanline_thickness == 1.0) ? 1 : 2;
        d.scan_strength = (float)p.scanline_strength;
        const double period = fmax(1e-6, p.scanline_period_px);
        d.scan_c32 = (float)(2.0 * M_PI / period);
        d.scan_tan = tan(p.scanline_angle * (M_PI / 180.0));
        d.scan_inv_period = 1.0 / period;
        d.scan_inv_sharp = (float)(1.0 / fmin(fmax(p.scanline_thickness, 0.1), 4.0));
    }
    // vignette
    d.vig_mode = p.vignette_on;
    if (p.vignette_on == 1) {
        d.vig_strength = (float)p.vignette_strength;
        d.vig_cx = (float)((W - 1) / 2.0); d.vig_cy = (float)((H - 1) / 2.0);
        d.vig_irx = (float)(1.0 / fmax(1.0, W / 2.0)); d.vig_iry = (float)(1.0 / fmax(1.0, H / 2.0));
    } else if (p.vignette_on == 2) {
        if (t.bytes[CRT_TABLE_VIGNETTE_PLANE] != (size_t)W * H * 4) { *err = "vignette_on=2 needs CRT_TABLE_VIGNETTE_PLANE [H][W] float32"; return CRT_ERR_INVALID; }
        d.vig_plane = (const float*)t.tab[CRT_TABLE_VIGNETTE_PLANE];
    }
    // noise
    d.noise_on = p.noise_strength > 0.0;
    d.noise_scale = (float)(p.noise_strength / 255.0);
    d.gh = H; d.gw = W;
    if (d.noise_on && p.grain_size > 1) {
        d.gh = H / p.grain_size > 1 ? H / p.grain_size : 1;
        d.gw = W / p.grain_size > 1 ? W / p.grain_size : 1;
        if (!t.nz_x || !t.nz_y) { *err = "grain tables missing"; return CRT_ERR_INVALID; }
        d.nz_x = t.nz_x; d.nz_y = t.nz_y;
    }
    // warp
    d.warp_on = p.warp_strength != 0.0;
    const double cx = (W - 1) / 2.0, cy = (H - 1) / 2.0;
    d.warp_cx = (float)cx; d.warp_cy = (float)cy;
    d.warp_dx = (float)fmax(1.0, cx); d.warp_dy = (float)fmax(1.0, cy);
    d.warp_k = (float)(p.warp_strength * 0.5);
    d.warp_mono = d.warp_k > -0.2f;      // d(map)/dx = 1 + k (3 x^2 + y^2) > 0 for |x|, |y| <= 1
    d.persist = (float)p.persistence; d.persist_q = (float)(1.0 - p.persistence);
    *out = d;
    return CRT_OK;
}

struct GlitchGeom { int y0, rows, seg_len, nseg; };
// Geometry of the glitch band (crt_filter.py:667-669, :843-844)
inline GlitchGeom glitch_geom(const crt_params& p, int W, int H) {
    GlitchGeom g{H, 0, W, 1};
    if (!glitch_active(p)) return g;
    int y0 = H - (int)(H * p.glitch_height_frac);
    y0 = y0 < 0 ? 0 : (y0 > H ? H : y0);
    g.y0 = y0; g.rows = H - y0;
    if (p.variant == CRT_VARIANT_EXPORT) {
        int s = W >= 120 ? W / 120 : 8;
        g.seg_len = s < 8 ? 8 : (s > 32 ? 32 : s);
        g.nseg = (W + g.seg_len - 1) / g.seg_len;
    }
    return g;
}

// Composite triad tables (Dev::triad_comp): detects the regular interior pattern of the mask
// (bright on x % 3 == channel, dim elsewhere, identical on every interior column) and
// pre-applies mask multiply + inverse LUT to every forward-LUT entry with the kernels' own
// float32 operations.  Returns false when the mask is not of that form.
inline bool build_triad_comp(const float* cols, int W, const float* fwd, const float* inv, std::vector<float>* comp, int* x0, int* x1) {
    if (W < 12) return false;
    const int xm = W / 2;
    const float bright = cols[xm * 3 + xm % 3], dim = cols[xm * 3 + (xm + 1) % 3];
    auto regular = [&](int x) {
        for (int c = 0; c < 3; ++c) if (cols[x * 3 + c] != (x % 3 == c ? bright : dim)) return false;
        return true;
    };
    int a = xm, b = xm;
    while (a > 0 && regular(a - 1)) --a;
    while (b < W - 1 && regular(b + 1)) ++b;
    if (!regular(xm) || b - a + 1 < W / 2) return false;
    comp->resize(2 * 1025);
    for (int t = 0; t < 2; ++t) {
        const float m = t == 0 ? bright : dim;
        for (int i = 0; i < 1025; ++i) (*comp)[t * 1025 + i] = sat(inv[lut_index(fmul(fwd[i], m))]);
    }
    *x0 = a; *x1 = b;
    return true;
}

// Per-frame scalars (flicker factor: crt_filter.py:632, evaluated in double).
inline FrameDev derive_frame(const crt_params& p, const crt_frame& fr) {
    FrameDev f{};
    f.phase = fr.phase_px; f.phase32 = (float)fr.phase_px;
    f.flicker_on = p.flicker_strength > 0.0 && p.flicker_hz > 0.0;
    f.flicker = f.flicker_on ? (float)(1.0 + 0.25 * p.flicker_strength * sin(2.0 * M_PI * p.flicker_hz * fr.time_sec)) : 1.0f;
    return f;
}

}  // namespace crt
