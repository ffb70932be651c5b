// crt_stages.cuh — per-pixel / per-cell bodies of the STAGED kernels (the general
// path that handles every parameter combination through global-memory
// intermediates).  The CUDA kernels in crt_kernels.cu are thin thread-mapping
// wrappers around these; tests/host_emu loops over the same functions on the CPU.
#pragma once
#include "crt_math.cuh"

namespace crt {

// Global-memory intermediates of the staged path for one frame.
struct Scratch {
    float* ds;      // fast bloom: half-size plane [hh][hw][3]
    float* bl;      // gaussian bloom: blurred plane [H][W][3]
    float* q;       // pre-warp processed image [H][W][3] (only when warp is on)
};

// ---- fast bloom: 2x down-scale cell (cv2.resize INTER_LINEAR, crt_filter.py:606) ----
CRT_HD void bloom_down_cell(const Dev& d, const uint8_t* __restrict__ in, float* __restrict__ ds, int j, int i) {
    Lerp1 cx = down_coord(d, d.dn_x, i), cy = down_coord(d, d.dn_y, j);
    F3 a = bloom_src(d, graded_input(d, in, cy.s0, cx.s0));
    F3 b = bloom_src(d, graded_input(d, in, cy.s0, cx.s1));
    F3 c = bloom_src(d, graded_input(d, in, cy.s1, cx.s0));
    F3 e = bloom_src(d, graded_input(d, in, cy.s1, cx.s1));
    float* o = ds + ((size_t)j * d.hw + i) * 3;
    o[0] = lerp_cv(lerp_cv(a.x, b.x, cx.w), lerp_cv(c.x, e.x, cx.w), cy.w);
    o[1] = lerp_cv(lerp_cv(a.y, b.y, cx.w), lerp_cv(c.y, e.y, cx.w), cy.w);
    o[2] = lerp_cv(lerp_cv(a.z, b.z, cx.w), lerp_cv(c.z, e.z, cx.w), cy.w);
}

// ---- fast bloom: 2x up-scale sample (crt_filter.py:607) ------------------------------
// `ds` may be any array indexed [row][col][3] with row stride `stride` floats and
// origin (j0, i0) — global plane (stride hw*3, origin 0) or a shared-memory window.
CRT_HD F3 bloom_up_at(const Dev& d, const float* __restrict__ ds, int stride, int j0, int i0, int y, int x) {
    Lerp1 cx = up_coord(d, d.up_x, x, d.hw), cy = up_coord(d, d.up_y, y, d.hh);
    const float* r0 = ds + (size_t)(cy.s0 - j0) * stride;
    const float* r1 = ds + (size_t)(cy.s1 - j0) * stride;
    int c0 = (cx.s0 - i0) * 3, c1 = (cx.s1 - i0) * 3;
    F3 o;
    o.x = lerp_cv(lerp_cv(r0[c0 + 0], r0[c1 + 0], cx.w), lerp_cv(r1[c0 + 0], r1[c1 + 0], cx.w), cy.w);
    o.y = lerp_cv(lerp_cv(r0[c0 + 1], r0[c1 + 1], cx.w), lerp_cv(r1[c0 + 1], r1[c1 + 1], cx.w), cy.w);
    o.z = lerp_cv(lerp_cv(r0[c0 + 2], r0[c1 + 2], cx.w), lerp_cv(r1[c0 + 2], r1[c1 + 2], cx.w), cy.w);
    return o;
}

// ---- gaussian bloom taps (cv2.GaussianBlur float32, oracle/cv_restated.py) --------------
// Row pass: `p` points at the tap for offset -r of one channel, taps are `step` floats apart.
CRT_HD float gauss_row(const float* __restrict__ p, int step, const float* __restrict__ k, int K) {
    if (K == 1) return fmul(p[0], k[0]);
    if (K == 3) return ffma(p[step], k[1], fmul(fadd(p[0], p[2 * step]), k[2]));
    if (K == 5) {
        float inner = ffma(p[2 * step], k[2], fmul(fadd(p[step], p[3 * step]), k[3]));
        return ffma(fadd(p[4 * step], p[0]), k[4], inner);
    }
    float s = fmul(p[0], k[0]);
    for (int i = 1; i < K; ++i) s = ffma(p[i * step], k[i], s);
    return s;
}
// Column pass: `c` points at the centre tap, rows are `step` floats apart.
CRT_HD float gauss_col(const float* __restrict__ c, int step, const float* __restrict__ k, int K) {
    int r = K >> 1;
    float s = fmul(c[0], k[r]);
    for (int i = 1; i <= r; ++i) s = ffma(fadd(c[i * step], c[-i * step]), k[r + i], s);
    return s;
}

// ---- stages 0-10 at one pixel (the image the warp samples) ------------------------------
CRT_HD F3 pre_warp_pixel(const Dev& d, const FrameDev& f, const uint8_t* __restrict__ in, const Scratch& s,
                         int y, int x, const float* __restrict__ fwd, const float* __restrict__ inv) {
    F3 v = graded_input(d, in, y, x);
    if (d.bloom_mode == 1) v = add_bloom(d, v, bloom_up_at(d, s.ds, d.hw * 3, 0, 0, y, x));
    else if (d.bloom_mode == 2) {
        const float* b = s.bl + ((size_t)y * d.W + x) * 3;
        v = add_bloom(d, v, mk3(b[0], b[1], b[2]));
    }
    float row_mask = d.scan_mode == 1 ? scan_row(d, f, y) : 1.0f;
    return after_bloom(d, f, v, y, x, fwd, inv, row_mask);
}

// ---- stages 11-13 at one output pixel: glitch shift, warp gather, text-after -----------------
// Returns the float image value that persistence blends (what apply_static_effects returns).
CRT_HD F3 post_pixel(const Dev& d, const FrameDev& f, const uint8_t* __restrict__ in, const Scratch& s,
                     int y, int x, const float* __restrict__ fwd, const float* __restrict__ inv) {
    int gx = glitch_src_x(d, f, y, x);
    F3 v;
    if (d.warp_on) {
        Taps t = warp_taps(d, y, gx);
        float a[4][3];
        for (int k = 0; k < 4; ++k) {
            int ty = t.iy + (k >> 1), tx = t.ix + (k & 1);
            bool ok = ty >= 0 && ty < d.H && tx >= 0 && tx < d.W;
            const float* q = s.q + ((size_t)(ok ? ty : 0) * d.W + (ok ? tx : 0)) * 3;
            a[k][0] = ok ? q[0] : 0.f; a[k][1] = ok ? q[1] : 0.f; a[k][2] = ok ? q[2] : 0.f;
        }
        v = mk3(gather4(a[0][0], a[1][0], a[2][0], a[3][0], t), gather4(a[0][1], a[1][1], a[2][1], a[3][1], t),
                gather4(a[0][2], a[1][2], a[2][2], a[3][2], t));
    } else {
        v = pre_warp_pixel(d, f, in, s, y, gx, fwd, inv);
    }
    if (d.text_mode == 2) v = text_blend(d, v, y, gx);
    return v;
}

// ---- stages 14-15: blend with the persistence state, quantise, store ---------------------------
CRT_HD void finish_pixel(const Dev& d, F3 v, int has_prev, float* __restrict__ state, uint8_t* __restrict__ out, int y, int x) {
    size_t o = ((size_t)y * d.W + x) * 3;
    if (has_prev) {
        float p = d.persist, q = d.persist_q;
        v.x = blend(state[o + 0], v.x, p, q); v.y = blend(state[o + 1], v.y, p, q); v.z = blend(state[o + 2], v.z, p, q);
    }
    if (state) { state[o + 0] = v.x; state[o + 1] = v.y; state[o + 2] = v.z; }
    out[o + 0] = quantise(v.x); out[o + 1] = quantise(v.y); out[o + 2] = quantise(v.z);
}

}  // namespace crt
