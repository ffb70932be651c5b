// crt_fused.cuh — the fused tile kernel: ONE launch per frame does the whole chain
// (u8 in -> colour -> bloom -> triad -> scanlines -> vignette -> flicker -> noise ->
// warp gather -> text -> persistence -> u8 out + float32 state).  Every byte of the
// frame, of the state and of the output crosses HBM once; the bloom halo and the
// warp footprint live in shared memory.
//
// Per CTA (256 threads, output tile 64 x TH):
//   phase 0  [warp only] every thread computes the cv2.remap taps of its output
//            pixels; a block-wide min/max gives the source footprint Q of the tile
//   phase 1  stages 0-4 (graded input, thresholded bloom source) for the region
//            P = Q grown by the bloom halo -> shared memory.  With pixelate on,
//            each distinct source pixel is graded once and replicated.
//   phase 2  bloom in shared memory: 2x2 down-scale cells (fast path) or the
//            gaussian row pass
//   phase 3  [warp only] stages 5-10 evaluated in place over Q
//   phase 4  per output pixel: (column pass +) stages 5-10, or the 4-tap gather
//            from Q; text; blend with the state; 16-byte state stores, packed
//            uint8 stores
// Everything that feeds the triad LUT uses crt_math.cuh's exact arithmetic
// (shared with the staged kernels and checked on the host by tests/host_emu).
// After the LUT only +-1 LSB matters, so the masks use per-tile row/column
// tables and fast intrinsics (mask_at below) instead of double precision.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "crt_stages.cuh"
#if defined(__CUDACC__)
#include "crt_launch.h"
#endif

namespace crt {

constexpr int FT = 256;                 // threads per CTA
constexpr int FTW = 64;                 // tile width (pixels); 4 pixels per thread per row
constexpr int FROW_THREADS = FTW / 4;   // 16 threads cover one tile row
constexpr int FROWS_PER_PASS = FT / FROW_THREADS;   // 16 rows per pass
constexpr int FMAX_ROWS = 192;          // rows of per-tile row tables
constexpr int FMAX_COLS = 256;          // columns of per-tile column tables

struct FusedPlan {
    bool ok = false;
    const char* why = "not planned";
    int bloom = 0, warp = 0;
    int th = 16;                        // tile height
    int nt = 256;                       // threads per CTA (512 for tall tiles: same shared memory, twice the resident warps)
    int cap_px = 0;                     // capacity of the P-region buffer in pixels
    int cap_aux = 0;                    // floats of the auxiliary buffer (ds cells / row pass / T1 tile)
    int ps2 = 0;                        // 1: the pixel_size-2 block kernel (crt_fused_ps2.cuh)
    int gauss_k = 0;                    // != 0: the packed-FP32 gaussian kernel (crt_fused_gauss.cuh) with this tap count
    size_t smem = 0;
};

// Shared-memory footprint / availability of k_fused_gauss<K> (crt_fused_gauss.cuh)
inline size_t fused_gauss_smem(int K, int th) {
    const int R = K / 2, PW = FTW + 2 * R, PH = th + 2 * R;
    return ((size_t)3 * PW * PH + (size_t)3 * PH * FTW + (size_t)th * FTW * 3) * sizeof(float);
}
inline bool fused_gauss_supported(int K) { return K == 5 || K == 7 || K == 9 || K == 11 || K == 13 || K == 25; }


struct FusedGeom { int th, cap_px, cap_aux; };

inline int env_int(const char* name, int dflt) { const char* v = getenv(name); return v && *v ? atoi(v) : dflt; }

#if defined(__CUDACC__)
// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// Consecutive frames depend on each other only through the persistence state, which a tile kernel
// touches in its last phase.  A kernel launched with launch_pdl(pdl = true) may therefore start while
// the previous kernel in the stream is still draining its last wave: the staging of tables, the graded
// input and the blur of its first tiles overlap that tail (and the launch latency); griddep_wait()
// then blocks until the previous kernel has completed and its writes are visible.  Every access to
// memory that another kernel of the stream writes (state, pre-warp image, noise plane) must come after
// griddep_wait(); the frame input and the parameter tables are never written by this library's kernels.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
// Cooperative launch: the whole grid is resident at once or the launch fails — what lets CTAs wait for each other (clip mode).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_coop(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

// ---- region helpers ------------------------------------------------------------------------
struct Box { int x0, y0, x1, y1; };     // inclusive
CRT_HD int box_w(const Box& b) { return b.x1 - b.x0 + 1; }
CRT_HD int box_h(const Box& b) { return b.y1 - b.y0 + 1; }

// Region of graded-input pixels needed to evaluate bloom on Q (cells = ds cells of the fast path).
CRT_HD Box grow_for_bloom(const Dev& d, const Box& q, Box* cells) {
    Box p = q;
    if (d.bloom_mode == 1) {
        Box c;
        c.x0 = up_coord(d, d.up_x, q.x0, d.hw).s0; c.x1 = up_coord(d, d.up_x, q.x1, d.hw).s1;
        c.y0 = up_coord(d, d.up_y, q.y0, d.hh).s0; c.y1 = up_coord(d, d.up_y, q.y1, d.hh).s1;
        *cells = c;
        p.x0 = imin(q.x0, down_coord(d, d.dn_x, c.x0).s0); p.x1 = imax(q.x1, down_coord(d, d.dn_x, c.x1).s1);
        p.y0 = imin(q.y0, down_coord(d, d.dn_y, c.y0).s0); p.y1 = imax(q.y1, down_coord(d, d.dn_y, c.y1).s1);
    } else if (d.bloom_mode == 2) {
        const int r = d.ksize >> 1;      // REPLICATE border: clamped coordinates stay inside the image
        p.x0 = imax(q.x0 - r, 0); p.x1 = imin(q.x1 + r, d.W - 1);
        p.y0 = imax(q.y0 - r, 0); p.y1 = imin(q.y1 + r, d.H - 1);
    }
    return p;
}

// Footprint of a tile from the min/max tap coordinates (ix, ix+1, iy, iy+1 over the sampled
// output pixels): clipped to the image (taps outside contribute 0 and are never read); when the
// extremes come from the perimeter only (warp_mono) one pixel of slack covers float rounding.
CRT_HD Box footprint_box(const Dev& d, int x0, int y0, int x1, int y1, bool* empty) {
    const int m = d.warp_mono ? 1 : 0;
    *empty = x0 > x1 || x0 > d.W - 1 || x1 < 0 || y0 > d.H - 1 || y1 < 0;
    Box q;
    q.x0 = imax(x0 - m, 0); q.x1 = imin(x1 + m, d.W - 1);
    q.y0 = imax(y0 - m, 0); q.y1 = imin(y1 + m, d.H - 1);
    return q;
}

// Warp map with the per-row / per-column normalised coordinates hoisted (same float32
// operations and order as warp_taps, crt_math.cuh).
CRT_HD float warp_norm(float v, float c, float dv) { return fdiv(fsub(v, c), dv); }
CRT_HD Taps warp_taps_n(const Dev& d, float xn, float yn) {
    float r2 = fadd(fmul(xn, xn), fmul(yn, yn));
    float fac = fadd(1.0f, fmul(d.warp_k, r2));
    float mx = fadd(fmul(fmul(xn, fac), d.warp_cx), d.warp_cx);
    float my = fadd(fmul(fmul(yn, fac), d.warp_cy), d.warp_cy);
    float qx = clampf(fmul(mx, 32.0f), -1.0e9f, 1.0e9f), qy = clampf(fmul(my, 32.0f), -1.0e9f, 1.0e9f);
    int sx = (int)rintf(qx), sy = (int)rintf(qy);
    Taps t;
    t.ix = sx >> 5; t.iy = sy >> 5;
    float fx = fmul((float)(sx & 31), 0.03125f), fy = fmul((float)(sy & 31), 0.03125f);
    float gx = fsub(1.0f, fx), gy = fsub(1.0f, fy);
    t.w00 = fmul(gy, gx); t.w01 = fmul(gy, fx); t.w10 = fmul(fy, gx); t.w11 = fmul(fy, fx);
    return t;
}

#if defined(__CUDACC__)

__device__ __forceinline__ void store_f3(float* p, F3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
__device__ __forceinline__ F3 load_f3(const float* p) { return mk3(p[0], p[1], p[2]); }
// q = n / dv with magic = 2^32 / dv + 1 computed once per CTA; exact for n * dv < 2^32 (dv >= 2)
__device__ __forceinline__ int fastdiv(int n, unsigned magic) { return magic ? (int)__umulhi((unsigned)n, magic) : n; }
__device__ __forceinline__ unsigned make_magic(int dv) { return dv <= 1 ? 0u : 0xffffffffu / (unsigned)dv + 1u; }   // 0 = divide by 1

// After the triad LUT only the final +-1 LSB matters: these use fused multiply-adds where the
// reference's float64 / OpenCV code uses separate operations.
__device__ __forceinline__ float gather4_fast(float a, float b, float c, float e, const Taps& t) {
    return __fmaf_rn(e, t.w11, __fmaf_rn(c, t.w10, __fmaf_rn(b, t.w01, a * t.w00)));
}
__device__ __forceinline__ float blend_fast(float prev, float v, float p, float q) { return __saturatef(__fmaf_rn(p, prev, q * v)); }
// Float -> integer conversions run on the quarter-rate conversion pipe; both conversions of the chain are
// done on the FMA pipe instead by adding a power of two whose ulp is 1:
//   quantise: values are already clipped to [0, 1] (an ulp above 1 after the gather still rounds to 255);
//             the low byte of v * 255 + 1.5 * 2^23 (round-to-nearest-even) is round-half-even(255 v)
//   LUT index: v * 1024 is exact in float32, so v * 1024 + 2^23 rounded toward zero is 2^23 + floor(1024 v),
//             the reference's truncating index (crt_filter.py:250), for v in [0, 1]
__device__ __forceinline__ uint32_t quantise_bits(float v) { return __float_as_uint(__fmaf_rn(v, 255.0f, 12582912.0f)); }
__device__ __forceinline__ uint32_t quantise_fast(float v) { return quantise_bits(v) & 0xffu; }
__device__ __forceinline__ uint32_t pack4(float a, float b, float c, float e) {
    return __byte_perm(__byte_perm(quantise_bits(a), quantise_bits(b), 0x0040), __byte_perm(quantise_bits(c), quantise_bits(e), 0x0040), 0x5410);
}
// the same with a table offset folded into the magic number (2^23 + offset, offset + 1024 < 4096)
__device__ __forceinline__ int lut_index_magic(float v01, float magic) { return (int)(__float_as_uint(__fmaf_rz(v01, 1024.0f, magic)) & 0xfffu); }
__device__ __forceinline__ int lut_index_fast(float v01) { return (int)(__float_as_uint(__fmaf_rz(v01, 1024.0f, 8388608.0f)) & 0x7ffu); }

// L2 prefetch of a tile's persistence-state rows (one 128-byte line per thread).  MEASURED (round 1, run 23):
// issuing it at kernel start made the kernels 2-4 % slower, so no kernel calls it; kept for experiments.
__device__ __forceinline__ void prefetch_state_tile(const float* __restrict__ state, int W, int x0, int y0, int tw, int rows, int tid, int nthreads) {
    const int lines_per_row = (tw * 12 + 127) >> 7;
    for (int i = tid; i < rows * lines_per_row; i += nthreads) {
        const int r = i / lines_per_row, sgm = i - r * lines_per_row;
        const char* p = reinterpret_cast<const char*>(state + ((size_t)(y0 + r) * W + x0) * 3) + (sgm << 7);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
}

// Per-tile tables for the masks applied after the triad LUT.
struct MaskTabs {
    float* row_scan;    // scan_mode 1: row mask;  scan_mode 2: phase fraction of the row
    float* col_scan;    // scan_mode 2: phase fraction of the column
    float* row_vig;     // ((y - cy) / ry)^2
    float* col_vig;     // ((x - cx) / rx)^2
};

// Stages 7-8 as one multiplier (after the LUT the clip between them is a no-op: both factors are <= 1).
__device__ __forceinline__ float mask_at(const Dev& d, const MaskTabs& m, int r, int c, int y, int x) {
    float mk = 1.0f;
    if (d.scan_mode == 1) mk = m.row_scan[r];
    else if (d.scan_mode == 2) {
        float t = m.row_scan[r] + m.col_scan[c];
        t = t >= 1.0f ? t - 1.0f : t;                                  // turns in [0, 1)
        float s = fmaxf(__fmaf_rn(-0.5f, __sinf(__fmaf_rn(6.283185307f, t, -3.14159265f)), 0.5f), 0.0f);   // 0.5 (1 + sin(2 pi t))
        float shaped = (d.scan_inv_sharp == 1.0f) ? s : __powf(s, d.scan_inv_sharp);
        mk = __fmaf_rn(-d.scan_strength, shaped, 1.0f);
    }
    if (d.vig_mode == 1) mk *= __fmaf_rn(-d.vig_strength, __saturatef(m.row_vig[r] + m.col_vig[c]), 1.0f);
    else if (d.vig_mode == 2) mk *= d.vig_plane[(size_t)y * d.W + x];
    return mk;
}

// Stages 6-10 with the mask tables (the triad LUT part is crt_math.cuh's exact code).
__device__ __forceinline__ F3 after_bloom_fast(const Dev& d, const FrameDev& f, F3 v, int y, int x, const float* fwd, const float* inv,
                                               const MaskTabs& m, int r, int c) {
    if (d.triad_mode) {
        if (d.triad_comp) {
            if (x >= d.comp_x0 && x <= d.comp_x1) {               // regular column: one composite look-up per channel
                const int ph = x - 3 * (int)__umulhi((unsigned)x, 0x55555556u);     // x % 3
                const int p0 = d.bgr ? 2 : 0;                     // mask phase of channel index 0 (R in RGB order)
                v.x = (ph == p0 ? fwd : inv)[lut_index_fast(__saturatef(v.x))];
                v.y = (ph == 1 ? fwd : inv)[lut_index_fast(__saturatef(v.y))];
                v.z = (ph == 2 - p0 ? fwd : inv)[lut_index_fast(__saturatef(v.z))];
            } else {
                v = triad(d, v, x, d.lut_fwd, d.lut_inv);         // mask edge columns: full path from global memory
            }
        } else {
            v = triad(d, v, x, fwd, inv);
        }
    }
    if (d.scan_mode | d.vig_mode) {
        const float mk = mask_at(d, m, r, c, y, x);
        v.x = __saturatef(v.x * mk); v.y = __saturatef(v.y * mk); v.z = __saturatef(v.z * mk);
    }
    if (f.flicker_on) { v.x = __saturatef(v.x * f.flicker); v.y = __saturatef(v.y * f.flicker); v.z = __saturatef(v.z * f.flicker); }
    if (d.noise_on) {
        const float n = noise_at(d, f, y, x);
        v.x = __saturatef(v.x + n); v.y = __saturatef(v.y + n); v.z = __saturatef(v.z + n);
    }
    return v;
}

// Blend four horizontally adjacent pixels with the persistence state, write state (16-byte
// stores) and packed uint8 output.  `pixel(y, x, k)` returns the float image value
// (what apply_static_effects returns) of pixel k of the quad.  With `q_out` set the float
// values are stored there instead (first pass of the two-pass path, crt_abi.cu).
template <typename PixelFn>
__device__ __forceinline__ void finish_quad(const Dev& d, float* __restrict__ state, uint8_t* __restrict__ out, float* __restrict__ q_out,
                                            int has_prev, int y, int xb, int npx, PixelFn&& pixel,
                                            bool prev_given = false, float4 ga = float4(), float4 gb = float4(), float4 gc = float4()) {
    const float pp = d.persist, pq = d.persist_q;
    const bool vec = (d.W & 3) == 0;
    const int o = (y * d.W + xb) * 3;               // < 2^31 (checked by plan_fused)
    float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa, pc = pa;       // previous state of the 4 pixels
    if (q_out) { has_prev = 0; state = q_out; }
    if (has_prev) {
        if (prev_given) {                           // the caller fetched the previous state early
            pa = ga; pb = gb; pc = gc;
        } else if (vec) {
            const float4* sp = reinterpret_cast<const float4*>(state + o);
            pa = sp[0]; pb = sp[1]; pc = sp[2];
        } else {
            float t[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) t[k] = (k < npx * 3) ? state[o + k] : 0.f;
            pa = make_float4(t[0], t[1], t[2], t[3]); pb = make_float4(t[4], t[5], t[6], t[7]); pc = make_float4(t[8], t[9], t[10], t[11]);
        }
    }
    const float prev[12] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w, pc.x, pc.y, pc.z, pc.w};
    float res[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        F3 v = mk3(0.f, 0.f, 0.f);
        if (k < npx) {
            v = pixel(y, xb + k, k);
            if (has_prev) { v.x = blend_fast(prev[k * 3], v.x, pp, pq); v.y = blend_fast(prev[k * 3 + 1], v.y, pp, pq); v.z = blend_fast(prev[k * 3 + 2], v.z, pp, pq); }
        }
        res[k * 3] = v.x; res[k * 3 + 1] = v.y; res[k * 3 + 2] = v.z;
    }
    if (vec) {
        if (state) {
            float4* sp = reinterpret_cast<float4*>(state + o);
            sp[0] = make_float4(res[0], res[1], res[2], res[3]);
            sp[1] = make_float4(res[4], res[5], res[6], res[7]);
            sp[2] = make_float4(res[8], res[9], res[10], res[11]);
        }
        if (q_out || !out) return;          // float image only (first pass of the two-pass path / crt_process_static)
        uint32_t w[3];
#pragma unroll
        for (int j = 0; j < 3; ++j)
            w[j] = pack4(res[j * 4], res[j * 4 + 1], res[j * 4 + 2], res[j * 4 + 3]);
        uint32_t* op = reinterpret_cast<uint32_t*>(out + o);
        op[0] = w[0]; op[1] = w[1]; op[2] = w[2];
    } else {
#pragma unroll
        for (int k = 0; k < 12; ++k)
            if (k < npx * 3) {
                if (state) state[o + k] = res[k];
                if (!q_out && out) out[o + k] = quantise(res[k]);
            }
    }
}

template <int BLOOM, bool WARP, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) k_fused(Dev d, FrameDev f, const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                              float* __restrict__ state, float* __restrict__ q_out, int has_prev, FusedGeom g) {
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(16) float s_fwd[1028], s_inv[1028];
    __shared__ float s_unit[256];
    __shared__ __align__(16) float s_pow[POW_TAB_FLOATS];
    __shared__ float s_rows[2 * FMAX_ROWS], s_cols[2 * FMAX_COLS];
    __shared__ int s_ci[2 * FMAX_COLS], s_ri[2 * FMAX_ROWS];       // fast bloom: up-scale tap offsets per Q column / row
    __shared__ float s_cw[FMAX_COLS], s_rw[FMAX_ROWS];              // ... and weights (cv2.resize coordinates)
    __shared__ int s_box[4], s_geo[24];
    float* T = sm;                                  // [ph][pw][3] graded input (bloom source when thresholded)
    float* A = sm + g.cap_px * 3;                   // auxiliary: ds cells | row pass (+ T1 tile)
    const int tid = threadIdx.x, lane = tid & 31;
    const int ox0 = blockIdx.x * FTW, oy0 = blockIdx.y * g.th;
    const int ox1 = imin(ox0 + FTW, d.W) - 1, oy1 = imin(oy0 + g.th, d.H) - 1;
    const int trow = tid / FROW_THREADS, xb = ox0 + (tid % FROW_THREADS) * 4;    // this thread's pixel quad

    // triad tables in shared memory: the composite (bright, dim) pair when the mask is regular, else (forward, inverse)
    const float* lut_a = d.triad_comp ? d.triad_comp : d.lut_fwd;
    const float* lut_b = d.triad_comp ? d.triad_comp + 1028 : d.lut_inv;
    if (d.triad_mode >= 2) {                        // 2 x 1025 floats, 16-byte loads
        for (int i = tid; i < 256; i += NT) {
            reinterpret_cast<float4*>(s_fwd)[i] = reinterpret_cast<const float4*>(lut_a)[i];
            reinterpret_cast<float4*>(s_inv)[i] = reinterpret_cast<const float4*>(lut_b)[i];
        }
        if (tid == 0) { s_fwd[1024] = lut_a[1024]; s_inv[1024] = lut_b[1024]; }
    }
    if (tid < 256) s_unit[tid] = __fdiv_rn((float)tid, 255.0f);
    if (d.col_gamma) for (int i = tid; i < POW_TAB_FLOATS; i += blockDim.x) s_pow[i] = d.pow_tab[i];
    if (WARP && tid < 4) s_box[tid] = (tid < 2) ? 0x7fffffff : -0x7fffffff;
    __syncthreads();

    // ---- phase 0: footprint of the tile in the pre-warp image ------------------------------
    float xn[4] = {0.f, 0.f, 0.f, 0.f};
    if (WARP) {
#pragma unroll
        for (int k = 0; k < 4; ++k) xn[k] = warp_norm((float)(xb + k), d.warp_cx, d.warp_dx);
        int bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -0x7fffffff, by1 = -0x7fffffff;
        if (d.warp_mono) {
            // monotone map: the extremes over the tile are attained on its perimeter
            const int tw = ox1 - ox0 + 1, thh = oy1 - oy0 + 1, nper = 2 * tw + 2 * thh;
            for (int i = tid; i < nper; i += NT) {
                int x, y;
                if (i < 2 * tw) { x = ox0 + (i < tw ? i : i - tw); y = i < tw ? oy0 : oy1; }
                else { const int j = i - 2 * tw; y = oy0 + (j < thh ? j : j - thh); x = j < thh ? ox0 : ox1; }
                const Taps t = warp_taps_n(d, warp_norm((float)x, d.warp_cx, d.warp_dx), warp_norm((float)y, d.warp_cy, d.warp_dy));
                bx0 = imin(bx0, t.ix); bx1 = imax(bx1, t.ix + 1); by0 = imin(by0, t.iy); by1 = imax(by1, t.iy + 1);
            }
        } else {
            for (int y = oy0 + trow; y <= oy1; y += NT / FROW_THREADS) {
                const float yn = warp_norm((float)y, d.warp_cy, d.warp_dy);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (xb + k <= ox1) {
                        const Taps t = warp_taps_n(d, xn[k], yn);
                        bx0 = imin(bx0, t.ix); bx1 = imax(bx1, t.ix + 1); by0 = imin(by0, t.iy); by1 = imax(by1, t.iy + 1);
                    }
                }
            }
        }
        bx0 = __reduce_min_sync(0xffffffffu, bx0); by0 = __reduce_min_sync(0xffffffffu, by0);
        bx1 = __reduce_max_sync(0xffffffffu, bx1); by1 = __reduce_max_sync(0xffffffffu, by1);
        if (lane == 0) { atomicMin(&s_box[0], bx0); atomicMin(&s_box[1], by0); atomicMax(&s_box[2], bx1); atomicMax(&s_box[3], by1); }
        __syncthreads();
    }
    // Tile geometry is the same for every thread: one thread derives it (integer divisions, cv2
    // coordinate look-ups) and publishes it through shared memory.
    if (tid == 0) {
        Box q0;
        bool empty = false;
        if (WARP) {
            q0 = footprint_box(d, s_box[0], s_box[1], s_box[2], s_box[3], &empty);
        } else { q0.x0 = ox0; q0.y0 = oy0; q0.x1 = ox1; q0.y1 = oy1; }
        Box c0{0, 0, -1, -1};
        Box p0 = q0;
        if (!empty) p0 = grow_for_bloom(d, q0, &c0);
        const int ps0 = (d.pix_uniform > 1 && d.text_mode != 1) ? d.pix_uniform : 1;
        int* gq = s_geo;
        gq[0] = q0.x0; gq[1] = q0.y0; gq[2] = q0.x1; gq[3] = q0.y1;
        gq[4] = p0.x0; gq[5] = p0.y0; gq[6] = p0.x1; gq[7] = p0.y1;
        gq[8] = c0.x0; gq[9] = c0.y0; gq[10] = c0.x1; gq[11] = c0.y1;
        gq[12] = empty; gq[13] = ps0;
        gq[14] = p0.x0 / ps0; gq[15] = p0.y0 / ps0;                       // first unique source column / row
        gq[16] = p0.x1 / ps0 - gq[14] + 1; gq[17] = p0.y1 / ps0 - gq[15] + 1;
        gq[18] = (int)make_magic(gq[16]);                                  // / nux
        gq[19] = (int)make_magic(box_w(c0));                               // / dw
        gq[20] = (int)make_magic(empty ? 1 : box_w(q0));                   // / qw
        gq[21] = (int)make_magic(BLOOM == 2 ? box_w(q0) * 3 : 1);          // / row-pass length
    }
    __syncthreads();
    Box q{s_geo[0], s_geo[1], s_geo[2], s_geo[3]};
    const Box p{s_geo[4], s_geo[5], s_geo[6], s_geo[7]};
    const Box cells{s_geo[8], s_geo[9], s_geo[10], s_geo[11]};
    const bool q_empty = s_geo[12] != 0;
    const int pw = q_empty ? 0 : box_w(p), ph = q_empty ? 0 : box_h(p);
    const int qw = q_empty ? 0 : box_w(q), qh = q_empty ? 0 : box_h(q);
    // planning guarantees the region fits; a violated bound must never corrupt memory
    if (pw * ph > g.cap_px || qh > FMAX_ROWS || qw > FMAX_COLS) {
        if (tid == 0) printf("crt_b200: fused tile region %dx%d exceeds capacity %d\n", pw, ph, g.cap_px);
        return;
    }
    MaskTabs mt{s_rows, s_cols, s_rows + FMAX_ROWS, s_cols + FMAX_COLS};

    // ---- phase 1: graded input over P (each distinct source pixel once) ------------------------
    const int cw = BLOOM == 2 ? qw : 0;             // row-pass columns
    float* T1 = A + ph * cw * 3;                    // [th][FTW][3], only BLOOM==2 && thr_on
    if (!q_empty) {
        const int ps = s_geo[13], ux0 = s_geo[14], uy0 = s_geo[15], nux = s_geo[16], nuy = s_geo[17];
        const unsigned magic = (unsigned)s_geo[18];
        for (int u = tid; u < nux * nuy; u += NT) {
            const int uy = fastdiv(u, magic), ux = u - uy * nux;
            const int xa = imax((ux0 + ux) * ps, p.x0), xe = imin((ux0 + ux) * ps + ps - 1, p.x1);
            const int ya = imax((uy0 + uy) * ps, p.y0), ye = imin((uy0 + uy) * ps + ps - 1, p.y1);
            const F3 v1 = ps > 1 ? graded_source_lut(d, in, (uy0 + uy) * ps, (ux0 + ux) * ps, ya, xa, s_unit, s_pow) : graded_input_lut(d, in, ya, xa, s_unit, s_pow);
            const F3 v = (BLOOM == 2 && d.thr_on) ? bloom_src(d, v1) : v1;
            float* t0 = T + ((ya - p.y0) * pw + (xa - p.x0)) * 3;
            if (ps == 2 && xe == xa + 1 && ye == ya + 1 && !(BLOOM == 2 && d.thr_on)) {      // whole 2x2 block inside the region
                store_f3(t0, v); store_f3(t0 + 3, v); store_f3(t0 + pw * 3, v); store_f3(t0 + pw * 3 + 3, v);
            } else {
                for (int y = ya; y <= ye; ++y)
                    for (int x = xa; x <= xe; ++x) {
                        if (BLOOM == 2 && d.thr_on && y >= oy0 && y <= oy1 && x >= ox0 && x <= ox1)
                            store_f3(T1 + ((y - oy0) * FTW + (x - ox0)) * 3, v1);
                        store_f3(t0 + ((y - ya) * pw + (x - xa)) * 3, v);
                    }
            }
        }
        // per-tile mask tables over Q
        for (int r = tid; r < qh; r += NT) {
            const int y = q.y0 + r;
            if (d.scan_mode == 1) mt.row_scan[r] = scan_row(d, f, y);
            else if (d.scan_mode == 2) { double t = ((double)y + f.phase) * d.scan_inv_period; mt.row_scan[r] = (float)(t - floor(t)); }
            if (d.vig_mode == 1) { const float ny = ((float)y - d.vig_cy) * d.vig_iry; mt.row_vig[r] = ny * ny; }
        }
        for (int c = tid; c < qw; c += NT) {
            const int x = q.x0 + c;
            if (d.scan_mode == 2) { double t = (d.scan_tan * (double)x) * d.scan_inv_period; mt.col_scan[c] = (float)(t - floor(t)); }
            if (d.vig_mode == 1) { const float nx = ((float)x - d.vig_cx) * d.vig_irx; mt.col_vig[c] = nx * nx; }
        }
        if (BLOOM == 1) {
            for (int c = tid; c < qw; c += NT) {
                const Lerp1 cx = up_coord(d, d.up_x, q.x0 + c, d.hw);
                s_ci[c] = (cx.s0 - cells.x0) * 3; s_ci[FMAX_COLS + c] = (cx.s1 - cells.x0) * 3; s_cw[c] = cx.w;
            }
            for (int r = tid; r < qh; r += NT) {
                const Lerp1 cy = up_coord(d, d.up_y, q.y0 + r, d.hh);
                s_ri[r] = (cy.s0 - cells.y0) * qw * 3; s_ri[FMAX_ROWS + r] = (cy.s1 - cells.y0) * qw * 3; s_rw[r] = cy.w;
            }
        }
    }
    __syncthreads();

    // ---- phase 2: bloom in shared memory -------------------------------------------------------
    const int dw = box_w(cells), dh = box_h(cells);
    float* HP = A + dw * dh * 3;                    // [dh][qw][3] fast bloom: cells up-scaled along x
    // vertical half of the up-scale at Q pixel (r, c)
    auto bloom_fast_at = [&](int r, int c) -> F3 {
        const float* h0 = HP + s_ri[r] + c * 3;
        const float* h1 = HP + s_ri[FMAX_ROWS + r] + c * 3;
        const float w = s_rw[r];
        return mk3(lerp_cv(h0[0], h1[0], w), lerp_cv(h0[1], h1[1], w), lerp_cv(h0[2], h1[2], w));
    };
    if (BLOOM == 1 && !q_empty) {
        const unsigned magic = (unsigned)s_geo[19];
        for (int u = tid; u < dw * dh; u += NT) {
            const int r = fastdiv(u, magic), c = u - r * dw;
            const Lerp1 cy = down_coord(d, d.dn_y, cells.y0 + r), cx = down_coord(d, d.dn_x, cells.x0 + c);
            const float* r0 = T + (cy.s0 - p.y0) * pw * 3;
            const float* r1 = T + (cy.s1 - p.y0) * pw * 3;
            const int a0 = (cx.s0 - p.x0) * 3, a1 = (cx.s1 - p.x0) * 3;
            const F3 a = bloom_src(d, load_f3(r0 + a0)), b = bloom_src(d, load_f3(r0 + a1));
            const F3 cc = bloom_src(d, load_f3(r1 + a0)), e = bloom_src(d, load_f3(r1 + a1));
            float* o = A + u * 3;
            o[0] = lerp_cv(lerp_cv(a.x, b.x, cx.w), lerp_cv(cc.x, e.x, cx.w), cy.w);
            o[1] = lerp_cv(lerp_cv(a.y, b.y, cx.w), lerp_cv(cc.y, e.y, cx.w), cy.w);
            o[2] = lerp_cv(lerp_cv(a.z, b.z, cx.w), lerp_cv(cc.z, e.z, cx.w), cy.w);
        }
        __syncthreads();
        // horizontal half of the 2x up-scale (cv2.resize works rows first), once per (cell row, Q column)
        const unsigned magic_q = (unsigned)s_geo[20];
        for (int u = tid; u < dh * qw; u += NT) {
            const int j = fastdiv(u, magic_q), c = u - j * qw;
            const float* dr = A + j * dw * 3;
            const int a0 = s_ci[c], a1 = s_ci[FMAX_COLS + c];
            const float w = s_cw[c];
            float* o = HP + u * 3;
            o[0] = lerp_cv(dr[a0], dr[a1], w); o[1] = lerp_cv(dr[a0 + 1], dr[a1 + 1], w); o[2] = lerp_cv(dr[a0 + 2], dr[a1 + 2], w);
        }
        __syncthreads();
    }
    if (BLOOM == 2) {
        // row pass over every region row, for the tile's columns; REPLICATE = clamped column index
        const int K = d.ksize, rad = K >> 1;
        const int rowlen = cw * 3;
        const unsigned magic = (unsigned)s_geo[21];
        for (int u = tid; u < ph * rowlen; u += NT) {
            const int r = fastdiv(u, magic), e = u - r * rowlen;
            const int c = fastdiv(e, 0x55555556u /* 2^32/3 + 1 */), ch = e - c * 3, x = q.x0 + c;
            const float* row = T + r * pw * 3;
            float acc;
            if (x - rad >= p.x0 && x + rad <= p.x1) {
                acc = gauss_row(row + (x - rad - p.x0) * 3 + ch, 3, d.taps, K);
            } else {                     // image border: REPLICATE = clamped column index
                auto tapv = [&](int i) { return row[(imin(imax(x - rad + i, 0), d.W - 1) - p.x0) * 3 + ch]; };
                if (K == 1) acc = fmul(tapv(0), d.taps[0]);
                else if (K == 3) acc = ffma(tapv(1), d.taps[1], fmul(fadd(tapv(0), tapv(2)), d.taps[2]));
                else if (K == 5) acc = ffma(fadd(tapv(4), tapv(0)), d.taps[4], ffma(tapv(2), d.taps[2], fmul(fadd(tapv(1), tapv(3)), d.taps[3])));
                else {
                    acc = fmul(tapv(0), d.taps[0]);
                    for (int i = 1; i < K; ++i) acc = ffma(tapv(i), d.taps[i], acc);
                }
            }
            A[u] = acc;
        }
        __syncthreads();
    }

    // ---- phase 3 [warp]: stages 5-10 in place over Q ----------------------------------------------
    if (WARP && !q_empty) {
        const unsigned magic = (unsigned)s_geo[20];
        for (int u = tid; u < qw * qh; u += NT) {
            const int r = fastdiv(u, magic), c = u - r * qw;
            const int y = q.y0 + r, x = q.x0 + c;
            float* tp = T + ((y - p.y0) * pw + (x - p.x0)) * 3;
            F3 v = load_f3(tp);
            if (BLOOM == 1) v = add_bloom(d, v, bloom_fast_at(r, c));
            v = after_bloom_fast(d, f, v, y, x, s_fwd, s_inv, mt, r, c);
            store_f3(tp, v);
        }
        __syncthreads();
    }

    // ---- phase 4: output pixels ---------------------------------------------------------------------
    if (xb > ox1) return;
    float yn = 0.f;
    auto pixel = [&](int y, int x, int k) -> F3 {
        F3 v = mk3(0.f, 0.f, 0.f);
        if (WARP) {
            if (!q_empty) {
                const Taps t = warp_taps_n(d, xn[k], yn);
                const float* base = T + ((t.iy - p.y0) * pw + (t.ix - p.x0)) * 3;
                F3 a[4];
                if (t.ix >= q.x0 && t.ix < q.x1 && t.iy >= q.y0 && t.iy < q.y1) {       // all four taps inside Q (hence inside the image)
                    a[0] = load_f3(base); a[1] = load_f3(base + 3); a[2] = load_f3(base + pw * 3); a[3] = load_f3(base + pw * 3 + 3);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int ty = t.iy + (j >> 1), tx = t.ix + (j & 1);
                        const bool ok = ty >= 0 && ty < d.H && tx >= 0 && tx < d.W;
                        a[j] = ok ? load_f3(T + ((ty - p.y0) * pw + (tx - p.x0)) * 3) : mk3(0.f, 0.f, 0.f);
                    }
                }
                v = mk3(gather4_fast(a[0].x, a[1].x, a[2].x, a[3].x, t), gather4_fast(a[0].y, a[1].y, a[2].y, a[3].y, t),
                        gather4_fast(a[0].z, a[1].z, a[2].z, a[3].z, t));
            }
        } else {
            const float* tp = T + ((y - p.y0) * pw + (x - p.x0)) * 3;
            v = (BLOOM == 2 && d.thr_on) ? load_f3(T1 + ((y - oy0) * FTW + (x - ox0)) * 3) : load_f3(tp);
            if (BLOOM == 1) v = add_bloom(d, v, bloom_fast_at(y - q.y0, x - q.x0));
            if (BLOOM == 2) {
                const int K = d.ksize, rad = K >> 1, c = x - q.x0;
                float bl[3];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    if (y - rad >= p.y0 && y + rad <= p.y1) {
                        bl[ch] = gauss_col(A + ((y - p.y0) * cw + c) * 3 + ch, cw * 3, d.taps, K);
                    } else {     // image border: REPLICATE = clamped row index
                        const float* col = A + c * 3 + ch;
                        float acc = fmul(col[(y - p.y0) * cw * 3], d.taps[rad]);
                        for (int i = 1; i <= rad; ++i) {
                            const int yu = imin(y + i, d.H - 1) - p.y0, yd = imax(y - i, 0) - p.y0;
                            acc = ffma(fadd(col[yu * cw * 3], col[yd * cw * 3]), d.taps[rad + i], acc);
                        }
                        bl[ch] = acc;
                    }
                }
                v = add_bloom(d, v, mk3(bl[0], bl[1], bl[2]));
            }
            v = after_bloom_fast(d, f, v, y, x, s_fwd, s_inv, mt, y - q.y0, x - q.x0);
        }
        if (d.text_mode == 2) v = text_blend(d, v, y, x);
        return v;
    };
    for (int y = oy0 + trow; y <= oy1; y += NT / FROW_THREADS) {
        if (WARP) yn = warp_norm((float)y, d.warp_cy, d.warp_dy);
        finish_quad(d, state, out, q_out, has_prev, y, xb, imin(4, ox1 - xb + 1), pixel);
    }
}

// Second pass of the two-pass path: glitch shift, 4-tap warp gather from the pre-warp image
// written by the first pass (global memory, L2-resident for the most part), text layer,
// persistence, quantise.  Handles every gather the single-pass kernel declines.
// ncu (run 41, 4K): 78 us, 187 thread-instructions per pixel, issue slots 60 % busy, 40 registers, 67 %
// occupancy, long-scoreboard stalls on the twelve 4-byte tap loads per pixel.  Tried and measured slower
// on cfg3 (runs 42-44, 7 560 frames/s as it stands): column coordinates hoisted over 4 rows per thread
// (58 registers, 7 170), an all-taps-inside fast path (48 registers, 7 400; capped at 40 registers 7 380),
// programmatic dependent launch of this kernel.  Next step: stage the tile's footprint in shared memory.
constexpr int GATHER_TH = 16;
template <bool WARP>
__global__ void __launch_bounds__(256) k_gather(Dev d, FrameDev f, const float* __restrict__ qimg, uint8_t* __restrict__ out,
                                                float* __restrict__ state, int has_prev) {
    const int tid = threadIdx.x;
    const int y = blockIdx.y * GATHER_TH + tid / FROW_THREADS, xb = blockIdx.x * FTW + (tid % FROW_THREADS) * 4;
    griddep_launch_dependents();        // the next frame's first pass may fill this kernel's last wave (it waits before it writes)
    if (y >= d.H || xb >= d.W) return;
    const float yn = WARP ? warp_norm((float)y, d.warp_cy, d.warp_dy) : 0.f;
    auto pixel = [&](int yy, int x, int k) -> F3 {
        const int gx = glitch_src_x(d, f, yy, x);
        F3 v = mk3(0.f, 0.f, 0.f);
        if (WARP) {
            const Taps t = warp_taps_n(d, warp_norm((float)gx, d.warp_cx, d.warp_dx), yn);
            F3 a[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ty = t.iy + (j >> 1), tx = t.ix + (j & 1);
                const bool ok = ty >= 0 && ty < d.H && tx >= 0 && tx < d.W;
                a[j] = ok ? load_f3(qimg + ((size_t)ty * d.W + tx) * 3) : mk3(0.f, 0.f, 0.f);
            }
            v = mk3(gather4_fast(a[0].x, a[1].x, a[2].x, a[3].x, t), gather4_fast(a[0].y, a[1].y, a[2].y, a[3].y, t),
                    gather4_fast(a[0].z, a[1].z, a[2].z, a[3].z, t));
        } else {
            v = load_f3(qimg + ((size_t)yy * d.W + gx) * 3);
        }
        if (d.text_mode == 2) v = text_blend(d, v, yy, gx);
        return v;
    };
    finish_quad(d, state, out, nullptr, has_prev, y, xb, imin(4, d.W - xb), pixel);
}

#if defined(CRT_TU_FUSED)
inline int run_gather(const Dev& d, const FrameDev& f, const float* qimg, uint8_t* out, float* state, int has_prev, cudaStream_t st, int* launches) {
    dim3 grid((d.W + FTW - 1) / FTW, (d.H + GATHER_TH - 1) / GATHER_TH);
    if (d.warp_on) k_gather<true><<<grid, 256, 0, st>>>(d, f, qimg, out, state, has_prev);
    else k_gather<false><<<grid, 256, 0, st>>>(d, f, qimg, out, state, has_prev);
    ++*launches;
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

#endif  // CRT_TU_FUSED

#endif  // __CUDACC__

// ---- planning (host) --------------------------------------------------------------------------
// `hd` is a copy of the device parameter block whose coordinate tables point at HOST copies.
// The worst-case region over all tiles is exact: it is evaluated with the same float32 warp map
// the kernel uses.  Tall tiles amortise the bloom/warp halo, so the tallest tile that leaves room
// for two CTAs per SM is chosen.
inline FusedPlan plan_fused(const Dev& hd, bool glitch_on) {
    const Dev& d = hd;
    FusedPlan pl;
    pl.bloom = d.bloom_mode; pl.warp = d.warp_on;
    if (glitch_on) { pl.why = "glitch gather is handled by the staged kernels"; return pl; }
    if (d.warp_on && d.bloom_mode == 2) { pl.why = "warp + gaussian bloom is handled by the staged kernels"; return pl; }
    if ((size_t)d.W * d.H * 3 >= ((size_t)1 << 31)) { pl.why = "frame too large for 32-bit indexing"; return pl; }
    if (d.pix_uniform == 2 && d.even_dims && (d.W & 3) == 0 && !d.warp_on && !glitch_on && d.text_mode == 0 && d.bloom_mode != 2 &&
        !env_int("CRT_NO_PS2", 0)) {            // == fused_ps2_supported (crt_fused_ps2.cuh): the default-chain block kernel
        pl.ok = true; pl.why = ""; pl.ps2 = 1; pl.th = 32; pl.nt = 256; pl.smem = 0;
        return pl;
    }
    if (d.bloom_mode == 2 && d.pix_uniform == 2 && d.even_dims && (d.W & 3) == 0 && !d.warp_on && d.text_mode == 0 &&
        fused_gauss_supported(d.ksize) && !env_int("CRT_NO_PS2", 0)) {     // == fused_gauss_ps2_supported: block-resolution gaussian
        pl.ok = true; pl.why = ""; pl.ps2 = 1; pl.gauss_k = d.ksize; pl.bloom = 2; pl.th = 32; pl.nt = 256; pl.smem = 0;
        return pl;
    }
    // grid sizing: at least two waves of CTAs over the 148 SMs when the frame allows it
    auto enough_tiles = [&](int th) { return (long long)((d.W + FTW - 1) / FTW) * ((d.H + th - 1) / th) >= 2 * 148; };
    if (d.bloom_mode == 2 && !d.warp_on && fused_gauss_supported(d.ksize) && d.W >= 4 && d.H >= 4) {
        FusedPlan cand;
        const int force_th = env_int("CRT_GAUSS_TH", 0);      // tuning knobs (tile height / threads per CTA)
        for (int th : {32, 16}) {
            if (force_th && th != force_th) continue;
            const size_t smem = fused_gauss_smem(d.ksize, th);
            const bool two_ctas = smem + 12 * 1024 <= 113 * 1024;   // + static shared memory and the per-CTA reserve
            if (two_ctas || (th == 16 && smem <= 200 * 1024)) {
                cand.ok = true; cand.why = ""; cand.bloom = 2; cand.th = th; cand.smem = smem; cand.gauss_k = d.ksize;
                cand.nt = env_int("CRT_GAUSS_NT", th >= 32 ? 512 : 256);
                if (enough_tiles(th)) return cand;
            }
        }
        if (cand.ok) return cand;        // smallest tile that fits: most CTAs
    }
    // warp: per-row / per-column normalised coordinates, as the kernel computes them
    std::vector<float> xn(d.warp_on ? d.W : 0), yn(d.warp_on ? d.H : 0);
    for (int x = 0; x < (int)xn.size(); ++x) xn[x] = warp_norm((float)x, d.warp_cx, d.warp_dx);
    for (int y = 0; y < (int)yn.size(); ++y) yn[y] = warp_norm((float)y, d.warp_cy, d.warp_dy);
    FusedPlan best, small;
    best.why = "tile footprint does not fit in shared memory";
    const int force_th = env_int("CRT_FUSED_TH", 0);
    for (int th : {64, 32, 16}) {
        if (force_th && th != force_th) continue;
        size_t best_px = 0, best_aux = 0;
        int max_qh = 0, max_qw = 0;
        const int tiles_x = (d.W + FTW - 1) / FTW, tiles_y = (d.H + th - 1) / th;
        for (int ty = 0; ty < tiles_y; ++ty)
            for (int tx = 0; tx < tiles_x; ++tx) {
                Box o{tx * FTW, ty * th, imin((tx + 1) * FTW, d.W) - 1, imin((ty + 1) * th, d.H) - 1};
                Box q = o;
                if (d.warp_on) {
                    int x0 = 0x7fffffff, y0 = 0x7fffffff, x1 = -0x7fffffff, y1 = -0x7fffffff;
                    for (int y = o.y0; y <= o.y1; ++y) {
                        const bool edge_row = y == o.y0 || y == o.y1;
                        for (int x = o.x0; x <= o.x1; ++x) {
                            if (d.warp_mono && !edge_row && x != o.x0 && x != o.x1) continue;      // same sampling as the kernel
                            Taps t = warp_taps_n(d, xn[x], yn[y]);
                            x0 = imin(x0, t.ix); x1 = imax(x1, t.ix + 1); y0 = imin(y0, t.iy); y1 = imax(y1, t.iy + 1);
                        }
                    }
                    bool empty = false;
                    q = footprint_box(d, x0, y0, x1, y1, &empty);
                    if (empty) continue;
                }
                Box cells{0, 0, -1, -1};
                Box p = grow_for_bloom(d, q, &cells);
                size_t px = (size_t)box_w(p) * box_h(p);
                size_t aux = 0;
                if (d.bloom_mode == 1) aux = ((size_t)box_w(cells) + box_w(q)) * box_h(cells) * 3;   // ds cells + x-up-scaled cells
                if (d.bloom_mode == 2) aux = (size_t)box_h(p) * box_w(q) * 3 + (d.thr_on ? (size_t)th * FTW * 3 : 0);
                if (px > best_px) best_px = px;
                if (aux > best_aux) best_aux = aux;
                if (box_h(q) > max_qh) max_qh = box_h(q);
                if (box_w(q) > max_qw) max_qw = box_w(q);
            }
        const size_t smem = (best_px * 3 + best_aux) * sizeof(float);
        const bool fits = smem <= 200 * 1024 && max_qh <= FMAX_ROWS && max_qw <= FMAX_COLS && best_px < 21000;
        if (!fits) continue;
        FusedPlan c;
        c.ok = true; c.why = ""; c.bloom = d.bloom_mode; c.warp = d.warp_on;
        c.th = th; c.cap_px = (int)best_px; c.cap_aux = (int)best_aux; c.smem = smem;
        c.nt = env_int("CRT_FUSED_NT", th >= 32 ? 512 : 256);
        if (!best.ok) best = c;                          // tallest tile that fits at all
        if (smem + 20 * 1024 <= 113 * 1024) {            // two CTAs per SM (+ ~19 KB static shared memory and reserve)
            if (enough_tiles(th)) return c;              // tallest such tile that still fills the GPU twice over
            small = c;
        }
    }
    return small.ok ? small : best;
}

#if defined(__CUDACC__) && defined(CRT_TU_FUSED)      // launchers: compiled only in crt_tu_fused.cu
template <int BLOOM, bool WARP, int NT>
inline int launch_fused_t(LaunchEnv& env, const FusedPlan& pl, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state,
                          float* q_out, int has_prev, cudaStream_t st) {
    auto kern = k_fused<BLOOM, WARP, NT>;
    if (env.raise((const void*)kern, (int)pl.smem) &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess) return 2;
    dim3 grid((d.W + FTW - 1) / FTW, (d.H + pl.th - 1) / pl.th);
    FusedGeom g{pl.th, pl.cap_px, pl.cap_aux};
    kern<<<grid, NT, pl.smem, st>>>(d, f, in, out, state, q_out, has_prev, g);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

inline int run_fused(LaunchEnv& env, const FusedPlan& pl, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, float* q_out,
                     int has_prev, cudaStream_t st, int* launches) {
    int rc;
#define CRT_LAUNCH(B, W) (pl.nt == 512 ? launch_fused_t<B, W, 512>(env, pl, d, f, in, out, state, q_out, has_prev, st) \
                                       : launch_fused_t<B, W, 256>(env, pl, d, f, in, out, state, q_out, has_prev, st))
    if (d.warp_on) rc = d.bloom_mode == 1 ? CRT_LAUNCH(1, true) : CRT_LAUNCH(0, true);
    else rc = d.bloom_mode == 2 ? CRT_LAUNCH(2, false) : d.bloom_mode == 1 ? CRT_LAUNCH(1, false) : CRT_LAUNCH(0, false);
#undef CRT_LAUNCH
    ++*launches;
    return rc;
}
#endif

}  // namespace crt
