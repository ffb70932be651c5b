// crt_fused_gauss.cuh — fused tile kernel for the gaussian-bloom chain without warp
// (BASELINE.json configs[1]): same single-launch structure as k_fused, with the
// separable blur written for Blackwell's packed FP32 pipe.
//
//   * cv2.GaussianBlur's float32 arithmetic is kept bit for bit (row pass
//     s = x0*k0; s = fma(x_i, k_i, s); column pass s = c*k0; s = fma(x_+i + x_-i, k_i, s)),
//     but two outputs are computed per instruction with FMUL2 / FFMA2 / FADD2
//     (__fmul2_rn, __ffma2_rn, __fadd2_rn: per-component IEEE round-to-nearest, sm_100+).
//   * The thresholded source is stored TRANSPOSED and planar, St[ch][x][y], so the row
//     pass reads aligned (y, y+1) pairs with LDS.64 while blocking four outputs along x;
//     the row-pass result is row-major planar, Rp[ch][y][x], so the column pass reads
//     aligned (x, x+1) pairs while blocking four outputs along y.  Each pass costs
//     (K + 3) LDS.64 + 4 K packed FMAs per 8 outputs.
//   * The REPLICATE border is materialised while filling St (clamped coordinates), so
//     neither pass has border logic.
#pragma once
#include "crt_fused.cuh"

namespace crt {

#if defined(__CUDACC__)

__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

// cv2 row pass for one pair of outputs: in[0..K-1] are the K taps (each a pair of rows)
template <int K>
__device__ __forceinline__ float2 gauss_row2(const float2* in, const float* k) {
    if (K == 3) return __ffma2_rn(in[1], splat2(k[1]), __fmul2_rn(__fadd2_rn(in[0], in[2]), splat2(k[2])));
    if (K == 5) {
        const float2 inner = __ffma2_rn(in[2], splat2(k[2]), __fmul2_rn(__fadd2_rn(in[1], in[3]), splat2(k[3])));
        return __ffma2_rn(__fadd2_rn(in[4], in[0]), splat2(k[4]), inner);
    }
    float2 s = __fmul2_rn(in[0], splat2(k[0]));
#pragma unroll
    for (int i = 1; i < K; ++i) s = __ffma2_rn(in[i], splat2(k[i]), s);
    return s;
}
// cv2 column pass for one pair of outputs: c points at the centre tap
template <int K>
__device__ __forceinline__ float2 gauss_col2(const float2* c, const float* k) {
    constexpr int R = K / 2;
    float2 s = __fmul2_rn(c[0], splat2(k[R]));
#pragma unroll
    for (int i = 1; i <= R; ++i) s = __ffma2_rn(__fadd2_rn(c[i], c[-i]), splat2(k[R + i]), s);
    return s;
}

template <int K, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) k_fused_gauss(Dev d, FrameDev f, const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                    float* __restrict__ state, float* __restrict__ q_out, int has_prev, int th) {
    constexpr int R = K / 2, PW = FTW + 2 * R;
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(16) float s_fwd[1028], s_inv[1028];
    __shared__ float s_unit[256];
    __shared__ __align__(16) float s_pow[POW_TAB_FLOATS];
    __shared__ float s_rows[2 * 64], s_cols[2 * FTW];
    __shared__ int s_geo[8];
    const int PH = th + 2 * R;                      // padded rows; th is even, so PH is even
    const int pitch = PH;                           // floats between consecutive x in St
    float* St = sm;                                 // [3][PW][pitch]  thresholded source, transposed
    float* Rp = St + 3 * PW * pitch;                // [3][PH][FTW]    row-pass result
    float* T1 = Rp + 3 * PH * FTW;                  // [th][FTW][3]    graded tile (bloom is added to this)
    float* Bl = St;                                 // [3][th][FTW]    blurred tile (St is dead after the row pass)
    const int tid = threadIdx.x;
    const int ox0 = blockIdx.x * FTW, oy0 = blockIdx.y * th;
    const int ox1 = imin(ox0 + FTW, d.W) - 1, oy1 = imin(oy0 + th, d.H) - 1;
    const int trow = tid / FROW_THREADS, xb = ox0 + (tid % FROW_THREADS) * 4;

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
    float taps[K];
#pragma unroll
    for (int i = 0; i < K; ++i) taps[i] = d.taps[i];
    if (tid == 0) {          // uniform integer divisions once per CTA
        const int ps0 = (d.pix_uniform > 1 && d.text_mode != 1) ? d.pix_uniform : 1;
        const int jx0 = imax(ox0 - K / 2, 0), jx1 = imin(ox1 + K / 2, d.W - 1), jy0 = imax(oy0 - K / 2, 0), jy1 = imin(oy1 + K / 2, d.H - 1);
        s_geo[0] = ps0; s_geo[1] = jx0 / ps0; s_geo[2] = jy0 / ps0;
        s_geo[3] = jx1 / ps0 - s_geo[1] + 1; s_geo[4] = jy1 / ps0 - s_geo[2] + 1;
        s_geo[5] = (int)make_magic(s_geo[3]);
        s_geo[6] = (int)make_magic((th + 2 * (K / 2)) >> 1);
        s_geo[7] = (int)make_magic(th >> 2);
    }
    __syncthreads();
    MaskTabs mt{s_rows, s_cols, s_rows + 64, s_cols + FTW};

    // ---- phase 1: graded input over the in-image part of tile + halo; REPLICATE border materialised ----
    const int px0 = ox0 - R, py0 = oy0 - R;                         // padded origin (may be negative)
    const int PWe = ox1 - ox0 + 1 + 2 * R, PHe = oy1 - oy0 + 1 + 2 * R;
    const int ix0 = imax(px0, 0), ix1 = imin(ox1 + R, d.W - 1), iy0 = imax(py0, 0), iy1 = imin(oy1 + R, d.H - 1);
    {
        const int ps = s_geo[0], ux0 = s_geo[1], uy0 = s_geo[2], nux = s_geo[3], nuy = s_geo[4];
        const unsigned magic = (unsigned)s_geo[5];
        for (int u = tid; u < nux * nuy; u += NT) {
            const int uy = fastdiv(u, magic), ux = u - uy * nux;
            const int xa = imax((ux0 + ux) * ps, ix0), xe = imin((ux0 + ux) * ps + ps - 1, ix1);
            const int ya = imax((uy0 + uy) * ps, iy0), ye = imin((uy0 + uy) * ps + ps - 1, iy1);
            const F3 v1 = ps > 1 ? graded_source_lut(d, in, (uy0 + uy) * ps, (ux0 + ux) * ps, ya, xa, s_unit, s_pow) : graded_input_lut(d, in, ya, xa, s_unit, s_pow);
            const F3 v = bloom_src(d, v1);
            // padded positions that replicate this block (image borders extend outwards)
            const int xx0 = (xa == 0) ? 0 : xa - px0, xx1 = (xe == d.W - 1) ? PWe - 1 : xe - px0;
            const int yy0 = (ya == 0) ? 0 : ya - py0, yy1 = (ye == d.H - 1) ? PHe - 1 : ye - py0;
            float* c0 = St + xx0 * pitch + yy0;
            if (xx1 == xx0 + 1 && yy1 == yy0 + 1) {            // interior 2x2 block (pixel_size 2): straight-line stores
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float val = ch == 0 ? v.x : (ch == 1 ? v.y : v.z);
                    float* cc = c0 + ch * PW * pitch;
                    cc[0] = val; cc[1] = val; cc[pitch] = val; cc[pitch + 1] = val;
                }
            } else {
                for (int xx = xx0; xx <= xx1; ++xx, c0 += pitch)
                    for (int yy = 0; yy <= yy1 - yy0; ++yy) { c0[yy] = v.x; c0[PW * pitch + yy] = v.y; c0[2 * PW * pitch + yy] = v.z; }
            }
            const int tx0 = imax(xa, ox0), tx1 = imin(xe, ox1), ty0 = imax(ya, oy0), ty1 = imin(ye, oy1);
            float* t1 = T1 + ((ty0 - oy0) * FTW + (tx0 - ox0)) * 3;
            if (tx1 == tx0 + 1 && ty1 == ty0 + 1) {
                store_f3(t1, v1); store_f3(t1 + 3, v1); store_f3(t1 + FTW * 3, v1); store_f3(t1 + FTW * 3 + 3, v1);
            } else {
                for (int y = ty0; y <= ty1; ++y)
                    for (int x = tx0; x <= tx1; ++x) store_f3(T1 + ((y - oy0) * FTW + (x - ox0)) * 3, v1);
            }
        }
        for (int r = tid; r <= oy1 - oy0; r += NT) {
            const int y = oy0 + r;
            if (d.scan_mode == 1) mt.row_scan[r] = scan_row(d, f, y);
            else if (d.scan_mode == 2) { double t = ((double)y + f.phase) * d.scan_inv_period; mt.row_scan[r] = (float)(t - floor(t)); }
            if (d.vig_mode == 1) { const float ny = ((float)y - d.vig_cy) * d.vig_iry; mt.row_vig[r] = ny * ny; }
        }
        for (int c = tid; c <= ox1 - ox0; c += NT) {
            const int x = ox0 + c;
            if (d.scan_mode == 2) { double t = (d.scan_tan * (double)x) * d.scan_inv_period; mt.col_scan[c] = (float)(t - floor(t)); }
            if (d.vig_mode == 1) { const float nx = ((float)x - d.vig_cx) * d.vig_irx; mt.col_vig[c] = nx * nx; }
        }
    }
    __syncthreads();

    // ---- phase 2a: row pass, 4 outputs along x for a pair of rows per task ------------------------------
    {
        const int nyp = PH >> 1;
        const unsigned magic = (unsigned)s_geo[6];
        const int hp = pitch >> 1;                                  // float2 stride between consecutive x
        for (int u = tid; u < 3 * (FTW / 4) * nyp; u += NT) {
            const int t = fastdiv(u, magic), yp = u - t * nyp;
            const int xblk = t & (FTW / 4 - 1), ch = t >> 4;        // FTW / 4 == 16
            const float2* src = reinterpret_cast<const float2*>(St + (ch * PW + xblk * 4) * pitch) + yp;
            float2 v[K + 3];
#pragma unroll
            for (int i = 0; i < K + 3; ++i) v[i] = src[i * hp];
            float* o0 = Rp + (ch * PH + 2 * yp) * FTW + xblk * 4;
            float2 r0 = gauss_row2<K>(v + 0, taps), r1 = gauss_row2<K>(v + 1, taps), r2 = gauss_row2<K>(v + 2, taps), r3 = gauss_row2<K>(v + 3, taps);
            *reinterpret_cast<float4*>(o0) = make_float4(r0.x, r1.x, r2.x, r3.x);
            *reinterpret_cast<float4*>(o0 + FTW) = make_float4(r0.y, r1.y, r2.y, r3.y);
        }
    }
    __syncthreads();

    // ---- phase 2b: column pass, 4 outputs along y for a pair of columns per task ----------------------------
    {
        const int nyb = th >> 2;
        const unsigned magic = (unsigned)s_geo[7];
        for (int u = tid; u < 3 * nyb * (FTW / 2); u += NT) {
            const int xp = u & (FTW / 2 - 1), t = u >> 5;           // FTW / 2 == 32
            const int ch = fastdiv(t, magic), yblk = t - ch * nyb;
            const float2* src = reinterpret_cast<const float2*>(Rp + (ch * PH + yblk * 4) * FTW) + xp;
            float2 v[K + 3];
#pragma unroll
            for (int i = 0; i < K + 3; ++i) v[i] = src[i * (FTW / 2)];
            float2* o0 = reinterpret_cast<float2*>(Bl + (ch * th + yblk * 4) * FTW) + xp;
#pragma unroll
            for (int j = 0; j < 4; ++j) o0[j * (FTW / 2)] = gauss_col2<K>(v + R + j, taps);
        }
    }
    __syncthreads();

    // ---- phase 4: output pixels ---------------------------------------------------------------------------------
    if (xb > ox1) return;
    auto pixel = [&](int y, int x, int k) -> F3 {
        const int ly = y - oy0, lx = x - ox0;
        F3 v = load_f3(T1 + (ly * FTW + lx) * 3);
        const float* b = Bl + ly * FTW + lx;
        v = add_bloom(d, v, mk3(b[0], b[th * FTW], b[2 * th * FTW]));
        v = after_bloom_fast(d, f, v, y, x, s_fwd, s_inv, mt, ly, lx);
        if (d.text_mode == 2) v = text_blend(d, v, y, x);
        return v;
    };
    for (int y = oy0 + trow; y <= oy1; y += NT / FROW_THREADS) finish_quad(d, state, out, q_out, has_prev, y, xb, imin(4, ox1 - xb + 1), pixel);
}

#if defined(CRT_TU_GAUSS)      // launcher: compiled only in the translation unit that owns these kernels (build.py)
template <int K, int GAUSS_NT>
inline int launch_fused_gauss_t(LaunchEnv& env, int th, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, float* q_out,
                                int has_prev, cudaStream_t st) {
    const size_t smem = fused_gauss_smem(K, th);
    auto kern = k_fused_gauss<K, GAUSS_NT>;
    if (env.raise((const void*)kern, (int)smem) &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
    dim3 grid((d.W + FTW - 1) / FTW, (d.H + th - 1) / th);
    kern<<<grid, GAUSS_NT, smem, st>>>(d, f, in, out, state, q_out, has_prev, th);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

inline int run_fused_gauss(LaunchEnv& env, int th, int nt, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, float* q_out,
                           int has_prev, cudaStream_t st, int* launches) {
    int rc = 4;
#define CRT_GAUSS_CASE(K) case K: rc = nt == 512 ? launch_fused_gauss_t<K, 512>(env, th, d, f, in, out, state, q_out, has_prev, st) \
                                                  : launch_fused_gauss_t<K, 256>(env, th, d, f, in, out, state, q_out, has_prev, st); break;
    switch (d.ksize) {
        CRT_GAUSS_CASE(5) CRT_GAUSS_CASE(7) CRT_GAUSS_CASE(9) CRT_GAUSS_CASE(11) CRT_GAUSS_CASE(13) CRT_GAUSS_CASE(25)
        default: break;
    }
#undef CRT_GAUSS_CASE
    if (rc != 4) ++*launches;
    return rc;
}

#endif  // CRT_TU_GAUSS

#endif  // __CUDACC__

}  // namespace crt
