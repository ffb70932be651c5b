// crt_fused_gauss_ps2.cuh — gaussian-bloom chain at pixel_size 2 (BASELINE.json configs[1]):
// the separable blur of crt_fused_gauss.cuh evaluated at BLOCK resolution.
//
// With pixel_size 2 and even frame dimensions the thresholded bloom source is constant over
// aligned 2x2 blocks, S(x, y) = Sb(x >> 1, y >> 1), and cv2's REPLICATE border is a clamp of
// the block index.  cv2.GaussianBlur's float32 arithmetic is kept bit for bit (same tap order,
// same fused multiply-adds); what changes is how much of it has to be evaluated:
//   * graded input + threshold: one value per block of tile + halo (36 x 20 blocks for a
//     64 x 32 tile and K = 9, instead of 72 x 40 pixels);
//   * row pass: its result depends on the block row only, so it runs over 20 block rows instead
//     of 40 pixel rows.  A task blocks four outputs along x for a pair of block rows; the K + 3
//     taps it needs are only (K + 3) / 2 + 1 distinct block columns (LDS.64 each, transposed
//     planar source Sb[ch][bx][by]), combined with FMUL2 / FFMA2 in cv2's order;
//   * column pass: merged into the output phase.  A thread owns a 4 x 2 pixel patch; the 2 K + 1
//     row-pass rows it needs are R + 1 distinct block rows, read as (x, x+1) pairs, and the
//     blurred values stay in registers (no blurred tile in shared memory, no extra barrier).
// The per-pixel tail (triad LUT, masks, persistence, stores) is ps2_patch_tail (crt_fused_ps2.cuh).
// Shared memory: ~31 KB dynamic + ~11 KB static for K = 9 -> three 256-thread CTAs per SM at 80 registers.
//
// Measured and NOT adopted (round 1, runs 47-49): grading the blocks in a kernel of its own, once per block of
// the frame (no halo recomputation; 12 us at 70 % issue rate, block images in L2) with this kernel reading the
// block images.  Correct, but 33 650 vs 37 060 frames/s on BASELINE configs[1]: the tile kernel keeps 150 of
// its 256 thread-instructions per pixel and, without the arithmetic of the grading phase to cover its loads,
// issues at 36 % instead of 50 %; the two kernels of a frame form a serial chain (the grading kernel only gets
// SM slots when the previous frame's tile kernel drains), so the sum, not the overlap, is what is paid.
#pragma once
#include "crt_fused_gauss.cuh"
#include "crt_fused_ps2.cuh"

namespace crt {

// geometry of the block-resolution tile for a K-tap kernel (radius R = K / 2)
struct GaussPs2Geo {
    int hb;        // halo blocks per side = ceil(R / 2)
    int off;       // R & 1: pixel offset of the halo origin inside its block
    int nbx, nby;  // blocks per tile incl. halo
    int pitch;     // floats between consecutive bx in Sb (even, == 2 mod 4: fewest bank conflicts)
};
CRT_HD GaussPs2Geo gauss_ps2_geo(int K) {
    GaussPs2Geo g;
    const int R = K / 2;
    g.hb = (R + 1) / 2; g.off = R & 1;
    g.nbx = P2_TW / 2 + 2 * g.hb; g.nby = P2_TH / 2 + 2 * g.hb;
    g.pitch = (g.nby & 3) == 2 ? g.nby : g.nby + 2;
    return g;
}
inline size_t fused_gauss_ps2_smem(int K, bool state_tile = false, bool input_tile = false) {
    const GaussPs2Geo g = gauss_ps2_geo(K);
    return sizeof(float) * ((size_t)3 * g.nbx * g.pitch + (size_t)3 * g.nby * P2_TW + (size_t)3 * (P2_TH / 2) * (P2_TW / 2) +
                            (state_tile ? (size_t)P2_TH * P2_TW * 3 : 0)) + (input_tile ? (size_t)g.nby * 256 : 0);
}
// input tile by TMA: one 256-byte box must cover the tile's blocks, the aberration shift and the 16-byte alignment slack
inline bool fused_gauss_ps2_tin_ok(const Dev& d, int K) {
    const GaussPs2Geo g = gauss_ps2_geo(K);
    const int a0 = d.aberr != 0 ? d.aberr_mod : 0, as = a0 > (d.W >> 1) ? a0 - d.W : a0, aa = as < 0 ? -as : as;
    return (d.W & 7) == 0 && 6 * g.nbx + 6 * aa + 15 <= 256;
}
CRT_HD bool fused_gauss_ps2_supported(const Dev& d, bool glitch_on) {
    return d.bloom_mode == 2 && d.pix_uniform == 2 && d.even_dims && (d.W & 3) == 0 && !d.warp_on && !glitch_on && d.text_mode == 0;
}

#if defined(__CUDACC__)

// TST: the tile's previous state arrives by ONE tiled tensor-map copy (TMA) issued at the top of the tile's iteration and is
// consumed two phases later from shared memory; the new state (or the pre-warp image of the two-pass path) goes back into the
// same tile and leaves with one TMA store — as in k_fused_ps2_pipe.  ncu (round 2, run 3) put 10.6 % of this kernel's
// warp-stall samples on the first use of the per-thread state loads (L2 latency: at 1080p the state lives in L2) and another
// 10 % on the table staging at kernel entry; with TST the tables also arrive by bulk copies that are only waited for at their
// first use, after the first tile's input loads are in flight.  24.5 KB more shared memory: still three CTAs per SM.
// TIN (with TST): the tile's input bytes — NBY even rows x 256 bytes around what its blocks read — arrive by one tensor-map
// copy per tile as well, issued one tile ahead into a single buffer (the copy for tile t + 1 starts when phase 1 of tile t has
// read the buffer; for the first tile at kernel entry, before the previous frame's kernel has finished).  ncu (run 25) put
// 9 % of the stall samples on the first use of the per-thread byte loads.  Tiles on the left / right frame edge, where the
// chromatic aberration wraps around (np.roll), keep the per-thread loads.
// CLIP (with TST): clip mode — a run of frames in one launch, (frame, tile) items from an atomic counter, a tile's frames chained
// through per-tile flags (ClipArgs, crt_fused_ps2.cuh); `in` / `out` / `frame` describe the first frame of the run
template <int K, bool FAST, int MINB, bool TST, int SPEC = 0, bool TIN = false, bool CLIP = false>
__global__ void __launch_bounds__(P2_NT, MINB) k_fused_gauss_ps2(Dev d_arg, FrameDev f_arg, const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                           float* __restrict__ state, float* __restrict__ q_out, int has_prev,
                                                           const __grid_constant__ CUtensorMap map_st, const __grid_constant__ CUtensorMap map_in,
                                                           int frame, int th, const __grid_constant__ ClipArgs ca) {
    Dev d = d_arg;
    FrameDev f = f_arg;
    specialise<SPEC>(d, f);             // SPEC != 0: feature flags become compile-time constants (crt_fused_ps2.cuh)
    constexpr int R = K / 2, HB = (R + 1) / 2, OFF = R & 1;
    constexpr int NBX = P2_TW / 2 + 2 * HB, NBY = P2_TH / 2 + 2 * HB;
    constexpr int PITCH = (NBY & 3) == 2 ? NBY : NBY + 2;
    constexpr int MR = ((K + 2 + OFF) >> 1) + 1;                  // distinct block columns under 4 outputs' taps
    constexpr int MC = ((2 * R + 1 + OFF) >> 1) + 1;              // distinct block rows under 2 output rows' taps
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(16) float s_lut[2 * 1028];
    __shared__ __align__(16) float s_sel[3][12];
    float* const s_fwd = s_lut;
    float* const s_inv = s_lut + 1028;
    __shared__ float s_unit[256];
    __shared__ __align__(16) float s_pow[POW_TAB_FLOATS];
    __shared__ float s_rows[2 * P2_TH], s_cols[2 * P2_TW];
    __shared__ __align__(8) uint64_t bar_tab, bar_st, bar_in;
    constexpr int RAW_BYTES = TIN ? NBY * P2_RAW_W : 0;
    float* const s_state = sm;                      // [TH][TW*3]        state tile (TST only; first: TMA wants 128-byte alignment)
    uint8_t* const s_raw = reinterpret_cast<uint8_t*>(sm + (TST ? P2_TH * P2_TW * 3 : 0));      // [NBY][256] input bytes (TIN only)
    float* Sb = sm + (TST ? P2_TH * P2_TW * 3 : 0) + RAW_BYTES / 4;      // [3][NBX][PITCH]   thresholded bloom source per block, transposed
    float* Rp = Sb + 3 * NBX * PITCH;               // [3][NBY][P2_TW]   row-pass result per block row
    float* T1 = Rp + 3 * NBY * P2_TW;               // [3][TH/2][TW/2]   graded block values of the tile
    const int tid = threadIdx.x;
    griddep_launch_dependents();        // the next frame's kernel may begin its state-independent phases (see launch_pdl)
    const float* lut_a = d.triad_comp ? d.triad_comp : d.lut_fwd;
    const float* lut_b = d.triad_comp ? d.triad_comp + 1028 : d.lut_inv;
    const bool use_state = has_prev && !q_out;      // a previous state to fetch
    const bool tile_out = TST && (state || q_out);  // the result leaves through the shared-memory tile
    if (TST) {
        if (tid == 0) {                             // tables by bulk copies (16-byte multiples; element 1024 of each LUT below)
            mbar_init(&bar_tab, 1); mbar_init(&bar_st, 1); mbar_init(&bar_in, 1);
            fence_mbar_init();
            const uint32_t bytes = (d.triad_mode >= 2 ? 2 * 4096 : 0) + (d.col_gamma ? POW_TAB_FLOATS * 4 : 0);
            mbar_expect_tx(&bar_tab, bytes);
            if (d.triad_mode >= 2) { bulk_g2s(s_fwd, lut_a, 4096, &bar_tab); bulk_g2s(s_inv, lut_b, 4096, &bar_tab); }
            if (d.col_gamma) bulk_g2s(s_pow, d.pow_tab, POW_TAB_FLOATS * 4, &bar_tab);
            if (d.triad_mode >= 2) { s_fwd[1024] = lut_a[1024]; s_inv[1024] = lut_b[1024]; }
        }
    } else {
        if (d.triad_mode >= 2) {
            reinterpret_cast<float4*>(s_fwd)[tid] = reinterpret_cast<const float4*>(lut_a)[tid];
            reinterpret_cast<float4*>(s_inv)[tid] = reinterpret_cast<const float4*>(lut_b)[tid];
            if (tid == 0) { s_fwd[1024] = lut_a[1024]; s_inv[1024] = lut_b[1024]; }
        }
        if (d.col_gamma) for (int i = tid; i < POW_TAB_FLOATS; i += blockDim.x) s_pow[i] = d.pow_tab[i];
    }
    s_unit[tid] = __fdiv_rn((float)tid, 255.0f);
    ps2_fill_sel(s_sel, tid, d.bgr);
    float taps[K];                                  // the kernel is symmetric: R + 1 registers
#pragma unroll
    for (int i = 0; i <= R; ++i) { taps[R + i] = d.taps[R + i]; taps[R - i] = taps[R + i]; }
    MaskTabs mt{s_rows, s_cols, s_rows + P2_TH, s_cols + P2_TW};
    __shared__ int s_item[2], s_prev[2], s_hint;    // clip mode: the item of the next iteration; tile and frame of the item being stored;
    const int rank = (CLIP && ca.per_sm > 0) ? ((int)blockIdx.x % ca.per_sm) * ((int)gridDim.x / ca.per_sm) + (int)blockIdx.x / ca.per_sm : (int)blockIdx.x;
    if (CLIP && tid == 0) { s_hint = 0; s_item[0] = ca.static_items ? rank : atomicAdd(ca.sync, 1); }      // the item's flag as seen earlier
    __syncthreads();
    const int tiles_x = (d.W + P2_TW - 1) / P2_TW, ntiles = tiles_x * ((d.H + th - 1) / th);
    const int nitems = CLIP ? ntiles * ca.nf : ntiles;
    const unsigned magic_tx = make_magic(tiles_x);
    int fr = 0;                                     // frame of the current item within the run
    int hint = 0;                                   // clip mode, thread 0: the current item's flag as seen one iteration ago
    auto split = [&](int item, int& ifr, int& iby, int& ibx) {      // ifr on entry: a frame not after the item's (items only grow)
        int t = item - ifr * ntiles;
        while (t >= ntiles) { t -= ntiles; ++ifr; }
        iby = fastdiv(t, magic_tx); ibx = t - iby * tiles_x;
    };
    const int nby_t = (th >> 1) + 2 * HB, nblk = NBX * nby_t;      // block rows / blocks of a tile of th rows + halo (th <= P2_TH: the buffers hold NBY rows)
    const uint32_t st_bytes = (uint32_t)th * P2_TW * 3 * 4;
    const int step_y = gridDim.x / tiles_x, step_x = gridDim.x - step_y * tiles_x;      // tile += gridDim.x without a division per tile
    int tby = blockIdx.x / tiles_x, tbx = blockIdx.x - tby * tiles_x;
    if (CLIP) split(s_item[0] < nitems ? s_item[0] : 0, fr, tby, tbx);
    const int a0 = d.aberr != 0 ? d.aberr_mod : 0;
    const int as = a0 > (d.W >> 1) ? a0 - d.W : a0;                 // signed aberration shift (aberr_mod is taken modulo W)
    const int aa = as < 0 ? -as : as;
    if (TIN && tid == 0 && (CLIP ? s_item[0] : (int)blockIdx.x) < nitems) {      // first tile's input: independent of the previous kernel
        mbar_expect_tx(&bar_in, RAW_BYTES);
        tma_load_2d_hint(s_raw, &map_in, (6 * ((tbx * P2_TW >> 1) - HB) - 3 * aa) & ~15, (frame + fr) * d.hh + (tby * th >> 1) - HB, &bar_in, L2_EVICT_FIRST);
    }
    int iter = 0;
    for (int tile = CLIP ? s_item[0] : (int)blockIdx.x; tile < nitems; ++iter) {    // persistent CTAs, tables staged once
        if (CLIP) {
            split(tile, fr, tby, tbx);
            const FrameVar fv = ca.fv[fr];
            f.phase32 = fv.phase32; f.phase = fv.phase; f.flicker = fv.flicker;
        }
        const int ox0 = tbx * P2_TW, oy0 = tby * th;
        const int ox1 = imin(ox0 + P2_TW, d.W) - 1, oy1 = imin(oy0 + th, d.H) - 1;
        const uint8_t* const in_f = CLIP ? in + fr * ca.frame_bytes : in;
        uint8_t* const out_f = CLIP ? out + fr * ca.frame_bytes : out;
        int next = tile + (int)gridDim.x, nfr = fr, nbx = tbx + step_x, nby = tby + step_y;      // next tile of this CTA
        if (nbx >= tiles_x) { nbx -= tiles_x; ++nby; }
        // clip mode's bookkeeping, split between thread 0 (state traffic) and thread CLIP_B (next item): see k_fused_ps2_pipe
        bool owed = false;                   // thread 0: the previous item's completion is still to be published
        const bool own = CLIP && ca.static_items == 2, own_one = own && rank + (int)gridDim.x >= ntiles;      // owned tiles
        if (CLIP && tid == CLIP_B) {
            if (own) {
                const int nt = tby * tiles_x + tbx + (int)gridDim.x;
                next = nt < ntiles ? fr * ntiles + nt : (fr + 1 < ca.nf ? (fr + 1) * ntiles + rank : nitems);
            } else next = ca.static_items ? tile + (int)gridDim.x : atomicAdd(ca.sync, 1);
        }
        if (own && tid == 0) {
            if (iter > 0) bulk_wait_read(); else griddep_wait();
            if (own_one && fr > 0) { bulk_wait_all(); clip_settle(ca.release); }
            mbar_expect_tx(&bar_st, st_bytes); tma_load_2d(s_state, &map_st, ox0 * 3, oy0, &bar_st);
        }
        if (CLIP && !own && tid == 0) {
            if (iter > 0) bulk_wait_read(); else griddep_wait();
            if (iter > 0) owed = true;
            if (fr > 0 && (iter == 0 || s_hint < fr || (ca.release & 16))) {
                if (owed) { bulk_wait_all(); clip_publish(ca.sync + 1 + s_prev[0], s_prev[1] + 1, ca.release); owed = false; }
                clip_wait(ca.sync + 1 + tby * tiles_x + tbx, fr);
            } else if (fr > 0) clip_acquire(ca.release);
            mbar_expect_tx(&bar_st, st_bytes); tma_load_2d(s_state, &map_st, ox0 * 3, oy0, &bar_st);
        }
        if (!CLIP && TST && tile_out && tid == 0) {
            // The previous tile's TMA store must have drained the tile buffer; then fetch this tile's state.  The first tile
            // waits for the previous kernel of the stream first (this thread only: the other warps start their grading).
            if (iter > 0) bulk_wait_read(); else griddep_wait();
            if (use_state) { mbar_expect_tx(&bar_st, st_bytes); tma_load_2d(s_state, &map_st, ox0 * 3, oy0, &bar_st); }
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

        // ---- phase 1: graded value + bloom source per block of tile + halo (clamped block index = REPLICATE) ----
        const int gbx0 = (ox0 >> 1) - HB, gby0 = (oy0 >> 1) - HB;
        {
            constexpr int NIT = (NBX * NBY + P2_NT - 1) / P2_NT;
            // blocks are dealt to the threads rotated by 224: the last, partial round skips warp 0 (and half of warp 7), which keeps
            // the state traffic of the tile (waits, fences) and so reaches the barrier with the others
            const int vt = (tid + 224) & (P2_NT - 1);
            uint32_t raw[NIT][3];                                 // 32-bit: a byte array would live in local memory
            // all loads first: their latencies overlap.  Tiles away from the left / right frame edge need neither the
            // block clamp nor the aberration wrap in x: one 32-bit offset per block, byte offsets per channel.
            const bool x_inside = 2 * gbx0 - aa >= 0 && 2 * (gbx0 + NBX - 1) + aa < d.W;       // tile-uniform
            if (TIN) mbar_wait(&bar_in, iter & 1);                // this tile's input bytes have landed
            if (TIN && x_inside) {
                // byte 0 of the buffer is byte (6 gbx0 - 3 aa) & ~15 of the frame row; buffer row = (clamped) block row - gby0
                const int xoff = 6 * gbx0 - ((6 * gbx0 - 3 * aa) & ~15);
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    if ((vt & ~31) + it * P2_NT >= nblk) break;
                    const int u = imin(vt + it * P2_NT, nblk - 1);
                    const int bj = u / NBX, bi = u - bj * NBX;
                    const uint8_t* p = s_raw + (imin(imax(gby0 + bj, 0), d.hh - 1) - gby0) * P2_RAW_W + 6 * bi + xoff;
                    raw[it][0] = p[-3 * as];
                    raw[it][1] = p[1];
                    raw[it][2] = p[3 * as + 2];
                }
            } else if (x_inside) {
                const int W3 = d.W * 3;
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    if ((vt & ~31) + it * P2_NT >= nblk) break;        // warp-uniform: whole surplus warps skip the iteration
                    const int u = imin(vt + it * P2_NT, nblk - 1);     // surplus lanes repeat the last block (no divergence)
                    const int bj = u / NBX, bi = u - bj * NBX;
                    const int sy = 2 * imin(imax(gby0 + bj, 0), d.hh - 1);
                    const uint8_t* p = in_f + (unsigned)(sy * W3 + 6 * (gbx0 + bi));      // < 2^31 (checked by plan_fused)
                    raw[it][0] = p[-3 * as];
                    raw[it][1] = p[1];
                    raw[it][2] = p[3 * as + 2];
                }
            } else {
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    if ((vt & ~31) + it * P2_NT >= nblk) break;
                    const int u = imin(vt + it * P2_NT, nblk - 1);
                    const int bj = u / NBX, bi = u - bj * NBX;
                    const int sx = 2 * imin(imax(gbx0 + bi, 0), d.hw - 1), sy = 2 * imin(imax(gby0 + bj, 0), d.hh - 1);
                    const uint8_t* row = in_f + (size_t)sy * d.W * 3;
                    raw[it][0] = row[wrap(sx - a0, d.W) * 3 + 0];
                    raw[it][1] = row[sx * 3 + 1];
                    raw[it][2] = row[wrap(sx + a0, d.W) * 3 + 2];
                }
            }
            if (TST && iter == 0) mbar_wait(&bar_tab, 0);   // the tables have landed (the loads above are already in flight)
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                if ((vt & ~31) + it * P2_NT >= nblk) break;
                const int u = imin(vt + it * P2_NT, nblk - 1);
                const int bj = u / NBX, bi = u - bj * NBX;
                const F3 v1 = colour(d, mk3(s_unit[raw[it][0]], s_unit[raw[it][1]], s_unit[raw[it][2]]), s_pow);
                const F3 sv = bloom_src(d, v1);
                float* s = Sb + bi * PITCH + bj;
                s[0] = sv.x; s[NBX * PITCH] = sv.y; s[2 * NBX * PITCH] = sv.z;
                const int ti = bi - HB, tj = bj - HB;
                if ((unsigned)ti < (unsigned)(P2_TW / 2) && (unsigned)tj < (unsigned)(P2_TH / 2)) {
                    float* t = T1 + tj * (P2_TW / 2) + ti;
                    t[0] = v1.x; t[(P2_TH / 2) * (P2_TW / 2)] = v1.y; t[2 * (P2_TH / 2) * (P2_TW / 2)] = v1.z;
                }
            }
        }
        if (CLIP && tid == 0) {      // (in the slack the rotation above leaves warp 0; the store was issued a phase ago)
            if (own) { if (iter > 0 && !own_one) { bulk_wait_all(); clip_settle(ca.release); } }
            else {
                if (owed) { bulk_wait_all(); clip_publish(ca.sync + 1 + s_prev[0], s_prev[1] + 1, ca.release); }
                s_prev[0] = tby * tiles_x + tbx; s_prev[1] = fr;
            }
        }
        __syncthreads();
        if (CLIP && tid == CLIP_B) {
            s_item[(iter & 1) ^ 1] = next; split(next < nitems ? next : 0, nfr, nby, nbx);
            if (next < nitems && nfr > 0 && !own) hint = clip_peek(ca.sync + 1 + nby * tiles_x + nbx);
        }
        if (TIN && tid == (CLIP ? CLIP_B : 0) && next < nitems) {      // the buffer has been read: fetch the next tile's input bytes
            mbar_expect_tx(&bar_in, RAW_BYTES);
            tma_load_2d_hint(s_raw, &map_in, (6 * ((nbx * P2_TW >> 1) - HB) - 3 * aa) & ~15, (frame + nfr) * d.hh + (nby * th >> 1) - HB, &bar_in, L2_EVICT_FIRST);
        }

        // ---- phase 2: row pass over block rows; a task = 4 outputs along x for a pair of block rows ----
        const int nyp = (nby_t + 1) >> 1;                                    // pairs of block rows
        for (int u = tid; u < 3 * nyp * (P2_TW / 4); u += P2_NT) {
            const int xblk = u & (P2_TW / 4 - 1), t = u >> 4;                // P2_TW / 4 == 16
            const int ch = (t >= nyp) + (t >= 2 * nyp), yp = t - ch * nyp;
            const float2* src = reinterpret_cast<const float2*>(Sb + (ch * NBX + 2 * xblk) * PITCH) + yp;
            float2 v[MR];
#pragma unroll
            for (int m = 0; m < MR; ++m) v[m] = src[m * (PITCH / 2)];
            float2 r[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                float2 a[K];
#pragma unroll
                for (int i = 0; i < K; ++i) a[i] = v[(jj + i + OFF) >> 1];
                r[jj] = gauss_row2<K>(a, taps);
            }
            float* o0 = Rp + (ch * NBY + 2 * yp) * P2_TW + xblk * 4;
            *reinterpret_cast<float4*>(o0) = make_float4(r[0].x, r[1].x, r[2].x, r[3].x);
            *reinterpret_cast<float4*>(o0 + P2_TW) = make_float4(r[0].y, r[1].y, r[2].y, r[3].y);
        }
        __syncthreads();
        griddep_wait();         // previous kernel of the stream complete: state / pre-warp image / noise may be touched from here on
        if (TST && use_state) mbar_wait(&bar_st, iter & 1);      // this tile's previous state has landed

        // ---- phase 3 + 4: column pass in registers, then the per-pixel tail for a 4 x 2 patch ----
        const int tx = tid & 15, ty = tid >> 4;
        const int xb = ox0 + 4 * tx, y0 = oy0 + 2 * ty;
        if (xb <= ox1 && y0 <= oy1) {
            float bl[2][4][3];
            float t1[2][3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float2 tv = *reinterpret_cast<const float2*>(T1 + (ch * (P2_TH / 2) + ty) * (P2_TW / 2) + 2 * tx);
                t1[0][ch] = tv.x; t1[1][ch] = tv.y;
#pragma unroll
                for (int h = 0; h < 2; ++h) {                                   // pixel pairs (xb, xb+1), (xb+2, xb+3)
                    const float2* src = reinterpret_cast<const float2*>(Rp + (ch * NBY + ty) * P2_TW + 4 * tx + 2 * h);
                    float2 v[MC];
#pragma unroll
                    for (int m = 0; m < MC; ++m) v[m] = src[m * (P2_TW / 2)];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        // pixel row 2 ty + r, tap e in [-R, R] -> block row ty + ((r + e + R + OFF) >> 1)
                        float2 s = __fmul2_rn(v[(r + R + OFF) >> 1], splat2(taps[R]));
#pragma unroll
                        for (int i = 1; i <= R; ++i)
                            s = __ffma2_rn(__fadd2_rn(v[(r + i + R + OFF) >> 1], v[(r - i + R + OFF) >> 1]), splat2(taps[R + i]), s);
                        bl[r][2 * h][ch] = s.x; bl[r][2 * h + 1][ch] = s.y;
                    }
                }
            }
            ps2_patch_tail<true, FAST>(d, f, mt, s_fwd, s_inv, s_sel, state, out_f, q_out, has_prev, ox0, oy0, ox1, oy1, xb, y0, t1,
                                       [&](int r, int k) { return mk3(bl[r][k][0], bl[r][k][1], bl[r][k][2]); },
                                       tile_out ? s_state + (y0 - oy0) * (P2_TW * 3) + 12 * tx : nullptr, tile_out);
        }
        if (CLIP && tid == CLIP_B) s_hint = hint;      // (the look's latency has passed behind the tail)
        if (tile_out) fence_proxy_async();      // the new state in shared memory -> visible to the TMA engine
        __syncthreads();        // everyone is done with this tile's tables (and state tile) before the next tile overwrites them
        if (tile_out && tid == 0) {             // one coalesced TMA store per tile (rows outside the frame are clipped); the pre-warp
            tma_store_2d_hint(&map_st, s_state, ox0 * 3, oy0, q_out ? L2_EVICT_LAST : L2_EVICT_NORMAL);      // image is read back by the next kernel
            bulk_commit();
        }
        if (CLIP) tile = s_item[(iter & 1) ^ 1];      // written by thread 0 before this iteration's barriers
        else { tile = next; tbx = nbx; tby = nby; }
    }
    if (tile_out && tid == 0) {
        bulk_wait_all();      // the last tile's store has completed before the CTA exits
        if (CLIP && iter > 0 && ca.static_items != 2) clip_publish(ca.sync + 1 + s_prev[0], s_prev[1] + 1, ca.release);
    }
}

#if defined(CRT_TU_GAUSS_PS2)      // launcher: compiled only in the translation unit that owns these kernels (build.py)
template <int K, int MINB>
inline int launch_fused_gauss_ps2_t(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, float* q_out,
                                    int has_prev, cudaStream_t st, bool pdl, const Ps2Maps* maps, const CUtensorMap* gmap_in) {
    const bool fast = d.triad_mode == 2 && d.triad_comp && d.vig_mode <= 1;
    // state tile by TMA: needs the tensor map of the state buffer (or of the pre-warp image) and the tile to fit beside three CTAs
    static const bool use_tst = env_int("CRT_GPS2_TMA_STATE", 1) != 0;
    // (measured, run 9: 30.1 -> 29.0 us per 1080p frame with a state to blend; the pre-warp image of the two-pass path is 2 % faster
    // through per-thread stores, as in round 1)
    const bool tst = maps && use_tst && state && !q_out && fused_gauss_ps2_smem(K, true) + 16 * 1024 <= 76 * 1024;
    static const bool use_tin = env_int("CRT_GPS2_TMA_INPUT", 1) != 0;
    const bool tin = tst && use_tin && gmap_in && fused_gauss_ps2_tin_ok(d, K) && fused_gauss_ps2_smem(K, true, true) + 14 * 1024 <= 75 * 1024;
    const size_t smem = fused_gauss_ps2_smem(K, tst, tin);
    auto kern = tst ? (fast ? k_fused_gauss_ps2<K, true, MINB, true> : k_fused_gauss_ps2<K, false, MINB, true>)
                    : (fast ? k_fused_gauss_ps2<K, true, MINB, false> : k_fused_gauss_ps2<K, false, MINB, false>);
    static const bool use_spec = env_int("CRT_SPEC", 1) != 0;
    if (tin) kern = fast ? k_fused_gauss_ps2<K, true, MINB, true, 0, true> : k_fused_gauss_ps2<K, false, MINB, true, 0, true>;
    if (K == 9 && use_spec && tst && fast && spec_matches(SPEC_GRADED, d, f.flicker_on != 0, fast))      // BASELINE configs[1]
        kern = tin ? k_fused_gauss_ps2<K == 9 ? 9 : K, true, MINB, true, K == 9 ? SPEC_GRADED : 0, true>
                   : k_fused_gauss_ps2<K == 9 ? 9 : K, true, MINB, true, K == 9 ? SPEC_GRADED : 0, false>;
    // first pass of the two-pass path with every stage on: BASELINE configs[3] (K = 9) and [4] (K = 25)
    if ((K == 9 || K == 25) && use_spec && !tst && q_out && fast && spec_matches(SPEC_FULL | SP_THR * (K == 9), d, f.flicker_on != 0, fast))
        kern = k_fused_gauss_ps2<(K == 9 || K == 25) ? K : 9, true, MINB, false, (K == 9 || K == 25) ? (SPEC_FULL | SP_THR * (K == 9)) : 0>;
    // per context and kernel: opt-in shared-memory size, then the number of CTAs the device holds (persistent grid)
    auto it = env.memo.find((const void*)kern);
    if (it == env.memo.end()) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, P2_NT, smem);
        it = env.memo.emplace((const void*)kern, env.sms * (per_sm > 0 ? per_sm : 1)).first;
    }
    static const bool persist = env_int("CRT_PS2_PERSIST", 1) != 0;
    const int resident = persist ? it->second : (1 << 30);
    const int th = maps ? maps->th : P2_TH;          // tile height of this call (the state map's box has th rows)
    const int ntiles = ((d.W + P2_TW - 1) / P2_TW) * ((d.H + th - 1) / th);
    const dim3 grid(ntiles < resident ? ntiles : resident);
    static const CUtensorMap no_map{};
    const cudaError_t e = launch_pdl(kern, grid, dim3(P2_NT), smem, st, pdl, d, f, in, out, state, q_out, has_prev, tst ? maps->st : no_map,
                                     tin ? *gmap_in : no_map, maps ? maps->frame : 0, th, ClipArgs{});
    return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : 2;
}

// Clip mode (ClipArgs, crt_fused_ps2.cuh): instantiated for the common variant only — composite-LUT tail, state and input tiles by TMA.
inline bool fused_gauss_ps2_clip_ok(const Dev& d, int K) {
    const bool fast = d.triad_mode == 2 && d.triad_comp && d.vig_mode <= 1;
    return fast && (K == 5 || K == 7 || K == 9) && fused_gauss_ps2_tin_ok(d, K) &&      // (wider kernels: no room for the input tile)
           fused_gauss_ps2_smem(K, true) + 16 * 1024 <= 76 * 1024 && fused_gauss_ps2_smem(K, true, true) + 14 * 1024 <= 75 * 1024 &&
           env_int("CRT_GPS2_TMA_STATE", 1) != 0 && env_int("CRT_GPS2_TMA_INPUT", 1) != 0;
}
template <int K, int MINB>
inline int launch_fused_gauss_ps2_clip_t(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, cudaStream_t st,
                                         const Ps2Maps* maps, const CUtensorMap* gmap_in, ClipArgs& ca) {
    const size_t smem = fused_gauss_ps2_smem(K, true, true);
    auto kern = k_fused_gauss_ps2<K, true, MINB, true, 0, true, true>;
    static const bool use_spec = env_int("CRT_SPEC", 1) != 0;
    if (K == 9 && use_spec && spec_matches(SPEC_GRADED, d, f.flicker_on != 0, true))      // BASELINE configs[1]
        kern = k_fused_gauss_ps2<K == 9 ? 9 : K, true, MINB, true, K == 9 ? SPEC_GRADED : 0, true, true>;
    auto it = env.memo.find((const void*)kern);
    if (it == env.memo.end()) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, P2_NT, smem);
        it = env.memo.emplace((const void*)kern, env.sms * (per_sm > 0 ? per_sm : 1)).first;
    }
    const long long nitems = (long long)((d.W + P2_TW - 1) / P2_TW) * ((d.H + maps->th - 1) / maps->th) * ca.nf;
    ca.frame_bytes = (unsigned long long)d.W * d.H * 3;
    const dim3 grid((unsigned)(nitems < it->second ? nitems : it->second));
    // items (see run_fused_ps2_clip): the atomic counter by default; it beats the fixed stride of a cooperative
    // launch by 12 % for this kernel (run 50: 52.8 k against 46.5 k frames/s on BASELINE configs[1] — its tiles differ in cost and a
    // lagging CTA then holds up the tiles that wait for its flags)
    const int items = env_int("CRT_CLIP_ITEMS", 0);
    cudaError_t e = cudaErrorNotSupported;
    if (items == 2 || items == 3) {      // 3: ranks permuted for a placement that puts consecutive CTAs on one SM
        ca.static_items = 2;
        ca.per_sm = 0;
        const long long ntl = nitems / ca.nf;
        if (items == 3 && ntl >= it->second && it->second % env.sms == 0) ca.per_sm = it->second / env.sms;
        e = launch_pdl(kern, dim3((unsigned)(ntl < it->second ? ntl : it->second)), dim3(P2_NT), smem, st, false, d, f, in, out, state, (float*)nullptr, 1,
                       maps->st, *gmap_in, maps->frame, maps->th, ca);
    } else if (items == 1 && env_int("CRT_CLIP_COOP_GAUSS", 0)) {
        ca.static_items = 1;
        e = launch_coop(kern, grid, dim3(P2_NT), smem, st, d, f, in, out, state, (float*)nullptr, 1, maps->st, *gmap_in, maps->frame, maps->th, ca);
        if (e != cudaSuccess) cudaGetLastError();
    }
    if (e != cudaSuccess && ca.static_items != 2) {
        ca.static_items = 0;
        e = launch_pdl(kern, grid, dim3(P2_NT), smem, st, false, d, f, in, out, state, (float*)nullptr, 1, maps->st, *gmap_in, maps->frame, maps->th, ca);
    }
    return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : 2;
}
inline int run_fused_gauss_ps2_clip(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, cudaStream_t st,
                                    int* launches, const Ps2Maps* maps, const CUtensorMap* gmap_in, ClipArgs& ca) {
    if (!maps || !gmap_in || !fused_gauss_ps2_clip_ok(d, d.ksize) || ca.nf < 1 || ca.nf > CLIP_MAX_FRAMES) return 4;
    int rc = 4;
    switch (d.ksize) {
        case 5: rc = launch_fused_gauss_ps2_clip_t<5, 3>(env, d, f, in, out, state, st, maps, gmap_in, ca); break;
        case 7: rc = launch_fused_gauss_ps2_clip_t<7, 3>(env, d, f, in, out, state, st, maps, gmap_in, ca); break;
        case 9: rc = launch_fused_gauss_ps2_clip_t<9, 3>(env, d, f, in, out, state, st, maps, gmap_in, ca); break;
        default: break;
    }
    if (rc != 4) ++*launches;
    return rc;
}

inline int run_fused_gauss_ps2(LaunchEnv& env, const Dev& d, const FrameDev& f, const uint8_t* in, uint8_t* out, float* state, float* q_out, int has_prev,
                               cudaStream_t st, int* launches, bool pdl = false, const Ps2Maps* maps = nullptr, const CUtensorMap* gmap_in = nullptr) {
    int rc = 4;
    switch (d.ksize) {
        case 5: rc = launch_fused_gauss_ps2_t<5, 3>(env, d, f, in, out, state, q_out, has_prev, st, pdl, maps, gmap_in); break;
        case 7: rc = launch_fused_gauss_ps2_t<7, 3>(env, d, f, in, out, state, q_out, has_prev, st, pdl, maps, gmap_in); break;
        case 9: rc = launch_fused_gauss_ps2_t<9, 3>(env, d, f, in, out, state, q_out, has_prev, st, pdl, maps, gmap_in); break;
        case 11: rc = launch_fused_gauss_ps2_t<11, 3>(env, d, f, in, out, state, q_out, has_prev, st, pdl, maps, gmap_in); break;
        case 13: rc = launch_fused_gauss_ps2_t<13, 3>(env, d, f, in, out, state, q_out, has_prev, st, pdl, maps, gmap_in); break;
        case 25: rc = launch_fused_gauss_ps2_t<25, 3>(env, d, f, in, out, state, q_out, has_prev, st, pdl, maps, gmap_in); break;
        default: break;
    }
    if (rc != 4) ++*launches;        // an unsupported tap count launches nothing
    return rc;
}

#endif  // CRT_TU_GAUSS_PS2

#endif  // __CUDACC__

}  // namespace crt
