/*
 * crt_b200.h — C ABI of the B200-native per-frame CRT effect chain.
 *
 * Drop-in boundary for the hot path of jaylikesbunda/PythonCRT.  The reference
 * has no FFI: its seam is two Python functions and one inline block
 * (SURVEY.md §8b).  Each entry point below names what it replaces
 * (file:line into /root/reference/crt_filter.py):
 *
 *   apply_crt_effect(frame, ...) -> (uint8 out, float state)      :531-699
 *   apply_static_effects(frame, ...) -> float image                :702-861
 *   persistence blend + convertScaleAbs in process_video           :1086-1098
 *
 * Conventions: extern "C", plain-old-data structs, no C++/torch types.  All
 * `d_*` pointers are DEVICE pointers owned by the caller (e.g.
 * torch.Tensor.data_ptr()); `h_*` pointers are HOST pointers.  Calls are
 * stream-ordered and asynchronous unless stated.  Every function returns 0 on
 * success and a non-zero crt_status on failure; crt_last_error() gives the
 * text.  One context per (device, host thread): the reference's export path
 * calls the chain from two worker threads (:1015-1017), each would own a ctx.
 * There is no CPU fallback: without a CUDA device crt_create fails.
 */
#ifndef CRT_B200_H
#define CRT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRT_B200_ABI_VERSION 2

typedef struct crt_ctx crt_ctx;

typedef enum crt_status {
    CRT_OK = 0,
    CRT_ERR_INVALID = 1,      /* bad argument / missing table            */
    CRT_ERR_CUDA = 2,         /* CUDA runtime error (see crt_last_error) */
    CRT_ERR_NO_DEVICE = 3,    /* no CUDA device: there is no CPU path    */
    CRT_ERR_UNSUPPORTED = 4
} crt_status;

/* Which of the reference's two chains is being replaced: they differ only in
 * the glitch offset generator (:664-679 vs :835-853). */
typedef enum crt_variant { CRT_VARIANT_GUI = 0, CRT_VARIANT_EXPORT = 1 } crt_variant;
typedef enum crt_channel_order { CRT_ORDER_RGB = 0, CRT_ORDER_BGR = 1 } crt_channel_order;

/*
 * Scalar effect parameters = the scalar arguments of apply_crt_effect
 * (:531-565), in double like the Python floats they replace; the library
 * narrows to float32 exactly where numpy does.  Masks the reference receives
 * as arrays (triad_mask, vignette_mask) are given as tables (crt_set_table) or
 * as the analytic strength below.  Flags `*_on` replace `mask is None` tests.
 */
typedef struct crt_params {
    /* colour grading, apply_color_adjustments :279-305 */
    double brightness, contrast, gamma, saturation, temperature;
    /* chromatic aberration :571-577 and pixelate :578-584 */
    int32_t aberration_px;
    int32_t pixel_size;
    /* bloom :599-612 */
    double bloom_sigma, bloom_strength, bloom_threshold;
    int32_t fast_bloom;
    /* triad :613-616, _apply_triad_mask :238-263 (column table: CRT_TABLE_TRIAD_COLS) */
    int32_t triad_on;
    double triad_gamma;
    int32_t triad_preserve_luma;
    /* scanlines :617-625, make_scanline_mask_dynamic :213-217, make_scanline_mask_2d :308-328 */
    double scanline_strength, scanline_period_px, scanline_angle, scanline_thickness;
    /* vignette :626-629, make_vignette :266-276.  vignette_on: 0 off, 1 analytic
     * (strength below), 2 arbitrary H*W float32 plane (CRT_TABLE_VIGNETTE_PLANE) */
    int32_t vignette_on;
    double vignette_strength;
    /* flicker :630-634 */
    double flicker_strength, flicker_hz;
    /* noise / grain :635-648.  noise_mode: 0 = draws injected per frame
     * (crt_frame.d_noise), 1 = counter-based generator keyed (seed, frame_index, cell) */
    double noise_strength;
    int32_t grain_size;
    int32_t noise_mode;
    uint64_t noise_seed;
    /* barrel warp :649-652, apply_barrel_warp :331-348 */
    double warp_strength;
    /* glitch :664-686 / :835-859.  glitch_mode: 0 = offsets injected per frame
     * (crt_frame.d_glitch_offs, made on the host from numpy's PCG64 exactly as the
     * reference does), 1 = generated on the device from the counter-based RNG, keyed on the reference's own seed
     * (int(|phase_px| * k) + (W << 10) + (H << 1), k = 0.05 gui :670 / 2.0 export :841) so that a pattern is held
     * for exactly as many frames as the reference holds it */
    int32_t glitch_amp_px;
    double glitch_height_frac;
    int32_t glitch_mode;
    /* text overlay layer (CRT_TABLE_TEXT_RGBA): 0 none, 1 before effects :588-598, 2 after :653-663 */
    int32_t text_mode;
    /* persistence :687-694 / :1086-1096 */
    double persistence;
    int32_t variant;          /* crt_variant */
    /* Channel order of the frames (and of the float state / image): the reference's arithmetic is defined by channel
     * INDEX with index 0 = R (luma weights :289, temperature gains :296-297, aberration :573-575, triad phase :224).
     * CRT_ORDER_RGB feeds index 0 to those rules like the reference; CRT_ORDER_BGR (e.g. OpenCV capture buffers before
     * the reference's COLOR_BGR2RGB, :1336) applies the R rules to index 2 and the B rules to index 0, so the result
     * equals the reference's on the channel-swapped frame, swapped back, bit for bit.  Host tables (triad columns, text
     * layer) are always given in the reference's RGB order. */
    int32_t channel_order;    /* crt_channel_order */
    int32_t reserved[6];
} crt_params;

/* Host-built tables.  The host side (pythoncrt_b200/tables.py) builds them with
 * the same numpy expressions the reference uses, so they are bit-identical to
 * what the reference indexes. */
typedef enum crt_table {
    CRT_TABLE_TRIAD_COLS = 0,     /* float32 [W][3]: row 0 of make_triad_mask :220-235      */
    CRT_TABLE_LUT_FWD = 1,        /* float32 [1025]: np.power(linspace, g)       :248-249   */
    CRT_TABLE_LUT_INV = 2,        /* float32 [1025]: np.power(linspace, 1/g)     :260       */
    CRT_TABLE_GAUSS_TAPS = 3,     /* float32 [k]: cv2.getGaussianKernel(k, sigma) :609-610  */
    CRT_TABLE_PIXELATE_X = 4,     /* int32 [W]: composite NEAREST index          :580-583   */
    CRT_TABLE_PIXELATE_Y = 5,     /* int32 [H]                                              */
    CRT_TABLE_VIGNETTE_PLANE = 6, /* float32 [H][W]: arbitrary caller-supplied vignette     */
    CRT_TABLE_TEXT_RGBA = 7,      /* uint8 [H][W][4]: rasterised text layer      :590-596   */
    CRT_TABLE_COUNT = 8
} crt_table;

/* Per-frame scalars: process_video derives them from the frame index
 * (phase :1043, time_sec :1064); the GUI passes speed*t and t (:1810-1852). */
typedef struct crt_frame {
    double phase_px;                /* scanline_phase_px                                   */
    double time_sec;                /* flicker time                                        */
    uint64_t frame_index;           /* global frame index: RNG counter for generated draws */
    const float* d_noise;           /* injected N(0,1) plane [gh][gw] (before grain up-scale), or NULL */
    const int32_t* d_glitch_offs;   /* injected offsets [rows][segments], or NULL          */
    int32_t glitch_y0, glitch_seg_len, glitch_segments, glitch_rows;
} crt_frame;

typedef struct crt_launch_info {
    int32_t kernels_launched;       /* kernels of this library launched by the call */
    int32_t fused;                  /* 1 = single fused tile kernel, 2 = two-pass (fused first pass + gather), 0 = staged kernels */
    int32_t reserved[6];            /* [0] = temporal shards the call ran concurrently (crt_set_shards), [1] = warm-up frames per shard,
                                     * [2] = frames that went through clip-mode launches (one launch per run of up to 64 frames) */
} crt_launch_info;

int crt_abi_version(void);

/* Create / destroy a context for frames of width x height on CUDA device `device`. */
int crt_create(int device, int width, int height, crt_ctx** out_ctx);
int crt_destroy(crt_ctx* ctx);
const char* crt_last_error(const crt_ctx* ctx);   /* ctx may be NULL: last creation error */

/* Scalar parameters and host-built tables (copied; synchronous, rarely called). */
int crt_set_params(crt_ctx* ctx, const crt_params* params);
int crt_set_table(crt_ctx* ctx, int table /* crt_table */, const void* h_data, size_t bytes);

/* Execution policy: 0 = auto (fused tile kernel when the parameters allow, else
 * staged kernels), 1 = force staged, 2 = force fused (error if not possible). */
int crt_set_policy(crt_ctx* ctx, int policy);

/*
 * Intra-GPU temporal shards for crt_process: a long clip is cut into up to `shards` contiguous pieces that run concurrently
 * on one GPU (own stream and context each), every piece after the first preceded by a persistence warm-up of
 * ceil(ln(1/2040) / ln(persistence)) frames whose outputs are discarded — the scheme process_video's frame order admits
 * (the only cross-frame dependency is the blend at :1092) and the one used across GPUs (pythoncrt_b200/clip.py).  The result
 * differs from the strictly serial run by at most persistence^warm-up <= 1/8 LSB before quantisation (identical when
 * persistence == 0).  1 = off (default: every call is strictly serial), 0 = automatic (up to 4 shards, each at least 48 frames
 * and 8 warm-ups long; none for frames of 24 Mpixel and more, which fill the GPU on their own), k = at most k.  Calls shorter than that, and crt_process_static / crt_process_host, stay serial.
 *
 * Clip mode.  A strictly serial call of two or more frames whose parameters select a single-pass block kernel (pixel_size 2,
 * fast / no / gaussian bloom up to 9 taps, persistence on, no noise plane or glitch table per frame, a frame of at least one
 * tile per resident CTA: 1080p and up) is launched as runs of up to 64 frames per kernel launch, the frames of a tile chained
 * on the device through per-tile flags: the same bytes and the same state as one launch per frame (the recurrence at :1092
 * exactly), without the per-frame launch and tail.  In automatic mode (0) the library takes clip mode instead of shards where it
 * measures faster.  crt_launch_info.reserved[2] reports the frames that went that way.  The caller sees no difference
 * other than that: the call is asynchronous on `stream` as before.
 */
int crt_set_shards(crt_ctx* ctx, int shards);

/*
 * Run n_frames consecutive frames through the chain, persistence included.
 *   d_in   uint8 [n_frames][H][W][3], channel index 0 treated as R like the reference
 *   d_out  uint8 [n_frames][H][W][3]
 *   d_state float32 [H][W][3]: persistence state (the blended float image the
 *          reference returns / keeps, :699 / :1096); read if *state_valid != 0,
 *          always written with the state after the last frame.  May be NULL
 *          when persistence == 0.
 *   state_valid  0 = "state_prev is None" (first frame, no blend, :687 / :1086)
 *   frames host array of n_frames per-frame records
 *   stream cudaStream_t (NULL = legacy default stream)
 * When W % 4 == 0, d_state must be 16-byte and d_out 4-byte aligned (CRT_ERR_INVALID
 * otherwise; cudaMalloc / torch allocations are); with d_in 16-byte aligned as well the
 * block kernels move their tiles with TMA.
 * Replaces n_frames calls of apply_crt_effect (:531-699), or of
 * apply_static_effects (:702-861) + the blend/quantise block (:1086-1098).
 */
int crt_process(crt_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, float* d_state, int state_valid,
                const crt_frame* frames, int n_frames, void* stream, crt_launch_info* info /* may be NULL */);

/*
 * Same, but the float image BEFORE persistence/quantise is written instead
 * (d_img float32 [n_frames][H][W][3]) — the return value of apply_static_effects (:861).
 */
int crt_process_static(crt_ctx* ctx, const uint8_t* d_in, float* d_img, const crt_frame* frames, int n_frames,
                       void* stream, crt_launch_info* info);

/*
 * Host-buffer entry point: frames come from and go back to HOST memory (pinned
 * or pageable), staged through internal pinned/device rings with copies
 * overlapped with compute; the persistence state stays on the device inside
 * the context (crt_reset_state = "state_prev = None", :1765).  Synchronous.
 * This is what a reference-side binding calls once per frame or per batch.
 */
int crt_process_host(crt_ctx* ctx, const uint8_t* h_in, uint8_t* h_out, const crt_frame* frames, int n_frames,
                     crt_launch_info* info);
int crt_reset_state(crt_ctx* ctx);

/*
 * Persistence state kept from frames of another size (the GUI's preview window was resized): apply_crt_effect resizes it
 * with cv2.resize(prev, (w, h), INTER_LINEAR) before the blend (:689-690).  d_src float32 [src_height][src_width][3] ->
 * d_dst float32 [H][W][3] of this context, OpenCV's bilinear arithmetic (half-pixel centres, edge clamp, rows first).
 * Synchronous on `stream`.
 */
int crt_resize_state(crt_ctx* ctx, const float* d_src, int src_width, int src_height, float* d_dst, void* stream);

/* Timing hooks for bench.py: while enabled, CUDA events are recorded (on the
 * launch stream) around ALL kernels of a frame (generators, first pass, gather),
 * up to max_samples frames.  crt_profile_end synchronises them and returns the
 * summed duration and the number of frames timed. */
int crt_profile_begin(crt_ctx* ctx, int max_samples);
int crt_profile_end(crt_ctx* ctx, double* total_ms, int* samples);
/* Time only one launch in `every` (default 1 = all).  An event between two frames' kernels
 * serialises them; with a stride the other launches keep their overlap (programmatic dependent
 * launch), so the timed region's throughput is not the profiler's. */
int crt_profile_sample_every(crt_ctx* ctx, int every);

/* Device-side generators (counter-based RNG), used when noise_mode/glitch_mode == 1;
 * exposed so a host can pre-generate and inspect the draws. */
int crt_generate_noise(crt_ctx* ctx, uint64_t frame_index, float* d_plane /* [gh][gw] */, void* stream);
int crt_generate_glitch(crt_ctx* ctx, const crt_frame* frame, int32_t* d_offs /* [rows][segments] */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRT_B200_H */
