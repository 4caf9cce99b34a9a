/*
 * nesr_b200.h -- C ABI of libnesr_b200.so: the B200 (sm_100a) implementation of NESR's
 * Real-ESRGAN x2plus upscaling stage.
 *
 * The reference (gddickinson/neural_enhanced_super_resolution) is pure Python and has no FFI layer;
 * its operator boundary for this path is duck-typing on two third-party classes plus three methods
 * of its own pipeline class.  Every entry point below names the reference interface it stands
 * behind (paths are into the reference tree).  The ctypes binding a maintainer would add is in
 * INTEGRATION.md; the shipped Python mirror is neural_enhanced_super_resolution_b200/_ffi.py.
 *
 * Conventions
 *   - every function returns 0 on success, a negative NESR_E_* code otherwise; the message is
 *     available from nesr_b200_last_error().  No exceptions or abort() cross this boundary.
 *   - the caller allocates all image buffers; the library owns weights, activation arenas, its
 *     CUDA stream and its tensor maps.  A handle is bound to one CUDA device and may be used from
 *     any ONE thread at a time (the reference calls from a QThread: nesr/gui/app.py:72-134).
 *   - NESR_PTR_* flags say whether an image pointer is host or device memory.  Host pointers are
 *     copied with cudaMemcpyAsync on the handle's stream (pinned memory makes that asynchronous).
 *   - there is NO CPU fallback: without a CUDA device nesr_b200_create() fails.
 */
#ifndef NESR_B200_H_
#define NESR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NESR_B200_ABI_VERSION 1

enum {
  NESR_OK = 0,
  NESR_E_INVALID = -1,   /* bad argument / unsupported geometry */
  NESR_E_CUDA = -2,      /* CUDA runtime or driver error */
  NESR_E_WEIGHTS = -3,   /* unknown / mis-shaped / missing tensor */
  NESR_E_STATE = -4,     /* call order (e.g. enhance before finalize) */
  NESR_E_NOMEM = -5
};

enum {                   /* 16-bit operand formats of tcgen05 kind::f16 */
  NESR_FMT_BF16 = 0,
  NESR_FMT_FP16 = 1
};

enum {                   /* pointer-kind flags */
  NESR_PTR_IN_DEVICE = 1,
  NESR_PTR_OUT_DEVICE = 2
};

typedef struct nesr_b200_handle nesr_b200_handle;

/* Architecture + precision.  Mirrors the constructor the reference calls:
 *   RRDBNet(num_in_ch, num_out_ch, scale, num_feat, num_block, num_grow_ch)
 *   -- nesr/nesr.py:216, standalone/direct_esrgan.py:104, standalone/superres_project.py:69.
 * x2plus is {3, 3, 2, 64, 23, 32}.  This build supports scale == 2, num_feat == 64,
 * num_grow_ch == 32, num_in_ch == num_out_ch == 3 and any num_block >= 1. */
typedef struct nesr_b200_config {
  int32_t abi_version;        /* NESR_B200_ABI_VERSION */
  int32_t device;             /* CUDA ordinal */
  int32_t num_in_ch;
  int32_t num_out_ch;
  int32_t scale;
  int32_t num_feat;
  int32_t num_block;
  int32_t num_grow_ch;
  int32_t body_format;        /* residual-dense-block convs: NESR_FMT_BF16 (default) | NESR_FMT_FP16 */
  int32_t edge_format;        /* conv_first/body/up1/up2/hr/last:  NESR_FMT_FP16 (default) | NESR_FMT_BF16 */
  int32_t conv_impl;          /* 0 = row-folded tcgen05/TMEM/TMA kernels (product): tiles are processed in
                                 L2-resident groups, each with ONE persistent launch for the 69 residual dense
                                 blocks (conv3x3_trunk.cu) and one launch per edge layer.  Test-only
                                 cross-checks, never selected implicitly: 1 = SIMT validation kernel (CUDA
                                 cores, no TMA / tcgen05), 3 = row-folded kernel with one launch per layer
                                 pass (stream order is the only synchronisation), 4 = whole-frame persistent
                                 trunk kernel with a grid-wide arrival counter (conv3x3_body.cu; also what a
                                 tile too large for the trunk kernel's TMEM row budget runs on).  2 is
                                 rejected (the first-generation per-tap kernel was removed in round 2) */
  int32_t feat_in_ch;         /* input channels of conv_first ON THE FEATURE GRID, 1..64; 0 = num_in_ch * 4, the x2plus un-shuffle
                                 (12).  Other values serve nesr_b200_forward_feat_f32 only: 3 = the scale-4 architecture
                                 (x4plus: no un-shuffle), 48 = the scale-1 architecture (un-shuffle by 4) */
  int64_t max_batch_pixels;   /* cap on feature-grid pixels per tile group (batch); 0 = default (200k for
                                 conv_impl 0: the group's dense-block activations stay in the 126 MB L2) */
} nesr_b200_config;

typedef struct nesr_b200_stats {
  int64_t kernel_launches;    /* kernels this handle launched since creation */
  int64_t conv_launches;      /* ... of which tcgen05 conv kernels */
  int64_t tiles_processed;
  double  last_device_ms;     /* CUDA-event time of the last enhance/forward call (device part only) */
  double  last_conv_ms;       /* ... conv kernels only (first to last conv of the call) */
  int64_t arena_bytes;        /* activation arena currently allocated */
  double  last_trunk_ms;      /* ... persistent trunk kernel(s) of the last call only (the dominant kernel; sum over tile groups) */
  int64_t last_trunk_launches;/* trunk kernel launches of the last call (= tile groups) */
} nesr_b200_stats;

/* Fills *cfg with the x2plus defaults for `device`. */
void nesr_b200_default_config(nesr_b200_config* cfg, int32_t device);

/* Replaces: RRDBNet(...) construction + RealESRGANer.__init__ device placement
 * (realesrgan/utils.py ctor; nesr/nesr.py:216-229). */
int nesr_b200_create(const nesr_b200_config* cfg, nesr_b200_handle** out);
int nesr_b200_destroy(nesr_b200_handle* h);

/* Message of the most recent failure on this handle (h may be NULL: failure of create). */
const char* nesr_b200_last_error(const nesr_b200_handle* h);

/* Replaces: model.load_state_dict(loadnet['params_ema'|'params'], strict=True)
 * (RealESRGANer.__init__; call site nesr/nesr.py:220-229).  `name` is the checkpoint key, e.g.
 * "body.7.rdb2.conv3.weight"; data is host fp32, conv weights OIHW.  Unknown names and shape
 * mismatches fail (strict).  finalize fails unless every tensor was supplied; it repacks the
 * weights into the 16-bit K-major tap/chunk layout the kernels read and uploads them. */
int nesr_b200_load_weight(nesr_b200_handle* h, const char* name, const float* data,
                          const int64_t* shape, int32_t ndim);
int nesr_b200_finalize_weights(nesr_b200_handle* h);

/* Replaces: RealESRGANer.enhance(img, outscale=None) for 8-bit 3-channel input
 * (standalone/superres_project.py:282, standalone/direct_esrgan.py:148, pre-HEAD nesr/nesr.py):
 * BGR HWC u8 -> /255 -> RGB -> reflect pre_pad / mod-pad -> pixel_unshuffle(2) -> RRDBNet ->
 * tile_process paste (tile > 0) -> post-crop -> clamp -> BGR -> round-half-even u8.
 * in: H x W x 3 (in_stride bytes per row); out: 2H x 2W x 3.  tile == 0 processes the image whole.
 * For scale 2 `tile` and `tile_pad` must be even (upstream's un-shuffle asserts the same). */
int nesr_b200_enhance_u8(nesr_b200_handle* h, const uint8_t* in_bgr, int32_t H, int32_t W,
                         int64_t in_stride, int32_t tile, int32_t tile_pad, int32_t pre_pad,
                         uint8_t* out_bgr, int64_t out_stride, int32_t flags);

/* Throughput mode (BASELINE config 4): n equally sized frames, each exactly as enhance_u8.
 * Frame f starts at in + f*in_frame_stride / out + f*out_frame_stride. */
int nesr_b200_enhance_batch_u8(nesr_b200_handle* h, const uint8_t* in_bgr, int32_t n_frames,
                               int32_t H, int32_t W, int64_t in_stride, int64_t in_frame_stride,
                               int32_t tile, int32_t tile_pad, int32_t pre_pad, uint8_t* out_bgr,
                               int64_t out_stride, int64_t out_frame_stride, int32_t flags);

/* Tile-sharded form of enhance_u8 (multi-GPU: BASELINE config 3).  Processes only tiles
 * [tile_first, tile_first + tile_count) of upstream's row-major tile_process grid and pastes them
 * into `out_bgr`, which addresses the FULL 2H x 2W output (pixels of other tiles are untouched).
 * nesr_b200_tile_count reports the size of the grid. */
int nesr_b200_tile_count(int32_t H, int32_t W, int32_t tile, int32_t pre_pad, int32_t scale);

/* (test hook, host only -- no CUDA call, works without a GPU) Builds the tile-group plan that an enhance call with these
 * arguments would use on a device with `num_sms` SMs and verifies its invariants: every pixel of every tile of every
 * resolution level owned by exactly one (CTA, band, lane); segment lanes inside the 128-lane MMA / 136-row slab; TMEM row
 * limit of the trunk kernel; and, pixel by pixel against an independently built owner map, that the trunk kernel's
 * row-dependency table makes every input slab row of every CTA wait for the rows of every CTA owning one of its pixels.
 * out[8]: groups, tiles, feature pixels, level-0 strip rows, max output rows per CTA, groups run by the TMEM-resident trunk
 * kernel, reserved (0), level-0 halo rows.  pairs / sets: reserved (planner variants removed in round 2), ignored. */
int nesr_b200_debug_plan(int32_t n_frames, int32_t H, int32_t W, int32_t tile, int32_t tile_pad, int32_t pre_pad, int32_t num_sms,
                         int32_t conv_impl, int64_t max_batch_pixels, int32_t pairs, int32_t sets, int64_t* out);
int nesr_b200_enhance_tiles_u8(nesr_b200_handle* h, const uint8_t* in_bgr, int32_t H, int32_t W,
                               int64_t in_stride, int32_t tile, int32_t tile_pad, int32_t pre_pad,
                               int32_t tile_first, int32_t tile_count, uint8_t* out_bgr,
                               int64_t out_stride, int32_t flags);

/* Tile-major form of enhance_tiles_u8 -- the multi-GPU exchange format ("each GPU stitches its tiles into its
 * slice of the output ... one gather of stitched u8 output rows").  Tile k of the range [tile_first, tile_first + tile_count)
 * is written into slot k of `slots` (DEVICE memory, slot_h rows of slot_w BGR pixels each, slots contiguous), its cropped
 * output rectangle at the slot's origin: a rank's whole contribution is ONE contiguous buffer, the operand of a single
 * ncclAllGather.  slot_w / slot_h >= tile * scale (the largest rectangle a tile pastes; reference tile_process:
 * output_tile = out[.., (in_start - in_start_pad) * s : .. + (in_end - in_start) * s]). */
int nesr_b200_enhance_tiles_packed_u8(nesr_b200_handle* h, const uint8_t* in_bgr, int32_t H, int32_t W, int64_t in_stride,
                                      int32_t tile, int32_t tile_pad, int32_t pre_pad, int32_t tile_first, int32_t tile_count,
                                      uint8_t* slots, int32_t slot_w, int32_t slot_h, int32_t flags);
/* The same for an arbitrary subset of the tile grid (slot k holds tile tile_ids[k]): tiles are independent forwards, so the ranks'
 * shares need not be contiguous ranges -- a longest-first deal of the tiles balances padded-pixel cost to a few percent where the best
 * contiguous cut of a 4K frame's 40 tiles over 8 ranks is 12 % off. */
int nesr_b200_enhance_tile_list_packed_u8(nesr_b200_handle* h, const uint8_t* in_bgr, int32_t H, int32_t W, int64_t in_stride,
                                          int32_t tile, int32_t tile_pad, int32_t pre_pad, const int32_t* tile_ids, int32_t tile_count,
                                          uint8_t* slots, int32_t slot_w, int32_t slot_h, int32_t flags);
/* The reference HEAD's ESRGAN stage in one call (`_apply_esrgan_12channel` / `_apply_esrgan_3channel`, nesr/nesr.py:845-986), for a handle
 * created with the x2plus weights: RGB H x W x 3 u8 -> the 12-channel tensor the reference builds (BGR / 255; x1.1 and x0.9 clamped;
 * cv2.GaussianBlur 3x3 -- or four copies with force_3channel), computed inside the input pack kernel -> model(x12) (the 12-channel scale-4
 * architecture is the x2plus network behind its un-shuffle) -> clip(out * 255, 0, 255) TRUNCATED to u8 -> RGB, 4H x 4W x 3. */
int nesr_b200_enhance_head_u8(nesr_b200_handle* h, const uint8_t* in_rgb, int32_t H, int32_t W, int64_t in_stride, int32_t force_3channel,
                              uint8_t* out_rgb, int64_t out_stride, int32_t flags);
/* Inverse placement: slots (device) of tiles [tile_first, tile_first + tile_count) -> their rectangles in the full 2H x 2W
 * frame `out_bgr` (device).  Called once per rank's slice of the gathered buffer; other pixels of the frame are untouched. */
int nesr_b200_unpack_tiles_u8(nesr_b200_handle* h, const uint8_t* slots, int32_t slot_w, int32_t slot_h, int32_t H, int32_t W,
                              int32_t tile, int32_t pre_pad, int32_t tile_first, int32_t tile_count, uint8_t* out_bgr,
                              int64_t out_stride);
/* The same for slots holding an arbitrary tile list: slot k holds tile tile_ids[k]; an id < 0 marks an empty (padding) slot.  One call
 * pastes the whole gathered buffer of every rank. */
int nesr_b200_unpack_tile_list_u8(nesr_b200_handle* h, const uint8_t* slots, int32_t slot_w, int32_t slot_h, int32_t H, int32_t W,
                                  int32_t tile, int32_t pre_pad, const int32_t* tile_ids, int32_t slot_count, uint8_t* out_bgr,
                                  int64_t out_stride);

/* Replaces: RRDBNet.forward / `upsampler.model(x)` (nesr/nesr.py:887-891,930-935): device fp32
 * NCHW [n, num_in_ch, H, W] in [0,1] -> device fp32 NCHW [n, num_out_ch, 2H, 2W], unclamped.
 * Both pointers are device memory; the work is enqueued on `stream` (a cudaStream_t, taken literally:
 * NULL is the legacy default stream -- pass torch.cuda.current_stream().cuda_stream) and the call
 * returns without synchronising it. */
int nesr_b200_forward_nchw_f32(nesr_b200_handle* h, const float* x, int32_t n, int32_t H,
                               int32_t W, float* y, void* stream);

/* Replaces: `self.models['esrgan'].model(img_12ch)` of the reference HEAD (nesr/nesr.py:887-891, 930-935), whose model is
 * RRDBNet(num_in_ch=12, num_out_ch=3) with the default scale=4 (nesr/nesr.py:216): the x2plus weights applied to a 12-channel
 * tensor at full resolution, no un-shuffle, x4 out.  Device fp32 NCHW [n, 12, H, W] -> device fp32 NCHW [n, num_out_ch, 4H, 4W],
 * unclamped; stream semantics as nesr_b200_forward_nchw_f32. */
int nesr_b200_forward_nchw12_f32(nesr_b200_handle* h, const float* x12, int32_t n, int32_t H,
                                 int32_t W, float* y, void* stream);

/* Replaces: RRDBNet.forward for the architectures whose first op is NOT the x2 un-shuffle (upstream rrdbnet_arch.py: scale 4 feeds x
 * itself to conv_first, scale 1 feeds pixel_unshuffle(x, 4); SURVEY.md 8f row f4 "scale 1/4 nets"): `feat` is that tensor --
 * DEVICE float32 [n, C, h, w] with C == the handle's feat_in_ch (12 when 0) -- and y is [n, num_out_ch, 4h, 4w]: conv_first .. conv_last
 * as for the x2plus network, whose nchw12 entry is the C == 12 case. */
int nesr_b200_forward_feat_f32(nesr_b200_handle* h, const float* feat, int32_t n, int32_t C, int32_t H, int32_t W, float* y, void* stream);

/* Replaces: SuperResolutionPipeline._ensemble_results (nesr/nesr.py:1033-1054) for K >= 2
 * equally sized H x W x 3 u8 members: acc(f32) += f64(img) * w[i] rounded to f32 per member,
 * truncated to u8.  weights == NULL means the reference's uniform 1/K. */
int nesr_b200_blend_u8(nesr_b200_handle* h, const uint8_t* const* members, int32_t K, int32_t H,
                       int32_t W, const double* weights, uint8_t* out, int32_t flags);

/* Replaces: SuperResolutionPipeline._postprocess_image with adaptive_sharpening on
 * (nesr/nesr.py:1056-1084): RGB (or BGR with bgr != 0) H x W x 3 u8 -> same layout, bit-exact
 * with cv2 4.13's fixed-point path. */
int nesr_b200_sharpen_u8(nesr_b200_handle* h, const uint8_t* in, int32_t H, int32_t W,
                         int32_t bgr, uint8_t* out, int32_t flags);

/* Replaces: the unsharp stage of SuperResolutionPipeline._segment_and_enhance (nesr/nesr.py:728-747; SURVEY.md 8f row f4,
 * "segmentation-masked unsharp reusing K6"): object_mask = cv2.dilate(object_mask, ones((3, 3))); blurred =
 * cv2.GaussianBlur(img, (0, 0), 3); sharpened = cv2.addWeighted(img, 1.5, blurred, -0.5, 0); out = where(object_mask == 1,
 * sharpened, img).  `object_mask` is the H x W u8 mask at image resolution BEFORE the dilation (the reference's
 * cv2.resize(object_mask, (W, H)), nesr/nesr.py:730), on the same side (host / device) as `in`; the dilation is done here.
 * Same kernel and fixed-point arithmetic as nesr_b200_sharpen_u8, bit-exact with cv2 4.13.  The segmentation model itself
 * (SegFormer via transformers) is out of scope. */
int nesr_b200_masked_unsharp_u8(nesr_b200_handle* h, const uint8_t* in, const uint8_t* object_mask, int32_t H, int32_t W,
                                int32_t bgr, uint8_t* out, int32_t flags);

/* Replaces: SuperResolutionPipeline._preprocess_image (nesr/nesr.py:668-689): RGB H x W x 3 u8 ->
 * same layout.  denoise_h > 0: cv2.fastNlMeansDenoisingColored(img, None, denoise_h, denoise_h_color, 7, 21) (the reference
 * passes 10 * config['denoise_level'] for both; denoise_h_color <= 0 means "same as denoise_h"); denoise_h <= 0 skips the
 * denoiser as the reference does at level 0.  Then CLAHE(clahe_clip, (tiles_x, tiles_y)) on L of cvtColor(RGB2LAB) and
 * cvtColor(LAB2RGB) -- the reference uses 2.0 and (8, 8).  Bit-exact with cv2 4.13 (integer Lab paths, integer NLM weights,
 * float32 CLAHE interpolation without contraction). */
int nesr_b200_preprocess_u8(nesr_b200_handle* h, const uint8_t* rgb, int32_t H, int32_t W, float denoise_h,
                            float denoise_h_color, float clahe_clip, int32_t tiles_x, int32_t tiles_y,
                            uint8_t* out, int32_t flags);

/* Test hooks (host only, no GPU): the committed 8-bit Lab tables (which: 0 sRGB gamma u16[256], 1 cube root u16[3072],
 * 2 L -> (y, fy) u16[512], 3 inverse sRGB gamma u8[4096]) and the NLM weight table for (h, channels); both return the size
 * (bytes resp. entries; out may be NULL to query) or -1. */
int nesr_b200_debug_lab_table(int32_t which, void* out, int32_t capacity_bytes);
int nesr_b200_debug_nlm_weights(float h, int32_t channels, int32_t* out, int32_t capacity);

int nesr_b200_get_stats(const nesr_b200_handle* h, nesr_b200_stats* out);
int nesr_b200_synchronize(nesr_b200_handle* h);

/* Test hook: run ONE 3x3 conv layer of the given geometry through the selected implementation
 * on caller-provided device buffers (used by tests/ to compare the tcgen05 kernel with the SIMT
 * validation kernel tap by tap).  See csrc/engine.cu. */
int nesr_b200_debug_conv(nesr_b200_handle* h, int32_t impl, int32_t fmt, int32_t H, int32_t W,
                         int32_t cin, int32_t cout, const float* weight_oihw, const float* bias,
                         const float* x_nchw_host, int32_t lrelu, float* y_nchw_host);

#ifdef __cplusplus
}
#endif
#endif /* NESR_B200_H_ */
