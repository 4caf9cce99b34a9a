// kernels.h -- host-callable launchers of every CUDA kernel in libnesr_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <vector>
#include <stdint.h>

#include "layout.h"

namespace nesr {

// --- conv3x3_fold.cu : row-folded tcgen05 conv, N = 3*Cout, one layer pass per launch ------------
cudaError_t conv3x3_fold_configure();
bool conv3x3_fold_fits(int cin16, int npad);
cudaError_t launch_conv3x3_fold(const CUtensorMap& amap136, const CUtensorMap& amap8, const CUtensorMap& wmap,
                                const ConvParams& p, int grid, cudaStream_t stream);

// --- conv3x3_body.cu : all RDB layer passes of a batch in one persistent cooperative launch ----------
cudaError_t conv3x3_body_configure();
cudaError_t launch_conv3x3_body(const CUtensorMap& d0_136, const CUtensorMap& d1_136, const CUtensorMap& d0_8,
                                const CUtensorMap& d1_8, const CUtensorMap& wmap96, const ConvParams* d_passes, int npass,
                                unsigned* d_gbar, int grid, cudaStream_t stream);

// --- conv3x3_trunk.cu : all RDB layer passes of an L2-resident tile group; TMEM-resident bands, chunk-major sweeps,
//     streamed weights, pass transitions overlapped (needs <= kTrunkMaxRows rows and <= kTrunkMaxBands bands per CTA)
// Tensor maps of the trunk kernel: the two dense-block buffers with a 136-pixel box (full strips) and with boxes of
// 8 / 16 / 32 / 64 pixels (segments of packed remainder strips are loaded with as few TMA operations as possible --
// with 8-pixel boxes only, the CTAs that own packed strips were producer-bound and set the pace of all their neighbours).
struct TrunkMaps {
  CUtensorMap full[2];
  CUtensorMap box[2][4];
  CUtensorMap half[2], hbox[2][4];   // 32-channel (64-byte rows, SWIZZLE_64B) boxes of the same buffers: the half-width slabs of sweeps S2 / S5
  CUtensorMap w192, w160;      // merged trunk weights, boxes of 192 / 160 rows
};
cudaError_t conv3x3_trunk_configure();
cudaError_t launch_conv3x3_trunk(const TrunkMaps& maps, const ConvParams* d_passes, int npass, const TrunkSweep* d_sweeps, int nsweep,
                                 unsigned* d_prog, int grid, cudaStream_t stream);

// --- conv3x3_simt.cu : plain CUDA-core conv over the same buffers (test-only cross-check) -------
cudaError_t launch_conv3x3_simt(const ConvParams& p, cudaStream_t stream);

// --- pixel_io.cu : image -> level-0 input features (BGR->RGB, /255, reflect pad, pixel_unshuffle) -
struct PackParams {
  const BlockRef* blocks;
  const TileGeom* tiles;
  int32_t nblk;
  const uint8_t* in_u8;        // BGR HWC frames, or null
  int64_t in_stride, in_frame_stride;
  const float* in_f32;         // RGB NCHW frames, or null
  const float* in_f32_12;      // frames already on the feature grid, [frame][feat_ch][H/2][W/2] (feat_ch 12: the reference HEAD's 12-channel tensor), or null
  int32_t feat_ch;             // in_f32_12: channels per frame (12: un-shuffled x2plus input; 3: scale-4 nets; 48: scale-1 nets)
  const uint8_t* in_u8_head;   // RGB HWC u8 image of H/2 x W/2 from which the reference HEAD builds its 12 channels (nesr/nesr.py:859-880), or null
  int32_t head_replicate;      // in_u8_head: force_3channel -- four copies of the image instead of image, x1.1, x0.9, 3x3 blur
  int32_t H, W;                // un-padded frame size
  int32_t pre_pad;             // reflect pad (right/bottom) applied before the mod pad
  void* x0;                    // [pixels][64] 16-bit, channels 0..11 written
  int32_t fmt;
};
cudaError_t launch_pack(const PackParams& p, cudaStream_t stream);
// tile-major slots (tile k of [first, first+count) in slot k, its output rectangle at the slot's origin) -> their places in the frame
constexpr int kUnpackMaxTiles = 64;     // tile ids per launch (kernel parameter space); longer lists are pasted in several launches
cudaError_t launch_unpack_tiles(const uint8_t* slots, int slot_w, int slot_h, int tiles_x, int tile_out, int out_h, int out_w, int first,
                                int count, const int32_t* tile_ids, uint8_t* out, int64_t out_stride, cudaStream_t stream);

// --- stencil.cu : bit-exact integer post-process kernels ----------------------------------------
constexpr int kMaxBlendMembers = 16;
struct BlendParams {
  const uint8_t* members[kMaxBlendMembers];
  double weights[kMaxBlendMembers];
  int32_t k;
  int64_t nbytes;
  uint8_t* out;
};
cudaError_t launch_blend(const BlendParams& p, cudaStream_t stream);
// ext_mask == null: the adaptive detail mask of _postprocess_image; else the segmentation-masked unsharp (H x W u8 object mask, dilated 3 x 3 here)
cudaError_t launch_sharpen(const uint8_t* in, uint8_t* out, int32_t H, int32_t W, int32_t bgr, const uint8_t* ext_mask, cudaStream_t stream);
// sharpen_mma.cu: the same stage with both Gaussian passes as banded-Toeplitz products on the tensor pipe (IMMA.16832.U8.U8) -- the
// product path; launch_sharpen dispatches to it unless NESR_B200_SHARPEN_IMPL=1 selects the dp4a kernel (cross-check, bit-identical)
cudaError_t launch_sharpen_mma(const uint8_t* in, uint8_t* out, int32_t H, int32_t W, int32_t bgr, const uint8_t* ext_mask, cudaStream_t stream);

// --- preprocess.cu : NLM denoise in Lab + CLAHE, bit-exact with cv2 (reference nesr/nesr.py:668-689) --------------
std::vector<int32_t> nlm_weight_table(float h, int channels);                    // non-zero prefix of cv2's almost_dist2weight_
const void* lab_table_host(int which, int* count, int* elem_bytes);              // the committed Lab tables (tests)
size_t preprocess_workspace_bytes(int H, int W, int tiles_x, int tiles_y);
cudaError_t launch_preprocess(const uint8_t* rgb, uint8_t* out, int H, int W, const int32_t* wtab_l, int n_wl, const int32_t* wtab_ab, int n_wab,
                              float clip, int tiles_x, int tiles_y, uint8_t* workspace, int* launches, cudaStream_t stream);

}  // namespace nesr
