// Parameter blocks + launchers of the non-GEMM detector kernels (detector_kernels.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace pe {

constexpr int kRpnLevels = 5;
constexpr int kRpnOutC = 16;   // 3 objectness logits | 1 pad | 12 anchor deltas, fp32, channels-last
constexpr int kTopkSlots = 1024;
constexpr int kMaxDet = 100;   // TEST.DETECTIONS_PER_IMAGE (config/defaults.py:560)
constexpr float kScaleClamp = 4.135166556742356f;  // log(1000/16), box_regression.py:11

struct StemNorm { float mean[8]; float std[8]; };

struct RpnLevels {
  const float* out[kRpnLevels];   // [B, H, W, kRpnOutC]
  const float* logit[kRpnLevels]; // optional dense copy of the objectness logits [B, H, W, 4] (3 logits | pad) or nullptr
  int H[kRpnLevels], W[kRpnLevels], stride[kRpnLevels];
  float anchor[kRpnLevels][3][4]; // cell anchors (anchor_generator.py:151-187), float32 of the double maths
};

struct RpnScratch {
  float4* cand_box;           // [B, 5, 1024]
  float* cand_score;          // [B, 5, 1024]
  unsigned char* cand_valid;  // [B, 5, 1024]
  int* cand_count;            // [B, 5]
  int* keep_idx;              // [B, 5, 1024]
  int* keep_count;            // [B, 5]
  unsigned* nms_mask;         // [B, 5, 1024, 32] suppression bitmask
};

struct RoiLevels {
  const __nv_bfloat16* feat[4];   // p2..p5, [B, H, W, C]
  int H[4], W[4];
  float scale[4];
};

struct HeadParams {
  float img_h, img_w;     // network input size (after resize), used for clipping
  float out_h, out_w;     // size the boxes are rescaled to (postprocessing.py)
  float scale_x, scale_y;
  float score_thresh, nms_thresh;
  int max_det;
};

struct DetOut {
  float4* boxes; float* scores; int* classes; float* logits; float* probs; float* vars; int* roi_index; int* count;
};

struct PackIn {
  const float4* boxes[4]; const float* scores[4]; const int* classes[4]; const float* probs[4]; const float* vars[4];
  const int* count[4];
};

int launch_stem_im2col(const float* img, void* canvas, void* A, int B, int Ctot, int c0, int C, int Hi, int Wi, int Hc, int Wc,
                       const StemNorm& nrm, cudaStream_t st);
int launch_stem_im2col_u8(const unsigned char* frames, void* canvas, void* A, int B, int Ctot, int c0, int C, int Hs, int Ws, int Hi,
                          int Wi, int Hc, int Wc, int round_u8, const StemNorm& nrm, cudaStream_t st, void* taps_ws = nullptr);
constexpr int kStemK = 224;  // 7 rows x (8 px x 4 ch): K of the stem GEMM (147 real taps, the rest meet zero weights)
int launch_resize_frames(const unsigned char* src, float* dst, int B, int C, int Hs, int Ws, int Hd, int Wd, int round_u8,
                         cudaStream_t st);
int launch_maxpool(const void* x, void* y, int B, int H, int W, int C, cudaStream_t st, int reverse = 0);
int launch_subsample2(const void* x, void* y, int B, int H, int W, int C, cudaStream_t st);
int launch_concat_channels(const void* a, const void* b, void* y, long long pixels, int C, cudaStream_t st);
int launch_rpn_proposals(const RpnLevels& lv, int B, int pre_topk, int post_topk, float nms_thr, float img_h, float img_w,
                         const RpnScratch& s, int max_props, float4* props, int* prop_count, cudaStream_t st);
int launch_roi_align(const RoiLevels& fl, const float4* props, const int* prop_count, int B, int max_props, int C, void* out,
                     cudaStream_t st);
int launch_head_post(const float* head, int npad, const float4* props, const int* prop_count, int B, int max_props, int K,
                     const HeadParams& hp, const DetOut& out, cudaStream_t st);
int launch_pack(const PackIn& in, int B, int M, int K, int* offsets, float4* boxes, float* scores, int* classes, float* probs,
                float* vars, cudaStream_t st);

}  // namespace pe
