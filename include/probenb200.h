/*
 * probenb200.h — C ABI of libprobenb200.so (sm_100a).
 *
 * Drop-in boundary for the RGB+thermal detection -> ProbEn hot path of
 * Jamie725/Multimodal-Object-Detection-via-Probabilistic-Ensembling (paths below are relative to the
 * reference checkout).  Every entry point:
 *   - is extern "C", takes plain pointers/sizes (no torch types),
 *   - takes BORROWED device pointers owned by the caller's allocator and a cudaStream_t (passed as void*),
 *   - is asynchronous on that stream (no internal device synchronisation, no host round trips),
 *   - returns PE_OK (0) or a negative pe_status; never throws.  pe_status_string() names the code.
 * There is no CPU fallback anywhere behind this header.
 */
#ifndef PROBENB200_H_
#define PROBENB200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define PE_API __attribute__((visibility("default")))
#else
#define PE_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pe_status {
  PE_OK = 0,
  PE_ERR_INVALID_ARGUMENT = -1,
  PE_ERR_UNSUPPORTED = -2,
  PE_ERR_CUDA = -3,
  PE_ERR_WORKSPACE_TOO_SMALL = -4,
  PE_ERR_NOT_INITIALISED = -5
} pe_status;

/* score fusion (demo/FLIR/demo_probEn.py:144-152) and box fusion (:155-167) selectors */
enum { PE_SCORE_PROBEN = 0, PE_SCORE_AVG = 1, PE_SCORE_MAX = 2 };
enum { PE_BOX_VAVG = 0, PE_BOX_SAVG = 1, PE_BOX_AVG = 2, PE_BOX_ARGMAX = 3 };

PE_API const char* pe_status_string(int status);
/* Library/ABI version; bumps whenever a signature changes. */
PE_API int pe_abi_version(void);
/* Text of the last CUDA error seen by this thread inside the library ("" if none). */
PE_API const char* pe_last_error_string(void);

/* ------------------------------------------------------------------------------------------------
 * ProbEn late fusion: replaces the per-image Python loop of demo/FLIR/demo_probEn.py
 *   apply_late_fusion_and_evaluate :236-267 (0 / 1 / >=2 non-empty models dispatch),
 *   fusion :189-196, prepare_data :79-90, nms_bayesian :92-187, bayesian_fusion_multiclass :32-42,
 *   weighted_box_fusion :73-77, avg_bbox_fusion :20-22, nms_1 :44-71 (+ detectron2/layers/nms.py:9-26).
 *
 * Layout (all device memory): detections of image b, model m are rows
 *   det_offsets[b*M+m] .. det_offsets[b*M+m+1]-1 of the SoA arrays (model order preserved, so the rows of
 *   one image are the reference's np.concatenate order).  boxes [N,4] xyxy (16-byte aligned), scores [N],
 *   classes [N] int32, probs [N,K] (first K softmax columns), vars [N].
 * Output: image b's fused detections are written, in the reference's output order (cluster heads by
 *   descending original score), to rows det_offsets[b*M] + 0..out_counts[b]-1 of out_boxes [N,4],
 *   out_scores [N], out_classes [N] (fused class may be K = background for score_mode PROBEN).
 *   out_counts[b] = 0 means "no model had detections" (the reference skips the image);
 *   out_counts[b] = -1 flags an image with more than pe_fuse_max_dets_per_image() detections.
 * workspace: pe_fuse_workspace_bytes(B) bytes of device scratch.
 * K must be 1 (binary, KAIST form demo_probEn.py:24-30) or 3 (FLIR).  M in [1, 31].
 */
PE_API size_t pe_fuse_workspace_bytes(int B);
PE_API int pe_fuse_max_dets_per_image(void);
PE_API int pe_fuse_batch(const float* boxes, const float* scores, const int32_t* classes, const float* probs,
                  const float* vars, const int32_t* det_offsets, int B, int M, int K, float iou_thr,
                  int score_mode, int box_mode, float img_w, float img_h, float* out_boxes,
                  float* out_scores, int32_t* out_classes, int32_t* out_counts, void* workspace,
                  size_t workspace_bytes, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Convolution / linear layer on tcgen05 tensor cores (implicit GEMM, TMA-fed, fp32 accumulate in TMEM).
 * Replaces the cuDNN/cuBLAS calls behind detectron2.layers.Conv2d + FrozenBatchNorm2d (folded into w/bias,
 * detectron2/layers/batch_norm.py:45-64) on the detector path: modeling/backbone/resnet.py:160-221,
 * modeling/backbone/fpn.py:127-137, modeling/proposal_generator/rpn.py:74-85, and nn.Linear in
 * modeling/roi_heads/box_head.py:73-81 / fast_rcnn.py:531-545 (H = 1, W = rows).
 *   x        [N, H, W, Cin]  bf16 NHWC, Cin % 64 == 0
 *   w        [Cout, KH, KW, Cin] bf16 (K-major), KH = KW in {1, 3}, zero padding KH/2
 *   bias     [Cout] fp32 or NULL;  Cout % 8 == 0
 *   stride   1, or 2 for 1x1 convs (stride_in_1x1, resnet.py:158)
 *   residual_mode 0: none; 1: += residual[N,Ho,Wo,Cout] (bottleneck shortcut, resnet.py:216-219);
 *                 2: += residual[N,ceil(Ho/2),ceil(Wo/2),Cout] nearest-2x upsampled (FPN top-down, fpn.py:131-133)
 *   relu     fused ReLU after the adds;  out_fp32: y is fp32 instead of bf16
 *   y        [N, Ho, Wo, Cout]
 */
typedef struct pe_conv_desc {
  int N, H, W, Cin, Cout, KH, KW, stride, relu, residual_mode, out_fp32;
} pe_conv_desc;
PE_API int pe_conv2d_fwd(const pe_conv_desc* desc, const void* x, const void* w, const float* bias,
                         const void* residual, void* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PROBENB200_H_ */
