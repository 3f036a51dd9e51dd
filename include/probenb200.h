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
  int in_fp16; /* x and w are IEEE fp16 instead of bf16 (stem: raw pixel range needs the 11-bit mantissa) */
} pe_conv_desc;
PE_API int pe_conv2d_fwd(const pe_conv_desc* desc, const void* x, const void* w, const float* bias,
                         const void* residual, void* y, void* stream);
/* 1x1 conv over TWO inputs as one GEMM: y = act([x | x2] . w + bias), w = [Cout][Cin + cin2] bf16, x2 = [N, h2, w2, cin2] bf16
 * read at `stride2` (1 | 2) so that it lands on x's H x W grid.  This is how the engine runs the first bottleneck block of a
 * stage: out = relu(conv3(t) + shortcut(x)) (modeling/backbone/resnet.py:205-221) = relu([t | x] . [W3 | Wsc] + b3 + bsc): the
 * projection-shortcut tensor is never written to HBM.  desc: KH = KW = 1, stride = 1, residual_mode = 0, Cin % 64 == 0. */
/* 1x1 conv with a CHAINED 1x1 conv on its own output: y = act([x | x2] . w + bias (+ residual)), y_c = act_c(y . wc + bias_c) with
 * wc = [n_c][Cout] bf16, n_c in {64, 128, 256}, y_c = [N, H, W, n_c] bf16.  The chained GEMM reads y's bf16 tiles from the shared-memory
 * staging buffers of the first epilogue (they are K-major 128B-swizzled A tiles already), so y is not re-read from HBM.  The engine
 * runs a bottleneck's conv3 and the NEXT block's conv1 this way (modeling/backbone/resnet.py:205-221).  x2 may be NULL (no second
 * input).  desc: KH = KW = 1, stride 1, bf16 output, Cout % 256 == 0, residual_mode 0 | 1. */
PE_API int pe_conv1x1_chain_fwd(const pe_conv_desc* desc, const void* x, const void* x2, int cin2, int h2, int w2, int stride2,
                                const void* w, const float* bias, const void* residual, void* y, const void* wc,
                                const float* bias_c, int n_c, int relu_c, void* y_c, void* stream);
/* RPN head in one kernel (proposal_generator/rpn.py:74-85): t = relu(conv3x3(x, w) + bias), y_c = t . wc + bias_c with wc = [16][256]
 * bf16 (3 objectness rows | 1 zero row | 12 delta rows, the engine's kind-3 parameter) and y_c = [N, H, W, 16] float32.  The hidden
 * tensor t exists only as bf16 sub-tiles in shared memory (the A operand of the second tcgen05 GEMM); it is never written to HBM.
 * desc: KH = KW = 3, stride 1, Cin % 8 == 0, Cout == 256, relu as given, residual_mode 0, out_fp32 0. */
PE_API int pe_conv_rpn_head_fwd(const pe_conv_desc* desc, const void* x, const void* w, const float* bias, const void* wc,
                                const float* bias_c, float* y_c, void* stream);
PE_API int pe_conv1x1_dual_fwd(const pe_conv_desc* desc, const void* x, const void* x2, int cin2, int h2, int w2, int stride2,
                               const void* w, const float* bias, void* y, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Detector engine: the inference path of detectron2/modeling/meta_arch/rcnn.py:219-267 (GeneralizedRCNN with
 * ResNet-50/101-FPN, RPN, StandardROIHeads + the fork's variance head, fast_rcnn.py:531-545) as a native
 * kernel sequence.  The engine owns no device memory:
 *   weights    one blob laid out per the manifest pe_detector_param_info() publishes (the host folds
 *              FrozenBatchNorm2d into conv weights/biases and repacks to [Cout][KH][KW][Cin] bf16; kinds:
 *              0 conv+BN, 1 conv+bias, 2 stem 7x7 as [64][Kpad] fp16, 3 RPN objectness|deltas rows,
 *              4 fc1 with (c,ph,pw)->(ph,pw,c) column permutation, 5 linear, 6 cls|bbox|var predictor rows,
 *              7 first-block conv3 with the block's projection shortcut appended along K, see pe_conv1x1_dual_fwd);
 *   workspace  pe_detector_workspace_bytes() of scratch; pe_detector_buffer_info() exposes named
 *              intermediates (p2..p6, rpn_out2..6, proposals, roi_feats, head_out, ...) for stage-wise parity tests;
 *   images     [B, in_channels, img_h, img_w] float32 (what GeneralizedRCNN.forward receives in
 *              batched_inputs[i]["image"], i.e. after ResizeShortestEdge), normalised + zero padded to the
 *              canvas on the fly (rcnn.py:269-286);
 *   out        fixed-stride records, PE_MAX_DETECTIONS per image, already rescaled to (out_h, out_w) by
 *              detector_postprocess (postprocessing.py:8-52); field names follow the fork's Instances fields.
 */
#define PE_MAX_DETECTIONS 100
typedef struct pe_detector pe_detector;
typedef struct pe_detector_config {
  int depth;           /* 50 | 101 (MODEL.RESNETS.DEPTH) */
  int in_channels;     /* 3 | 4 (early fusion BGRT) | 6 with middle_fusion (BGRTTT) */
  int middle_fusion;   /* rcnn.py:240-248: shared backbone on both 3-channel halves, features concatenated */
  int num_classes;     /* K: 3 (FLIR) | 1 (KAIST) */
  int max_batch;
  int canvas_h, canvas_w;            /* padded network input, multiples of 32 (fpn.py:102) */
  float pixel_mean[8], pixel_std[8];
  float score_thresh, nms_thresh, rpn_nms_thresh;
  int pre_nms_topk, post_nms_topk, detections_per_image;
} pe_detector_config;
typedef struct pe_param_info {
  char name[96];
  int kind, cout, kh, kw, cin;
  size_t weight_offset, bias_offset;
} pe_param_info;
typedef struct pe_detections {
  float* boxes;          /* [B, 100, 4] xyxy in the (out_h, out_w) frame */
  float* scores;         /* [B, 100] */
  int32_t* classes;      /* [B, 100] */
  float* class_logits;   /* [B, 100, K+1] */
  float* probs;          /* [B, 100, K]   (prob_score) */
  float* vars;           /* [B, 100]      (vars, with the reference's candidate-index quirk) */
  int32_t* roi_index;    /* [B, 100] proposal row each detection came from */
  int32_t* counts;       /* [B] */
} pe_detections;
PE_API int pe_detector_create(const pe_detector_config* cfg, pe_detector** out);
PE_API void pe_detector_destroy(pe_detector* d);
PE_API int pe_detector_num_params(const pe_detector* d);
PE_API int pe_detector_param_info(const pe_detector* d, int i, pe_param_info* info);
PE_API size_t pe_detector_weight_bytes(const pe_detector* d);
PE_API size_t pe_detector_workspace_bytes(const pe_detector* d);
PE_API int pe_detector_buffer_info(const pe_detector* d, const char* name, size_t* offset, int* dims4, int* elem_bytes);
PE_API int pe_detector_forward(pe_detector* d, const void* weights, const float* images, int B, int img_h, int img_w,
                               float out_h, float out_w, const pe_detections* out, void* workspace, size_t workspace_bytes,
                               void* stream);
/* Module seams of the reference's registries (modeling/backbone/build.py:20-33, proposal_generator/build.py, roi_heads/roi_heads.py,
 * meta_arch/build.py:12-19): the same plan, restricted to the stages in `stages`; the stages exchange their tensors through the
 * named workspace buffers (p2..p6 / pout*_0, proposals + prop_count, head_out), which the host may read or overwrite in between.
 * prenormalized != 0: `images` already went through rcnn.py:269-286 (what Backbone.forward receives). */
#define PE_STAGE_BACKBONE 1
#define PE_STAGE_RPN 2
#define PE_STAGE_ROI_HEADS 4
#define PE_STAGE_ALL 7
PE_API int pe_detector_forward_stages(pe_detector* d, const void* weights, const float* images, int B, int img_h, int img_w,
                                      float out_h, float out_w, const pe_detections* out, void* workspace, size_t workspace_bytes,
                                      int stages, int prenormalized, void* stream);

/* Same, from raw uint8 HWC frames [B, src_h, src_w, in_channels]: DefaultPredictor's resize to (img_h, img_w)
 * (pe_resize_frames arithmetic) is fused into the stem's input staging, so no float32 image is materialised. */
PE_API int pe_detector_forward_frames(pe_detector* d, const void* weights, const uint8_t* frames, int B, int src_h, int src_w,
                                      int img_h, int img_w, int round_u8, float out_h, float out_w, const pe_detections* out,
                                      void* workspace, size_t workspace_bytes, void* stream);
/* Instrumentation for bench.py: enabled = 1: CUDA events around every tensor-core GEMM launch of the next forwards (per-layer table);
 * enabled = 2: one event pair around every RUN of back-to-back GEMM launches (no event in between, so the launch gaps and the
 * programmatic-dependent-launch overlap of consecutive layers count exactly as in a real forward); 0 = off. */
PE_API int pe_detector_set_profiling(pe_detector* d, int enabled);
/* Cross-detector stagger: every following forward records `event` (a cudaEvent_t; NULL switches it off) on its stream once
 * `after_launches` kernels have been launched (at the latest when the forward ends).  A second detector whose stream waits for
 * that event starts that much later, so the latency-bound stages of the two (RPN top-k / NMS, a few CTAs each) do not coincide
 * and each runs under the other one's GEMMs.  Works under stream capture (the record / wait pair becomes a graph edge). */
PE_API int pe_detector_set_stagger_event(pe_detector* d, void* event, int after_launches);
PE_API int pe_detector_last_profile(pe_detector* d, float* gemm_ms, float* span_ms, int* launches, int* gemm_launches);
/* Per GEMM launch of the last profiled forward: device ms, algorithmic FLOPs, algorithmic bytes (every operand once).
 * Fills at most `capacity` entries; returns the number of launches recorded (0 when profiling is off). */
PE_API int pe_detector_profile_launches(pe_detector* d, float* ms, double* flops, double* bytes, int capacity);
/* Non-GEMM launch groups of the last forward profiled with mode 1 (canvas staging, max-pool, RPN selection + NMS, ROIAlign,
 * head post-processing ...): `names` receives one 32-byte NUL-terminated launcher name per entry, `ms` the device time.
 * Fills at most `capacity` entries; returns the number recorded. */
PE_API int pe_detector_profile_kernels(pe_detector* d, char* names, float* ms, int capacity);
/* DefaultPredictor's ResizeShortestEdge (engine/defaults.py:186-190): uint8 HWC frames [B,src_h,src_w,C] ->
 * float32 CHW [B,C,dst_h,dst_w], bilinear with half-pixel centres; round_u8 mimics PIL's uint8 output. */
PE_API int pe_resize_frames(const uint8_t* frames, float* out, int B, int C, int src_h, int src_w, int dst_h, int dst_w,
                            int round_u8, void* stream);
/* Gathers M models' detections into pe_fuse_batch's packed layout (det_offsets [B*M+1] + SoA rows). */
PE_API int pe_pack_detections(const pe_detections* models, int M, int B, int K, int32_t* det_offsets, float* boxes,
                              float* scores, int32_t* classes, float* probs, float* vars, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Op-level entry points of the detector's non-GEMM stages (the engine launches the same kernels).
 *
 * pe_rpn_proposals: find_top_rpn_proposals (modeling/proposal_generator/rpn_outputs.py:52-162) + anchors
 *   (anchor_generator.py:130-199, sizes 32..512 x ratios .5/1/2, strides 4..64) + apply_deltas
 *   (box_regression.py:78-115).  rpn_out[l]: [B, H[l], W[l], 16] fp32 channels-last = 3 objectness logits,
 *   1 pad, 12 anchor deltas (a*4+j).  proposals [B, 1000, 4], proposal_counts [B].
 * pe_roi_align_fwd: ROIPooler + ROIAlign(7x7, sampling_ratio 0, aligned) (modeling/poolers.py:180-235,
 *   layers/csrc/ROIAlign/ROIAlign_cuda.cu:65-139; pybind seam detectron2._C.roi_align_forward, csrc/vision.cpp:89).
 *   features[l]: p2..p5 [B, H[l], W[l], C] bf16; out [B*max_props, 49, C] bf16 (rows past the count are zero).
 * pe_head_postprocess: FastRCNNOutputs.inference + fast_rcnn_inference_single_image (fast_rcnn.py:86-147,
 *   345-360,417-452) + detector_postprocess (postprocessing.py:8-52).  head_out [B*max_props, npad] fp32 rows =
 *   K+1 class logits | 4K box deltas | 1 log-variance | pad.
 */
PE_API size_t pe_rpn_proposals_workspace_bytes(int B);
PE_API int pe_rpn_proposals(const float* const* rpn_out, const int* H, const int* W, int B, int pre_nms_topk,
                            int post_nms_topk, float nms_thresh, float img_h, float img_w, float* proposals,
                            int32_t* proposal_counts, void* workspace, size_t workspace_bytes, void* stream);
PE_API int pe_roi_align_fwd(const void* const* features, const int* H, const int* W, int C, const float* proposals,
                            const int32_t* proposal_counts, int B, int max_props, void* out, void* stream);
PE_API int pe_head_postprocess(const float* head_out, int npad, const float* proposals, const int32_t* proposal_counts,
                               int B, int max_props, int K, float img_h, float img_w, float out_h, float out_w,
                               float score_thresh, float nms_thresh, int detections_per_image,
                               const pe_detections* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Operator seams of detectron2/layers, for callers written against the reference's ops.
 *
 * pe_batched_nms: detectron2/layers/nms.py:9-26 batched_nms -> torchvision.ops.boxes.batched_nms.  boxes [n,4]
 *   xyxy float32 (16-byte aligned), scores [n], idxs [n] int64 category ids (NULL = plain torchvision.ops.nms).
 *   mode 0 = torchvision's coordinate-offset trick (boxes + float(idx) * (max coordinate + 1) in float32, the
 *   path the reference takes for <= 20000 box elements on CUDA), mode 1 = suppression inside a category only
 *   (_batched_nms_vanilla, and the reference's own loop for >= 40000 boxes).  keep [n] int64 receives the kept
 *   ORIGINAL indices in descending score order (ties: lower index first), *n_keep their number.
 * pe_roi_align_forward: detectron2._C.roi_align_forward (layers/csrc/vision.cpp:89, ROIAlign/ROIAlign.h:54-84,
 *   ROIAlign_cuda.cu:65-139).  input [N,C,H,W] float32, rois [num_rois,5] = (batch index, x1, y1, x2, y2),
 *   out [num_rois, C, pooled_h, pooled_w] float32 (caller-allocated; the reference op allocates it itself).
 */
PE_API size_t pe_batched_nms_workspace_bytes(int n);
PE_API int pe_batched_nms_max_boxes(void);
PE_API int pe_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int n, float iou_thr, int mode,
                          int64_t* keep, int32_t* n_keep, void* workspace, size_t workspace_bytes, void* stream);
PE_API int pe_roi_align_forward(const float* input, int N, int C, int H, int W, const float* rois, int num_rois,
                                float spatial_scale, int pooled_h, int pooled_w, int sampling_ratio, int aligned,
                                float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * COCO bbox evaluation, matching stage: COCOeval.evaluateImg + computeIoU of the reference's vendored
 * pycocotools (detectron2/pycocotools/cocoeval.py:124-320; `_mask.iou` = maskApi bbIou) as FLIREvaluator drives them
 * (detectron2/evaluation/FLIR_evaluation.py:496-563).  Groups p = (image, category) pairs with any ground truth or detection:
 *   gt_boxes [n_gt,4] xywh float64, gt_area [n_gt], gt_iscrowd [n_gt] (= the 'ignore' flag, cocoeval.py:246), gt_offsets [P+1];
 *   dt_boxes [n_dt,4] xywh float64 sorted by descending score inside each group (stable) and capped at maxDets, dt_area, dt_offsets;
 *   iou_thrs [T], area_rng [A][2]; T * A <= 64.
 * Outputs (uint8): dt_matched [A][T][n_dt], dt_ignore [A][T][n_dt], gt_ignore [A][n_gt].  The accumulate stage stays on the host
 * (probenb200/evaluation.py).  max_gt_per_group sizes the per-block scratch (<= pe_coco_match_max_gt()).
 */
PE_API int pe_coco_match_max_gt(void);
PE_API int pe_coco_match(const double* gt_boxes, const double* gt_area, const uint8_t* gt_iscrowd, const int32_t* gt_offsets,
                         const double* dt_boxes, const double* dt_area, const int32_t* dt_offsets, int P, const double* iou_thrs, int T,
                         const double* area_rng, int A, long long n_dt, long long n_gt, int max_gt_per_group, uint8_t* dt_matched,
                         uint8_t* dt_ignore, uint8_t* gt_ignore, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Input side (demo/FLIR/demo_FLIR_save_predictions.py:93-121: cv2.imread of the RGB / thermal JPEGs, cv2.resize of
 * the RGB frame to the thermal size, 3-/4-/6-channel assembly).
 *
 * pe_jpeg_*: nvJPEG decode (libnvjpeg of the CUDA toolkit, loaded on first use; PE_ERR_UNSUPPORTED if absent) of n
 *   JPEG byte strings held in HOST memory into frames [n, height, width, 3] interleaved BGR uint8 in DEVICE memory -
 *   cv2.imread's layout; grey-scale JPEGs (FLIR thermal_8_bit) are replicated into the three channels.  Every image
 *   must have the stated size.  The entropy decode runs on the calling host thread, the rest on `stream`, which is
 *   synchronised after every image (the decoder state's staging buffers are reused) - the only entry point that
 *   blocks on its stream.
 * pe_resize_u8_cv: cv2.resize(src, (dst_w, dst_h)) for uint8, i.e. INTER_LINEAR with OpenCV's 11-bit fixed-point
 *   weights, bit for bit.  Reads channels [src_c0, src_c0+channels) of src [B, src_h, src_w, src_channels] and writes
 *   channels [dst_c0, dst_c0+channels) of dst [B, dst_h, dst_w, dst_channels]: with equal sizes it is a channel
 *   copy, so two calls assemble the early-fusion BGRT or middle-fusion BGRTTT input (:104-121).
 */
typedef struct pe_jpeg_decoder pe_jpeg_decoder;
PE_API int pe_jpeg_create(pe_jpeg_decoder** out);
PE_API void pe_jpeg_destroy(pe_jpeg_decoder* d);
PE_API int pe_jpeg_image_info(pe_jpeg_decoder* d, const uint8_t* data, size_t bytes, int* height, int* width, int* components);
PE_API int pe_jpeg_decode_batch(pe_jpeg_decoder* d, const uint8_t* const* data, const size_t* bytes, int n, uint8_t* frames,
                                int height, int width, void* stream);
PE_API int pe_resize_u8_cv(const uint8_t* src, int src_channels, int src_c0, uint8_t* dst, int dst_channels, int dst_c0,
                           int channels, int B, int src_h, int src_w, int dst_h, int dst_w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PROBENB200_H_ */
