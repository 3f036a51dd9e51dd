// COCO bbox evaluation, the per-(image, category) matching stage on the GPU.
//
// Replaces COCOeval.evaluateImg / computeIoU of the reference's vendored pycocotools (detectron2/pycocotools/cocoeval.py:124-320,
// driven by FLIREvaluator, detectron2/evaluation/FLIR_evaluation.py:496-563) and the C `bbIou` behind `_mask.iou`
// (pycocotools is not installed).  The accumulate stage (sort all detections of a category by score,
// cumulative TP / FP, precision envelope, 101 recall samples) stays on the host: it is one O(D log D) pass.
//
// One block per (image, category) group that has ground truth or detections.  Thread (t, a) owns IoU threshold t and area range
// a and runs the reference's greedy loop: detections in descending score order (the host passes them sorted, capped at maxDets),
// ground truth with the ignored boxes last (ignored = iscrowd or area outside the range; two passes over the original order
// reproduce the stable argsort of cocoeval.py:254-255), a detection takes the best still-free ground truth with IoU >= the
// threshold, crowd boxes can be taken repeatedly, a non-ignored match is never traded for an ignored one.  IoUs are float64 in
// the operation order of the numpy restatement (probenb200/evaluation.py bbox_iou), without FMA contraction, so every decision is
// bit-identical to the host evaluator.
#include "common.cuh"

namespace pe {
namespace {

constexpr int kEvalThreads = 64;   // >= T * A (10 x 4)
constexpr int kMaxGtPerGroup = 1024;

__device__ __forceinline__ double box_iou_f64(const double* d, const double* g, bool crowd) {
  const double da = __dmul_rn(d[2], d[3]), ga = __dmul_rn(g[2], g[3]);
  const double w = __dsub_rn(fmin(__dadd_rn(d[0], d[2]), __dadd_rn(g[0], g[2])), fmax(d[0], g[0]));
  const double h = __dsub_rn(fmin(__dadd_rn(d[1], d[3]), __dadd_rn(g[1], g[3])), fmax(d[1], g[1]));
  const double inter = __dmul_rn(fmax(w, 0.0), fmax(h, 0.0));
  const double uni = crowd ? da : __dsub_rn(__dadd_rn(da, ga), inter);
  return __ddiv_rn(inter, uni);
}

__global__ void __launch_bounds__(kEvalThreads) coco_match_kernel(
    const double* __restrict__ gt_box, const double* __restrict__ gt_area, const unsigned char* __restrict__ gt_crowd,
    const int* __restrict__ gt_off, const double* __restrict__ dt_box, const double* __restrict__ dt_area,
    const int* __restrict__ dt_off, int P, const double* __restrict__ iou_thrs, int T, const double* __restrict__ area_rng, int A,
    long long Dtot, long long Gtot, unsigned char* __restrict__ dt_matched, unsigned char* __restrict__ dt_ignore,
    unsigned char* __restrict__ gt_ignore) {
  extern __shared__ unsigned char s_taken[];  // [T * A][G]: detection index + 1 that took the ground truth (0 = free) -> only a flag is needed
  const int p = blockIdx.x;
  if (p >= P) return;
  const int g0 = gt_off[p], G = gt_off[p + 1] - g0;
  const int d0 = dt_off[p], D = dt_off[p + 1] - d0;
  const int tid = threadIdx.x;
  const int a = tid / T, t = tid - a * T;
  // ground-truth ignore flags per area range (cocoeval.py:246-250)
  for (int i = tid; i < A * G; i += blockDim.x) {
    const int aa = i / G, g = i - aa * G;
    const double ar = gt_area[g0 + g];
    gt_ignore[(size_t)aa * Gtot + g0 + g] = (gt_crowd[g0 + g] || ar < area_rng[2 * aa] || ar > area_rng[2 * aa + 1]) ? 1 : 0;
  }
  __syncthreads();
  if (tid >= T * A) return;
  unsigned char* taken = s_taken + (size_t)tid * G;
  for (int g = 0; g < G; ++g) taken[g] = 0;
  const unsigned char* gig = gt_ignore + (size_t)a * Gtot + g0;
  const double thr = fmin(iou_thrs[t], 1.0 - 1e-10);
  const double lo = area_rng[2 * a], hi = area_rng[2 * a + 1];
  for (int d = 0; d < D; ++d) {
    const double* db = dt_box + 4 * (size_t)(d0 + d);
    double iou = thr;
    int m = -1;
    bool m_ign = false, done = false;
    // pass 0: non-ignored ground truth in original order; pass 1: ignored ground truth (= the stable sort by the ignore flag)
    for (int pass = 0; pass < 2 && !done; ++pass) {
      for (int g = 0; g < G; ++g) {
        const bool ig = gig[g] != 0;
        if ((int)ig != pass) continue;
        const bool crowd = gt_crowd[g0 + g] != 0;
        if (taken[g] && !crowd) continue;                      // already matched, and not a crowd
        if (m > -1 && !m_ign && ig) { done = true; break; }    // a regular match is never traded for an ignored box
        const double v = box_iou_f64(db, gt_box + 4 * (size_t)(g0 + g), crowd);
        if (v < iou) continue;
        iou = v;
        m = g;
        m_ign = ig;
      }
    }
    const size_t o = ((size_t)a * T + t) * Dtot + d0 + d;
    bool ign = false;
    if (m >= 0) {
      taken[m] = 1;
      ign = m_ign;
    } else {
      const double ar = dt_area[d0 + d];
      ign = ar < lo || ar > hi;                                // unmatched detections outside the area range are ignored
    }
    dt_matched[o] = m >= 0 ? 1 : 0;
    dt_ignore[o] = ign ? 1 : 0;
  }
}

}  // namespace
}  // namespace pe

extern "C" PE_API int pe_coco_match_max_gt(void) { return pe::kMaxGtPerGroup; }

extern "C" PE_API int pe_coco_match(const double* gt_boxes, const double* gt_area, const uint8_t* gt_iscrowd, const int32_t* gt_offsets,
                                    const double* dt_boxes, const double* dt_area, const int32_t* dt_offsets, int P,
                                    const double* iou_thrs, int T, const double* area_rng, int A, long long n_dt, long long n_gt,
                                    int max_gt_per_group, uint8_t* dt_matched, uint8_t* dt_ignore, uint8_t* gt_ignore, void* stream) {
  if (P < 0 || T < 1 || A < 1 || T * A > pe::kEvalThreads) return PE_ERR_INVALID_ARGUMENT;
  if (P == 0) return PE_OK;
  if (!gt_offsets || !dt_offsets || !iou_thrs || !area_rng || !dt_matched || !dt_ignore || !gt_ignore) return PE_ERR_INVALID_ARGUMENT;
  if (max_gt_per_group < 0 || max_gt_per_group > pe::kMaxGtPerGroup) return PE_ERR_UNSUPPORTED;
  const size_t smem = (size_t)T * A * (size_t)(max_gt_per_group > 0 ? max_gt_per_group : 1);
  static pe::DeviceOnce once;
  if (smem > 48 * 1024) {
    if (once.needed()) {
      PE_CUDA_CHECK(cudaFuncSetAttribute(pe::coco_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(pe::kEvalThreads * pe::kMaxGtPerGroup)));
      once.mark();
    }
  }
  pe::coco_match_kernel<<<P, pe::kEvalThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      gt_boxes, gt_area, gt_iscrowd, gt_offsets, dt_boxes, dt_area, dt_offsets, P, iou_thrs, T, area_rng, A, n_dt, n_gt, dt_matched,
      dt_ignore, gt_ignore);
  PE_LAUNCH_CHECK();
  return PE_OK;
}
