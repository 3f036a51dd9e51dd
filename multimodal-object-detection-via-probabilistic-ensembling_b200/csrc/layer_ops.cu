// Stand-alone operator seams of the reference's detectron2/layers package:
//   pe_batched_nms       <- detectron2/layers/nms.py:9-26 batched_nms (-> torchvision.ops.boxes.batched_nms / nms)
//   pe_roi_align_forward <- detectron2._C.roi_align_forward (layers/csrc/vision.cpp:89, ROIAlign/ROIAlign.h:54-84,
//                           ROIAlign_cuda.cu:10-139), NCHW float32, any pooled size / sampling ratio / aligned flag
// The detector engine does not go through these: it launches fused, layout-specialised kernels
// (detector_kernels.cu).  These entry points exist so that code written against the reference's operators -
// its ROIAlign layer, its batched_nms - can switch to this library call for call.
#include <math.h>
#include "common.cuh"

namespace pe {
namespace {

// ---- batched NMS -------------------------------------------------------------------------------------------
// 1. rank by counting: position of every box in the stable descending-score order (ties: lower index first,
//    torchvision's CPU kernel sorts stably), 2. gather boxes in that order, with torchvision's coordinate-offset
//    trick (mode 0: box + float(idx) * (max coordinate + 1), float32) or with the class id kept aside (mode 1,
//    _batched_nms_vanilla: suppression only inside a class), 3. upper-triangular suppression bitmask,
//    4. one-block scan that walks the order, keeps a box iff no kept box suppressed it, and emits original indices.

constexpr int kNmsBlock = 256;

__device__ __forceinline__ unsigned ordered_key(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void nms_max_coord_kernel(const float4* __restrict__ boxes, int n, unsigned* __restrict__ max_key) {
  float m = -INFINITY;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 b = boxes[i];
    m = fmaxf(m, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
  }
  unsigned k = ordered_key(m);
  k = __reduce_max_sync(kFullMask, k);
  if ((threadIdx.x & 31) == 0) atomicMax(max_key, k);
}

__global__ void __launch_bounds__(kNmsBlock) nms_rank_kernel(const float* __restrict__ scores, int n, int* __restrict__ order) {
  __shared__ float s_tile[kNmsBlock];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // total order on the raw bits (NaN sorts first like torch.sort(descending=True)); equal floats have equal keys
  const unsigned ki = ordered_key(i < n ? scores[i] : 0.f);
  int rank = 0;
  for (int j0 = 0; j0 < n; j0 += kNmsBlock) {
    __syncthreads();
    s_tile[threadIdx.x] = j0 + threadIdx.x < n ? scores[j0 + threadIdx.x] : -INFINITY;
    __syncthreads();
    const int lim = min(kNmsBlock, n - j0);
    for (int t = 0; t < lim; ++t) {
      const unsigned kj = ordered_key(s_tile[t]);
      rank += (kj > ki) || (kj == ki && j0 + t < i);
    }
  }
  if (i < n) order[rank] = i;
}

__global__ void nms_gather_kernel(const float4* __restrict__ boxes, const long long* __restrict__ idxs, const int* __restrict__ order,
                                  int n, int mode, const unsigned* __restrict__ max_key, float4* __restrict__ sorted_boxes,
                                  long long* __restrict__ sorted_cls) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int i = order[r];
  float4 b = boxes[i];
  const long long c = idxs ? idxs[i] : 0;
  if (mode == 0 && idxs) {
    const unsigned k = *max_key;
    const float mc = __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
    const float off = __fmul_rn((float)c, __fadd_rn(mc, 1.f));
    b = make_float4(__fadd_rn(b.x, off), __fadd_rn(b.y, off), __fadd_rn(b.z, off), __fadd_rn(b.w, off));
  }
  sorted_boxes[r] = b;
  sorted_cls[r] = c;
}

// word (i, w) = bits j in [32w, 32w+32) with j > i and IoU(i, j) > thr (torchvision's float32 expression)
__global__ void __launch_bounds__(kNmsBlock) nms_bitmask_kernel(const float4* __restrict__ sb, const long long* __restrict__ sc, int n,
                                                                 int words, float thr, int per_class, unsigned* __restrict__ mask) {
  __shared__ float4 s_box[32];
  __shared__ long long s_cls[32];
  const int w = blockIdx.x;           // column word
  const int i0 = blockIdx.y * kNmsBlock;
  if (i0 >= (w + 1) * 32) {           // whole row block lies past the column word: nothing above the diagonal
    const int i = i0 + threadIdx.x;
    if (i < n) mask[(size_t)i * words + w] = 0u;
    return;
  }
  if (threadIdx.x < 32) {
    const int j = w * 32 + threadIdx.x;
    s_box[threadIdx.x] = j < n ? sb[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    s_cls[threadIdx.x] = j < n ? sc[j] : -1;
  }
  __syncthreads();
  const int i = i0 + threadIdx.x;
  if (i >= n) return;
  const float4 a = sb[i];
  const long long ca = sc[i];
  const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  unsigned bits = 0u;
  const int jn = min(32, n - w * 32);
  for (int t = 0; t < jn; ++t) {
    const int j = w * 32 + t;
    if (j <= i) continue;
    if (per_class && s_cls[t] != ca) continue;
    const float4 b = s_box[t];
    const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    const float iw = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    const float ih = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    const float inter = __fmul_rn(iw, ih);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
    bits |= (unsigned)(ovr > thr) << t;
  }
  mask[(size_t)i * words + w] = bits;
}

__global__ void __launch_bounds__(1024) nms_keep_kernel(const unsigned* __restrict__ mask, const int* __restrict__ order, int n, int words,
                                                        long long* __restrict__ keep, int* __restrict__ n_keep) {
  extern __shared__ unsigned s_removed[];  // [words]
  for (int w = threadIdx.x; w < words; w += blockDim.x) s_removed[w] = 0u;
  __syncthreads();
  int kept = 0;
  for (int i = 0; i < n; ++i) {
    if ((s_removed[i >> 5] >> (i & 31)) & 1u) continue;  // block-uniform
    if (threadIdx.x == 0) keep[kept] = order[i];
    ++kept;
    __syncthreads();  // everyone has read the removed word before it changes
    const unsigned* row = mask + (size_t)i * words;
    for (int w = (i >> 5) + threadIdx.x; w < words; w += blockDim.x) s_removed[w] |= row[w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_keep = kept;
}

struct NmsWorkspace {
  unsigned* max_key;
  int* order;
  float4* sorted_boxes;
  long long* sorted_cls;
  unsigned* mask;
};

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

size_t nms_layout(int n, unsigned char* base, NmsWorkspace* ws) {
  const size_t words = (size_t)(n + 31) / 32;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
  const size_t o_key = take(256), o_order = take((size_t)n * 4), o_box = take((size_t)n * 16), o_cls = take((size_t)n * 8);
  const size_t o_mask = take((size_t)n * words * 4);
  if (ws) {
    ws->max_key = reinterpret_cast<unsigned*>(base + o_key);
    ws->order = reinterpret_cast<int*>(base + o_order);
    ws->sorted_boxes = reinterpret_cast<float4*>(base + o_box);
    ws->sorted_cls = reinterpret_cast<long long*>(base + o_cls);
    ws->mask = reinterpret_cast<unsigned*>(base + o_mask);
  }
  return off;
}

// ---- ROIAlign forward, NCHW float32 ------------------------------------------------------------------------
// One block per ROI.  The sampling grid of an ROI is separable: the block first tabulates, once, the row taps
// (pooled_h x grid_h entries: two source rows + two weights) and the column taps (pooled_w x grid_w) in shared
// memory; every thread then owns (channel, bin) outputs and only gathers and accumulates.  The reference kernel
// recomputes the taps for every output element of every channel (ROIAlign_cuda.cu:65-139).  ROIs whose adaptive
// grid does not fit the tables (> kTapCap entries per axis) take the direct path.

constexpr int kRoiThreads = 256;
constexpr int kTapCap = 1024;

struct Tap {
  int lo, hi;     // source index pair, or lo = -1 for a sample outside [-1, size]
  float wl, wh;   // weights of lo / hi
};

__device__ __forceinline__ Tap make_tap(float v, int size) {
  Tap t;
  if (v < -1.f || v > (float)size) { t.lo = -1; t.hi = 0; t.wl = 0.f; t.wh = 0.f; return t; }
  if (v <= 0.f) v = 0.f;
  int lo = (int)v, hi;
  if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; } else { hi = lo + 1; }
  const float l = v - (float)lo;
  t.lo = lo; t.hi = hi; t.wl = 1.f - l; t.wh = l;
  return t;
}

__global__ void __launch_bounds__(kRoiThreads) roi_align_nchw_kernel(const float* __restrict__ input, int N, int C, int H, int W,
                                                                      const float* __restrict__ rois, float scale, int ph, int pw,
                                                                      int sampling_ratio, int aligned, float* __restrict__ out) {
  __shared__ Tap s_ty[kTapCap], s_tx[kTapCap];
  const int r = blockIdx.x;
  const float* roi = rois + (size_t)r * 5;
  const int b = (int)roi[0];
  const float offset = aligned ? 0.5f : 0.f;
  const float x0 = roi[1] * scale - offset, y0 = roi[2] * scale - offset;
  const float x1 = roi[3] * scale - offset, y1 = roi[4] * scale - offset;
  float rw = x1 - x0, rh = y1 - y0;
  if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
  const float bin_h = rh / (float)ph, bin_w = rw / (float)pw;
  const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)ph);
  const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)pw);
  const float count = (float)max(gh * gw, 1);
  float* o = out + (size_t)r * C * ph * pw;
  const int bins = ph * pw;
  if (gh <= 0 || gw <= 0 || b < 0 || b >= N) {  // empty box (or a batch index outside the input): zeros
    for (int e = threadIdx.x; e < C * bins; e += blockDim.x) o[e] = 0.f;
    return;
  }
  const bool tabulated = (long long)ph * gh <= kTapCap && (long long)pw * gw <= kTapCap;
  if (tabulated) {
    for (int e = threadIdx.x; e < ph * gh; e += blockDim.x) {
      const int p = e / gh, iy = e - p * gh;
      s_ty[e] = make_tap(y0 + (float)p * bin_h + ((float)iy + .5f) * bin_h / (float)gh, H);
    }
    for (int e = threadIdx.x; e < pw * gw; e += blockDim.x) {
      const int p = e / gw, ix = e - p * gw;
      s_tx[e] = make_tap(x0 + (float)p * bin_w + ((float)ix + .5f) * bin_w / (float)gw, W);
    }
    __syncthreads();
  }
  const float* plane0 = input + (size_t)b * C * H * W;
  for (int e = threadIdx.x; e < C * bins; e += blockDim.x) {
    const int c = e / bins, bin = e - c * bins;
    const int py = bin / pw, px = bin - py * pw;
    const float* plane = plane0 + (size_t)c * H * W;
    float acc = 0.f;
    for (int iy = 0; iy < gh; ++iy) {
      const Tap ty = tabulated ? s_ty[py * gh + iy]
                               : make_tap(y0 + (float)py * bin_h + ((float)iy + .5f) * bin_h / (float)gh, H);
      if (ty.lo < 0) continue;
      const float* row_lo = plane + (size_t)ty.lo * W;
      const float* row_hi = plane + (size_t)ty.hi * W;
      for (int ix = 0; ix < gw; ++ix) {
        const Tap tx = tabulated ? s_tx[px * gw + ix]
                                 : make_tap(x0 + (float)px * bin_w + ((float)ix + .5f) * bin_w / (float)gw, W);
        if (tx.lo < 0) continue;
        // same association as the reference: w1*v1 + w2*v2 + w3*v3 + w4*v4, weights formed as products first
        const float w1 = ty.wl * tx.wl, w2 = ty.wl * tx.wh, w3 = ty.wh * tx.wl, w4 = ty.wh * tx.wh;
        acc += w1 * __ldg(row_lo + tx.lo) + w2 * __ldg(row_lo + tx.hi) + w3 * __ldg(row_hi + tx.lo) + w4 * __ldg(row_hi + tx.hi);
      }
    }
    o[e] = acc / count;
  }
}

}  // namespace
}  // namespace pe

extern "C" PE_API size_t pe_batched_nms_workspace_bytes(int n) {
  if (n <= 0) return 256;
  return pe::nms_layout(n, nullptr, nullptr);
}

extern "C" PE_API int pe_batched_nms_max_boxes(void) { return 65536; }

extern "C" PE_API int pe_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int n, float iou_thr, int mode,
                                     int64_t* keep, int32_t* n_keep, void* workspace, size_t workspace_bytes, void* stream) {
  if (n < 0 || !n_keep || (mode != 0 && mode != 1)) return PE_ERR_INVALID_ARGUMENT;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n == 0) {
    PE_CUDA_CHECK(cudaMemsetAsync(n_keep, 0, sizeof(int32_t), st));
    return PE_OK;
  }
  if (!boxes || !scores || !keep || !workspace) return PE_ERR_INVALID_ARGUMENT;
  if (reinterpret_cast<uintptr_t>(boxes) & 15) return PE_ERR_INVALID_ARGUMENT;
  if (n > pe_batched_nms_max_boxes()) return PE_ERR_UNSUPPORTED;
  if (workspace_bytes < pe_batched_nms_workspace_bytes(n)) return PE_ERR_WORKSPACE_TOO_SMALL;
  pe::NmsWorkspace ws;
  pe::nms_layout(n, reinterpret_cast<unsigned char*>(workspace), &ws);
  const int words = (n + 31) / 32;
  const int blocks = (n + pe::kNmsBlock - 1) / pe::kNmsBlock;
  PE_CUDA_CHECK(cudaMemsetAsync(ws.max_key, 0, sizeof(unsigned), st));
  if (mode == 0 && idxs) {
    const int g = blocks < 4 * pe::sm_count() ? blocks : 4 * pe::sm_count();
    pe::nms_max_coord_kernel<<<g, pe::kNmsBlock, 0, st>>>(reinterpret_cast<const float4*>(boxes), n, ws.max_key);
    PE_LAUNCH_CHECK();
  }
  pe::nms_rank_kernel<<<blocks, pe::kNmsBlock, 0, st>>>(scores, n, ws.order);
  PE_LAUNCH_CHECK();
  pe::nms_gather_kernel<<<blocks, pe::kNmsBlock, 0, st>>>(reinterpret_cast<const float4*>(boxes),
                                                          reinterpret_cast<const long long*>(idxs), ws.order, n, mode, ws.max_key,
                                                          ws.sorted_boxes, ws.sorted_cls);
  PE_LAUNCH_CHECK();
  pe::nms_bitmask_kernel<<<dim3(words, blocks), pe::kNmsBlock, 0, st>>>(ws.sorted_boxes, ws.sorted_cls, n, words, iou_thr,
                                                                        (mode == 1 && idxs) ? 1 : 0, ws.mask);
  PE_LAUNCH_CHECK();
  pe::nms_keep_kernel<<<1, 1024, words * sizeof(unsigned), st>>>(ws.mask, ws.order, n, words, reinterpret_cast<long long*>(keep),
                                                                  n_keep);
  PE_LAUNCH_CHECK();
  return PE_OK;
}

extern "C" PE_API int pe_roi_align_forward(const float* input, int N, int C, int H, int W, const float* rois, int num_rois,
                                           float spatial_scale, int pooled_h, int pooled_w, int sampling_ratio, int aligned,
                                           float* out, void* stream) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || num_rois < 0 || pooled_h <= 0 || pooled_w <= 0) return PE_ERR_INVALID_ARGUMENT;
  if (num_rois == 0 || C == 0) return PE_OK;
  if (!rois || !out || (!input && N > 0)) return PE_ERR_INVALID_ARGUMENT;
  if (H == 0 || W == 0) return PE_ERR_INVALID_ARGUMENT;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  pe::roi_align_nchw_kernel<<<num_rois, pe::kRoiThreads, 0, st>>>(input, N, C, H, W, rois, spatial_scale, pooled_h, pooled_w,
                                                                   sampling_ratio, aligned ? 1 : 0, out);
  PE_LAUNCH_CHECK();
  return PE_OK;
}
