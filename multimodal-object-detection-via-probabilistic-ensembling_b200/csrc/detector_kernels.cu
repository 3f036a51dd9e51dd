// Non-GEMM kernels of the detector path (sm_100a): pre-processing + stem im2col, max-pool, RPN top-k /
// decode / NMS / merge, multi-level ROIAlign, Fast R-CNN head post-processing, and the packer that turns the
// per-model detections into pe_fuse_batch's layout.  All batched, no host synchronisation.
//
// Reference semantics restated (paths relative to the reference checkout):
//   stem_im2col      detectron2/modeling/meta_arch/rcnn.py:269-286 (normalise, zero pad to /32) feeding
//                    modeling/backbone/resnet.py:369-384 (7x7/2 conv as a GEMM over K = 49*C)
//   maxpool3x3s2     resnet.py:383 (F.max_pool2d(k=3, s=2, p=1))
//   subsample2       modeling/backbone/fpn.py:166-178 (LastLevelMaxPool = max_pool2d(k=1, s=2))
//   rpn_topk         modeling/proposal_generator/rpn_outputs.py:95-124,409-451 + anchor_generator.py:130-199 +
//                    box_regression.py:78-115 + clip / non-empty test of rpn_outputs.py:131-146
//   rpn_nms/merge    rpn_outputs.py:147-160 via detectron2/layers/nms.py:9-26 (per-level NMS, IoU 0.7, top-1000)
//   roi_align        modeling/poolers.py:13-81,180-235 + layers/csrc/ROIAlign/ROIAlign_cuda.cu:65-139
//                    (aligned=True, sampling_ratio=0)
//   head_post        modeling/roi_heads/fast_rcnn.py:86-147,345-360,417-452 + modeling/postprocessing.py:8-52
#include <cuda_bf16.h>
#include <stdlib.h>
#include <cuda_fp16.h>
#include <math.h>
#include "common.cuh"
#include "detector_kernels.cuh"

namespace pe {
namespace {

__device__ __forceinline__ float bflo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bfhi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
// Eight fp32 accumulators as four packed f32x2 registers: sm_100's FFMA2 (fma.rn.f32x2) does two IEEE fp32 FMAs per
// instruction, which halves the accumulate work of the gather kernels (ROIAlign is instruction-bound).
struct Acc8 {
  unsigned long long v[4];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = 0ull;
  }
};
__device__ __forceinline__ unsigned long long pack2f(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// acc += w * (8 bf16 channels of v)
__device__ __forceinline__ void acc_bf16x8(Acc8& a, unsigned long long w2, uint4 v) {
  a.v[0] = ffma2(pack2f(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u)), w2, a.v[0]);
  a.v[1] = ffma2(pack2f(__uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u)), w2, a.v[1]);
  a.v[2] = ffma2(pack2f(__uint_as_float(v.z << 16), __uint_as_float(v.z & 0xffff0000u)), w2, a.v[2]);
  a.v[3] = ffma2(pack2f(__uint_as_float(v.w << 16), __uint_as_float(v.w & 0xffff0000u)), w2, a.v[3]);
}
// acc += w * t
__device__ __forceinline__ void acc_axpy(Acc8& a, unsigned long long w2, const Acc8& t) {
#pragma unroll
  for (int i = 0; i < 4; ++i) a.v[i] = ffma2(t.v[i], w2, a.v[i]);
}
__device__ __forceinline__ uint32_t packbf2(unsigned long long v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
// (acc * scale) rounded to 8 bf16 channels
__device__ __forceinline__ uint4 acc_store_bf16(const Acc8& a, unsigned long long s2) {
  return make_uint4(packbf2(fmul2(a.v[0], s2)), packbf2(fmul2(a.v[1], s2)), packbf2(fmul2(a.v[2], s2)), packbf2(fmul2(a.v[3], s2)));
}
__device__ __forceinline__ uint32_t packbf(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ bool finite4(float4 b) { return isfinite(b.x) && isfinite(b.y) && isfinite(b.z) && isfinite(b.w); }

// One bilinear sample of DefaultPredictor's resize (half-pixel centres).  Written with explicit round-to-nearest
// intrinsics so that every kernel that resizes (stand-alone or fused into the stem staging) produces the same bits.
__device__ __forceinline__ float resize_sample(const unsigned char* __restrict__ im, int Hs, int Ws, int C, int c, int y, int x,
                                               float sy, float sx) {
  const float fy = __fsub_rn(__fmul_rn(__fadd_rn((float)y, 0.5f), sy), 0.5f);
  const float fx = __fsub_rn(__fmul_rn(__fadd_rn((float)x, 0.5f), sx), 0.5f);
  int y0 = (int)floorf(fy), x0 = (int)floorf(fx);
  float ly = __fsub_rn(fy, (float)y0), lx = __fsub_rn(fx, (float)x0);
  int y1 = y0 + 1, x1 = x0 + 1;
  if (y0 < 0) { y0 = 0; y1 = 0; ly = 0.f; }
  if (x0 < 0) { x0 = 0; x1 = 0; lx = 0.f; }
  if (y1 >= Hs) { y1 = Hs - 1; if (y0 >= Hs) y0 = Hs - 1; }
  if (x1 >= Ws) { x1 = Ws - 1; if (x0 >= Ws) x0 = Ws - 1; }
  const float v00 = im[((size_t)y0 * Ws + x0) * C + c], v01 = im[((size_t)y0 * Ws + x1) * C + c];
  const float v10 = im[((size_t)y1 * Ws + x0) * C + c], v11 = im[((size_t)y1 * Ws + x1) * C + c];
  const float hx = __fsub_rn(1.f, lx), hy = __fsub_rn(1.f, ly);
  const float top = __fadd_rn(__fmul_rn(hx, v00), __fmul_rn(lx, v01));
  const float bot = __fadd_rn(__fmul_rn(hx, v10), __fmul_rn(lx, v11));
  float v = __fadd_rn(__fmul_rn(hy, top), __fmul_rn(ly, bot));
  return v;
}

// Pillow's BILINEAR resize of 8-bit images, bit for bit (the path DefaultPredictor takes for 3-channel uint8 frames:
// data/transforms/transform.py:92-96 -> Image.resize).  Pillow (libImaging/Resample.c, ImagingResample with the
// bilinear filter: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc, ...Vertical_8bpc; the
// reference environment pins Pillow 9.2) filters in two passes - rows first, then columns - with double-precision
// triangle weights normalised to sum 1, converted to 22-bit fixed point, and a rounded uint8 intermediate image.
// The weights of one output coordinate are recomputed here in double with the library's exact operation order.
constexpr int kPilPrecisionBits = 32 - 8 - 2;

template <int KMAX>
struct PilTaps {
  int lo, n;
  int k[KMAX];
};

template <int KMAX>
__device__ __forceinline__ PilTaps<KMAX> pil_taps(int xx, int in_size, int out_size) {
  PilTaps<KMAX> t;
  const double scale = (double)((float)in_size - 0.f) / (double)out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;  // bilinear support 1.0 * filterscale
  const double center = __dadd_rn(0.0, __dmul_rn(__dadd_rn((double)xx, 0.5), scale));
  const double ss = __ddiv_rn(1.0, filterscale);
  int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  if (xmax > KMAX) xmax = KMAX;  // cannot happen for scale <= (KMAX - 1) / 2, which the launcher checks
  double w[KMAX];
  double ww = 0.0;
#pragma unroll
  for (int x = 0; x < KMAX; ++x) {
    double v = 0.0;
    if (x < xmax) {
      double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
      if (a < 0.0) a = -a;
      v = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
      ww = __dadd_rn(ww, v);
    }
    w[x] = v;
  }
#pragma unroll
  for (int x = 0; x < KMAX; ++x) {
    double v = w[x];
    if (x < xmax && ww != 0.0) v = __ddiv_rn(v, ww);
    t.k[x] = x < xmax ? (int)__dadd_rn(0.5, __dmul_rn(v, (double)(1 << kPilPrecisionBits))) : 0;
  }
  t.lo = xmin;
  t.n = xmax;
  return t;
}

__device__ __forceinline__ int pil_clip8(int v) {
  v >>= kPilPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// All C (<= 4) channels c0.. of one output pixel.
template <int KMAX>
__device__ __forceinline__ void pil_resize_pixel(const unsigned char* __restrict__ im, int Ws, int Ctot, int c0, int C,
                                                 const PilTaps<KMAX>& ty, const PilTaps<KMAX>& tx, int out[4]) {
  int acc[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) acc[c] = 1 << (kPilPrecisionBits - 1);
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    if (j < ty.n) {
      const unsigned char* rowp = im + ((size_t)(ty.lo + j) * Ws + tx.lo) * Ctot + c0;
      int h[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) h[c] = 1 << (kPilPrecisionBits - 1);
#pragma unroll
      for (int i = 0; i < KMAX; ++i) {
        if (i < tx.n) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < C) h[c] += (int)rowp[(size_t)i * Ctot + c] * tx.k[i];
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c] += pil_clip8(h[c]) * ty.k[j];
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c] = pil_clip8(acc[c]);
}

// ---------------------------------------------------------------------------------------------- stem im2col
// Stage 1: normalise (rcnn.py:269-286) into a zero-bordered fp16 HWC4 canvas [B, Hc+6, Wc+8, 4]: image pixel (y, x)
// sits at (y+3, x+3); the border supplies both the conv's 3-pixel zero padding and the /32 canvas padding.
__global__ void stem_canvas_kernel(const float* __restrict__ img, __half* __restrict__ canvas, int B, int Ctot, int c0, int C,
                                   int Hi, int Wi, int Hp, int Wp, StemNorm nrm) {
  const long long total = (long long)B * Hp * Wp;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int xp = (int)(t % Wp);
    long long r = t / Wp;
    const int yp = (int)(r % Hp), b = (int)(r / Hp);
    const int y = yp - 3, x = xp - 3;
    __align__(8) __half v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float val = 0.f;
      if (c < C && y >= 0 && y < Hi && x >= 0 && x < Wi)
        val = __fdiv_rn(__ldg(img + (((size_t)b * Ctot + c0 + c) * Hi + y) * Wi + x) - nrm.mean[c], nrm.std[c]);
      v[c] = __float2half_rn(val);
    }
    *reinterpret_cast<uint2*>(canvas + (size_t)t * 4) = *reinterpret_cast<const uint2*>(v);
  }
}

// Stage 1': the same canvas straight from raw uint8 HWC frames: bilinear resize (resize_frames_kernel's arithmetic, i.e.
// DefaultPredictor's ResizeShortestEdge) + normalisation fused, so the float32 network input never exists in HBM.
// Upscales need at most 3 taps per axis: the engine tabulates them once per forward (row table [Hi] then column table
// [Wi], int4 = first source index, three 22-bit weights) so that the per-pixel work is integer only.
__global__ void pil_taps_table_kernel(int4* __restrict__ table, int Hs, int Ws, int Hi, int Wi, __half* __restrict__ lut, int C,
                                      StemNorm nrm) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (lut && t < 4 * 256) {  // normalisation table of the tiled staging kernel: lut[c][v] = fp16((v - mean_c) / std_c), 0 for c >= C
    const int c = t >> 8;
    lut[t] = __float2half_rn(c < C ? __fdiv_rn((float)(t & 255) - nrm.mean[c], nrm.std[c]) : 0.f);
  }
  if (t >= Hi + Wi) return;
  const PilTaps<3> tp = t < Hi ? pil_taps<3>(t, Hs, Hi) : pil_taps<3>(t - Hi, Ws, Wi);
  table[t] = make_int4(tp.lo | (tp.n << 24), tp.k[0], tp.k[1], tp.k[2]);
}

__device__ __forceinline__ PilTaps<3> pil_taps_from_table(int4 e) {
  PilTaps<3> t;
  t.lo = e.x & 0xffffff; t.n = e.x >> 24; t.k[0] = e.y; t.k[1] = e.z; t.k[2] = e.w;
  return t;
}

template <int KMAX>  // KMAX > 0: Pillow-exact uint8 resize (round_u8), KMAX = 0: float32 bilinear
__global__ void stem_canvas_u8_kernel(const unsigned char* __restrict__ frames, __half* __restrict__ canvas, int B, int Ctot, int c0,
                                      int C, int Hs, int Ws, int Hi, int Wi, int Hp, int Wp, StemNorm nrm,
                                      const int4* __restrict__ taps) {
  const long long total = (long long)B * Hp * Wp;
  const float sy = (float)Hs / (float)Hi, sx = (float)Ws / (float)Wi;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int xp = (int)(t % Wp);
    long long r = t / Wp;
    const int yp = (int)(r % Hp), b = (int)(r / Hp);
    const int y = yp - 3, x = xp - 3;
    __align__(8) __half v[4];
    const bool inside = y >= 0 && y < Hi && x >= 0 && x < Wi;
    const unsigned char* im = frames + (size_t)b * Hs * Ws * Ctot;
    if (KMAX > 0) {
      int px[4] = {0, 0, 0, 0};
      if (inside) {
        constexpr int KM = KMAX > 0 ? KMAX : 1;
        if (KMAX == 3 && taps) {
          const PilTaps<3> ty = pil_taps_from_table(__ldg(taps + y)), tx = pil_taps_from_table(__ldg(taps + Hi + x));
          pil_resize_pixel<3>(im, Ws, Ctot, c0, C, ty, tx, px);
        } else {
          const PilTaps<KM> ty = pil_taps<KM>(y, Hs, Hi), tx = pil_taps<KM>(x, Ws, Wi);
          pil_resize_pixel<KM>(im, Ws, Ctot, c0, C, ty, tx, px);
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        v[c] = __float2half_rn((c < C && inside) ? __fdiv_rn((float)px[c] - nrm.mean[c], nrm.std[c]) : 0.f);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float val = 0.f;
        if (c < C && inside) val = __fdiv_rn(resize_sample(im, Hs, Ws, Ctot, c0 + c, y, x, sy, sx) - nrm.mean[c], nrm.std[c]);
        v[c] = __float2half_rn(val);
      }
    }
    *reinterpret_cast<uint2*>(canvas + (size_t)t * 4) = *reinterpret_cast<const uint2*>(v);
  }
}

// Stage 1'', upscales with Pillow rounding (the FLIR / KAIST case, 512x640 -> 800x1000): Pillow's own two-pass order, tiled.
// A block owns kCanvasRows canvas rows of one image.  Pass 1 filters the source rows those canvas rows touch horizontally
// (<= kCanvasRows + 3 of them for any upscale; tap windows are monotone in y) into shared memory as rounded uint8 pixels,
// 4 channels to a word - exactly Pillow's intermediate image; pass 2 filters vertically from shared memory and maps the
// 256 possible pixel values of each channel through a (v - mean) / std table.  Same arithmetic as stem_canvas_u8_kernel<3>
// (bit-identical canvas), but every horizontally filtered pixel is computed once per block instead of once per output
// pixel and row tap, and the per-pixel 64-bit index arithmetic and IEEE divisions are gone: 203 -> see profiles/README.md.
constexpr int kCanvasRows = 8;
constexpr int kCanvasSpan = kCanvasRows + 4;

__global__ void __launch_bounds__(256) stem_canvas_u8_tiled_kernel(const unsigned char* __restrict__ frames, __half* __restrict__ canvas,
                                                                    int Ctot, int c0, int C, int Hs, int Ws, int Hi, int Wi, int Hp, int Wp,
                                                                    const int4* __restrict__ taps, const __half* __restrict__ lut) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  uint32_t* s_h = reinterpret_cast<uint32_t*>(s_dyn);                         // [kCanvasSpan][Wi] filtered rows
  unsigned char* s_raw = s_dyn + (size_t)kCanvasSpan * Wi * sizeof(uint32_t);  // [kCanvasSpan][Ws * Ctot] source rows
  __shared__ __half s_lut[4][256];
  const int tid = threadIdx.x, b = blockIdx.y, yp0 = blockIdx.x * kCanvasRows;
  for (int i = tid; i < 4 * 256 / 2; i += 256)  // (v - mean) / std per channel, tabulated once per forward by pil_taps_table_kernel
    reinterpret_cast<uint32_t*>(&s_lut[0][0])[i] = __ldg(reinterpret_cast<const uint32_t*>(lut) + i);
  const int ya = max(yp0 - 3, 0), yb = min(yp0 + kCanvasRows - 1 - 3, Hi - 1);  // image rows under this block's canvas rows
  int s0 = 0, ns = 0;
  if (ya <= yb) {
    const int4 ea = __ldg(taps + ya), eb = __ldg(taps + yb);
    s0 = ea.x & 0xffffff;
    ns = min((eb.x & 0xffffff) + (eb.x >> 24) - s0, kCanvasSpan);
  }
  // the source rows of this block are one contiguous byte range: staged with 16-byte loads when aligned, so the 9 byte reads
  // per filtered pixel hit shared memory (from global memory they were 40 % of the kernel's stall samples)
  const int row_bytes = Ws * Ctot;
  const unsigned char* src = frames + ((size_t)b * Hs + s0) * row_bytes;
  const int nbytes = ns * row_bytes;
  if (((reinterpret_cast<uintptr_t>(src) | (uintptr_t)nbytes) & 15) == 0) {
    for (int i = tid; i < nbytes / 16; i += 256) reinterpret_cast<uint4*>(s_raw)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
  } else {
    for (int i = tid; i < nbytes; i += 256) s_raw[i] = __ldg(src + i);
  }
  __syncthreads();
  for (int x = tid; x < Wi; x += 256) {
    const PilTaps<3> tx = pil_taps_from_table(__ldg(taps + Hi + x));
    const unsigned char* rowp = s_raw + tx.lo * Ctot + c0;
    for (int r = 0; r < ns; ++r, rowp += row_bytes) {
      int h[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) h[c] = 1 << (kPilPrecisionBits - 1);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (i < tx.n) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < C) h[c] += (int)rowp[i * Ctot + c] * tx.k[i];
        }
      }
      s_h[r * Wi + x] = (uint32_t)pil_clip8(h[0]) | ((uint32_t)pil_clip8(h[1]) << 8) | ((uint32_t)pil_clip8(h[2]) << 16) |
                        ((uint32_t)pil_clip8(h[3]) << 24);
    }
  }
  __syncthreads();
  for (int ry = 0; ry < kCanvasRows; ++ry) {
    const int yp = yp0 + ry, y = yp - 3;
    if (yp >= Hp) break;
    uint2* dst = reinterpret_cast<uint2*>(canvas + ((size_t)b * Hp + yp) * Wp * 4);
    const bool row_in = y >= 0 && y < Hi;
    PilTaps<3> ty;
    ty.lo = 0; ty.n = 0; ty.k[0] = ty.k[1] = ty.k[2] = 0;
    if (row_in) ty = pil_taps_from_table(__ldg(taps + y));
    const uint32_t* hrow = s_h + (ty.lo - s0) * Wi;
    for (int xp = tid; xp < Wp; xp += 256) {
      const int x = xp - 3;
      uint2 o = make_uint2(0u, 0u);
      if (row_in && x >= 0 && x < Wi) {
        int acc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] = 1 << (kPilPrecisionBits - 1);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (j < ty.n) {
            const uint32_t w = hrow[j * Wi + x];
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c] += (int)((w >> (8 * c)) & 0xffu) * ty.k[j];
          }
        }
        const unsigned short v0 = __half_as_ushort(s_lut[0][pil_clip8(acc[0])]), v1 = __half_as_ushort(s_lut[1][pil_clip8(acc[1])]);
        const unsigned short v2 = __half_as_ushort(s_lut[2][pil_clip8(acc[2])]), v3 = __half_as_ushort(s_lut[3][pil_clip8(acc[3])]);
        o = make_uint2((uint32_t)v0 | ((uint32_t)v1 << 16), (uint32_t)v2 | ((uint32_t)v3 << 16));
      }
      dst[xp] = o;
    }
  }
}

// Stage 2: A[pixel][kh*32 + kw*4 + c] = canvas[2*ho + kh][2*wo + kw][c] for kh < 7, kw < 8 (kw = 7 and c >= C meet
// zero weights): every (pixel, kh) is one contiguous 64-byte run of the canvas -> four 16-byte copies.
__global__ void stem_im2col_kernel(const __half* __restrict__ canvas, __half* __restrict__ A, int B, int Ho, int Wo, int Hp, int Wp) {
  const long long total = (long long)B * Ho * Wo * 28;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int part = (int)(t % 28);
    long long pix = t / 28;
    const int kh = part >> 2, quad = part & 3;
    const int wo = (int)(pix % Wo);
    const long long r = pix / Wo;
    const int ho = (int)(r % Ho), b = (int)(r / Ho);
    const uint4* src = reinterpret_cast<const uint4*>(canvas + (((size_t)b * Hp + 2 * ho + kh) * Wp + 2 * wo) * 4) + quad;
    reinterpret_cast<uint4*>(A)[t] = __ldg(src);
  }
}

// ---------------------------------------------------------------------------------------------- frame resize
// DefaultPredictor's ResizeShortestEdge (engine/defaults.py:186-190, data/transforms/transform.py:81-99): bilinear with
// half-pixel centres (cv2.INTER_LINEAR / PIL BILINEAR geometry), uint8 HWC frames -> float32 CHW network input.
// round_u8 reproduces the uint8 output quantisation of the PIL path used for 3-channel uint8 images.
template <int KMAX>
__global__ void resize_frames_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, int B, int C, int Hs, int Ws,
                                     int Hd, int Wd) {
  const float sy = (float)Hs / (float)Hd, sx = (float)Ws / (float)Wd;
  if (KMAX > 0) {  // one thread per output pixel, channels in groups of 4
    constexpr int KM = KMAX > 0 ? KMAX : 1;
    const long long total = (long long)B * Hd * Wd;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
      const int x = (int)(t % Wd);
      const long long r = t / Wd;
      const int y = (int)(r % Hd), b = (int)(r / Hd);
      const PilTaps<KM> ty = pil_taps<KM>(y, Hs, Hd), tx = pil_taps<KM>(x, Ws, Wd);
      for (int c0 = 0; c0 < C; c0 += 4) {
        int px[4];
        const int cn = min(4, C - c0);
        pil_resize_pixel<KM>(src + (size_t)b * Hs * Ws * C, Ws, C, c0, cn, ty, tx, px);
        for (int c = 0; c < cn; ++c) dst[(((size_t)b * C + c0 + c) * Hd + y) * Wd + x] = (float)px[c];
      }
    }
    return;
  }
  const long long total = (long long)B * C * Hd * Wd;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t % Wd);
    long long r = t / Wd;
    const int y = (int)(r % Hd);
    r /= Hd;
    const int c = (int)(r % C), b = (int)(r / C);
    dst[t] = resize_sample(src + (size_t)b * Hs * Ws * C, Hs, Ws, C, c, y, x, sy, sx);
  }
}

// ---------------------------------------------------------------------------------------------- pooling
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}

// reverse: walk the outputs from the last to the first, so the kernel starts on the part of x its producer (the stem conv, which
// walks forward) wrote last and that is still in L2; the consumer of y then walks forward for the same reason
__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int H, int W,
                                    int C, int Ho, int Wo, int reverse) {
  const int cg = C >> 3;
  const long long total = (long long)B * Ho * Wo * cg;
  for (long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; t0 < total; t0 += (long long)gridDim.x * blockDim.x) {
    const long long t = reverse ? total - 1 - t0 : t0;
    const int g = (int)(t % cg);
    long long pix = t / cg;
    const int wo = (int)(pix % Wo);
    pix /= Wo;
    const int ho = (int)(pix % Ho), b = (int)(pix / Ho);
    // max is exact in any precision: compare the packed bf16 pairs directly (HMNMX2.BF16) instead of unpacking to fp32
    // (the kernel was instruction-bound: 400 instructions per thread, most of them unpack + fmaxf)
    uint4 m = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);  // -inf
    const uint4* xb = reinterpret_cast<const uint4*>(x + (size_t)b * H * W * C) + g;
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = 2 * ho - 1 + dy;
      if (yy < 0 || yy >= H) continue;
      const uint4* xr = xb + (size_t)yy * W * cg;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = 2 * wo - 1 + dx;
        if (xx < 0 || xx >= W) continue;
        const uint4 r = __ldg(xr + (size_t)xx * cg);
        m.x = bf16x2_max(m.x, r.x); m.y = bf16x2_max(m.y, r.y); m.z = bf16x2_max(m.z, r.z); m.w = bf16x2_max(m.w, r.w);
      }
    }
    reinterpret_cast<uint4*>(y)[t] = m;
  }
}

__global__ void subsample2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int H, int W, int C,
                                  int Ho, int Wo) {
  const int cg = C >> 3;
  const long long total = (long long)B * Ho * Wo * cg;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % cg);
    long long pix = t / cg;
    const int wo = (int)(pix % Wo);
    pix /= Wo;
    const int ho = (int)(pix % Ho), b = (int)(pix / Ho);
    reinterpret_cast<uint4*>(y)[t] = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)b * H + 2 * ho) * W + 2 * wo) * C) + g);
  }
}

// concat two NHWC tensors along channels (middle fusion, rcnn.py:245-247)
__global__ void concat_channels_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                       __nv_bfloat16* __restrict__ y, long long pixels, int C) {
  const int cg = C >> 3;
  const long long total = pixels * cg * 2;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % (2 * cg));
    const long long pix = t / (2 * cg);
    const uint4* src = g < cg ? reinterpret_cast<const uint4*>(a + pix * C) + g : reinterpret_cast<const uint4*>(b + pix * C) + (g - cg);
    reinterpret_cast<uint4*>(y)[t] = __ldg(src);
  }
}

// ---------------------------------------------------------------------------------------------- RPN top-k + decode
__device__ __forceinline__ uint32_t sort_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

constexpr int kTopkThreads = 1024;
constexpr int kTopkCap = kTopkSlots;  // pre_nms_topk <= 1024

// Finds, scanning bins from the highest down, the bin holding the kth largest element.
// hist has nbins (<= 2048) entries; returns through smem result[0]=bin, result[1]=count in higher bins.
__device__ void find_kth_bin(const unsigned* hist, int nbins, unsigned kth, unsigned* result, unsigned* warp_tot) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // each thread owns 2 descending-ordered bins
  const int i0 = nbins - 1 - 2 * tid, i1 = i0 - 1;
  const unsigned h0 = i0 >= 0 ? hist[i0] : 0u, h1 = i1 >= 0 ? hist[i1] : 0u;
  unsigned s = h0 + h1;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned o = __shfl_up_sync(kFullMask, s, d);
    if (lane >= d) s += o;
  }
  if (lane == 31) warp_tot[wid] = s;
  __syncthreads();
  if (wid == 0) {
    unsigned w = warp_tot[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned o = __shfl_up_sync(kFullMask, w, d);
      if (lane >= d) w += o;
    }
    warp_tot[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  const unsigned before = (wid ? warp_tot[wid - 1] : 0u) + s - (h0 + h1);  // count in bins above i0
  if (before < kth && kth <= before + h0) { result[0] = (unsigned)i0; result[1] = before; }
  else if (before + h0 < kth && kth <= before + h0 + h1) { result[0] = (unsigned)i1; result[1] = before + h0; }
  __syncthreads();
}

constexpr int kTopkCandCap = 4096;  // elements of the threshold bin (+ everything above it) kept in shared memory
constexpr size_t kTopkSmemBytes = 2048 * sizeof(unsigned) + (size_t)kTopkCandCap * 8 + (size_t)kTopkCap * 8;

__global__ void __launch_bounds__(kTopkThreads) rpn_topk_kernel(const RpnLevels lv, int pre_topk, float img_h, float img_w,
                                                                float4* __restrict__ cand_box, float* __restrict__ cand_score,
                                                                unsigned char* __restrict__ cand_valid, int* __restrict__ cand_count) {
  extern __shared__ __align__(16) unsigned char topk_smem[];
  unsigned* hist = reinterpret_cast<unsigned*>(topk_smem);
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(hist + 2048);
  unsigned long long* sel = cand + kTopkCandCap;
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned res[2];
  __shared__ unsigned sel_n, eq_run, cand_n;
  const int level = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int H = lv.H[level], W = lv.W[level], A = 3;
  const int n = H * W * A;
  const int k = pre_topk < n ? pre_topk : n;
  const float* base = lv.out[level] + (size_t)b * H * W * kRpnOutC;
  // The select passes read one float4 per pixel (3 logits | pad): from the dense logit plane the RPN-head kernel writes next to its
  // 64-byte records when there is one (4x fewer sectors: a p2 block streams 51 200 pixels twice and is bound by its SM's L2
  // bandwidth), else from the first 16 bytes of the records.
  const float* src = lv.logit[level] ? lv.logit[level] + (size_t)b * H * W * 4 : base;
  const int sstride = lv.logit[level] ? 4 : kRpnOutC;
  const int npix = H * W;
  auto logit = [&](int e) { return __ldg(src + (size_t)(e / A) * sstride + (e % A)); };
  auto logit3 = [&](int p) { return __ldg(reinterpret_cast<const float4*>(src + (size_t)p * sstride)); };
  auto pack = [](uint32_t key, int e) { return ((unsigned long long)key << 32) | (unsigned)(0xffffffffu - (unsigned)e); };

  // ---- pass 0 over global memory: histogram of the top 11 key bits -> bin of the k-th largest logit
  for (int i = tid; i < 2048; i += blockDim.x) hist[i] = 0;
  if (tid == 0) { sel_n = 0; eq_run = 0; cand_n = 0; }
  __syncthreads();
  for (int p = tid; p < npix; p += blockDim.x) {
    const float4 v = logit3(p);
    atomicAdd(&hist[sort_key(v.x) >> 21], 1u);
    atomicAdd(&hist[sort_key(v.y) >> 21], 1u);
    atomicAdd(&hist[sort_key(v.z) >> 21], 1u);
  }
  __syncthreads();
  find_kth_bin(hist, 2048, (unsigned)k, res, warp_tot);
  const unsigned bin0 = res[0], above0 = res[1];
  const unsigned in_bin0 = hist[bin0];
  __syncthreads();
  uint32_t T;
  unsigned need_eq, eq_total;
  const bool small = above0 + in_bin0 <= (unsigned)kTopkCandCap;  // block-uniform
  if (small) {
    // ---- pass 1 over global memory: keep everything from the threshold bin upwards in shared memory
    for (int p = tid; p < npix; p += blockDim.x) {
      const float4 v = logit3(p);
      const uint32_t k0 = sort_key(v.x), k1 = sort_key(v.y), k2 = sort_key(v.z);
      if ((k0 >> 21) >= bin0) cand[atomicAdd(&cand_n, 1u)] = pack(k0, p * A);
      if ((k1 >> 21) >= bin0) cand[atomicAdd(&cand_n, 1u)] = pack(k1, p * A + 1);
      if ((k2 >> 21) >= bin0) cand[atomicAdd(&cand_n, 1u)] = pack(k2, p * A + 2);
    }
    __syncthreads();
    const int nc = (int)cand_n;
    uint32_t prefix = bin0 << 21, prefix_mask = 0x7ffu << 21;
    unsigned kth = (unsigned)k - above0;
    const int shifts[2] = {10, 0};
    const int widths[2] = {11, 10};
    for (int pass = 0; pass < 2; ++pass) {  // remaining radix passes run on the shared-memory list
      for (int i = tid; i < 2048; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      const int sh = shifts[pass], nb = 1 << widths[pass];
      for (int i = tid; i < nc; i += blockDim.x) {
        const uint32_t key = (uint32_t)(cand[i] >> 32);
        if ((key & prefix_mask) == prefix) atomicAdd(&hist[(key >> sh) & (nb - 1)], 1u);
      }
      __syncthreads();
      find_kth_bin(hist, nb, kth, res, warp_tot);
      prefix |= res[0] << sh;
      prefix_mask |= (uint32_t)(nb - 1) << sh;
      kth -= res[1];
      __syncthreads();
    }
    T = prefix;
    need_eq = kth;
    eq_total = hist[T & 1023u];
    for (int i = tid; i < nc; i += blockDim.x) {
      const unsigned long long c = cand[i];
      const uint32_t key = (uint32_t)(c >> 32);
      bool take = key > T;
      if (key == T) {
        if (eq_total == need_eq) take = true;
        else {  // ties at the threshold: the need_eq lowest indices (= largest packed low words) win
          unsigned ahead = 0;
          for (int j = 0; j < nc; ++j) ahead += ((uint32_t)(cand[j] >> 32) == T) && (cand[j] > c);
          take = ahead < need_eq;
        }
      }
      if (take) sel[atomicAdd(&sel_n, 1u)] = c;
    }
  } else {
    // ---- many logits share the threshold bin: stay on global memory for the remaining passes
    uint32_t prefix = bin0 << 21, prefix_mask = 0x7ffu << 21;
    unsigned kth = (unsigned)k - above0;
    const int shifts[2] = {10, 0};
    const int widths[2] = {11, 10};
    for (int pass = 0; pass < 2; ++pass) {
      for (int i = tid; i < 2048; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      const int sh = shifts[pass], nb = 1 << widths[pass];
      for (int e = tid; e < n; e += blockDim.x) {
        const uint32_t key = sort_key(logit(e));
        if ((key & prefix_mask) == prefix) atomicAdd(&hist[(key >> sh) & (nb - 1)], 1u);
      }
      __syncthreads();
      find_kth_bin(hist, nb, kth, res, warp_tot);
      prefix |= res[0] << sh;
      prefix_mask |= (uint32_t)(nb - 1) << sh;
      kth -= res[1];
      __syncthreads();
    }
    T = prefix;       // key of the kth largest element
    need_eq = kth;    // how many elements equal to T are selected (lowest index first)
    eq_total = hist[T & 1023u];
    if (eq_total == need_eq) {
      for (int e = tid; e < n; e += blockDim.x) {
        const uint32_t key = sort_key(logit(e));
        if (key >= T) sel[atomicAdd(&sel_n, 1u)] = pack(key, e);
      }
    } else {
      // ties at the threshold: take them in index order with a block-wide running count
      for (int e0 = 0; e0 < n; e0 += blockDim.x) {
        const int e = e0 + tid;
        uint32_t key = 0;
        bool gt = false, eq = false;
        if (e < n) { key = sort_key(logit(e)); gt = key > T; eq = key == T; }
        const unsigned bal = __ballot_sync(kFullMask, eq);
        const unsigned before_lane = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        unsigned before = eq_run;
        for (int w2 = 0; w2 < wid; ++w2) before += warp_tot[w2];
        if (gt || (eq && before + before_lane < need_eq)) sel[atomicAdd(&sel_n, 1u)] = pack(key, e);
        __syncthreads();
        if (tid == 0) { unsigned tot = 0; for (int w2 = 0; w2 < 32; ++w2) tot += warp_tot[w2]; eq_run += tot; }
        __syncthreads();
      }
    }
  }
  __syncthreads();
  for (int i = k + tid; i < kTopkCap; i += blockDim.x) sel[i] = 0ull;
  __syncthreads();
  // bitonic sort, descending
  for (int size = 2; size <= kTopkCap; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int i = tid;
      const int j = i ^ stride;
      if (j > i) {
        const unsigned long long a = sel[i], c = sel[j];
        const bool desc = (i & size) == 0;
        if (desc ? a < c : a > c) { sel[i] = c; sel[j] = a; }
      }
      __syncthreads();
    }
  }
  // decode the selected anchors
  const size_t obase = ((size_t)b * kRpnLevels + level) * kTopkCap;
  if (tid < k) {
    const unsigned long long s = sel[tid];
    const int e = (int)(0xffffffffu - (unsigned)(s & 0xffffffffu));
    const float score = key_to_float((uint32_t)(s >> 32));
    const int pix = e / A, an = e - pix * A;
    const int y = pix / W, x = pix - y * W;
    const float sx = (float)(x * lv.stride[level]), sy = (float)(y * lv.stride[level]);
    const float ax1 = __fadd_rn(sx, lv.anchor[level][an][0]), ay1 = __fadd_rn(sy, lv.anchor[level][an][1]);
    const float ax2 = __fadd_rn(sx, lv.anchor[level][an][2]), ay2 = __fadd_rn(sy, lv.anchor[level][an][3]);
    const float4 d = __ldg(reinterpret_cast<const float4*>(base + (size_t)pix * kRpnOutC + 4 + an * 4));
    const float wdt = __fsub_rn(ax2, ax1), hgt = __fsub_rn(ay2, ay1);
    const float cx = __fadd_rn(ax1, __fmul_rn(0.5f, wdt)), cy = __fadd_rn(ay1, __fmul_rn(0.5f, hgt));
    const float dw = fminf(d.z, kScaleClamp), dh = fminf(d.w, kScaleClamp);
    const float pcx = __fadd_rn(__fmul_rn(d.x, wdt), cx), pcy = __fadd_rn(__fmul_rn(d.y, hgt), cy);
    const float pw = __fmul_rn(expf(dw), wdt), ph = __fmul_rn(expf(dh), hgt);
    float4 bx = make_float4(__fsub_rn(pcx, __fmul_rn(0.5f, pw)), __fsub_rn(pcy, __fmul_rn(0.5f, ph)),
                            __fadd_rn(pcx, __fmul_rn(0.5f, pw)), __fadd_rn(pcy, __fmul_rn(0.5f, ph)));
    const bool fin = finite4(bx) && isfinite(score);
    bx.x = fminf(fmaxf(bx.x, 0.f), img_w); bx.z = fminf(fmaxf(bx.z, 0.f), img_w);
    bx.y = fminf(fmaxf(bx.y, 0.f), img_h); bx.w = fminf(fmaxf(bx.w, 0.f), img_h);
    const bool ok = fin && (bx.z - bx.x > 0.f) && (bx.w - bx.y > 0.f);
    cand_box[obase + tid] = bx;
    cand_score[obase + tid] = score;
    cand_valid[obase + tid] = ok ? 1 : 0;
  }
  if (tid == 0) cand_count[b * kRpnLevels + level] = k;
}

// ---------------------------------------------------------------------------------------------- NMS (bitmask)
__device__ __forceinline__ bool iou_gt(float4 a, float aa, float4 b, float ab, float thr) {
  const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
  const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
  const float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter)) > thr;
}

constexpr int kNmsThreads = 256;
constexpr int kNmsRowsPerBlock = 64;

// "inter / (area_a + area_b - inter) > thr" with torchvision's float32 rounding: decided without the division
// unless the margin is within float32 round-off of the threshold.
__device__ __forceinline__ bool iou_gt_fast(float4 a, float aa, float4 b, float ab, float thr) {
  const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
  const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
  const float inter = __fmul_rn(w, h);
  const float uni = __fsub_rn(__fadd_rn(aa, ab), inter);
  const float d = inter - thr * uni;
  if (fabsf(d) > 4e-7f * fabsf(uni) && uni > 0.f) return d > 0.f;
  return __fdiv_rn(inter, uni) > thr;
}

// Suppression bitmask of score-sorted segments of <= 1024 boxes: block (seg, slice) fills rows
// [slice*64, slice*64+64) of mask[seg][1024][32] (word w of row i: boxes j = 32w.. with j > i, iou(i, j) > thr).
__global__ void __launch_bounds__(kNmsThreads) nms_mask_kernel(const float4* __restrict__ boxes, const int* __restrict__ counts,
                                                               int seg_stride, float thr, unsigned* __restrict__ mask) {
  __shared__ float4 sb[1024];
  __shared__ float sa[1024];
  const int seg = blockIdx.x, tid = threadIdx.x;
  const int n = counts[seg];
  const int r0 = blockIdx.y * kNmsRowsPerBlock;
  if (r0 >= n) return;
  const size_t base = (size_t)seg * seg_stride;
  for (int i = r0 + tid; i < n; i += blockDim.x) {  // only columns j > r0 are ever compared
    const float4 b = boxes[base + i];
    sb[i] = b;
    sa[i] = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  }
  __syncthreads();
  const int Wn = (n + 31) >> 5;
  const int rows = min(kNmsRowsPerBlock, n - r0);
  unsigned* mrow = mask + (size_t)seg * 1024 * 32;
  for (int item = tid; item < rows * Wn; item += blockDim.x) {
    const int i = r0 + item / Wn, w = item % Wn;
    unsigned bits = 0;
    if (w >= (i >> 5)) {
      const float4 bi = sb[i];
      const float ai = sa[i];
      const int j0 = w << 5;
      const int rot = threadIdx.x & 31;  // neighbouring lanes own neighbouring words (512 B apart): rotate the column
#pragma unroll 4                         // order so that a warp's shared-memory reads fall into distinct banks
      for (int t0 = 0; t0 < 32; ++t0) {
        const int t = (t0 + rot) & 31;
        const int j = j0 + t;
        if (j > i && j < n && iou_gt_fast(bi, ai, sb[j], sa[j], thr)) bits |= 1u << t;
      }
    }
    mrow[i * 32 + w] = bits;
  }
}

// Greedy NMS over the bitmask without a serial walk.  "j is kept iff it is valid and no EARLIER KEPT box suppresses it" is
// iterated as a fixed point over all boxes at once: keep' = valid & ~OR(rows of the boxes in keep).  Rows only hold bits j > i, so
// after t rounds the first t boxes are final and the iteration ends in the greedy solution (unique); dense proposal sets settle
// in a few dozen rounds of ~100 instructions per warp (warp w ORs the kept rows 32w.., lane <-> word, conflict-free) where the
// serial walk of warp 0 needed <= 1000 dependent shuffle + shared-memory steps (68 us per segment).  A set that has not settled
// after kNmsMaxRounds rounds falls back to that walk.
constexpr int kNmsMaxRounds = 96;

__global__ void __launch_bounds__(1024) nms_scan_kernel(const unsigned* __restrict__ mask, const unsigned char* __restrict__ valid,
                                                        const int* __restrict__ counts, int seg_stride, int* __restrict__ keep_idx,
                                                        int* __restrict__ keep_count) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  unsigned* sm = reinterpret_cast<unsigned*>(nms_smem);  // [n][32]
  __shared__ unsigned s_keep[2][32], s_valid[32], s_removed[32];
  __shared__ int s_prefix[33];
  const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = counts[seg];
  const size_t base = (size_t)seg * seg_stride;
  const int Wn = (n + 31) >> 5;
  const uint4* src = reinterpret_cast<const uint4*>(mask + (size_t)seg * 1024 * 32);
  for (int i = tid; i < n * 8; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = __ldg(src + i);
  {  // validity words: warp w <-> boxes 32w..
    const int j = tid;
    const unsigned v = __ballot_sync(kFullMask, j < n && valid[base + j] != 0);
    if (lane == 0) { s_valid[wid] = v; s_keep[0][wid] = v; }
  }
  __syncthreads();
  int cur = 0;
  bool settled = false;
  for (int round = 0; round < kNmsMaxRounds; ++round) {
    if (tid < 32) s_removed[tid] = 0u;
    __syncthreads();
    unsigned kw = s_keep[cur][wid], acc = 0u;
    while (kw) {  // warp-uniform
      const int r = __ffs(kw) - 1;
      kw &= kw - 1u;
      if (lane < Wn) acc |= sm[((wid << 5) + r) * 32 + lane];
    }
    if (acc) atomicOr(&s_removed[lane], acc);
    __syncthreads();
    bool changed = false;
    if (tid < 32) {
      const unsigned nk = s_valid[tid] & ~s_removed[tid];
      changed = nk != s_keep[cur][tid];
      s_keep[cur ^ 1][tid] = nk;
    }
    cur ^= 1;
    if (!__syncthreads_or(changed)) { settled = true; break; }
  }
  if (!settled) {  // block-uniform: the serial greedy walk (lane w owns word w of the removed set)
    if (tid < 32) {
      unsigned removed = ~s_valid[lane];
      for (int i = 0; i < n; ++i) {
        const unsigned rw = __shfl_sync(kFullMask, removed, i >> 5);
        if ((rw >> (i & 31)) & 1u) continue;
        if (lane < Wn) removed |= sm[i * 32 + lane];
      }
      s_keep[cur][lane] = ~removed;
    }
    __syncthreads();
  }
  // survivors in index (= score) order
  if (tid < 32) {
    const int c = __popc(s_keep[cur][lane]);
    int incl = c;
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(kFullMask, incl, d);
      if (lane >= d) incl += o;
    }
    s_prefix[lane + 1] = incl;
    if (lane == 0) s_prefix[0] = 0;
    if (lane == 31) keep_count[seg] = incl;
  }
  __syncthreads();
  if (tid < n) {
    const unsigned kwd = s_keep[cur][wid];
    if ((kwd >> lane) & 1u) keep_idx[base + s_prefix[wid] + __popc(kwd & ((1u << lane) - 1u))] = tid;
  }
}

// Merge the per-level survivors of one image by descending score (ties: lower level, lower position) and keep
// the first post_topk: rpn_outputs.py:147-156.
__global__ void __launch_bounds__(1024) rpn_merge_kernel(const float4* __restrict__ cand_box, const float* __restrict__ cand_score,
                                                         const int* __restrict__ keep_idx, const int* __restrict__ keep_count,
                                                         int post_topk, int max_props, float4* __restrict__ props,
                                                         int* __restrict__ prop_count) {
  __shared__ float s_score[kRpnLevels][kTopkCap];
  __shared__ int s_n[kRpnLevels];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < kRpnLevels) s_n[tid] = keep_count[b * kRpnLevels + tid];
  __syncthreads();
  for (int l = 0; l < kRpnLevels; ++l) {
    const size_t base = ((size_t)b * kRpnLevels + l) * kTopkCap;
    for (int i = tid; i < s_n[l]; i += blockDim.x) s_score[l][i] = cand_score[base + keep_idx[base + i]];
  }
  __syncthreads();
  int total = 0;
  for (int l = 0; l < kRpnLevels; ++l) total += s_n[l];
  for (int l = 0; l < kRpnLevels; ++l) {
    const size_t base = ((size_t)b * kRpnLevels + l) * kTopkCap;
    for (int i = tid; i < s_n[l]; i += blockDim.x) {
      const float s = s_score[l][i];
      int rank = i;
      for (int m = 0; m < kRpnLevels; ++m) {
        if (m == l) continue;
        // scores in level m are descending: count entries > s (and == s when m < l)
        int lo = 0, hi = s_n[m];
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const float v = s_score[m][mid];
          if (v > s || (m < l && v == s)) lo = mid + 1; else hi = mid;
        }
        rank += lo;
      }
      if (rank < post_topk && rank < max_props) props[(size_t)b * max_props + rank] = cand_box[base + keep_idx[base + i]];
    }
  }
  if (tid == 0) prop_count[b] = total < post_topk ? (total < max_props ? total : max_props) : (post_topk < max_props ? post_topk : max_props);
}

// ---------------------------------------------------------------------------------------------- ROIAlign
// One block per ROI, warp ph handles bin row ph (7 bins), lanes cover the channels with 128-bit loads.  The
// per-sample geometry (ROIAlign_cuda.cu:14-61 bilinear_interpolate) is separable: the x terms of all 7 x gw sample
// columns are tabulated once per ROI in shared memory, the y terms once per warp in registers.
constexpr int kRoiGmax = 8;
constexpr int kRoiColMax = 64;  // pixel columns of a ROI window the fast sweep tabulates

struct AxisSample { int lo, hi; float frac; int valid; };

__device__ __forceinline__ AxisSample axis_sample(float start, float bin, int p, int i, int g, int size) {
  AxisSample a;
  const float v0 = start + p * bin + (i + 0.5f) * bin / (float)g;
  a.valid = !(v0 < -1.f || v0 > (float)size);
  float v = fmaxf(v0, 0.f);
  int lo = (int)v, hi;
  if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; } else hi = lo + 1;
  a.lo = lo; a.hi = hi; a.frac = v - lo;
  return a;
}

// Fast column sweep of one bin row (see roi_align_kernel): NR row loads per pixel column with zero-padded row weights (rows
// beyond ny re-read row 0), software-pipelined one column ahead; e = (weight in the current bin, weight in the next bin,
// current bin ends here).
template <int NR>
__device__ __forceinline__ void roi_fast_sweep(const uint4* col0, uint4* dst0, const float4* s_col, int ncols, int cgroups,
                                               size_t row_stride, int ny, float wy, float inv_count, int lane) {
  const unsigned long long inv2 = pack2f(inv_count, inv_count);
  unsigned long long wy2[NR];
  size_t roff[NR];
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) {
    const float w = __shfl_sync(kFullMask, wy, rr);
    wy2[rr] = rr < ny ? pack2f(w, w) : 0ull;
    roff[rr] = (size_t)(rr < ny ? rr : 0) * row_stride;
  }
  for (int g = lane; g < cgroups; g += 32) {
    const uint4* col = col0 + g;
    uint4* dst = dst0 + g;
    Acc8 cur, nxt;
    cur.zero(); nxt.zero();
    uint4 v[NR];
#pragma unroll
    for (int rr = 0; rr < NR; ++rr) v[rr] = __ldg(col + roff[rr]);
    for (int ci = 0; ci < ncols; ++ci) {
      const float4 e = s_col[ci];
      uint4 a[NR];
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) a[rr] = v[rr];
      if (ci + 1 < ncols) {  // next column's rows are requested before this column is reduced
        col += cgroups;
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) v[rr] = __ldg(col + roff[rr]);
      }
      Acc8 t;
      t.zero();
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) acc_bf16x8(t, wy2[rr], a[rr]);
      acc_axpy(cur, pack2f(e.x, e.x), t);
      acc_axpy(nxt, pack2f(e.y, e.y), t);
      if (e.z != 0.f) {  // warp-uniform: the column table is shared by the block
        *dst = acc_store_bf16(cur, inv2);
        dst += cgroups;
        cur = nxt;
        nxt.zero();
      }
    }
  }
}

// The same sweep with the pixel columns prefetched through cp.async into a per-warp shared-memory ring instead of registers:
// kRoiRingSlots / NR columns are in flight per warp (2-4 instead of 1), no registers are held while a load is outstanding (so
// a fourth block fits on the SM), and the consumer reads its own 16 bytes back (no cross-lane hand-off: cp.async.wait_group is
// the only synchronisation).  The kernel is latency-bound (633 vs 800 us per 16 000 ROIs at 21 vs 14 resident warps with the
// same instruction count); the arithmetic and its order are unchanged, so the outputs are bit-identical.
constexpr int kRoiRingSlots = 12;  // 512-byte warp rows per warp: 6 KB per warp, 42 KB per block (below the 48 KB a launch gets without opting in)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NR>
__device__ __forceinline__ void roi_fast_sweep_async(const uint4* col0, uint4* dst0, const float4* s_col, int ncols, int cgroups,
                                                     size_t row_stride, int ny, float wy, float inv_count, int lane, uint4* ring) {
  constexpr int D = kRoiRingSlots / NR;  // columns in flight
  const unsigned long long inv2 = pack2f(inv_count, inv_count);
  unsigned long long wy2[NR];
  size_t roff[NR];
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) {
    const float w = __shfl_sync(kFullMask, wy, rr);
    wy2[rr] = rr < ny ? pack2f(w, w) : 0ull;
    roff[rr] = (size_t)(rr < ny ? rr : 0) * row_stride;
  }
  uint4* my = ring + lane;  // slot (c, rr) of this lane: my[(c * NR + rr) * 32]
  for (int g = lane; g < cgroups; g += 32) {
    const uint4* col = col0 + g;  // next column to request
    uint4* dst = dst0 + g;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      if (c < ncols) {
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) cp_async16(my + (c * NR + rr) * 32, col + roff[rr]);
        col += cgroups;
      }
      cp_async_commit();
    }
    Acc8 cur, nxt;
    cur.zero(); nxt.zero();
    int slot = 0;
    for (int ci = 0; ci < ncols; ++ci) {
      cp_async_wait<D - 1>();  // column ci has landed (every iteration commits exactly one group)
      const float4 e = s_col[ci];
      uint4 a[NR];
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) a[rr] = my[(slot * NR + rr) * 32];
      if (ci + D < ncols) {  // refill the slot just read with column ci + D
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) cp_async16(my + (slot * NR + rr) * 32, col + roff[rr]);
        col += cgroups;
      }
      cp_async_commit();
      if (++slot == D) slot = 0;
      Acc8 t;
      t.zero();
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) acc_bf16x8(t, wy2[rr], a[rr]);
      acc_axpy(cur, pack2f(e.x, e.x), t);
      acc_axpy(nxt, pack2f(e.y, e.y), t);
      if (e.z != 0.f) {  // warp-uniform: the column table is shared by the block
        *dst = acc_store_bf16(cur, inv2);
        dst += cgroups;
        cur = nxt;
        nxt.zero();
      }
    }
    cp_async_wait<0>();
  }
}

// Small ROIs (bins narrower than a pixel: a column may lie in several bins of the row; 65 % of the ROIs of the benchmark's
// seeded-random detector, whose proposals stay close to the 32-pixel anchors): the column sweep with one accumulator set per
// bin, columns prefetched through the same cp.async ring.  The register version loaded row after row inside a run-time loop
// (shuffle for the weight, branch, load, use): ny x ncols EXPOSED L2 latencies per warp - a third of the kernel's stall samples sat
// on the first FMA after that load.
template <int NR, int PW0, int NPW>  // bins [PW0, PW0 + NPW) of the row: two passes (4 + 3 bins) keep the accumulators in registers
__device__ __forceinline__ void roi_small_sweep_async(const uint4* col0, uint4* dst0, const int (&bx0)[7], const int (&bnx)[7],
                                                      const float (*s_wx)[kRoiGmax + 2], int cgroups, int row_stride, int ny, float wy,
                                                      float inv_count, int lane, uint4* ring) {
  constexpr int D = kRoiRingSlots / NR;
  float wyr[NR];
  int roff[NR];
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) {
    const float w = __shfl_sync(kFullMask, wy, rr);
    wyr[rr] = rr < ny ? w : 0.f;
    roff[rr] = (rr < ny ? rr : 0) * row_stride;
  }
  int xs = 0x7fffffff, xe = -1;  // the columns these bins touch
#pragma unroll
  for (int pw = PW0; pw < PW0 + NPW; ++pw) { xs = min(xs, bx0[pw]); xe = max(xe, bx0[pw] + bnx[pw] - 1); }
  uint4* my = ring + lane;
  const int ncols = xe - xs + 1;
  for (int g = lane; g < cgroups; g += 32) {
    const uint4* col = col0 + (size_t)xs * cgroups + g;  // next column to request
#pragma unroll
    for (int c = 0; c < D; ++c) {
      if (c < ncols) {
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) cp_async16(my + (c * NR + rr) * 32, col + roff[rr]);
        col += cgroups;
      }
      cp_async_commit();
    }
    Acc8 acc[NPW];
#pragma unroll
    for (int pw = 0; pw < NPW; ++pw) acc[pw].zero();
    int slot = 0;
    for (int ci = 0; ci < ncols; ++ci) {
      cp_async_wait<D - 1>();
      uint4 a[NR];
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) a[rr] = my[(slot * NR + rr) * 32];
      if (ci + D < ncols) {
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) cp_async16(my + (slot * NR + rr) * 32, col + roff[rr]);
        col += cgroups;
      }
      cp_async_commit();
      if (++slot == D) slot = 0;
      Acc8 t;
      t.zero();
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) acc_bf16x8(t, pack2f(wyr[rr], wyr[rr]), a[rr]);
      const int x = xs + ci;
#pragma unroll
      for (int pw = 0; pw < NPW; ++pw) {
        const int c = x - bx0[PW0 + pw];
        if ((unsigned)c < (unsigned)bnx[PW0 + pw]) {  // warp-uniform
          const float w = s_wx[PW0 + pw][c];
          acc_axpy(acc[pw], pack2f(w, w), t);
        }
      }
    }
    cp_async_wait<0>();
    const unsigned long long inv2 = pack2f(inv_count, inv_count);
#pragma unroll
    for (int pw = 0; pw < NPW; ++pw) dst0[(size_t)(PW0 + pw) * cgroups + g] = acc_store_bf16(acc[pw], inv2);
  }
}

// The bilinear samples of one bin overlap heavily (sample spacing <= 1 feature pixel), so the bin average is
// evaluated in its separable form  sum_rows sum_cols Wy[row] * Wx[col] * f(row, col)  with per-pixel weights
// Wy/Wx accumulated from the reference's per-sample terms: every touched feature pixel is loaded once per bin
// instead of once per neighbouring sample (up to 4x fewer 128-bit loads).
// kAsync: the fast sweep prefetches through the cp.async ring (dynamic shared memory: 7 * kRoiRingSlots * 512 B)
template <int kMinBlocks, bool kAsync>  // 3 blocks/SM caps the kernel at 96 registers (a few spills), 2 blocks/SM runs spill-free at 127
__global__ void __launch_bounds__(224, kMinBlocks) roi_align_kernel(const RoiLevels fl, const float4* __restrict__ props, const int* __restrict__ prop_count,
                                                        int B, int max_props, int C, __nv_bfloat16* __restrict__ out) {
  extern __shared__ uint4 s_ring[];
  __shared__ float s_wx[7][kRoiGmax + 2];
  __shared__ int s_x0[7], s_nx[7], s_ny[7];
  __shared__ float4 s_col[kRoiColMax];
  const int lane = threadIdx.x & 31, ph = threadIdx.x >> 5;
  const int b = blockIdx.y, r = blockIdx.x;  // grid (max_props, B)
  const long long roi = (long long)b * max_props + r;
  const int cgroups = C >> 3;
  uint4* dst_roi = reinterpret_cast<uint4*>(out + (size_t)roi * 49 * C);
  if (r >= prop_count[b]) {
    for (int i = threadIdx.x; i < 49 * cgroups; i += blockDim.x) dst_roi[i] = make_uint4(0, 0, 0, 0);
    return;
  }
  const float4 box = __ldg(props + roi);
  // poolers.py:13-44: level = clamp(floor(4 + log2(sqrt(area)/224 + eps)), 2, 5) - 2
  const float area = __fmul_rn(__fsub_rn(box.z, box.x), __fsub_rn(box.w, box.y));
  float lvf = floorf(__fadd_rn(4.f, log2f(__fadd_rn(__fdiv_rn(sqrtf(area), 224.f), 2.220446049250313e-16f))));
  lvf = fminf(fmaxf(lvf, 2.f), 5.f);
  int lvl = (int)lvf - 2;
  if (!(lvf == lvf)) lvl = 0;
  const int H = fl.H[lvl], W = fl.W[lvl];
  const float scale = fl.scale[lvl];
  const __nv_bfloat16* feat = fl.feat[lvl] + (size_t)b * H * W * C;
  const float rsw = __fsub_rn(__fmul_rn(box.x, scale), 0.5f), rsh = __fsub_rn(__fmul_rn(box.y, scale), 0.5f);
  const float rew = __fsub_rn(__fmul_rn(box.z, scale), 0.5f), reh = __fsub_rn(__fmul_rn(box.w, scale), 0.5f);
  const float roi_w = __fsub_rn(rew, rsw), roi_h = __fsub_rn(reh, rsh);
  const float bin_h = __fdiv_rn(roi_h, 7.f), bin_w = __fdiv_rn(roi_w, 7.f);
  const int gh = (int)ceilf(__fdiv_rn(roi_h, 7.f)), gw = (int)ceilf(__fdiv_rn(roi_w, 7.f));
  const float count = (float)max(gh * gw, 1);
  const bool separable = gw >= 1 && gh >= 1 && gw <= kRoiGmax && gh <= kRoiGmax;  // block-uniform

  if (separable) {
    // Per-pixel weights of one axis from the bin's samples.  Lane i evaluates sample i ONCE (the reference's float expression
    // with its IEEE division); the accumulation then walks the samples in the reference's order through shuffles - every lane
    // used to re-evaluate all g samples plus the first and the last one (10 % of the kernel's instructions).
    auto axis_weights = [&](float start, float bin, int p, int g, int size, int& first_lo, int& span) -> float {
      const AxisSample mine = axis_sample(start, bin, p, lane < g ? lane : g - 1, g, size);
      const unsigned valid = __ballot_sync(kFullMask, mine.valid != 0);
      first_lo = __shfl_sync(kFullMask, mine.lo, 0);
      span = __shfl_sync(kFullMask, mine.hi, g - 1) - first_lo + 1;
      float w = 0.f;
      for (int i = 0; i < g; ++i) {
        const int lo = __shfl_sync(kFullMask, mine.lo, i), hi = __shfl_sync(kFullMask, mine.hi, i);
        const float fr = __shfl_sync(kFullMask, mine.frac, i);
        if (!((valid >> i) & 1u)) continue;
        if (lo == first_lo + lane) w += 1.f - fr;
        if (hi == first_lo + lane) w += fr;
      }
      return w;
    };
    // x weights of bin column pw: warp pw builds them (lane c <-> column x0 + c)
    {
      const int pw = ph;
      int x0, nx;
      const float wx = axis_weights(rsw, bin_w, pw, gw, W, x0, nx);
      if (lane < nx && lane < kRoiGmax + 2) s_wx[pw][lane] = wx;
      if (lane == 0) { s_x0[pw] = x0; s_nx[pw] = nx < kRoiGmax + 2 ? nx : kRoiGmax + 2; }
    }
    // y weights of this warp's bin row (lane r <-> row y0 + r), kept in registers
    int y0, ny;
    const float wy = axis_weights(rsh, bin_h, ph, gh, H, y0, ny);
    if (ny > kRoiGmax + 2) ny = kRoiGmax + 2;
    if (lane == 0) s_ny[ph] = ny;
    __syncthreads();
    // ---- fast column sweep (the usual case: <= 6 feature rows per bin row, bin ends strictly increasing) ------------------
    // One table entry per pixel column of the ROI window, built once per block: the weights of the (at most two) bins that
    // contain the column and whether the first of them ends there.  The sweep itself is then branch-light: NR row loads with
    // zero-padded weights (rows beyond ny re-read row 0), t = sum_r Wy[r] f, cur += wa t, nxt += wb t, and a warp-uniform
    // "bin complete" step.  (The general sweep below spends ~70 instructions per 128-bit load on bookkeeping.)
    {
      // the per-bin conditions are evaluated one bin per lane and combined with a vote (every warp reaches the same verdict):
      // reading the 21 table entries in every thread cost 6 % of the kernel's stall samples
      const int bp = lane < 7 ? lane : 6;
      const int bx = s_x0[bp], bn = s_nx[bp], by = s_ny[bp];
      const int bx1 = __shfl_down_sync(kFullMask, bx, 1), bn1 = __shfl_down_sync(kFullMask, bn, 1);
      const int bx2 = __shfl_down_sync(kFullMask, bx, 2);
      bool ok = by <= 6 && bn >= 1;
      if (lane < 6) ok &= bx1 >= bx && bx1 + bn1 > bx + bn && bx1 <= bx + bn;
      if (lane < 5) ok &= bx2 > bx + bn - 1;  // a column lies in at most two bins
      const int xs = __shfl_sync(kFullMask, bx, 0), xe = __shfl_sync(kFullMask, bx + bn - 1, 6);
      const int ncols = xe - xs + 1;
      const bool fast = __all_sync(kFullMask, ok) && ncols <= kRoiColMax;
      if (fast) {  // block-uniform
        if ((int)threadIdx.x < ncols) {
          const int x = xs + threadIdx.x;
          int pa = 0;
#pragma unroll
          for (int p2 = 1; p2 < 7; ++p2) pa += x > s_x0[p2 - 1] + s_nx[p2 - 1] - 1;  // first bin whose last column is >= x
          const float wa = s_wx[pa][x - s_x0[pa]];
          const float wb = (pa < 6 && x >= s_x0[pa + 1]) ? s_wx[pa + 1][x - s_x0[pa + 1]] : 0.f;
          s_col[threadIdx.x] = make_float4(wa, wb, x == s_x0[pa] + s_nx[pa] - 1 ? 1.f : 0.f, 0.f);
        }
        __syncthreads();
        const float inv_count = 1.f / count;
        const int nymax = __reduce_max_sync(kFullMask, by);
        const uint4* col = reinterpret_cast<const uint4*>(feat + ((size_t)y0 * W + xs) * C);
        uint4* dst = dst_roi + (size_t)(ph * 7) * cgroups;
        // NR = the block's tallest bin row: rows beyond a warp's own ny are zero-weight re-reads of row 0, so every NR above
        // the need costs a load + unpack + FMA group per pixel column (3- and 5-row bins are the common case at 14 / 28 px)
        if (kAsync) {
          uint4* ring = s_ring + ph * (kRoiRingSlots * 32);
          if (nymax <= 3) roi_fast_sweep_async<3>(col, dst, s_col, ncols, cgroups, (size_t)W * cgroups, ny, wy, inv_count, lane, ring);
          else if (nymax <= 4) roi_fast_sweep_async<4>(col, dst, s_col, ncols, cgroups, (size_t)W * cgroups, ny, wy, inv_count, lane, ring);
          else if (nymax <= 5) roi_fast_sweep_async<5>(col, dst, s_col, ncols, cgroups, (size_t)W * cgroups, ny, wy, inv_count, lane, ring);
          else roi_fast_sweep_async<6>(col, dst, s_col, ncols, cgroups, (size_t)W * cgroups, ny, wy, inv_count, lane, ring);
          return;
        }
        if (nymax <= 3) roi_fast_sweep<3>(col, dst, s_col, ncols, cgroups, (size_t)W * cgroups, ny, wy, inv_count, lane);
        else if (nymax <= 4) roi_fast_sweep<4>(col, dst, s_col, ncols, cgroups, (size_t)W * cgroups, ny, wy, inv_count, lane);
        else if (nymax <= 5) roi_fast_sweep<5>(col, dst, s_col, ncols, cgroups, (size_t)W * cgroups, ny, wy, inv_count, lane);
        else roi_fast_sweep<6>(col, dst, s_col, ncols, cgroups, (size_t)W * cgroups, ny, wy, inv_count, lane);
        return;
      }
    }
    // Column sweep: walk the bin row's pixel columns once, reduce each column over the rows (t = sum_y Wy f), and add
    // Wx * t to the bins that contain the column.  Adjacent bins share their border columns, so this loads ~20 % fewer
    // pixels than bin-by-bin and does the row reduction once per column instead of once per (bin, column).  Needs every
    // column to lie in at most two bins (block-uniform test; tiny ROIs take the bin-by-bin loop below).
    bool sweep = true;
#pragma unroll
    for (int pw = 0; pw < 5; ++pw) sweep &= s_x0[pw + 2] > s_x0[pw] + s_nx[pw] - 1;
#pragma unroll
    for (int pw = 0; pw < 6; ++pw) sweep &= s_x0[pw + 1] >= s_x0[pw] && s_x0[pw + 1] + s_nx[pw + 1] >= s_x0[pw] + s_nx[pw];
    if (sweep) {
      const float inv_count = 1.f / count;
      // row weights to registers once; with <= 4 rows (the usual case: ROIs span 7..14 pixels at their level) the rows of
      // the NEXT column are requested before the current column is reduced, so 4..8 128-bit loads are in flight per lane
      constexpr int kPre = 4;
      float wyr[kPre];
#pragma unroll
      for (int rr = 0; rr < kPre; ++rr) wyr[rr] = rr < ny ? __shfl_sync(kFullMask, wy, rr) : 0.f;
      const bool pipelined = ny <= kPre;
      for (int g = lane; g < cgroups; g += 32) {
        Acc8 cur, nxt;
        cur.zero(); nxt.zero();
        int pw = 0;
        const int xe = s_x0[6] + s_nx[6] - 1;
        const uint4* col0 = reinterpret_cast<const uint4*>(feat + (size_t)y0 * W * C) + g;
        const size_t row_stride = (size_t)W * cgroups, col_stride = (size_t)cgroups;
        uint4 pre[kPre];
        auto fetch = [&](int x) {
#pragma unroll
          for (int rr = 0; rr < kPre; ++rr)
            pre[rr] = (rr < ny && wyr[rr] != 0.f) ? __ldg(col0 + (size_t)rr * row_stride + (size_t)x * col_stride) : make_uint4(0, 0, 0, 0);
        };
        if (pipelined) fetch(s_x0[0]);
        for (int x = s_x0[0]; x <= xe && pw < 7; ++x) {
          Acc8 t;
          t.zero();
          if (pipelined) {
            uint4 v[kPre];
#pragma unroll
            for (int rr = 0; rr < kPre; ++rr) v[rr] = pre[rr];
            if (x < xe) fetch(x + 1);
#pragma unroll
            for (int rr = 0; rr < kPre; ++rr) acc_bf16x8(t, pack2f(wyr[rr], wyr[rr]), v[rr]);
          } else {
            const __nv_bfloat16* colp = feat + ((size_t)y0 * W + x) * C;
            for (int rr = 0; rr < ny; ++rr) {
              const float w = __shfl_sync(kFullMask, wy, rr);
              if (w == 0.f) continue;
              const uint4 v = __ldg(reinterpret_cast<const uint4*>(colp + (size_t)rr * W * C) + g);
              acc_bf16x8(t, pack2f(w, w), v);
            }
          }
          const int c0 = x - s_x0[pw];
          if (c0 >= 0 && c0 < s_nx[pw]) {
            const float w = s_wx[pw][c0];
            acc_axpy(cur, pack2f(w, w), t);
          }
          if (pw < 6) {
            const int c1 = x - s_x0[pw + 1];
            if (c1 >= 0 && c1 < s_nx[pw + 1]) {
              const float w = s_wx[pw + 1][c1];
              acc_axpy(nxt, pack2f(w, w), t);
            }
          }
          while (pw < 7 && x >= s_x0[pw] + s_nx[pw] - 1) {  // bin pw is complete (two bins may end on the same column)
            dst_roi[(size_t)(ph * 7 + pw) * cgroups + g] = acc_store_bf16(cur, pack2f(inv_count, inv_count));
            cur = nxt;
            nxt.zero();
            ++pw;
          }
        }
      }
      return;
    }
    // Small ROIs (bins narrower than a pixel: a column may lie in up to all 7 bins of the row): same column sweep
    // with one accumulator set per bin.  Each pixel of the bin row's window is still loaded exactly once.
    {
      const float inv_count = 1.f / count;
      int bx0[7], bnx[7];
      int xs = 0x7fffffff, xe = -1;
#pragma unroll
      for (int pw = 0; pw < 7; ++pw) {
        bx0[pw] = s_x0[pw]; bnx[pw] = s_nx[pw];
        xs = min(xs, bx0[pw]); xe = max(xe, bx0[pw] + bnx[pw] - 1);
      }
      if (kAsync && ny <= 6) {  // warp-uniform: each bin row picks its own depth
        const uint4* c0 = reinterpret_cast<const uint4*>(feat + (size_t)y0 * W * C);
        uint4* d0 = dst_roi + (size_t)(ph * 7) * cgroups;
        uint4* ring = s_ring + ph * (kRoiRingSlots * 32);
        const int rs = W * cgroups;
        if (ny <= 3) {
          roi_small_sweep_async<3, 0, 4>(c0, d0, bx0, bnx, s_wx, cgroups, rs, ny, wy, inv_count, lane, ring);
          roi_small_sweep_async<3, 4, 3>(c0, d0, bx0, bnx, s_wx, cgroups, rs, ny, wy, inv_count, lane, ring);
        } else if (ny <= 4) {
          roi_small_sweep_async<4, 0, 4>(c0, d0, bx0, bnx, s_wx, cgroups, rs, ny, wy, inv_count, lane, ring);
          roi_small_sweep_async<4, 4, 3>(c0, d0, bx0, bnx, s_wx, cgroups, rs, ny, wy, inv_count, lane, ring);
        } else {
          roi_small_sweep_async<6, 0, 4>(c0, d0, bx0, bnx, s_wx, cgroups, rs, ny, wy, inv_count, lane, ring);
          roi_small_sweep_async<6, 4, 3>(c0, d0, bx0, bnx, s_wx, cgroups, rs, ny, wy, inv_count, lane, ring);
        }
        return;
      }
      constexpr int kPre7 = 3;
      float wy7[kPre7];
#pragma unroll
      for (int rr = 0; rr < kPre7; ++rr) wy7[rr] = rr < ny ? __shfl_sync(kFullMask, wy, rr) : 0.f;
      const bool few_rows = ny <= kPre7;
      for (int g = lane; g < cgroups; g += 32) {
        Acc8 acc[7];
#pragma unroll
        for (int pw = 0; pw < 7; ++pw) acc[pw].zero();
        const uint4* col0 = reinterpret_cast<const uint4*>(feat + (size_t)y0 * W * C) + g;
        const size_t row_stride = (size_t)W * cgroups;
        for (int x = xs; x <= xe; ++x) {
          Acc8 t;
          t.zero();
          const uint4* colp = col0 + (size_t)x * cgroups;
          if (few_rows) {  // block-uniform: all rows of the column requested at once
            uint4 v[kPre7];
#pragma unroll
            for (int rr = 0; rr < kPre7; ++rr)
              v[rr] = (rr < ny && wy7[rr] != 0.f) ? __ldg(colp + (size_t)rr * row_stride) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int rr = 0; rr < kPre7; ++rr) acc_bf16x8(t, pack2f(wy7[rr], wy7[rr]), v[rr]);
          } else {
            for (int rr = 0; rr < ny; ++rr) {
              const float w = __shfl_sync(kFullMask, wy, rr);
              if (w == 0.f) continue;
              const uint4 v = __ldg(colp + (size_t)rr * row_stride);
              acc_bf16x8(t, pack2f(w, w), v);
            }
          }
#pragma unroll
          for (int pw = 0; pw < 7; ++pw) {
            const int c = x - bx0[pw];
            if (c >= 0 && c < bnx[pw]) {
              const float w = s_wx[pw][c];
              acc_axpy(acc[pw], pack2f(w, w), t);
            }
          }
        }
#pragma unroll
        for (int pw = 0; pw < 7; ++pw)
          dst_roi[(size_t)(ph * 7 + pw) * cgroups + g] = acc_store_bf16(acc[pw], pack2f(inv_count, inv_count));
      }
    }
    return;
  }

  // general path (very large or degenerate ROIs): sample by sample, as the reference kernel does
  for (int g = lane; g < cgroups; g += 32) {
    for (int pw = 0; pw < 7; ++pw) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int iy = 0; iy < gh; ++iy) {
        const AxisSample ay = axis_sample(rsh, bin_h, ph, iy, gh, H);
        if (!ay.valid) continue;
        const float ly = ay.frac, hy = 1.f - ly;
        const __nv_bfloat16* rowl = feat + (size_t)ay.lo * W * C;
        const __nv_bfloat16* rowh = feat + (size_t)ay.hi * W * C;
        for (int ix = 0; ix < gw; ++ix) {
          const AxisSample ax = axis_sample(rsw, bin_w, pw, ix, gw, W);
          if (!ax.valid) continue;
          const float lx = ax.frac, hx = 1.f - lx;
          const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
          const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(rowl + (size_t)ax.lo * C) + g);
          const uint4 v2 = __ldg(reinterpret_cast<const uint4*>(rowl + (size_t)ax.hi * C) + g);
          const uint4 v3 = __ldg(reinterpret_cast<const uint4*>(rowh + (size_t)ax.lo * C) + g);
          const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(rowh + (size_t)ax.hi * C) + g);
          acc[0] += w1 * bflo(v1.x) + w2 * bflo(v2.x) + w3 * bflo(v3.x) + w4 * bflo(v4.x);
          acc[1] += w1 * bfhi(v1.x) + w2 * bfhi(v2.x) + w3 * bfhi(v3.x) + w4 * bfhi(v4.x);
          acc[2] += w1 * bflo(v1.y) + w2 * bflo(v2.y) + w3 * bflo(v3.y) + w4 * bflo(v4.y);
          acc[3] += w1 * bfhi(v1.y) + w2 * bfhi(v2.y) + w3 * bfhi(v3.y) + w4 * bfhi(v4.y);
          acc[4] += w1 * bflo(v1.z) + w2 * bflo(v2.z) + w3 * bflo(v3.z) + w4 * bflo(v4.z);
          acc[5] += w1 * bfhi(v1.z) + w2 * bfhi(v2.z) + w3 * bfhi(v3.z) + w4 * bfhi(v4.z);
          acc[6] += w1 * bflo(v1.w) + w2 * bflo(v2.w) + w3 * bflo(v3.w) + w4 * bflo(v4.w);
          acc[7] += w1 * bfhi(v1.w) + w2 * bfhi(v2.w) + w3 * bfhi(v3.w) + w4 * bfhi(v4.w);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] /= count;
      dst_roi[(size_t)(ph * 7 + pw) * cgroups + g] =
          make_uint4(packbf(acc[0], acc[1]), packbf(acc[2], acc[3]), packbf(acc[4], acc[5]), packbf(acc[6], acc[7]));
    }
  }
}

// ---------------------------------------------------------------------------------------------- head post-processing
constexpr int kHeadThreads = 1024;  // up to 1000 ROIs per image: one row per thread, wide rank-by-counting
constexpr int kMaxCand = 1024;  // score_thresh >= 0.5 admits at most one class per ROI, so <= 1000 candidates

// KT = compile-time class count (1: KAIST, 3: FLIR); KT = 0: any class count given at run time (`k_rt`, e.g. the 80
// COCO classes of the reference's rgb_only model, demo_FLIR_save_predictions.py:58-60).  The run-time variant admits
// one candidate per ROI - the only possibility when score_thresh >= 0.5, which its launcher requires.
template <int KT>
__global__ void __launch_bounds__(kHeadThreads) head_post_kernel(const float* __restrict__ head /*[B*max_props][npad]*/, int npad,
                                                                 const float4* __restrict__ props, const int* __restrict__ prop_count,
                                                                 int max_props, HeadParams hp, DetOut out, int k_rt) {
  constexpr int K = KT ? KT : 1;       // static array extents
  const int Kn = KT ? KT : k_rt;       // class count
  __shared__ float4 c_box[kMaxCand];
  __shared__ float c_score[kMaxCand];
  __shared__ unsigned short c_row[kMaxCand], c_cls[kMaxCand], c_order[kMaxCand];
  __shared__ unsigned char c_dead[kMaxCand];
  __shared__ unsigned long long c_key[kMaxCand];
  __shared__ int warp_cnt[kHeadThreads / 32];
  __shared__ int s_ncand, s_run, s_keep[kMaxDet], s_nkeep;
  __shared__ float s_red[kHeadThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int R = prop_count[b];
  const float* hb = head + (size_t)b * max_props * npad;
  if (tid == 0) { s_ncand = 0; s_run = 0; s_nkeep = 0; }
  __syncthreads();
  // 1. candidates (row-major over (roi, class)) with score > thresh, ordered compaction
  for (int r0 = 0; r0 < R; r0 += blockDim.x) {
    const int r = r0 + tid;
    float pr[K + 1];
    float4 bx[K];
    int nc = 0;
    bool above[K];
#pragma unroll
    for (int k2 = 0; k2 < K; ++k2) above[k2] = false;
    if (KT == 0 && r < R) {
      // run-time class count: softmax in the same operation order (max, exp(x - max), sum in class order, divide),
      // then only the (unique) class above the threshold is decoded
      const float* row = hb + (size_t)r * npad;
      float mx = -INFINITY;
      for (int k2 = 0; k2 <= Kn; ++k2) mx = fmaxf(mx, row[k2]);
      float den = 0.f;
      for (int k2 = 0; k2 <= Kn; ++k2) den += expf(row[k2] - mx);
      bool fin = true;
      int best = -1;
      float bestp = 0.f;
      for (int k2 = 0; k2 <= Kn; ++k2) {
        const float pk = __fdiv_rn(expf(row[k2] - mx), den);
        fin &= isfinite(pk);
        if (k2 < Kn && pk > hp.score_thresh) { best = k2; bestp = pk; }
      }
      const float4 p = props[(size_t)b * max_props + r];
      const float wdt = __fsub_rn(p.z, p.x), hgt = __fsub_rn(p.w, p.y);
      const float cx = __fadd_rn(p.x, __fmul_rn(0.5f, wdt)), cy = __fadd_rn(p.y, __fmul_rn(0.5f, hgt));
      for (int k2 = 0; k2 < Kn; ++k2) {  // fast_rcnn.py:100-103 drops the ROI if ANY class box is non-finite
        const float* d = row + (Kn + 1) + 4 * k2;
        const float dx = __fdiv_rn(d[0], 10.f), dy = __fdiv_rn(d[1], 10.f);
        const float dw = fminf(__fdiv_rn(d[2], 5.f), kScaleClamp), dh = fminf(__fdiv_rn(d[3], 5.f), kScaleClamp);
        const float pcx = __fadd_rn(__fmul_rn(dx, wdt), cx), pcy = __fadd_rn(__fmul_rn(dy, hgt), cy);
        const float pw = __fmul_rn(expf(dw), wdt), ph = __fmul_rn(expf(dh), hgt);
        float4 q = make_float4(__fsub_rn(pcx, __fmul_rn(0.5f, pw)), __fsub_rn(pcy, __fmul_rn(0.5f, ph)),
                               __fadd_rn(pcx, __fmul_rn(0.5f, pw)), __fadd_rn(pcy, __fmul_rn(0.5f, ph)));
        fin &= finite4(q);
        if (k2 == best) {
          q.x = fminf(fmaxf(q.x, 0.f), hp.img_w); q.z = fminf(fmaxf(q.z, 0.f), hp.img_w);
          q.y = fminf(fmaxf(q.y, 0.f), hp.img_h); q.w = fminf(fmaxf(q.w, 0.f), hp.img_h);
          bx[0] = q;
        }
      }
      if (fin && best >= 0) { above[0] = true; nc = 1; pr[0] = bestp; pr[1] = __int_as_float(best); }
    }
    if (KT != 0 && r < R) {
      const float* row = hb + (size_t)r * npad;
      float mx = -INFINITY;
#pragma unroll
      for (int k2 = 0; k2 <= K; ++k2) { pr[k2] = row[k2]; mx = fmaxf(mx, pr[k2]); }
      float den = 0.f;
#pragma unroll
      for (int k2 = 0; k2 <= K; ++k2) { pr[k2] = expf(pr[k2] - mx); den += pr[k2]; }
      bool fin = true;
#pragma unroll
      for (int k2 = 0; k2 <= K; ++k2) { pr[k2] = __fdiv_rn(pr[k2], den); fin &= isfinite(pr[k2]); }
      const float4 p = props[(size_t)b * max_props + r];
      const float wdt = __fsub_rn(p.z, p.x), hgt = __fsub_rn(p.w, p.y);
      const float cx = __fadd_rn(p.x, __fmul_rn(0.5f, wdt)), cy = __fadd_rn(p.y, __fmul_rn(0.5f, hgt));
#pragma unroll
      for (int k2 = 0; k2 < K; ++k2) {
        const float* d = row + (K + 1) + 4 * k2;
        const float dx = __fdiv_rn(d[0], 10.f), dy = __fdiv_rn(d[1], 10.f);
        const float dw = fminf(__fdiv_rn(d[2], 5.f), kScaleClamp), dh = fminf(__fdiv_rn(d[3], 5.f), kScaleClamp);
        const float pcx = __fadd_rn(__fmul_rn(dx, wdt), cx), pcy = __fadd_rn(__fmul_rn(dy, hgt), cy);
        const float pw = __fmul_rn(expf(dw), wdt), ph = __fmul_rn(expf(dh), hgt);
        float4 q = make_float4(__fsub_rn(pcx, __fmul_rn(0.5f, pw)), __fsub_rn(pcy, __fmul_rn(0.5f, ph)),
                               __fadd_rn(pcx, __fmul_rn(0.5f, pw)), __fadd_rn(pcy, __fmul_rn(0.5f, ph)));
        fin &= finite4(q);
        q.x = fminf(fmaxf(q.x, 0.f), hp.img_w); q.z = fminf(fmaxf(q.z, 0.f), hp.img_w);
        q.y = fminf(fmaxf(q.y, 0.f), hp.img_h); q.w = fminf(fmaxf(q.w, 0.f), hp.img_h);
        bx[k2] = q;
      }
      if (fin) {
#pragma unroll
        for (int k2 = 0; k2 < K; ++k2) { above[k2] = pr[k2] > hp.score_thresh; nc += above[k2]; }
      }
    }
    // exclusive scan of nc over the block
    int inc = nc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(kFullMask, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) warp_cnt[wid] = inc;
    __syncthreads();
    int pos = s_run + inc - nc;
    for (int w2 = 0; w2 < wid; ++w2) pos += warp_cnt[w2];
    if (nc) {
#pragma unroll
      for (int k2 = 0; k2 < K; ++k2)
        if (above[k2]) {
          if (pos < kMaxCand) {
            c_box[pos] = bx[k2]; c_score[pos] = pr[k2]; c_row[pos] = (unsigned short)r;
            c_cls[pos] = (unsigned short)(KT ? k2 : __float_as_int(pr[1]));
          }
          ++pos;
        }
    }
    __syncthreads();
    if (tid == 0) { int tot = 0; for (int w2 = 0; w2 < kHeadThreads / 32; ++w2) tot += warp_cnt[w2]; s_run += tot; }
    __syncthreads();
  }
  const int n = s_run < kMaxCand ? s_run : kMaxCand;
  // 2. order by score (descending, ties lower candidate index first) and max coordinate for the offset trick
  // (bitonic sort of (score key, ~index) words, one element per thread: 55 barrier steps instead of n^2 / 1024 compares per thread,
  //  which cost ~30 us per image at n = 1000)
  float mc = -INFINITY;
  static_assert(kMaxCand == kHeadThreads, "one sort element per thread");
  c_key[tid] = tid < n ? ((unsigned long long)sort_key(c_score[tid]) << 32) | (unsigned)(0xffffffffu - (unsigned)tid) : 0ull;
  __syncthreads();
  for (int size = 2; size <= kMaxCand; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int j = tid ^ stride;
      if (j > tid) {
        const unsigned long long a = c_key[tid], c = c_key[j];
        const bool desc = (tid & size) == 0;
        if (desc ? a < c : a > c) { c_key[tid] = c; c_key[j] = a; }
      }
      __syncthreads();
    }
  }
  if (tid < n) {
    c_order[tid] = (unsigned short)(0xffffffffu - (unsigned)(c_key[tid] & 0xffffffffu));
    c_dead[tid] = 0;
    const float4 q = c_box[tid];
    mc = fmaxf(mc, fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)));
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) mc = fmaxf(mc, __shfl_xor_sync(kFullMask, mc, s));
  if (lane == 0) s_red[wid] = mc;
  __syncthreads();
  mc = s_red[0];
  for (int w2 = 1; w2 < kHeadThreads / 32; ++w2) mc = fmaxf(mc, s_red[w2]);
  const float off1 = __fadd_rn(mc, 1.f);
  // 3. greedy per-class NMS in score order, stop after detections_per_image survivors
  // (the survivor count lives in a register of every thread - all threads take the same decisions - so one barrier per SURVIVOR
  // is enough: it orders this round's c_dead writes, which only touch later candidates, before the next round's reads)
  int nkeep = 0;
  for (int a = 0; a < n; ++a) {
    const int i = c_order[a];
    if (c_dead[i]) continue;        // uniform: shared memory, synchronised below
    if (nkeep >= hp.max_det) break;
    if (tid == 0) { s_keep[nkeep] = i; }
    const float oi = __fmul_rn((float)c_cls[i], off1);
    const float4 q = c_box[i];
    const float4 bi = make_float4(__fadd_rn(q.x, oi), __fadd_rn(q.y, oi), __fadd_rn(q.z, oi), __fadd_rn(q.w, oi));
    const float ai = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
    for (int c = a + 1 + tid; c < n; c += blockDim.x) {
      const int j = c_order[c];
      if (c_dead[j]) continue;
      const float oj = __fmul_rn((float)c_cls[j], off1);
      const float4 p = c_box[j];
      const float4 bj = make_float4(__fadd_rn(p.x, oj), __fadd_rn(p.y, oj), __fadd_rn(p.z, oj), __fadd_rn(p.w, oj));
      const float aj = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
      if (iou_gt(bi, ai, bj, aj, hp.nms_thresh)) c_dead[j] = 1;
    }
    ++nkeep;
    __syncthreads();
  }
  if (tid == 0) s_nkeep = nkeep;
  __syncthreads();
  // 4. emit: postprocess scale/clip/non-empty (postprocessing.py:8-52), fork fields
  const int nk = s_nkeep;
  if (wid == 0) {
    int emitted = 0;
    for (int i0 = 0; i0 < nk; i0 += 32) {
      const int i = i0 + lane;
      bool ok = false;
      float4 q = make_float4(0, 0, 0, 0);
      int ci = 0;
      if (i < nk) {
        ci = s_keep[i];
        q = c_box[ci];
        q.x = __fmul_rn(q.x, hp.scale_x); q.z = __fmul_rn(q.z, hp.scale_x);
        q.y = __fmul_rn(q.y, hp.scale_y); q.w = __fmul_rn(q.w, hp.scale_y);
        q.x = fminf(fmaxf(q.x, 0.f), hp.out_w); q.z = fminf(fmaxf(q.z, 0.f), hp.out_w);
        q.y = fminf(fmaxf(q.y, 0.f), hp.out_h); q.w = fminf(fmaxf(q.w, 0.f), hp.out_h);
        ok = (q.z - q.x > 0.f) && (q.w - q.y > 0.f);
      }
      const unsigned bal = __ballot_sync(kFullMask, ok);
      if (ok) {
        const int o = emitted + __popc(bal & ((1u << lane) - 1u));
        const size_t ob = (size_t)b * kMaxDet + o;
        const int r = c_row[ci];
        const float* row = hb + (size_t)r * npad;
        out.boxes[ob] = q;
        out.scores[ob] = c_score[ci];
        out.classes[ob] = (int)c_cls[ci];
        float mx = -INFINITY;
        for (int k2 = 0; k2 <= Kn; ++k2) { out.logits[ob * (Kn + 1) + k2] = row[k2]; mx = fmaxf(mx, row[k2]); }
        float den = 0.f;
        for (int k2 = 0; k2 <= Kn; ++k2) den += expf(row[k2] - mx);
        for (int k2 = 0; k2 < Kn; ++k2) out.probs[ob * Kn + k2] = __fdiv_rn(expf(row[k2] - mx), den);
        // sic: the reference indexes the per-ROI variance with the candidate-list index (quirk 1)
        const int vrow = ci < R ? ci : R - 1;
        out.vars[ob] = expf(hb[(size_t)vrow * npad + (Kn + 1) + 4 * Kn]);
        out.roi_index[ob] = r;
      }
      emitted += __popc(bal);
    }
    if (lane == 0) out.count[b] = emitted;
  }
}

// ---------------------------------------------------------------------------------------------- pack for ProbEn
__global__ void pack_offsets_kernel(PackIn in, int B, int M, int* __restrict__ offsets) {
  // single block: exclusive scan of counts[b*M + m]
  __shared__ int carry;
  __shared__ int wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  const int n = B * M;
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int i = i0 + tid;
    int c = 0;
    if (i < n) { const int b = i / M, m = i - b * M; c = in.count[m][b]; }
    int inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(kFullMask, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    int pre = carry;
    for (int w = 0; w < wid; ++w) pre += wsum[w];
    if (i < n) offsets[i] = pre + inc - c;
    __syncthreads();
    if (tid == blockDim.x - 1) carry = pre + inc;
    __syncthreads();
  }
  if (tid == 0) offsets[n] = carry;
}

__global__ void pack_copy_kernel(PackIn in, int B, int M, int K, const int* __restrict__ offsets, float4* __restrict__ boxes,
                                 float* __restrict__ scores, int* __restrict__ classes, float* __restrict__ probs, float* __restrict__ vars) {
  const int seg = blockIdx.x, b = seg / M, m = seg - b * M;
  const int n = in.count[m][b], o = offsets[seg];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const size_t s = (size_t)b * kMaxDet + i;
    boxes[o + i] = in.boxes[m][s];
    scores[o + i] = in.scores[m][s];
    classes[o + i] = in.classes[m][s];
    vars[o + i] = in.vars[m][s];
    for (int k = 0; k < K; ++k) probs[(size_t)(o + i) * K + k] = in.probs[m][s * K + k];
  }
}

inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)sm_count() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

// ---------------------------------------------------------------------------------------------- launchers
int launch_stem_im2col(const float* img, void* canvas, void* A, int B, int Ctot, int c0, int C, int Hi, int Wi, int Hc, int Wc,
                       const StemNorm& nrm, cudaStream_t st) {
  if (C > 4) return PE_ERR_UNSUPPORTED;
  const int Ho = Hc / 2, Wo = Wc / 2, Hp = Hc + 6, Wp = Wc + 8;
  stem_canvas_kernel<<<grid_for((long long)B * Hp * Wp, 256), 256, 0, st>>>(img, reinterpret_cast<__half*>(canvas), B, Ctot, c0, C, Hi, Wi,
                                                                              Hp, Wp, nrm);
  PE_LAUNCH_CHECK();
  if (A) {  // optional explicit im2col matrix (kept for the op-level GEMM tests; the engine reads the canvas through TMA)
    const long long total = (long long)B * Ho * Wo * 28;
    stem_im2col_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const __half*>(canvas), reinterpret_cast<__half*>(A), B, Ho,
                                                              Wo, Hp, Wp);
    PE_LAUNCH_CHECK();
  }
  return PE_OK;
}

// Taps per output coordinate of the Pillow-exact resize: 3 cover any upscale, 9 a downscale by up to 4x.
static int pil_kmax(int Hs, int Ws, int Hd, int Wd) {
  const double s = fmax((double)Hs / Hd, (double)Ws / Wd);
  return s <= 1.0 ? 3 : (s <= 4.0 ? 9 : -1);
}

int launch_stem_im2col_u8(const unsigned char* frames, void* canvas, void* A, int B, int Ctot, int c0, int C, int Hs, int Ws, int Hi,
                          int Wi, int Hc, int Wc, int round_u8, const StemNorm& nrm, cudaStream_t st, void* taps_ws) {
  if (C > 4) return PE_ERR_UNSUPPORTED;
  const int Ho = Hc / 2, Wo = Wc / 2, Hp = Hc + 6, Wp = Wc + 8;
  const int kmax = round_u8 ? pil_kmax(Hs, Ws, Hi, Wi) : 0;
  if (kmax < 0) return PE_ERR_UNSUPPORTED;
  const int grid = grid_for((long long)B * Hp * Wp, 256);
  __half* cv = reinterpret_cast<__half*>(canvas);
  int4* taps = kmax == 3 ? reinterpret_cast<int4*>(taps_ws) : nullptr;
  __half* lut = nullptr;
  if (taps) {
    // the (v - mean) / std table of the tiled kernel (4 x 256 fp16 = 128 int4) sits behind the tap tables: taps_ws holds
    // canvas_h + canvas_w + 128 int4 (engine.cu "pil_taps")
    lut = reinterpret_cast<__half*>(taps + Hc + Wc);
    pil_taps_table_kernel<<<ceil_div(max(Hi + Wi, 1024), 256), 256, 0, st>>>(taps, Hs, Ws, Hi, Wi, lut, C, nrm);
    PE_LAUNCH_CHECK();
  }
  // upscales: the tiled two-pass kernel (PE_STEM_TILED=0 keeps the per-pixel kernel, A/B switch)
  static const int tiled_env = [] { const char* e = getenv("PE_STEM_TILED"); return e ? atoi(e) : 1; }();
  const size_t tiled_smem = (size_t)kCanvasSpan * (Wi * sizeof(uint32_t) + (size_t)Ws * Ctot);
  if (kmax == 3 && taps && lut && tiled_env && tiled_smem <= 160 * 1024) {
    static DeviceOnce attr_once;
    if (attr_once.needed()) {
      PE_CUDA_CHECK(cudaFuncSetAttribute(stem_canvas_u8_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      attr_once.mark();
    }
    stem_canvas_u8_tiled_kernel<<<dim3((unsigned)ceil_div(Hp, kCanvasRows), (unsigned)B), 256, tiled_smem, st>>>(frames, cv, Ctot, c0, C, Hs, Ws,
                                                                                                         Hi, Wi, Hp, Wp, taps, lut);
  } else if (kmax == 0) stem_canvas_u8_kernel<0><<<grid, 256, 0, st>>>(frames, cv, B, Ctot, c0, C, Hs, Ws, Hi, Wi, Hp, Wp, nrm, nullptr);
  else if (kmax == 3) stem_canvas_u8_kernel<3><<<grid, 256, 0, st>>>(frames, cv, B, Ctot, c0, C, Hs, Ws, Hi, Wi, Hp, Wp, nrm, taps);
  else stem_canvas_u8_kernel<9><<<grid, 256, 0, st>>>(frames, cv, B, Ctot, c0, C, Hs, Ws, Hi, Wi, Hp, Wp, nrm, nullptr);
  PE_LAUNCH_CHECK();
  if (A) {  // optional explicit im2col matrix (kept for the op-level GEMM tests; the engine reads the canvas through TMA)
    const long long total = (long long)B * Ho * Wo * 28;
    stem_im2col_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const __half*>(canvas), reinterpret_cast<__half*>(A), B, Ho,
                                                              Wo, Hp, Wp);
    PE_LAUNCH_CHECK();
  }
  return PE_OK;
}

int launch_resize_frames(const unsigned char* src, float* dst, int B, int C, int Hs, int Ws, int Hd, int Wd, int round_u8,
                         cudaStream_t st) {
  const int kmax = round_u8 ? pil_kmax(Hs, Ws, Hd, Wd) : 0;
  if (kmax < 0) return PE_ERR_UNSUPPORTED;
  const long long total = (long long)B * (kmax ? 1 : C) * Hd * Wd;
  const int grid = grid_for(total, 256);
  if (kmax == 0) resize_frames_kernel<0><<<grid, 256, 0, st>>>(src, dst, B, C, Hs, Ws, Hd, Wd);
  else if (kmax == 3) resize_frames_kernel<3><<<grid, 256, 0, st>>>(src, dst, B, C, Hs, Ws, Hd, Wd);
  else resize_frames_kernel<9><<<grid, 256, 0, st>>>(src, dst, B, C, Hs, Ws, Hd, Wd);
  PE_LAUNCH_CHECK();
  return PE_OK;
}

int launch_maxpool(const void* x, void* y, int B, int H, int W, int C, cudaStream_t st, int reverse) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = (long long)B * Ho * Wo * (C / 8);
  maxpool3x3s2_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y),
                                                              B, H, W, C, Ho, Wo, reverse);
  PE_LAUNCH_CHECK();
  return PE_OK;
}

int launch_subsample2(const void* x, void* y, int B, int H, int W, int C, cudaStream_t st) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = (long long)B * Ho * Wo * (C / 8);
  subsample2_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y),
                                                           B, H, W, C, Ho, Wo);
  PE_LAUNCH_CHECK();
  return PE_OK;
}

int launch_concat_channels(const void* a, const void* b, void* y, long long pixels, int C, cudaStream_t st) {
  const long long total = pixels * (C / 8) * 2;
  concat_channels_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(a),
                                                                reinterpret_cast<const __nv_bfloat16*>(b),
                                                                reinterpret_cast<__nv_bfloat16*>(y), pixels, C);
  PE_LAUNCH_CHECK();
  return PE_OK;
}

int launch_rpn_proposals(const RpnLevels& lv, int B, int pre_topk, int post_topk, float nms_thr, float img_h, float img_w,
                         const RpnScratch& s, int max_props, float4* props, int* prop_count, cudaStream_t st) {
  if (pre_topk > kTopkCap || pre_topk < 1) return PE_ERR_UNSUPPORTED;
  static DeviceOnce topk_once;
  if (topk_once.needed()) {
    PE_CUDA_CHECK(cudaFuncSetAttribute(rpn_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTopkSmemBytes));
    topk_once.mark();
  }
  rpn_topk_kernel<<<dim3(kRpnLevels, B), kTopkThreads, kTopkSmemBytes, st>>>(lv, pre_topk, img_h, img_w, s.cand_box, s.cand_score, s.cand_valid,
                                                                             s.cand_count);
  PE_LAUNCH_CHECK();
  const size_t smem = 1024 * 32 * sizeof(unsigned);
  static DeviceOnce scan_once;
  if (scan_once.needed()) {
    PE_CUDA_CHECK(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    scan_once.mark();
  }
  nms_mask_kernel<<<dim3(B * kRpnLevels, 1024 / kNmsRowsPerBlock), kNmsThreads, 0, st>>>(s.cand_box, s.cand_count, kTopkCap, nms_thr, s.nms_mask);
  PE_LAUNCH_CHECK();
  nms_scan_kernel<<<B * kRpnLevels, 1024, smem, st>>>(s.nms_mask, s.cand_valid, s.cand_count, kTopkCap, s.keep_idx, s.keep_count);
  PE_LAUNCH_CHECK();
  rpn_merge_kernel<<<B, 1024, 0, st>>>(s.cand_box, s.cand_score, s.keep_idx, s.keep_count, post_topk, max_props, props, prop_count);
  PE_LAUNCH_CHECK();
  return PE_OK;
}

int launch_roi_align(const RoiLevels& fl, const float4* props, const int* prop_count, int B, int max_props, int C, void* out,
                     cudaStream_t st) {
  if (C % 256) return PE_ERR_UNSUPPORTED;  // lanes cover the channels 8 at a time, whole warps per pass
  // PE_ROI_OCC: resident blocks per SM the kernel is compiled for.  Register prefetch (PE_ROI_ASYNC=0): 633 us (3) vs 800 us (2) per
  // 16 000 ROIs; cp.async ring (default): 3 | 4 blocks, see profiles/README.md
  static const int async_env = [] { const char* e = getenv("PE_ROI_ASYNC"); return e ? atoi(e) : 1; }();
  static const int occ = [] { const char* e = getenv("PE_ROI_OCC"); return e ? atoi(e) : 3; }();
  const dim3 grid((unsigned)max_props, (unsigned)B);
  const size_t ring_bytes = (size_t)7 * kRoiRingSlots * 32 * sizeof(uint4);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (async_env && occ >= 4) roi_align_kernel<4, true><<<grid, 224, ring_bytes, st>>>(fl, props, prop_count, B, max_props, C, o);
  else if (async_env && occ == 2) roi_align_kernel<2, true><<<grid, 224, ring_bytes, st>>>(fl, props, prop_count, B, max_props, C, o);
  else if (async_env) roi_align_kernel<3, true><<<grid, 224, ring_bytes, st>>>(fl, props, prop_count, B, max_props, C, o);
  else if (occ == 3) roi_align_kernel<3, false><<<grid, 224, 0, st>>>(fl, props, prop_count, B, max_props, C, o);
  else roi_align_kernel<2, false><<<grid, 224, 0, st>>>(fl, props, prop_count, B, max_props, C, o);
  PE_LAUNCH_CHECK();
  return PE_OK;
}

int launch_head_post(const float* head, int npad, const float4* props, const int* prop_count, int B, int max_props, int K,
                     const HeadParams& hp, const DetOut& out, cudaStream_t st) {
  if (hp.max_det > kMaxDet) return PE_ERR_UNSUPPORTED;
  if (K == 3) head_post_kernel<3><<<B, kHeadThreads, 0, st>>>(head, npad, props, prop_count, max_props, hp, out, 3);
  else if (K == 1) head_post_kernel<1><<<B, kHeadThreads, 0, st>>>(head, npad, props, prop_count, max_props, hp, out, 1);
  else if (K >= 2 && K <= 1000 && hp.score_thresh >= 0.5f)  // e.g. the 80 COCO classes of the rgb_only zoo model
    head_post_kernel<0><<<B, kHeadThreads, 0, st>>>(head, npad, props, prop_count, max_props, hp, out, K);
  else return PE_ERR_UNSUPPORTED;
  PE_LAUNCH_CHECK();
  return PE_OK;
}

int launch_pack(const PackIn& in, int B, int M, int K, int* offsets, float4* boxes, float* scores, int* classes, float* probs,
                float* vars, cudaStream_t st) {
  pack_offsets_kernel<<<1, 1024, 0, st>>>(in, B, M, offsets);
  PE_LAUNCH_CHECK();
  pack_copy_kernel<<<B * M, 128, 0, st>>>(in, B, M, K, offsets, boxes, scores, classes, probs, vars);
  PE_LAUNCH_CHECK();
  return PE_OK;
}

}  // namespace pe
