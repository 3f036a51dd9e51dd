// Input side of the detector path: what demo/FLIR/demo_FLIR_save_predictions.py:93-121 does
// on the CPU for every validation pair - cv2.imread of the RGB and thermal JPEGs, cv2.resize of the RGB frame to
// the thermal frame's size, and the 3-/4-/6-channel assembly - as device work:
//   pe_jpeg_*          nvJPEG decode of a batch of JPEG byte strings into interleaved BGR uint8 frames in HBM
//                      (third-party: libnvjpeg from the CUDA toolkit, loaded with dlopen so that the rest of the
//                      library does not depend on it; the reference decodes with OpenCV/libjpeg-turbo)
//   pe_resize_u8_cv    cv2.resize(src, (w, h)) for uint8 images, i.e. INTER_LINEAR with OpenCV's 11-bit fixed-point
//                      coefficients, bit for bit; reads a channel range of the source and writes a channel range of
//                      the destination, so the same call assembles the BGRT / BGRTTT inputs
#include <dlfcn.h>
#include <math.h>
#include <nvjpeg.h>
#include <stdlib.h>
#include "common.cuh"

namespace pe {
namespace {

// ---- cv2.resize, uint8, INTER_LINEAR -----------------------------------------------------------------------
// OpenCV (modules/imgproc/src/resize.cpp: resize -> resizeGeneric_<HResizeLinear<uchar,int,short,2048>,
// VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>>): per output column the source index and the weights
//   fx = float((dx + 0.5) * scale_x - 0.5); sx = floor(fx); fx -= sx; sx < 0 -> (0, fx = 0); sx >= W-1 -> (W-1, fx = 0)
//   alpha = (short)lrint((1 - fx) * 2048), (short)lrint(fx * 2048)
// per output row the same without the zeroing (the two source rows are clamped instead), then
//   row value  D = S[sx] * alpha0 + S[sx+1] * alpha1                       (int, scale 2^11)
//   output     = (((beta0 * (D0 >> 4)) >> 16) + ((beta1 * (D1 >> 4)) >> 16) + 2) >> 2
struct CvTap { int i0, i1, w0, w1; };

__device__ __forceinline__ CvTap cv_tap_x(int d, int ssize, double scale) {
  float f = (float)(__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5));
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (s < 0) { f = 0.f; s = 0; }
  if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
  CvTap t;
  t.i0 = s;
  t.i1 = min(s + 1, ssize - 1);
  t.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  t.w1 = __float2int_rn(__fmul_rn(f, 2048.f));
  return t;
}

__device__ __forceinline__ CvTap cv_tap_y(int d, int ssize, double scale) {
  float f = (float)(__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5));
  const int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  CvTap t;
  t.i0 = min(max(s, 0), ssize - 1);
  t.i1 = min(max(s + 1, 0), ssize - 1);
  t.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  t.w1 = __float2int_rn(__fmul_rn(f, 2048.f));
  return t;
}

__global__ void resize_u8_cv_kernel(const unsigned char* __restrict__ src, int src_C, int src_c0, unsigned char* __restrict__ dst,
                                    int dst_C, int dst_c0, int nC, int B, int Hs, int Ws, int Hd, int Wd) {
  // resize(): inv_scale = dsize / ssize, scale = 1. / inv_scale
  const double sx = __ddiv_rn(1.0, __ddiv_rn((double)Wd, (double)Ws)), sy = __ddiv_rn(1.0, __ddiv_rn((double)Hd, (double)Hs));
  const long long total = (long long)B * Hd * Wd;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t % Wd);
    const long long r = t / Wd;
    const int y = (int)(r % Hd), b = (int)(r / Hd);
    const CvTap tx = cv_tap_x(x, Ws, sx), ty = cv_tap_y(y, Hs, sy);
    const unsigned char* im = src + (size_t)b * Hs * Ws * src_C + src_c0;
    const unsigned char* r0 = im + (size_t)ty.i0 * Ws * src_C;
    const unsigned char* r1 = im + (size_t)ty.i1 * Ws * src_C;
    unsigned char* o = dst + ((size_t)(b * (long long)Hd + y) * Wd + x) * dst_C + dst_c0;
    for (int c = 0; c < nC; ++c) {
      const int d0 = (int)r0[(size_t)tx.i0 * src_C + c] * tx.w0 + (int)r0[(size_t)tx.i1 * src_C + c] * tx.w1;
      const int d1 = (int)r1[(size_t)tx.i0 * src_C + c] * tx.w0 + (int)r1[(size_t)tx.i1 * src_C + c] * tx.w1;
      const int v = (((ty.w0 * (d0 >> 4)) >> 16) + ((ty.w1 * (d1 >> 4)) >> 16) + 2) >> 2;
      o[c] = (unsigned char)min(max(v, 0), 255);
    }
  }
}

// ---- nvJPEG through dlopen ------------------------------------------------------------------------------------
struct NvjpegApi {
  void* lib = nullptr;
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
  nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*,
                           cudaStream_t) = nullptr;
  bool ok = false;
};

const NvjpegApi& nvjpeg_api() {
  static NvjpegApi api = [] {
    NvjpegApi a;
    const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so"};
    for (const char* n : names) {
      a.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (a.lib) break;
    }
    if (!a.lib) return a;
    auto sym = [&](const char* n) { return dlsym(a.lib, n); };
    a.CreateSimple = reinterpret_cast<decltype(a.CreateSimple)>(sym("nvjpegCreateSimple"));
    a.Destroy = reinterpret_cast<decltype(a.Destroy)>(sym("nvjpegDestroy"));
    a.JpegStateCreate = reinterpret_cast<decltype(a.JpegStateCreate)>(sym("nvjpegJpegStateCreate"));
    a.JpegStateDestroy = reinterpret_cast<decltype(a.JpegStateDestroy)>(sym("nvjpegJpegStateDestroy"));
    a.GetImageInfo = reinterpret_cast<decltype(a.GetImageInfo)>(sym("nvjpegGetImageInfo"));
    a.Decode = reinterpret_cast<decltype(a.Decode)>(sym("nvjpegDecode"));
    a.ok = a.CreateSimple && a.Destroy && a.JpegStateCreate && a.JpegStateDestroy && a.GetImageInfo && a.Decode;
    return a;
  }();
  return api;
}

}  // namespace
}  // namespace pe

struct pe_jpeg_decoder {
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
};

extern "C" PE_API int pe_resize_u8_cv(const uint8_t* src, int src_channels, int src_c0, uint8_t* dst, int dst_channels, int dst_c0,
                                      int channels, int B, int src_h, int src_w, int dst_h, int dst_w, void* stream) {
  if (B < 0 || channels < 0 || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0) return PE_ERR_INVALID_ARGUMENT;
  if (src_c0 < 0 || dst_c0 < 0 || src_c0 + channels > src_channels || dst_c0 + channels > dst_channels) return PE_ERR_INVALID_ARGUMENT;
  if (B == 0 || channels == 0) return PE_OK;
  if (!src || !dst) return PE_ERR_INVALID_ARGUMENT;
  const long long total = (long long)B * dst_h * dst_w;
  const long long want = (total + 255) / 256;
  const long long cap = (long long)pe::sm_count() * 16;
  const int grid = (int)(want < cap ? want : cap);
  pe::resize_u8_cv_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, src_channels, src_c0, dst, dst_channels, dst_c0,
                                                                                   channels, B, src_h, src_w, dst_h, dst_w);
  PE_LAUNCH_CHECK();
  return PE_OK;
}

extern "C" PE_API int pe_jpeg_create(pe_jpeg_decoder** out) {
  if (!out) return PE_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  const pe::NvjpegApi& api = pe::nvjpeg_api();
  if (!api.ok) return PE_ERR_UNSUPPORTED;  // libnvjpeg not found
  pe_jpeg_decoder* d = new pe_jpeg_decoder();
  if (api.CreateSimple(&d->handle) != NVJPEG_STATUS_SUCCESS) { delete d; return PE_ERR_CUDA; }
  if (api.JpegStateCreate(d->handle, &d->state) != NVJPEG_STATUS_SUCCESS) { api.Destroy(d->handle); delete d; return PE_ERR_CUDA; }
  *out = d;
  return PE_OK;
}

extern "C" PE_API void pe_jpeg_destroy(pe_jpeg_decoder* d) {
  if (!d) return;
  const pe::NvjpegApi& api = pe::nvjpeg_api();
  if (api.ok) {
    if (d->state) api.JpegStateDestroy(d->state);
    if (d->handle) api.Destroy(d->handle);
  }
  delete d;
}

extern "C" PE_API int pe_jpeg_image_info(pe_jpeg_decoder* d, const uint8_t* data, size_t bytes, int* height, int* width, int* components) {
  if (!d || !data || !height || !width) return PE_ERR_INVALID_ARGUMENT;
  const pe::NvjpegApi& api = pe::nvjpeg_api();
  int comps = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
  nvjpegChromaSubsampling_t ss;
  if (api.GetImageInfo(d->handle, data, bytes, &comps, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS) return PE_ERR_INVALID_ARGUMENT;
  *height = hs[0];
  *width = ws[0];
  if (components) *components = comps;
  return PE_OK;
}

extern "C" PE_API int pe_jpeg_decode_batch(pe_jpeg_decoder* d, const uint8_t* const* data, const size_t* bytes, int n, uint8_t* frames,
                                           int height, int width, void* stream) {
  if (!d || n < 0 || height <= 0 || width <= 0) return PE_ERR_INVALID_ARGUMENT;
  if (n == 0) return PE_OK;
  if (!data || !bytes || !frames) return PE_ERR_INVALID_ARGUMENT;
  const pe::NvjpegApi& api = pe::nvjpeg_api();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int i = 0; i < n; ++i) {
    int h = 0, w = 0;
    const int s = pe_jpeg_image_info(d, data[i], bytes[i], &h, &w, nullptr);
    if (s != PE_OK) return s;
    if (h != height || w != width) return PE_ERR_INVALID_ARGUMENT;  // every frame of a batch has the stated size
    nvjpegImage_t img;
    for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) { img.channel[c] = nullptr; img.pitch[c] = 0; }
    img.channel[0] = frames + (size_t)i * height * width * 3;
    img.pitch[0] = (size_t)width * 3;
    // interleaved BGR, the layout cv2.imread returns; a grey-scale JPEG is replicated into the three channels
    if (api.Decode(d->handle, d->state, data[i], bytes[i], NVJPEG_OUTPUT_BGRI, &img, st) != NVJPEG_STATUS_SUCCESS) return PE_ERR_CUDA;
    // One nvjpegJpegState owns the pinned staging buffers of the entropy-decoded coefficients; the next nvjpegDecode
    // on the same state refills them on the host while the previous image's copies may still be in flight (seen as
    // rare corrupt frames), so the state is drained after every image.  This is the one entry point of the library that
    // synchronises its stream; the decode is host-synchronous (CPU Huffman stage) anyway.
    PE_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  return PE_OK;
}
