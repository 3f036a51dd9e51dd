// Shared helpers for libprobenb200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/probenb200.h"

namespace pe {

void set_last_error(cudaError_t e, const char* where);

#define PE_CUDA_CHECK(expr)                                    \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) {                                   \
      ::pe::set_last_error(_e, #expr);                         \
      return PE_ERR_CUDA;                                      \
    }                                                          \
  } while (0)

#define PE_LAUNCH_CHECK()                                      \
  do {                                                         \
    cudaError_t _e = cudaGetLastError();                       \
    if (_e != cudaSuccess) {                                   \
      ::pe::set_last_error(_e, "kernel launch");               \
      return PE_ERR_CUDA;                                      \
    }                                                          \
  } while (0)

constexpr unsigned kFullMask = 0xffffffffu;

int sm_count();  // of the CURRENT device (cached per device)

// Per-device, thread-safe "do once" guard for cudaFuncSetAttribute(MaxDynamicSharedMemorySize): the attribute belongs to the
// (function, device) pair, so a process that drives several GPUs - or two host threads racing through the first launch -
// must set it once per device.  Returns true when the caller has to perform the setup for the current device.
struct DeviceOnce {
  unsigned long long done[2] = {0ull, 0ull};  // bit per device ordinal (up to 128 devices)
  bool needed() const {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 128) return true;
    return !((__atomic_load_n(&done[dev >> 6], __ATOMIC_ACQUIRE) >> (dev & 63)) & 1ull);
  }
  void mark() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 128) return;
    __atomic_fetch_or(&done[dev >> 6], 1ull << (dev & 63), __ATOMIC_RELEASE);
  }
};

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

// conv_gemm.cu
// optional second A operand of a 1x1 conv: y = act([x | x2] . w + bias) with x2 read at `stride` (bottleneck conv3 + projection
// shortcut of the first block of a stage as one GEMM, resnet.py:205-221)
struct ConvSecondInput { const void* x; int Cin, H, W, stride; };
// optional chained 1x1 conv on the layer's own output tile: y2 = act(y . w2 + bias2), w2 = [N][Cout] bf16, y2 = [.., N] bf16
// (a bottleneck's conv3 followed by the NEXT block's conv1, resnet.py:205-221, in one kernel)
// bn: main tile width, 256 (one main accumulator stage) | 128 (two stages, deferred chained GEMM / epilogue) | 0 = default
// out_fp32: narrow fp32 chain (N = 16, y2 = [.., 16] fp32; the RPN head's objectness | deltas on top of its 3x3 conv): the main
// output y is NOT stored
// y2 (narrow fp32 chain only, optional): dense float4 copy of the first three chained outputs per pixel (the RPN objectness
// logits | 0), the stream rpn_topk_kernel selects from
struct ConvChain { const void* w; const float* bias; void* y; int N; int relu; int bn; int out_fp32; void* y2; };
int conv2d_launch(const pe_conv_desc& d, const void* x, const void* w, const float* bias, const void* residual, void* y,
                  cudaStream_t st, const ConvSecondInput* x2 = nullptr, int reverse = 0, const ConvChain* chain = nullptr);

int conv_stem_launch(const void* canvas, const void* w, const float* bias, void* y, int B, int Hc, int Wc, cudaStream_t st, int reverse = 0);

}  // namespace pe
