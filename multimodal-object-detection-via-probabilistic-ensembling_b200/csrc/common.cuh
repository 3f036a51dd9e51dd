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

int sm_count();

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

// conv_gemm.cu
// optional second A operand of a 1x1 conv: y = act([x | x2] . w + bias) with x2 read at `stride` (bottleneck conv3 + projection
// shortcut of the first block of a stage as one GEMM, resnet.py:205-221)
struct ConvSecondInput { const void* x; int Cin, H, W, stride; };
int conv2d_launch(const pe_conv_desc& d, const void* x, const void* w, const float* bias, const void* residual, void* y,
                  cudaStream_t st, const ConvSecondInput* x2 = nullptr);

int conv_stem_launch(const void* canvas, const void* w, const float* bias, void* y, int B, int Hc, int Wc, cudaStream_t st);

}  // namespace pe
