// Implicit-GEMM convolution / linear layer on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Replaces the cuDNN / cuBLAS calls behind the reference's detectron2.layers.Conv2d (+FrozenBatchNorm2d,
// folded offline) and nn.Linear on the detector path: ResNet bottlenecks (modeling/backbone/resnet.py:160-221),
// FPN lateral/output convs (modeling/backbone/fpn.py:127-137), the RPN head (proposal_generator/rpn.py:74-85)
// and the box head FCs (roi_heads/box_head.py:73-81, fast_rcnn.py:531-545).
//
// D[pixels, Cout] = sum_{taps, Cin} A[pixels(+tap), Cin] * W[Cout, tap, Cin]; activations NHWC bf16,
// weights [Cout][KH][KW][Cin] bf16, fp32 accumulation in TMEM.
//   * M tile = a TH x TW spatial patch of 128 output pixels of one image; for every filter tap the
//     producer warp issues ONE 4-D TMA box load (64 ch, TW, TH, 1) at the shifted coordinate, so zero
//     padding, image borders and ragged tiles all come from TMA out-of-bounds zero fill.  Stride-2 1x1
//     convs read through a strided tensor map.  Linear layers are the H=1 case.
//   * B tile = (64 k, BLOCK_N) box of the K-major weight matrix.  Both land in 128B-swizzled K-major smem
//     tiles that tcgen05.mma consumes directly through shared-memory descriptors.
//   * warp 0: TMA producer, warp 1: MMA issuer (one elected lane), warp 2: TMEM allocator,
//     warps 4-7: epilogue (tcgen05.ld -> +bias, +residual / nearest-2x-upsampled residual, ReLU -> bf16/fp32
//     NHWC stores).  Two TMEM accumulator stages overlap the epilogue of tile i with the mainloop of i+1.
//   * persistent grid: min(#tiles, #SMs) CTAs, static round-robin tile schedule.
#include <cuda.h>
#include <stdlib.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace pe {
namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // one 128-byte swizzle atom of bf16
constexpr int kUmmaK = 16;
constexpr int kHaloTH = 16, kHaloTW = 8;                   // halo-mode patch: 16 rows of 8 pixels (one 8-row UMMA atom per patch row)
constexpr int kHaloRows = (kHaloTH + 2) * (kHaloTW + 2);  // 180 pixels
constexpr int kHaloBytes = 24 * 1024;                      // 180 * 128 B = 23040 B, rounded up so that every ring slot stays 1024 B aligned
constexpr int kMaxBStages = 20;
constexpr int kStemRowBytes = (2 * kBlockM + 6) * 8;      // canvas bytes feeding 128 neighbouring outputs of one filter row: 2096
constexpr int kStemRowPitch = 2112;                         // smem pitch of those segments (16 B aligned)
constexpr int kStemRowStages = 6;
constexpr int kStemWBytes = 7 * 64 * 64;                   // resident stem filter bank: 7 rows x [64 x 32] fp16, 64B-swizzled
constexpr int kEpilogueWarp0 = 4;
constexpr int kEpilogueWarps = 8;   // two warps per TMEM lane quarter (one per scheduler pair): each owns half of the columns
constexpr int kThreads = 32 * (kEpilogueWarp0 + kEpilogueWarps);
constexpr uint32_t kWatchdogPolls = 1u << 27;  // mbarrier polls before trapping (debug safety net)

struct ConvArgs {
  int N, Ho, Wo, Cin, Cout;
  int KH, KW, pad;
  int TH, TW, tiles_h, tiles_w, tiles_n;
  uint32_t mul_tiles_n, mul_tiles_w, mul_tiles_h;  // fast_div multipliers of the tile decomposition
  int k_chunks;  // ceil(Cin / 64)
  // chained 1x1 (bottleneck conv3 -> the NEXT block's conv1 in the same kernel): the bf16 output sub-tiles that the epilogue
  // stages in shared memory for the TMA store are exactly K-major 128B-swizzled A tiles, so a second tcgen05 GEMM consumes them
  // in place: acc2[128 px, chain_n] += out_tile[:, 64-channel chunk] . W2[chunk]; its epilogue (bias2, ReLU) is stored through
  // map_out2.  The next block's conv1 launch and its re-read of this layer's output from HBM disappear.
  int chain_n;        // 0 = off; 64 | 128 | 256 output channels of the chained conv
  int chain_relu;
  int b2_stages;      // ring depth of the chained weight chunks ([chain_n x 64] bf16 each)
  const float* chain_bias;
  // narrow fp32 chain (RPN head: 3x3 conv + ReLU -> objectness | deltas, rpn.py:74-85): chain_n = 16, the chained output goes
  // straight to chain_out [pixels][16] fp32, the MAIN output is never stored (nobody else reads the hidden tensor), the chained
  // weights stay resident, the chained accumulator ALIASES the first columns of the tile's own main TMEM stage (free once the
  // epilogue has staged sub-tile 0), and the MMA warp issues the chained MMAs of tile i opportunistically between the
  // k-iterations of tile i + 1 - both main stages stay available and the mainloop never waits for an epilogue.
  int chain_fp32;
  float* chain_out;
  float* chain_out2;  // optional dense [pixels][4] copy of chained outputs 0..2 | 0 (the objectness logits)
  int reverse;   // walk the tiles from the last to the first: consecutive layers alternate direction, so a layer starts with
                 // the part of its input that the previous layer wrote last and that is still in the 126 MB L2
  int k_chunks1; // dual-input 1x1 (bottleneck conv3 + projection shortcut as ONE GEMM over K = [t2 | x]): chunks [0, k_chunks1) come
                 // from map_a, the rest from map_a2 (the block input, strided for stride-2 stages); == k_chunks otherwise
  int relu, residual_mode, out_fp32, in_fp16;
  int halo;             // 3x3: one (TH+2)x(TW+2) halo tile per K chunk in smem, the 9 taps are shifted UMMA descriptors
  int a_stages, b_stages, b_resident;  // halo mode: A-halo ring / weight-tile ring depths; weights stay resident if they fit
  int stem_mode;        // 7x7/2 stem straight from the padded HWC4 canvas: 1 = 5-D TMA windows, 2 = row segments (see conv_stem_launch)
  const unsigned char* canvas;  // stem_mode 2
  int canvas_hp, canvas_wp;
  int stages, io_bufs;  // smem pipeline depth / number of 16 KB epilogue staging buffers (runtime split of the smem budget)
  int res_H, res_W;  // residual spatial size (mode 2: the coarser map)
  const float* bias;
  const __nv_bfloat16* residual;
  void* out;
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Suspend-time hint (ns) of mbarrier.try_wait: a waiting thread is parked by the hardware until the phase completes or the hint
// expires instead of re-issuing the poll every few cycles (the polls of 4 role warps + 8 epilogue warps cost issue slots and,
// on a power-capped part, clock).  0 = plain try_wait.  Set once per device from PE_CONV_WAIT_HINT (default below).
__constant__ uint32_t g_wait_hint_ns;

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  const uint32_t hint = g_wait_hint_ns;
  if (hint) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t polls = 0;
  const uint32_t limit = g_wait_hint_ns ? (1u << 23) : kWatchdogPolls;  // hinted polls last up to the hint: same wall-clock bound
  while (!mbar_try_wait(bar, parity)) {
    if (++polls > limit) __trap();  // a lost arrive would otherwise hang the GPU box
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// One lane of a converged warp (the same lane every time).  The producer and MMA warps run their loops warp-uniformly and
// predicate only the TMA / tcgen05 instructions with this: loop state then lives in uniform registers, and the issue path of
// an MMA is a handful of instructions (with the whole role inside `if (lane == 0)` the compiler emitted an
// ELECT / R2UR.BROADCAST sequence per operand and ~75 instructions per 4 MMAs, which made N <= 128 layers issue-bound).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// n / d for n * d < 2^32 as one multiply-high (mul = ceil(2^32 / d); d == 1 is encoded as mul == 0)
__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint32_t mul) { return mul ? __umulhi(n, mul) : n; }

// K-major, 128B-swizzled smem tile: rows of 128 B, 8-row groups 1024 B apart (cute UMMA::SmemDescriptor).
// Un-swizzled K-major operand whose rows are only 16 B apart (stem row mode): a core matrix is 8 rows x 16 B =
// 128 contiguous bytes, the next 16-byte K chunk of the same rows lies 16 B further (leading byte offset), the next
// 8 rows 128 B further (stride byte offset).  Consecutive rows therefore OVERLAP in memory - exactly the overlap of
// neighbouring 7x7/2 windows on one image row.
__device__ __forceinline__ uint64_t umma_smem_desc_rows16(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(16 >> 4) << 16;    // leading byte offset: K chunk to K chunk
  d |= (uint64_t)(128 >> 4) << 32;   // stride byte offset: 8-row group to 8-row group
  d |= (uint64_t)1 << 46;
  return d;                          // layout type 0 = SWIZZLE_NONE
}
// sw64: rows of 64 B, 8-row groups 512 B apart, SWIZZLE_64B (the stem's 32-element K chunks).
// Halo-mode A operand: the 128 rows are 16 groups of 8 consecutive halo pixels (128 B apart); successive groups
// start (TW+2)*128 = 1280 B apart and the window origin is only 128-byte aligned.  Measured on B200: the tensor
// core applies the 128B swizzle to the ABSOLUTE shared-memory address bits (exactly what TMA wrote), so the
// descriptor's base_offset field must stay 0 (setting it to the origin's row phase gives wrong results).
__device__ __forceinline__ uint64_t umma_smem_desc_halo(uint32_t smem_addr, int use_base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(((kHaloTW + 2) * 128) >> 4) << 32;      // stride byte offset between 8-row groups = 1280 B
  d |= (uint64_t)1 << 46;
  if (use_base_offset) d |= (uint64_t)((smem_addr >> 7) & 7) << 49;  // base offset
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, bool sw64 = false) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);            // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                                 // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)((sw64 ? 512 : 1024) >> 4) << 32;        // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                                 // descriptor version (Blackwell)
  d |= (uint64_t)(sw64 ? 4 : 2) << 61;                    // SWIZZLE_64B : SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: (bf16 | fp16) x same -> fp32, both operands K-major, M x N tile.
__host__ __device__ constexpr uint32_t umma_instr_desc(int M, int N, bool fp16) {
  return (1u << 4) | ((fp16 ? 0u : 1u) << 7) | ((fp16 ? 0u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// ---------------------------------------------------------------------------------------------- configuration
constexpr int kIoBytes = kBlockM * 128;  // one 128-row x 64-channel bf16 staging tile (128B-swizzled)
constexpr int kCoarseBytes = 32 * 128;   // 32 coarse pixels x 64 channels: the FPN top-down residual of one sub-tile
constexpr int kMaxIoBufs = 8;

template <int BLOCK_N, bool kStaged>
struct TileCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kMaxStages = BLOCK_N >= 256 ? 4 : (BLOCK_N >= 128 ? 6 : 8);
  // 128-wide staged tiles may run chained (two main stages in columns 0..255, the chained accumulator in 256..511)
  static constexpr int kTmemCols = (kStaged && BLOCK_N == 128) ? 512 : (2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N);
  // dynamic smem budget shared by the operand pipeline and (staged epilogue) the io buffers
  static constexpr int kBudget = kMaxStages * kStageBytes + (kStaged ? 2 * kIoBytes : 0);
  static constexpr int kSmemBytes = kBudget + 1024 /*alignment slack*/ + 1024 /*barriers*/;
};

// ---------------------------------------------------------------------------------------------- epilogue helpers
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// ---------------------------------------------------------------------------------------------- kernel
// kStaged (bf16 outputs, BLOCK_N >= 64): the epilogue works on 64-channel sub-tiles staged in two 16 KB
// 128B-swizzled smem buffers: the residual sub-tile is TMA-loaded into the buffer (one sub-tile ahead), each
// thread adds its accumulator row in place, and one elected thread TMA-stores the buffer (OOB rows/columns of
// ragged tiles are clipped by the TMA unit).  Otherwise (fp32 / narrow outputs) rows are written directly.
template <int BLOCK_N, bool kStaged>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
                 const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b2,
                 const __grid_constant__ CUtensorMap map_out2, const ConvArgs a) {
  using Cfg = TileCfg<BLOCK_N, kStaged>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int n_stages = a.stages;
  unsigned char* halo_b = smem + a.a_stages * kHaloBytes;  // halo mode: weight ring behind the A-halo ring
  unsigned char* stem_w = smem + kStemRowStages * Cfg::kStageBytes;  // stem row mode: resident filter bank
  unsigned char* io_stage = a.halo ? halo_b + a.b_stages * Cfg::kBBytes
                                   : (a.stem_mode == 2 ? stem_w + kStemWBytes : smem + n_stages * Cfg::kStageBytes);
  unsigned char* coarse_stage = io_stage + (kStaged ? a.io_bufs * kIoBytes : 0);
  unsigned char* b2_stage = coarse_stage + (a.residual_mode == 2 ? a.io_bufs * kCoarseBytes : 0);  // chained weight chunks
  const int b2_bytes = a.chain_n * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBudget);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + 8;
  uint64_t* tmem_full = bars + 16;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* io_ready = tmem_empty + 2;
  uint64_t* io_written = io_ready + kMaxIoBufs;
  uint64_t* a_full = io_written + kMaxIoBufs;   // halo mode: A-halo ring
  uint64_t* a_empty = a_full + 4;
  uint64_t* b_full = a_empty + 4;               // halo mode: weight-tile ring
  uint64_t* b_empty = b_full + kMaxBStages;
  uint64_t* chain_full = b_empty + kMaxBStages;   // chained accumulator stage complete (MMA -> epilogue)
  uint64_t* chain_empty = chain_full + 2;         // ... drained (epilogue -> MMA)
  uint64_t* chain_read = chain_empty + 2;         // staging buffer no longer read by the chained MMA / its own store issued
  uint64_t* b2_full = chain_read + kMaxIoBufs;
  uint64_t* b2_empty = b2_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b2_empty + 4);
  const bool chain = kStaged && (BLOCK_N == 256 || BLOCK_N == 128) && a.chain_n > 0;
  // 256-wide chained tiles: the chained accumulator takes TMEM columns 256.., ONE main stage is left and the chained GEMM /
  // epilogue of a tile follow it immediately.  128-wide chained tiles keep TWO main stages; the chained GEMM of tile i is then
  // issued AFTER the mainloop of tile i + 1 (so the mainloop overlaps tile i's epilogue as usual) and the chained epilogue of
  // an m-tile group after the first epilogue of the following tile ("deferred" order, same for every role).
  const bool defer = chain && BLOCK_N == 128;
  const bool rpn = chain && BLOCK_N == 256 && a.chain_fp32 != 0;  // narrow fp32 chain, see ConvArgs::chain_fp32
  const int n_acc = (chain && !defer && !rpn) ? 1 : 2;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = a.N * a.tiles_h * a.tiles_w;
  const int num_tiles = tiles_m * a.tiles_n;
  const int k_iters = a.stem_mode == 1 ? 7 : a.KH * a.KW * a.k_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (a.k_chunks1 < a.k_chunks) tma_prefetch_desc(&map_a2);
    if (a.chain_n) { tma_prefetch_desc(&map_b2); if (!a.chain_fp32) tma_prefetch_desc(&map_out2); }
    if (kStaged) {
      if (!a.chain_fp32) tma_prefetch_desc(&map_out);
      if (a.residual_mode) tma_prefetch_desc(&map_res);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpilogueWarps);  // one arrive per epilogue warp
    }
    for (int s = 0; s < kMaxIoBufs; ++s) {
      mbar_init(&io_ready[s], 1);
      mbar_init(&io_written[s], 32 * kEpilogueWarps);
    }
    for (int s = 0; s < 4; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kMaxBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&chain_full[s], 1); mbar_init(&chain_empty[s], kEpilogueWarps); }
    for (int s = 0; s < kMaxIoBufs; ++s) mbar_init(&chain_read[s], 1);
    for (int s = 0; s < 4; ++s) { mbar_init(&b2_full[s], 1); mbar_init(&b2_empty[s], 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: this grid may have been scheduled while the previous kernel of the stream was still
  // draining (its CTAs retire one by one; ours take the freed SMs and run the prologue above).  Everything below reads or
  // writes tensors of the layer chain, so it waits for the previous grid to complete and flush; our own dependents may be
  // scheduled as soon as every CTA of this grid got here.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // tile -> (n tile, patch column, patch row, image) without integer division
  auto decompose = [&](int tile, int& nt, int& tw, int& th, int& img) {
    if (a.reverse) tile = num_tiles - 1 - tile;
    const uint32_t mt = fast_div((uint32_t)tile, a.mul_tiles_n);
    nt = tile - (int)mt * a.tiles_n;
    const uint32_t r = fast_div(mt, a.mul_tiles_w);
    tw = (int)mt - (int)r * a.tiles_w;
    img = (int)fast_div(r, a.mul_tiles_h);
    th = (int)r - img * a.tiles_h;
  };

  // Tile schedule of this CTA: static round-robin over the (m, n) tiles; chained layers keep all n tiles of an m tile on the
  // same CTA, back to back (the chained GEMM accumulates over them), and round-robin over the m tiles instead.
  auto tile_at = [&](int it) -> int {
    if (!chain) { const int t = (int)blockIdx.x + it * (int)gridDim.x; return t < num_tiles ? t : -1; }
    const int grp = it / a.tiles_n;
    const int mt = (int)blockIdx.x + grp * (int)gridDim.x;
    return mt < tiles_m ? mt * a.tiles_n + (it - grp * a.tiles_n) : -1;
  };
  constexpr int kSubMain = BLOCK_N / 64;
  const int nsub2 = a.chain_n >> 6;                // 64-channel sub-tiles of the chained output
  auto group_end = [&](int it) { return it % a.tiles_n == a.tiles_n - 1; };
  // do the chained-output sub-tiles of a group follow the main sub-tiles of tile `it` in the CTA's staging-buffer sequence?
  auto chain_after = [&](int it) { return chain && (defer ? (it >= 1 && group_end(it - 1)) : group_end(it)); };

  if (warp == 0) {
    // ================================ TMA producer (warp-uniform loops, one elected lane issues) ================================
    const bool leader = elect_one();
    int stage = 0, hstage = 0, bs = 0, prev_n0 = 0;
    uint32_t phase = 0, hphase = 0, bphase = 0;
    for (int it = 0, tile; (tile = tile_at(it)) >= 0; ++it) {
      int nt, tw, th, img;
      decompose(tile, nt, tw, th, img);
      const int h0 = th * a.TH - a.pad, w0 = tw * a.TW - a.pad, n0 = nt * BLOCK_N;
      const bool first = it == 0;
      if (a.halo) {  // one halo tile per K chunk, then the 9 weight tiles of that chunk
        for (int kc = 0; kc < a.k_chunks; ++kc) {
          mbar_wait(&a_empty[hstage], hphase ^ 1);
          if (leader) {
            mbar_expect_tx(&a_full[hstage], kHaloRows * 128);
            tma_load_4d(&map_a, &a_full[hstage], smem + hstage * kHaloBytes, kc * kBlockK, w0, h0, img);
          }
          if (++hstage == a.a_stages) { hstage = 0; hphase ^= 1; }
          if (a.b_resident && !first) continue;  // weights were loaded with the first tile and stay
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_empty[stage], phase ^ 1);
            if (leader) {
              mbar_expect_tx(&b_full[stage], Cfg::kBBytes);
              tma_load_2d(&map_b, &b_full[stage], halo_b + stage * Cfg::kBBytes, tap * a.Cin + kc * kBlockK, n0);
            }
            if (++stage == a.b_stages) { stage = 0; phase ^= 1; }
          }
        }
        continue;
      }
      if (a.stem_mode == 2) {  // one stage per tile: the 7 canvas row segments under this run of 128 outputs
        if (first && leader) {
          mbar_expect_tx(&b_full[0], kStemWBytes);
          for (int kh = 0; kh < 7; ++kh) tma_load_2d(&map_b, &b_full[0], stem_w + kh * 4096, kh * 32, 0);
        }
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          unsigned char* sa = smem + stage * Cfg::kStageBytes;
          mbar_expect_tx(&full_bar[stage], 7 * kStemRowBytes);
          const unsigned char* src = a.canvas + (((size_t)img * a.canvas_hp + 2 * (th * a.TH)) * a.canvas_wp + 2 * (tw * a.TW)) * 8;
          const size_t row_pitch = (size_t)a.canvas_wp * 8;
#pragma unroll
          for (int kh = 0; kh < 7; ++kh) bulk_load_1d(sa + kh * kStemRowPitch, src + kh * row_pitch, kStemRowBytes, &full_bar[stage]);
        }
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
        continue;
      }
      if (a.stem_mode) {  // 7 row taps, each one 64-byte chunk per pixel: A rows (2*ho + kh) of the canvas
        for (int kh = 0; kh < 7; ++kh) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) {
            unsigned char* sa = smem + stage * Cfg::kStageBytes;
            unsigned char* sb = sa + Cfg::kABytes;
            mbar_expect_tx(&full_bar[stage], (kBlockM + BLOCK_N) * 64);
            tma_load_5d(&map_a, &full_bar[stage], sa, 0, w0, kh & 1, h0 + (kh >> 1), img);
            tma_load_2d(&map_b, &full_bar[stage], sb, kh * 32, n0);
          }
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      if (rpn && first && leader) {  // the whole chained weight matrix [16][Cout] stays resident: one chunk per 64 main channels
        for (int s2 = 0; s2 < kSubMain; ++s2) {
          mbar_expect_tx(&b2_full[s2], (uint32_t)b2_bytes);
          tma_load_2d(&map_b2, &b2_full[s2], b2_stage + s2 * b2_bytes, s2 * 64, 0);
        }
      }
      for (int kh = 0; kh < a.KH; ++kh)
        for (int kw = 0; kw < a.KW; ++kw)
          for (int kc = 0; kc < a.k_chunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (leader) {
              unsigned char* sa = smem + stage * Cfg::kStageBytes;
              unsigned char* sb = sa + Cfg::kABytes;
              mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
              if (kc < a.k_chunks1) tma_load_4d(&map_a, &full_bar[stage], sa, kc * kBlockK, w0 + kw, h0 + kh, img);
              else tma_load_4d(&map_a2, &full_bar[stage], sa, (kc - a.k_chunks1) * kBlockK, w0, h0, img);
              tma_load_2d(&map_b, &full_bar[stage], sb, (kh * a.KW + kw) * a.Cin + kc * kBlockK, n0);
            }
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
          }
      // the chained weights of an n tile's 64-channel K chunks: W2[:, n0 + 64 s .. +64].  Deferred order: tile it - 1's chunks
      // are requested after tile it's operands (that is when the MMA warp will consume them)
      auto load_b2 = [&](int n0c) {
        for (int s2 = 0; s2 < kSubMain; ++s2) {
          mbar_wait(&b2_empty[bs], bphase ^ 1);
          if (leader) {
            mbar_expect_tx(&b2_full[bs], (uint32_t)b2_bytes);
            tma_load_2d(&map_b2, &b2_full[bs], b2_stage + bs * b2_bytes, n0c + s2 * 64, 0);
          }
          if (++bs == a.b2_stages) { bs = 0; bphase ^= 1; }
        }
      };
      if (chain && !defer && !rpn) load_b2(n0);
      if (defer) {
        if (it >= 1) load_b2(prev_n0);
        prev_n0 = n0;
        if (tile_at(it + 1) < 0) load_b2(n0);  // last tile of this CTA: its chunks follow at once
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (warp-uniform loops, one elected lane issues) ================================
    const bool leader = elect_one();
    const uint32_t idesc = umma_instr_desc(kBlockM, BLOCK_N, a.in_fp16 != 0);
    const uint32_t smem0 = smem_u32(smem);
    int stage = 0, hstage = 0;
    uint32_t phase = 0, hphase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    // chained GEMM state: staging-buffer cursor (the same running sub-tile sequence the epilogue / io warp walk), weight ring,
    // chained accumulator stage
    const uint32_t idesc2 = umma_instr_desc(kBlockM, a.chain_n > 0 ? a.chain_n : 64, false);
    const int n_cacc = a.chain_n <= 128 ? 2 : 1;
    int cp = 0, bs = 0, cs = 0, prev_p = 0;
    uint32_t cpar = 0, bphase = 0, cphase = 0, prev_par = 0;
    // narrow fp32 chain (rpn): running count of chained sub-tile GEMMs issued (sub-tile g belongs to tile g / kSubMain of this
    // CTA and sits in staging buffer g % io_bufs), issued as soon as the epilogue has staged the sub-tile
    int rpn_issued = 0, rpn_p = 0, n_done = 0;
    uint32_t rpn_par = 0;
    auto rpn_step = [&](bool blocking) -> bool {
      if (blocking) mbar_wait(&io_written[rpn_p], rpn_par);
      else if (!__all_sync(0xffffffffu, mbar_try_wait(&io_written[rpn_p], rpn_par))) return false;
      constexpr int kSubDiv = kSubMain > 0 ? kSubMain : 1;  // (narrow instantiations never take this path)
      const int t = rpn_issued / kSubDiv, s2 = rpn_issued - t * kSubDiv;
      mbar_wait(&b2_full[s2], 0);  // resident weights: completes once
      tc_fence_after();
      if (leader) {
        const uint32_t d2 = tmem_base + (uint32_t)((t & 1) * BLOCK_N);  // columns [0, chain_n) of the tile's own main stage
        const uint64_t da = umma_smem_desc(smem_u32(io_stage + rpn_p * kIoBytes));
        const uint64_t db = umma_smem_desc(smem_u32(b2_stage + s2 * b2_bytes));
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k)
          umma_bf16(d2, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2, (s2 | k) != 0);
        umma_commit(&chain_read[rpn_p]);  // the io warp may hand the buffer to the next sub-tile
        if (s2 == kSubMain - 1) umma_commit(&chain_full[t & 1]);
      }
      ++rpn_issued;
      if (++rpn_p == a.io_bufs) { rpn_p = 0; rpn_par ^= 1; }
      return true;
    };
    for (int it = 0, tile; (tile = tile_at(it)) >= 0; ++it) {
      const bool first = it == 0;
      // the stage this tile needs was used by tile it - 2 and is released by ITS chained epilogue: all its chained GEMMs must be out
      if (rpn) while (rpn_issued < kSubMain * (it - 1)) rpn_step(true);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      if (a.stem_mode == 2) {
        if (first) mbar_wait(&b_full[0], 0);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader) {
          const uint64_t da0 = umma_smem_desc_rows16(smem0 + stage * Cfg::kStageBytes);
          const uint64_t db0 = umma_smem_desc(smem_u32(stem_w), true);
#pragma unroll
          for (int kh = 0; kh < 7; ++kh) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma_bf16(d_tmem, da0 + (uint64_t)((kh * kStemRowPitch + 32 * k) >> 4), db0 + (uint64_t)((kh * 4096) >> 4) + (uint64_t)(2 * k),
                        idesc, (kh | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          umma_commit(&tmem_full[acc]);
        }
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
        if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      if (a.halo) {
        for (int kc = 0; kc < a.k_chunks; ++kc) {
          mbar_wait(&a_full[hstage], hphase);
          tc_fence_after();
          const uint64_t da0 = umma_smem_desc_halo(smem0 + hstage * kHaloBytes, a.halo & 2);
          if (a.b_resident) {
            // the whole filter bank is resident (slot = kc * 9 + tap): 36 MMAs back to back, no barrier in between
            if (first && kc == 0) {
              for (int t = 0; t < a.b_stages; ++t) mbar_wait(&b_full[t], 0);
              tc_fence_after();
            }
            if (leader) {
              const uint64_t db0 = umma_smem_desc(smem_u32(halo_b + kc * 9 * Cfg::kBBytes));
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const uint64_t da = da0 + (uint64_t)((((tap / 3) * (kHaloTW + 2) + (tap % 3)) * 128) >> 4);
                const uint64_t db = db0 + (uint64_t)((tap * Cfg::kBBytes) >> 4);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k)
                  umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kc | tap | k) != 0);
              }
              umma_commit(&a_empty[hstage]);
              if (kc == a.k_chunks - 1) umma_commit(&tmem_full[acc]);
            }
          } else {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&b_full[stage], phase);
              tc_fence_after();
              if (leader) {
                const uint64_t da = da0 + (uint64_t)((((tap / 3) * (kHaloTW + 2) + (tap % 3)) * 128) >> 4);
                const uint64_t db = umma_smem_desc(smem_u32(halo_b + stage * Cfg::kBBytes));
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k)
                  umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kc | tap | k) != 0);
                umma_commit(&b_empty[stage]);
              }
              if (++stage == a.b_stages) { stage = 0; phase ^= 1; }
            }
            if (leader) {
              umma_commit(&a_empty[hstage]);
              if (kc == a.k_chunks - 1) umma_commit(&tmem_full[acc]);
            }
          }
          if (++hstage == a.a_stages) { hstage = 0; hphase ^= 1; }
        }
        if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      const bool sw64 = a.stem_mode == 1;
      for (int ki = 0; ki < k_iters; ++ki) {
        if (rpn && rpn_issued < kSubMain * it) rpn_step(false);  // a staged sub-tile of the previous tile, if one is ready
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader) {
          const uint32_t sa = smem0 + stage * Cfg::kStageBytes;
          const uint64_t da = umma_smem_desc(sa, sw64), db = umma_smem_desc(sa + Cfg::kABytes, sw64);
          const int n_mma = sw64 ? 2 : kBlockK / kUmmaK;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 32 B (= 2 x 16 B) along K inside the swizzle atom
            if (k < n_mma) umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (ki | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
          if (ki == k_iters - 1) umma_commit(&tmem_full[acc]);
        }
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
      if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
      n_done = it + 1;
      // ---- chained GEMM: acc2 += staged out sub-tiles (A, in place) x W2 chunks; t = the tile whose sub-tiles are consumed,
      //      (p0, par0) = staging-buffer cursor at its first sub-tile
      auto chain_part = [&](int t, int p0, uint32_t par0) {
        const int grp_pos = t % a.tiles_n;            // position of the n tile inside its m-tile group
        const uint32_t d2 = tmem_base + 256u + (uint32_t)(cs * a.chain_n);
        if (grp_pos == 0) {                           // chained accumulator stage drained by the chained epilogue
          mbar_wait(&chain_empty[cs], cphase ^ 1);
          tc_fence_after();
        }
        for (int s2 = 0; s2 < kSubMain; ++s2) {
          mbar_wait(&io_written[p0], par0);           // the epilogue finished this 128 x 64 bf16 sub-tile (swizzled, K-major)
          mbar_wait(&b2_full[bs], bphase);
          tc_fence_after();
          if (leader) {
            const uint64_t da = umma_smem_desc(smem_u32(io_stage + p0 * kIoBytes));
            const uint64_t db = umma_smem_desc(smem_u32(b2_stage + bs * b2_bytes));
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma_bf16(d2, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2, (grp_pos | s2 | k) != 0);
            umma_commit(&b2_empty[bs]);
            umma_commit(&chain_read[p0]);             // the io warp may recycle the buffer once these MMAs retired
          }
          if (++bs == a.b2_stages) { bs = 0; bphase ^= 1; }
          if (++p0 == a.io_bufs) { p0 = 0; par0 ^= 1; }
        }
        if (grp_pos == a.tiles_n - 1) {
          if (leader) umma_commit(&chain_full[cs]);
          if (++cs == n_cacc) { cs = 0; cphase ^= 1; }
        }
      };
      auto advance = [&](int n) { for (int i = 0; i < n; ++i) if (++cp == a.io_bufs) { cp = 0; cpar ^= 1; } };
      if (chain && !rpn) {
        const int p_main = cp;                        // cursor at this tile's first main sub-tile
        const uint32_t par_main = cpar;
        advance(kSubMain);
        if (chain_after(it)) advance(nsub2);          // the buffers the chained epilogue fills
        if (!defer) {
          chain_part(it, p_main, par_main);
        } else {
          if (it >= 1) chain_part(it - 1, prev_p, prev_par);
          prev_p = p_main; prev_par = par_main;
          if (tile_at(it + 1) < 0) chain_part(it, p_main, par_main);  // last tile of this CTA
        }
      }
    }
    if (rpn) while (rpn_issued < kSubMain * n_done) rpn_step(true);  // the last tile's (and any left-over) chained GEMMs
  } else if (kStaged && warp == 3) {
    // ================================ io warp (staged epilogue) ================================
    // Owns the 16 KB staging buffers: hands them to the epilogue warps (io_ready: free, or - on residual layers -
    // filled with the TMA-loaded residual sub-tile, running up to io_bufs sub-tiles ahead) and TMA-stores them once
    // all epilogue threads have written their rows (io_written).  The epilogue warps never wait on a store.
    // The sub-tiles of a CTA form one running sequence; chained layers append the chained output's sub-tiles after the last
    // n tile of every m-tile group.  A buffer is recycled when its store has read it AND (chained layers) the chained MMA has.
    const bool leader = elect_one();
    const int rmode = a.residual_mode;
    const int R = a.io_bufs;
    if (rpn) {  // nothing is stored: a staging buffer goes back to the epilogue as soon as the chained MMA has read it
      int n_tiles = 0;
      while (tile_at(n_tiles) >= 0) ++n_tiles;
      const int total = n_tiles * kSubMain;
      for (int i = 0; i < R && i < total; ++i)
        if (leader) mbar_arrive(&io_ready[i]);
      int p = 0;
      uint32_t par = 0;
      for (int g = 0; g + R < total; ++g) {
        mbar_wait(&chain_read[p], par);
        if (leader) mbar_arrive(&io_ready[p]);
        if (++p == R) { p = 0; par ^= 1; }
      }
    } else {
    const uint32_t res_bytes = rmode == 2 ? kCoarseBytes : kIoBytes;
    auto sub_count = [&](int n0) { const int left = (a.Cout - n0) >> 6; return left < kSubMain ? left : kSubMain; };
    // ---- cursor of the next sub-tile to be made ready
    int rd_it = 0, rd_tile = tile_at(0), rd_sub = 0, rd_chain = 0, rd_p = 0;
    bool rd_tail = false;  // deferred order: the last group's chained sub-tiles follow the last tile
    int rd_n0 = 0, rd_w0 = 0, rd_h0 = 0, rd_img = 0;
    auto rd_coords = [&]() {
      int nt, tw, th;
      decompose(rd_tile, nt, tw, th, rd_img);
      rd_n0 = nt * BLOCK_N; rd_w0 = tw * a.TW; rd_h0 = th * a.TH;
    };
    if (rd_tile >= 0) rd_coords();
    auto make_ready = [&]() {
      if (rd_chain > 0) {          // a chained-output sub-tile: no residual, the buffer only has to be free
        if (leader) mbar_arrive(&io_ready[rd_p]);
        if (++rd_p == R) rd_p = 0;
        --rd_chain;
        return;
      }
      if (rd_tile < 0) {
        if (defer && !rd_tail && rd_it > 0) {  // past the last tile: the last group's chained sub-tiles
          rd_tail = true;
          rd_chain = nsub2 - 1;
          if (leader) mbar_arrive(&io_ready[rd_p]);
          if (++rd_p == R) rd_p = 0;
        }
        return;
      }
      if (leader) {
        if (rmode == 0) {
          mbar_arrive(&io_ready[rd_p]);
        } else {
          mbar_expect_tx(&io_ready[rd_p], res_bytes);
          if (rmode == 1) tma_load_4d(&map_res, &io_ready[rd_p], io_stage + rd_p * kIoBytes, rd_n0 + rd_sub * 64, rd_w0, rd_h0, rd_img);
          else tma_load_4d(&map_res, &io_ready[rd_p], coarse_stage + rd_p * kCoarseBytes, rd_n0 + rd_sub * 64, rd_w0 >> 1, rd_h0 >> 1, rd_img);
        }
      }
      if (++rd_p == R) rd_p = 0;
      if (++rd_sub == sub_count(rd_n0)) {
        rd_sub = 0;
        if (chain_after(rd_it)) rd_chain = nsub2;  // a group's chained sub-tiles come next in the sequence
        rd_tile = tile_at(++rd_it);
        if (rd_tile >= 0) rd_coords();
      }
    };
    for (int i = 0; i < R; ++i) make_ready();
    int p = 0, pprev = 0;
    uint32_t par = 0, parprev = 0;
    bool any = false;
    auto store_one = [&](const CUtensorMap* map, int c0, int w0, int h0, int img, bool chained_out) {
      mbar_wait(&io_written[p], par);
      if (leader) {
        tma_store_4d(map, io_stage + p * kIoBytes, c0, w0, h0, img);
        tma_store_commit();
        if (chain && chained_out) mbar_arrive(&chain_read[p]);  // nobody else reads a chained-output buffer
        if (any) tma_store_wait_read<1>();  // the previous store has drained its buffer: recycle it for R sub-tiles ahead
      }
      __syncwarp();
      if (any) {
        if (chain) mbar_wait(&chain_read[pprev], parprev);      // ... and the chained MMA is done with it as well
        make_ready();
      }
      any = true;
      pprev = p; parprev = par;
      if (++p == R) { p = 0; par ^= 1; }
    };
    int pw0 = 0, ph0 = 0, pimg = 0, n_its = 0;  // patch of the previous tile (deferred order: its group's chained output)
    for (int it = 0, tile; (tile = tile_at(it)) >= 0; ++it) {
      int nt, tw, th, img;
      decompose(tile, nt, tw, th, img);
      const int n0 = nt * BLOCK_N, w0 = tw * a.TW, h0 = th * a.TH;
      const int nsub = sub_count(n0);
      for (int s2 = 0; s2 < nsub; ++s2) store_one(&map_out, n0 + s2 * 64, w0, h0, img, false);
      if (chain_after(it))
        for (int c2 = 0; c2 < nsub2; ++c2)
          store_one(&map_out2, c2 * 64, defer ? pw0 : w0, defer ? ph0 : h0, defer ? pimg : img, true);
      pw0 = w0; ph0 = h0; pimg = img;
      n_its = it + 1;
    }
    if (defer && n_its > 0)
      for (int c2 = 0; c2 < nsub2; ++c2) store_one(&map_out2, c2 * 64, pw0, ph0, pimg, true);
    if (leader) tma_store_wait_all();
    }
  } else if (warp >= kEpilogueWarp0) {
    // ================================ epilogue ================================
    // warps 4-7 and 8-11: warp w may only touch TMEM lanes [32 * (w % 4), +32); the two warps of a lane quarter split
    // the columns (half 0: channels 0-31 of every 64-channel sub-tile, half 1: channels 32-63), so each scheduler has two
    // epilogue warps to interleave and the dependent TMEM -> bias -> residual -> ReLU -> smem chain of one hides the other's
    const int q = (warp - kEpilogueWarp0) & 3;  // TMEM lane quarter owned by this warp
    const int half = (warp - kEpilogueWarp0) >> 2;
    const int row = q * 32 + lane;
    const int ph = row / a.TW, pw = row - ph * a.TW;
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (kStaged) {
      const int rmode = a.residual_mode;  // 1: same-size residual tile, 2: coarser map (nearest 2x); both arrive by TMA
      const int R = a.io_bufs;
      const int crow = (ph >> 1) * (a.TW >> 1) + (pw >> 1);  // this thread's pixel in the coarse (mode 2) tile
      int p = 0;          // staging buffer of the running sub-tile
      uint32_t par = 0;   // its barrier parity
      // One 128 x 64 sub-tile: this warp's 32 accumulator columns (TMEM column tcol + half * 32) + bias (+ residual) -> ReLU ->
      // bf16 -> the swizzled staging buffer.  full_bar: accumulator-complete barrier to wait for first (or null);
      // release: barrier to arrive on once the TMEM columns have been read (or null).
      auto sub_tile = [&](uint32_t tcol, const float* bias_ptr, int rm, int relu, uint64_t* full_bar_w, uint32_t full_par,
                          uint64_t* release) {
        unsigned char* io = io_stage + p * kIoBytes;
        // this warp's 32 bias values are fetched before any wait, so their latency hides behind the barrier / TMEM loads
        float4 bv[8];
        if (bias_ptr) {
          const float4* bp = reinterpret_cast<const float4*>(bias_ptr + half * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) bv[j] = __ldg(bp + j);
        }
        if (full_bar_w) {
          mbar_wait(full_bar_w, full_par);
          tc_fence_after();
        }
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + tcol + (uint32_t)(half * 32);
        tmem_ld_32x32b_x16(taddr, v);
        tmem_ld_32x32b_x16(taddr + 16, v + 16);
        tmem_ld_wait();
        if (release) {  // accumulator fully read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(release);
        }
        mbar_wait(&io_ready[p], par);  // buffer free and (residual layers) its residual sub-tile landed
        uint4* myrow = reinterpret_cast<uint4*>(io + row * 128);
        const uint4* crs = reinterpret_cast<const uint4*>(coarse_stage + p * kCoarseBytes + crow * 128);
        // this warp's four residual chunks of the row are read up front: inside the loop every load would have to wait
        // for the previous chunk's store (same buffer, the compiler cannot prove the swizzled slots distinct)
        uint4 res[4];
        if (rm) {
#pragma unroll
          for (int j = 0; j < 4; ++j) res[j] = rm == 1 ? myrow[(half * 4 + j) ^ (row & 7)] : crs[(half * 4 + j) ^ (crow & 7)];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = half * 4 + j;
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[j * 8 + i]);
          if (bias_ptr) {
            const float4 b0 = bv[2 * j], b1 = bv[2 * j + 1];
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
          }
          uint4* slot = myrow + (c ^ (row & 7));  // 128B swizzle: 16-byte chunk index XOR (row mod 8)
          if (rm) {
            const uint4 r = res[j];
            f[0] += bf16_lo(r.x); f[1] += bf16_hi(r.x); f[2] += bf16_lo(r.y); f[3] += bf16_hi(r.y);
            f[4] += bf16_lo(r.z); f[5] += bf16_hi(r.z); f[6] += bf16_lo(r.w); f[7] += bf16_hi(r.w);
          }
          if (relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
          }
          *slot = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
        }
        fence_proxy_async();            // make the generic-proxy row writes visible to the TMA store (and the chained MMA)
        if (rpn) tc_fence_before();     // ... and order this thread's TMEM reads before the chained MMA that overwrites the columns
        mbar_arrive(&io_written[p]);    // 256 arrivals release the buffer to the io warp
        if (++p == R) { p = 0; par ^= 1; }
      };
      const int n_cacc = a.chain_n <= 128 ? 2 : 1;
      int cs = 0, n_its = 0;
      uint32_t cphase = 0;
      auto chained_epilogue = [&]() {  // chained accumulator + bias2 -> ReLU -> bf16 -> staging -> map_out2
        for (int c2 = 0; c2 < nsub2; ++c2)
          sub_tile(256u + (uint32_t)(cs * a.chain_n + c2 * 64), a.chain_bias ? a.chain_bias + c2 * 64 : nullptr, 0, a.chain_relu,
                   c2 == 0 ? &chain_full[cs] : nullptr, cphase, c2 == nsub2 - 1 ? &chain_empty[cs] : nullptr);
        if (++cs == n_cacc) { cs = 0; cphase ^= 1; }
      };
      // narrow fp32 chain: this thread's pixel x 8 of the 16 chained outputs (half 0: objectness | pad, half 1... columns 8-15)
      // from the first columns of the tile's main stage -> + bias -> fp32 row of chain_out; then the stage goes back to the MMA warp
      auto rpn_epilogue = [&](int tile, int acc_s, uint32_t par_s) {
        int nt, tw, th, img;
        decompose(tile, nt, tw, th, img);
        const int h = th * a.TH + ph, w = tw * a.TW + pw;
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
        if (a.chain_bias) {
          b0 = __ldg(reinterpret_cast<const float4*>(a.chain_bias + half * 8));
          b1 = __ldg(reinterpret_cast<const float4*>(a.chain_bias + half * 8 + 4));
        }
        mbar_wait(&chain_full[acc_s], par_s);
        tc_fence_after();
        uint32_t v[8];
        tmem_ld_32x32b_x8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc_s * BLOCK_N + half * 8), v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc_s]);
        if (h < a.Ho && w < a.Wo) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[i]);
          f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
          f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
          if (a.chain_relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
          }
          const size_t opix = ((size_t)img * a.Ho + h) * a.Wo + w;
          float4* op = reinterpret_cast<float4*>(a.chain_out + opix * 16 + half * 8);
          op[0] = make_float4(f[0], f[1], f[2], f[3]);
          op[1] = make_float4(f[4], f[5], f[6], f[7]);
          if (half == 0 && a.chain_out2) reinterpret_cast<float4*>(a.chain_out2)[opix] = make_float4(f[0], f[1], f[2], 0.f);
        }
      };
      for (int it = 0, tile; (tile = tile_at(it)) >= 0; ++it) {
        const int ptile = a.reverse ? num_tiles - 1 - tile : tile;
        const int n0 = (ptile - (int)fast_div((uint32_t)ptile, a.mul_tiles_n) * a.tiles_n) * BLOCK_N;
        const int left = (a.Cout - n0) >> 6;
        const int nsub = left < kSubMain ? left : kSubMain;
        for (int s2 = 0; s2 < nsub; ++s2)
          sub_tile((uint32_t)(acc * BLOCK_N + s2 * 64), a.bias ? a.bias + n0 + s2 * 64 : nullptr, rmode, a.relu,
                   s2 == 0 ? &tmem_full[acc] : nullptr, acc_phase, (s2 == nsub - 1 && !rpn) ? &tmem_empty[acc] : nullptr);
        if (rpn) rpn_epilogue(tile, acc, acc_phase);  // waits for the chained GEMMs the MMA warp issues during the next mainloop
        if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
        if (!rpn && chain_after(it)) chained_epilogue();  // an m-tile group is complete (deferred order: the one before this tile)
        n_its = it + 1;
      }
      if (defer && n_its > 0) chained_epilogue();
    } else {
      for (int it = 0, tile; (tile = tile_at(it)) >= 0; ++it) {
        int nt, tw, th, img;
        decompose(tile, nt, tw, th, img);
        const int h = th * a.TH + ph, w = tw * a.TW + pw, n0 = nt * BLOCK_N;
        const bool pix_ok = h < a.Ho && w < a.Wo;
        const size_t opix = ((size_t)img * a.Ho + h) * a.Wo + w;
        size_t rpix = opix;
        if (a.residual_mode == 2) rpix = ((size_t)img * a.res_H + (h >> 1)) * a.res_W + (w >> 1);
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = half * 16; c0 < BLOCK_N; c0 += 32) {  // the two warps of a lane quarter take alternate 16-column chunks
          uint32_t v[16];
          tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0), v);
          tmem_ld_wait();
          const int ch = n0 + c0;
          if (pix_ok && ch < a.Cout) {  // Cout is a multiple of 8; a 16-wide chunk may be half valid
            const int nvalid = a.Cout - ch >= 16 ? 16 : 8;
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
            if (a.bias) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                if (i < nvalid) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + ch + i));
                  f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
                }
              }
            }
            if (a.residual_mode) {
              const uint4* rp = reinterpret_cast<const uint4*>(a.residual + rpix * a.Cout + ch);
#pragma unroll
              for (int i = 0; i < 16; i += 8) {
                if (i < nvalid) {
                  const uint4 r = __ldg(rp + i / 8);
                  f[i] += bf16_lo(r.x); f[i + 1] += bf16_hi(r.x); f[i + 2] += bf16_lo(r.y); f[i + 3] += bf16_hi(r.y);
                  f[i + 4] += bf16_lo(r.z); f[i + 5] += bf16_hi(r.z); f[i + 6] += bf16_lo(r.w); f[i + 7] += bf16_hi(r.w);
                }
              }
            }
            if (a.relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            if (a.out_fp32) {
              float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + opix * a.Cout + ch);
#pragma unroll
              for (int i = 0; i < 16; i += 4)
                if (i < nvalid) op[i / 4] = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
            } else {
              uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + opix * a.Cout + ch);
#pragma unroll
              for (int i = 0; i < 16; i += 8)
                if (i < nvalid)
                  op[i / 8] = make_uint4(pack_bf16(f[i], f[i + 1]), pack_bf16(f[i + 2], f[i + 3]),
                                         pack_bf16(f[i + 4], f[i + 5]), pack_bf16(f[i + 6], f[i + 7]));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 tensor map with 128B swizzle; dims/strides innermost first, strides[i] = byte stride of dim i+1.
bool make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
              const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

void pick_patch(int Ho, int Wo, int max_tw, int* TH, int* TW) {
  long best = -1;
  int bh = 8, bw = 16;
  for (int tw = 8; tw <= max_tw; tw <<= 1) {
    const int th = kBlockM / tw;
    const long cover = (long)ceil_div(Ho, th) * th * ceil_div(Wo, tw) * tw;
    // prefer least padded work, then the squarer patch
    const long score = cover * 1024 + (th > tw ? th / tw : tw / th);
    if (best < 0 || score < best) { best = score; bh = th; bw = tw; }
  }
  *TH = bh;
  *TW = bw;
}

uint32_t div_mul(int d) { return d <= 1 ? 0u : (uint32_t)(((1ull << 32) + (uint64_t)d - 1) / (uint64_t)d); }

template <int BLOCK_N, bool kStaged>
int launch_conv(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const CUtensorMap& mr, const CUtensorMap& ma2,
                const ConvArgs& a_in, cudaStream_t st, const CUtensorMap* mb2 = nullptr, const CUtensorMap* mo2 = nullptr) {
  using Cfg = TileCfg<BLOCK_N, kStaged>;
  ConvArgs a = a_in;
  a.mul_tiles_n = div_mul(a.tiles_n);
  a.mul_tiles_w = div_mul(a.tiles_w);
  a.mul_tiles_h = div_mul(a.tiles_h);
  if ((long long)a.N * a.tiles_h * a.tiles_w * a.tiles_n * 2048 >= (1ll << 32)) return PE_ERR_UNSUPPORTED;  // fast_div range
  static DeviceOnce hint_once;
  if (hint_once.needed()) {
    const char* e = getenv("PE_CONV_WAIT_HINT");
    const uint32_t hint = e ? (uint32_t)atoi(e) : 0u;
    PE_CUDA_CHECK(cudaMemcpyToSymbol(g_wait_hint_ns, &hint, sizeof(hint)));
    hint_once.mark();
  }
  static DeviceOnce attr_once;  // one per template instantiation
  if (attr_once.needed()) {
    PE_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N, kStaged>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));
    attr_once.mark();
  }
  // chained layers keep the n tiles of an m tile on one CTA: the grid is sized in m tiles
  const int tiles = a.N * a.tiles_h * a.tiles_w * (a.chain_n ? 1 : a.tiles_n);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  // PE_CONV_PDL=0 launches the layers fully serialised (A/B switch)
  static const int pdl = [] { const char* e = getenv("PE_CONV_PDL"); return e ? atoi(e) : 1; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  PE_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BLOCK_N, kStaged>, ma, mb, mo, mr, ma2, mb2 ? *mb2 : ma, mo2 ? *mo2 : ma, a));
  return PE_OK;
}

}  // namespace

int conv2d_launch(const pe_conv_desc& d, const void* x, const void* w, const float* bias, const void* residual, void* y,
                  cudaStream_t st, const ConvSecondInput* x2, int reverse, const ConvChain* ch) {
  if (ch && ch->w) {  // chained 1x1 on this layer's output tile (see ConvArgs::chain_n)
    if (ch->out_fp32) {  // narrow fp32 chain: one 256-wide n tile, no residual, 16 chained outputs
      if (d.out_fp32 || d.Cout != 256 || d.residual_mode || ch->N != 16 || !ch->y || d.stride != 1) return PE_ERR_UNSUPPORTED;
    } else if (d.out_fp32 || d.Cout % 256 || d.residual_mode == 2 || d.KH != 1 || (ch->N != 64 && ch->N != 128 && ch->N != 256) || !ch->y)
      return PE_ERR_UNSUPPORTED;
    if (ch->bn != 0 && ch->bn != 128 && ch->bn != 256) return PE_ERR_INVALID_ARGUMENT;
  }
  if (x2 && x2->x) {  // dual-input 1x1: y = act([x | x2(strided)] . w + bias), w = [Cout][Cin + Cin2]
    if (d.KH != 1 || d.KW != 1 || d.stride != 1 || d.Cin % 64 || x2->Cin % 8 || (x2->stride != 1 && x2->stride != 2)) return PE_ERR_UNSUPPORTED;
    if ((x2->H - 1) / x2->stride + 1 != d.H || (x2->W - 1) / x2->stride + 1 != d.W) return PE_ERR_INVALID_ARGUMENT;
  }
  if (d.N < 1 || d.H < 1 || d.W < 1 || d.Cin < 8 || d.Cin % 8 || d.Cout < 8 || d.Cout % 8) return PE_ERR_INVALID_ARGUMENT;
  if (!((d.KH == 1 && d.KW == 1) || (d.KH == 3 && d.KW == 3))) return PE_ERR_UNSUPPORTED;
  if (d.stride != 1 && !(d.stride == 2 && d.KH == 1)) return PE_ERR_UNSUPPORTED;
  if (d.residual_mode < 0 || d.residual_mode > 2 || (d.residual_mode && !residual)) return PE_ERR_INVALID_ARGUMENT;
  if (!x || !w || (!y && !(ch && ch->w && ch->out_fp32))) return PE_ERR_INVALID_ARGUMENT;  // the narrow fp32 chain stores no main output
  ConvArgs a = {};
  a.N = d.N;
  a.Ho = d.stride == 2 ? (d.H - 1) / 2 + 1 : d.H;
  a.Wo = d.stride == 2 ? (d.W - 1) / 2 + 1 : d.W;
  a.Cin = d.Cin;
  a.Cout = d.Cout;
  a.KH = d.KH;
  a.KW = d.KW;
  a.pad = d.KH / 2;
  pick_patch(a.Ho, a.Wo, d.residual_mode == 2 ? 64 : 128, &a.TH, &a.TW);  // mode 2 needs even TH, TW
  // 3x3 convs on <= 256 channels are bound by the L2->SM fabric when every tap re-reads its A tile (9 x 16 KB per
  // K chunk): halo mode loads one 18x10-pixel tile per chunk and feeds the 9 taps from it (PE_CONV_HALO=0 disables).
  static const int halo_env = [] { const char* e = getenv("PE_CONV_HALO"); return e ? atoi(e) : 1; }();
  // Measured (profiles/README.md): a win only where the whole 3x3 filter bank also stays resident in shared memory
  // (res2: 64 -> 64 channels, 72 KB of weights; 80 -> 67 us per 8 frames); with streamed weights the fabric traffic is
  // dominated by the weight tiles and the shorter operand queue costs more than the saved A re-reads.
  const bool halo = halo_env && d.KH == 3 && d.stride == 1 && !d.out_fp32 && !d.residual_mode && d.Cin % 64 == 0 && d.Cout % 64 == 0 &&
                    (halo_env > 1 || (d.Cin <= 64 && d.Cout <= 64)) && !(ch && ch->w);  // chained layers use the generic operand pipeline
  if (halo) { a.TH = kHaloTH; a.TW = kHaloTW; }
  a.tiles_h = ceil_div(a.Ho, a.TH);
  a.tiles_w = ceil_div(a.Wo, a.TW);
  // Widest N tile that fits: measured on B200, narrowing N to fill more SMs on small maps (res5, p5/p6) LOSES - those
  // layers are bound by L2->SM operand traffic and every extra n-tile re-reads the A patch (profiles/README.md).
  int bn = d.Cout >= 256 ? 256 : (d.Cout >= 128 ? 128 : (d.Cout >= 64 ? 64 : (d.Cout > 16 ? 32 : 16)));
  // chained layers: 128-wide main tiles keep two accumulator stages (deferred order, see the kernel); PE_CONV_CHAIN_BN overrides
  static const int chain_bn_env = [] { const char* e = getenv("PE_CONV_CHAIN_BN"); return e ? atoi(e) : 256; }();
  if (ch && ch->w && !ch->out_fp32 && (ch->bn ? ch->bn : chain_bn_env) == 128) bn = 128;
  a.tiles_n = ceil_div(d.Cout, bn);
  a.k_chunks = ceil_div(d.Cin, kBlockK);  // a ragged last chunk is zero-filled by TMA (A and W alike)
  a.k_chunks1 = a.k_chunks;
  a.reverse = reverse ? 1 : 0;
  const bool chained = ch && ch->w;
  a.chain_n = chained ? ch->N : 0;
  a.chain_relu = chained ? ch->relu : 0;
  a.chain_bias = chained ? ch->bias : nullptr;
  a.chain_fp32 = chained && ch->out_fp32 ? 1 : 0;
  a.chain_out = a.chain_fp32 ? reinterpret_cast<float*>(ch->y) : nullptr;
  a.chain_out2 = a.chain_fp32 ? reinterpret_cast<float*>(ch->y2) : nullptr;
  a.b2_stages = 0;
  const bool dual = x2 && x2->x;
  if (dual) a.k_chunks += ceil_div(x2->Cin, kBlockK);
  a.relu = d.relu;
  a.residual_mode = d.residual_mode;
  a.out_fp32 = d.out_fp32;
  a.in_fp16 = d.in_fp16;
  a.stem_mode = 0;
  static const int halo_bo = [] { const char* e = getenv("PE_HALO_BASEOFF"); return e ? atoi(e) : 0; }();
  a.halo = halo ? (halo_bo ? 3 : 1) : 0;
  a.a_stages = a.b_stages = a.b_resident = 0;
  a.res_H = (a.Ho + 1) / 2;
  a.res_W = (a.Wo + 1) / 2;
  a.bias = bias;
  a.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  a.out = y;

  CUtensorMap ma, mb, ma2;
  if (dual) {
    const cuuint64_t s2 = (cuuint64_t)x2->stride;
    cuuint64_t dims[4] = {(cuuint64_t)x2->Cin, (cuuint64_t)a.Wo, (cuuint64_t)a.Ho, (cuuint64_t)d.N};
    cuuint64_t strides[3] = {s2 * x2->Cin * 2, s2 * (cuuint64_t)x2->W * x2->Cin * 2, (cuuint64_t)x2->H * x2->W * x2->Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)a.TW, (cuuint32_t)a.TH, 1};
    if (!make_map(&ma2, x2->x, 4, dims, strides, box)) return PE_ERR_CUDA;
  }
  {
    const cuuint64_t s = (cuuint64_t)d.stride;
    cuuint64_t dims[4] = {(cuuint64_t)d.Cin, (cuuint64_t)a.Wo, (cuuint64_t)a.Ho, (cuuint64_t)d.N};
    cuuint64_t strides[3] = {s * d.Cin * 2, s * (cuuint64_t)d.W * d.Cin * 2, (cuuint64_t)d.H * d.W * d.Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)(halo ? a.TW + 2 : a.TW), (cuuint32_t)(halo ? a.TH + 2 : a.TH), 1};
    if (!make_map(&ma, x, 4, dims, strides, box)) return PE_ERR_CUDA;
  }
  {
    const cuuint64_t ktot = (cuuint64_t)d.KH * d.KW * d.Cin + (dual ? (cuuint64_t)x2->Cin : 0);
    cuuint64_t dims[2] = {ktot, (cuuint64_t)d.Cout};
    cuuint64_t strides[1] = {ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)bn};
    if (!make_map(&mb, w, 2, dims, strides, box)) return PE_ERR_CUDA;
  }
  // bf16 outputs with whole 64-channel groups take the smem-staged TMA-store epilogue
  const bool staged = !d.out_fp32 && bn >= 64 && d.Cout % 64 == 0;
  CUtensorMap mo = ma, mr = ma;
  if (!dual) ma2 = ma;
  if (staged && !a.chain_fp32) {
    cuuint64_t dims[4] = {(cuuint64_t)d.Cout, (cuuint64_t)a.Wo, (cuuint64_t)a.Ho, (cuuint64_t)d.N};
    cuuint64_t strides[3] = {(cuuint64_t)d.Cout * 2, (cuuint64_t)a.Wo * d.Cout * 2, (cuuint64_t)a.Ho * a.Wo * d.Cout * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)a.TW, (cuuint32_t)a.TH, 1};
    if (!make_map(&mo, y, 4, dims, strides, box)) return PE_ERR_CUDA;
    if (d.residual_mode == 1 && !make_map(&mr, residual, 4, dims, strides, box)) return PE_ERR_CUDA;
    if (d.residual_mode == 2) {  // coarser FPN level: one (TH/2 x TW/2) box supplies the whole patch
      cuuint64_t cdims[4] = {(cuuint64_t)d.Cout, (cuuint64_t)a.res_W, (cuuint64_t)a.res_H, (cuuint64_t)d.N};
      cuuint64_t cstr[3] = {(cuuint64_t)d.Cout * 2, (cuuint64_t)a.res_W * d.Cout * 2, (cuuint64_t)a.res_H * a.res_W * d.Cout * 2};
      cuuint32_t cbox[4] = {64, (cuuint32_t)(a.TW / 2), (cuuint32_t)(a.TH / 2), 1};
      if (!make_map(&mr, residual, 4, cdims, cstr, cbox)) return PE_ERR_CUDA;
    }
  }
  CUtensorMap mb2, mo2;
  if (chained) {
    if (!staged || (bn != 256 && bn != 128)) return PE_ERR_UNSUPPORTED;
    cuuint64_t wd[2] = {(cuuint64_t)d.Cout, (cuuint64_t)ch->N};           // W2 [N2][Cout], K-major
    cuuint64_t ws[1] = {(cuuint64_t)d.Cout * 2};
    cuuint32_t wb[2] = {(cuuint32_t)kBlockK, (cuuint32_t)ch->N};
    if (!make_map(&mb2, ch->w, 2, wd, ws, wb)) return PE_ERR_CUDA;
    cuuint64_t od[4] = {(cuuint64_t)ch->N, (cuuint64_t)a.Wo, (cuuint64_t)a.Ho, (cuuint64_t)d.N};
    cuuint64_t os[3] = {(cuuint64_t)ch->N * 2, (cuuint64_t)a.Wo * ch->N * 2, (cuuint64_t)a.Ho * a.Wo * ch->N * 2};
    cuuint32_t ob[4] = {64, (cuuint32_t)a.TW, (cuuint32_t)a.TH, 1};
    if (ch->out_fp32) mo2 = mb2;  // written with plain stores
    else if (!make_map(&mo2, ch->y, 4, od, os, ob)) return PE_ERR_CUDA;
  }
  {  // split the smem budget: short K loops need few operand stages and profit from a deep residual/store queue
    const int stage_bytes = kBlockM * kBlockK * 2 + bn * kBlockK * 2;
    const int max_stages = bn >= 256 ? 4 : (bn >= 128 ? 6 : 8);
    const int k_iters = d.KH * d.KW * a.k_chunks;
    a.stages = max_stages;
    a.io_bufs = 2;
    if (staged) {
      const int budget = max_stages * stage_bytes + 2 * kIoBytes;
      int st_want;
      if (d.residual_mode) {  // the residual queue needs the smem more than a deep operand pipeline does
        st_want = k_iters / 2 + 1;
        if (st_want < 2) st_want = 2;
        if (st_want > max_stages) st_want = max_stages;
        if (st_want == max_stages && max_stages > 3) st_want = max_stages - 1;
      } else {
        st_want = k_iters + 1 < max_stages ? k_iters + 1 : max_stages;
        if (st_want < 2) st_want = 2;
        if (st_want < max_stages) st_want = (st_want + max_stages + 1) / 2;
      }
      const int io_bytes = kIoBytes + (d.residual_mode == 2 ? kCoarseBytes : 0);
      int b2_total = 0;
      if (chained && ch->out_fp32) {  // resident chained weights (4 x 2 KB); the long-K mainloop keeps its operand stages and the
        // staging buffers take what is left (1 with 4 stages: a sub-tile is consumed by the chained MMA before the next is staged)
        static const int rpn_stages = [] { const char* e = getenv("PE_RPN_CHAIN_STAGES"); return e ? atoi(e) : 4; }();
        a.b2_stages = 4;
        b2_total = a.b2_stages * ch->N * 128;
        st_want = rpn_stages < 2 ? 2 : (rpn_stages > max_stages ? max_stages : rpn_stages);
        if (k_iters + 1 < st_want) st_want = k_iters + 1 < 2 ? 2 : k_iters + 1;
      } else if (chained) {  // chained weight ring: [N2 x 64] bf16 chunks; the operand pipeline gives way (these layers are HBM-bound)
        a.b2_stages = ch->N <= 64 ? 4 : (ch->N <= 128 ? (bn == 128 ? 4 : 3) : 2);
        b2_total = a.b2_stages * ch->N * 128;
        if (st_want > 2) st_want = 2;
      }
      int io = (budget - b2_total - st_want * stage_bytes) / io_bytes;
      if (io > kMaxIoBufs) io = kMaxIoBufs;
      if (io < 2 && !(chained && ch->out_fp32)) { io = 2; st_want = (budget - b2_total - 2 * io_bytes) / stage_bytes; }
      if (chained && !ch->out_fp32 && (io < 3 || st_want < 2)) return PE_ERR_UNSUPPORTED;
      if (chained && ch->out_fp32 && (io < 1 || st_want < 2)) return PE_ERR_UNSUPPORTED;
      a.stages = st_want;
      a.io_bufs = io;
      if (halo) {
        const int b_bytes = bn * kBlockK * 2;
        const int all_b = 9 * a.k_chunks;
        a.a_stages = 3;
        a.io_bufs = 2;
        int room = budget - a.a_stages * kHaloBytes - a.io_bufs * kIoBytes;
        if (all_b <= kMaxBStages && all_b * b_bytes <= room && a.tiles_n == 1) {  // whole filter bank fits: load it once per CTA
          a.b_resident = 1;
          a.b_stages = all_b;
          room -= all_b * b_bytes;
          int more = room / kIoBytes;
          if (more > 2) more = 2;
          a.io_bufs += more;
        } else {
          a.b_stages = room / b_bytes;
          if (a.b_stages > kMaxBStages) a.b_stages = kMaxBStages;
          if (a.b_stages < 3) { a.a_stages = 2; a.b_stages = (budget - 2 * kHaloBytes - 2 * kIoBytes) / b_bytes; }
        }
      }
    }
  }
  if (staged) {
    switch (bn) {
      case 256: return launch_conv<256, true>(ma, mb, mo, mr, ma2, a, st, chained ? &mb2 : nullptr, chained ? &mo2 : nullptr);
      case 128: return launch_conv<128, true>(ma, mb, mo, mr, ma2, a, st, chained ? &mb2 : nullptr, chained ? &mo2 : nullptr);
      default: return launch_conv<64, true>(ma, mb, mo, mr, ma2, a, st);
    }
  }
  switch (bn) {
    case 256: return launch_conv<256, false>(ma, mb, mo, mr, ma2, a, st);
    case 128: return launch_conv<128, false>(ma, mb, mo, mr, ma2, a, st);
    case 64: return launch_conv<64, false>(ma, mb, mo, mr, ma2, a, st);
    case 32: return launch_conv<32, false>(ma, mb, mo, mr, ma2, a, st);
    default: return launch_conv<16, false>(ma, mb, mo, mr, ma2, a, st);
  }
}

// 7x7 stride-2 stem (resnet.py:369-384) straight from the zero-bordered fp16 HWC4 canvas [B, Hc+6, Wc+8, 4]
// (detector_kernels.cu: stem_canvas*): out[b][ho][wo][co] = sum_kh sum_j canvas[b][2ho+kh][8wo + j] * w[co][kh][j],
// j = kw*4 + c over an 8-pixel window (kw = 7 and c >= C carry zero weights).  The A operand is read through a
// 5-D tensor map whose wo dimension has a 16-byte stride (overlapping 64-byte windows) and whose row dimension
// is split into (parity, row/2), so no im2col matrix is ever written.
int conv_stem_launch(const void* canvas, const void* w, const float* bias, void* y, int B, int Hc, int Wc, cudaStream_t st, int reverse) {
  if (!canvas || !w || !y || B < 1 || Hc % 32 || Wc % 32) return PE_ERR_INVALID_ARGUMENT;
  const int Hp = Hc + 6, Wp = Wc + 8;
  ConvArgs a = {};
  a.N = B; a.Ho = Hc / 2; a.Wo = Wc / 2; a.Cin = 224; a.Cout = 64;
  a.KH = 7; a.KW = 1; a.pad = 0;
  static const int stem_mode_env = [] { const char* e = getenv("PE_STEM_MODE"); return e ? atoi(e) : 2; }();
  const int mode = stem_mode_env == 1 ? 1 : 2;
  if (mode == 2) { a.TH = 1; a.TW = kBlockM; }  // a run of 128 outputs on one row: rows of the A operand are 16 B apart
  else pick_patch(a.Ho, a.Wo, 128, &a.TH, &a.TW);
  a.tiles_h = ceil_div(a.Ho, a.TH);
  a.tiles_w = ceil_div(a.Wo, a.TW);
  a.tiles_n = 1;
  a.k_chunks = 1;
  a.k_chunks1 = 1;
  a.reverse = reverse ? 1 : 0;
  a.relu = 1; a.residual_mode = 0; a.out_fp32 = 0; a.in_fp16 = 1; a.stem_mode = mode;
  a.canvas = reinterpret_cast<const unsigned char*>(canvas);
  a.canvas_hp = Hp;
  a.canvas_wp = Wp;
  a.res_H = a.res_W = 0;
  a.bias = bias; a.residual = nullptr; a.out = y;
  a.stages = mode == 2 ? kStemRowStages : 8;
  a.io_bufs = 2;
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return PE_ERR_CUDA;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMap ma, mb, mo;
  {
    const cuuint64_t pitch = (cuuint64_t)Wp * 8;
    cuuint64_t dims[5] = {32, (cuuint64_t)a.Wo, 2, (cuuint64_t)(Hp / 2), (cuuint64_t)B};
    cuuint64_t strides[4] = {16, pitch, 2 * pitch, (cuuint64_t)Hp * pitch};
    cuuint32_t box[5] = {32, (cuuint32_t)a.TW, 1, (cuuint32_t)a.TH, 1};
    if (fn(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(canvas), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return PE_ERR_CUDA;
  }
  {
    cuuint64_t dims[2] = {224, 64};
    cuuint64_t strides[1] = {448};
    cuuint32_t box[2] = {32, 64};
    if (fn(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return PE_ERR_CUDA;
  }
  {
    cuuint64_t dims[4] = {64, (cuuint64_t)a.Wo, (cuuint64_t)a.Ho, (cuuint64_t)B};
    cuuint64_t strides[3] = {128, (cuuint64_t)a.Wo * 128, (cuuint64_t)a.Ho * a.Wo * 128};
    cuuint32_t box[4] = {64, (cuuint32_t)a.TW, (cuuint32_t)a.TH, 1};
    if (!make_map(&mo, y, 4, dims, strides, box)) return PE_ERR_CUDA;
  }
  return launch_conv<64, true>(ma, mb, mo, mo, mo, a, st);
}

}  // namespace pe

extern "C" PE_API int pe_conv2d_fwd(const pe_conv_desc* desc, const void* x, const void* w, const float* bias,
                                    const void* residual, void* y, void* stream) {
  if (!desc) return PE_ERR_INVALID_ARGUMENT;
  return pe::conv2d_launch(*desc, x, w, bias, residual, y, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" PE_API int pe_conv1x1_chain_fwd(const pe_conv_desc* desc, const void* x, const void* x2, int cin2, int h2, int w2s, int stride2,
                                           const void* w, const float* bias, const void* residual, void* y, const void* wc,
                                           const float* bias_c, int n_c, int relu_c, void* y_c, void* stream) {
  if (!desc || !wc || !y_c) return PE_ERR_INVALID_ARGUMENT;
  pe::ConvSecondInput s2 = {x2, cin2, h2, w2s, stride2};
  pe::ConvChain ch = {wc, bias_c, y_c, n_c, relu_c, 0};
  return pe::conv2d_launch(*desc, x, w, bias, residual, y, reinterpret_cast<cudaStream_t>(stream), x2 ? &s2 : nullptr, 0, &ch);
}

extern "C" PE_API int pe_conv_rpn_head_fwd(const pe_conv_desc* desc, const void* x, const void* w, const float* bias, const void* wc,
                                           const float* bias_c, float* y_c, void* stream) {
  if (!desc || !wc || !y_c) return PE_ERR_INVALID_ARGUMENT;
  pe::ConvChain ch = {wc, bias_c, y_c, 16, 0, 0, 1, nullptr};
  return pe::conv2d_launch(*desc, x, w, bias, nullptr, nullptr, reinterpret_cast<cudaStream_t>(stream), nullptr, 0, &ch);
}

extern "C" PE_API int pe_conv1x1_dual_fwd(const pe_conv_desc* desc, const void* x, const void* x2, int cin2, int h2, int w2, int stride2,
                                          const void* w, const float* bias, void* y, void* stream) {
  if (!desc || !x2) return PE_ERR_INVALID_ARGUMENT;
  pe::ConvSecondInput s2 = {x2, cin2, h2, w2, stride2};
  return pe::conv2d_launch(*desc, x, w, bias, nullptr, y, reinterpret_cast<cudaStream_t>(stream), &s2);
}
