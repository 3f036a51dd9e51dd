// Native detector engine: builds the layer plan of the reference's GeneralizedRCNN inference path
// (detectron2/modeling/meta_arch/rcnn.py:219-267: ResNet-50/101-FPN backbone, RPN, StandardROIHeads with the
// fork's variance head) and runs it as a fixed sequence of sm_100a kernels on one stream - no host
// synchronisation, no Python in the loop.  Weights live in a caller-owned blob whose layout the engine
// publishes as a manifest (pe_detector_param_*); activations live in a caller-owned workspace.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "common.cuh"
#include "detector_kernels.cuh"

namespace pe {
namespace {

constexpr int kMaxProps = 1000;  // RPN POST_NMS_TOPK_TEST (configs/Base-RCNN-FPN.yaml:19)

// kConv3Shortcut: [conv3 | shortcut] of a stage's first bottleneck block concatenated along K (both FrozenBN-folded, biases summed):
// out = relu(conv3(t) + shortcut(x)) runs as one dual-input GEMM (pe_conv1x1_dual_fwd); the manifest entry carries conv3's name, the
// host derives the shortcut's (".conv3" -> ".shortcut")
enum ParamKind { kConvBN = 0, kConvBias = 1, kStem = 2, kRpnPred = 3, kFc1 = 4, kLinear = 5, kPredictor = 6, kConv3Shortcut = 7 };

struct Param {
  std::string name;
  int kind, Cout, KH, KW, Cin;
  size_t w_off, b_off;
};

struct Buf {
  std::string name;
  size_t off, bytes;
  int d[4];  // N, H, W, C
  int elem;  // bytes per element
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace
}  // namespace pe

struct pe_detector {
  pe_detector_config cfg;
  // optional per-forward instrumentation (pe_detector_set_profiling): CUDA events around every GEMM launch
  bool profiling = false;
  // profiling mode 2: ONE event pair per run of back-to-back GEMM launches (no event between them, so programmatic dependent launch
  // and the launch gaps count exactly as they do in a real forward); mode 1: one pair per launch (per-layer table)
  bool profile_runs = false;
  bool run_open = false;
  std::vector<cudaEvent_t> ev;
  int ev_used = 0;
  std::vector<double> prof_flops, prof_bytes;  // algorithmic work of the GEMM launch behind each event pair
  // mode 1 also brackets every NON-GEMM launch group (staging, pooling, RPN selection / NMS, ROIAlign, head post-processing)
  std::vector<cudaEvent_t> aux_ev;
  int aux_used = 0;
  std::vector<std::string> aux_names;
  int last_launches = 0, last_gemm_launches = 0;
  // optional cross-detector stagger (pe_detector_set_stagger_event): recorded on the forward's stream once `stagger_after` kernels
  // have been launched, so that a SECOND detector's stream can start that much later and the two never sit in their latency-bound
  // stages (RPN selection / NMS: a few CTAs) at the same time
  cudaEvent_t stagger_ev = nullptr;
  int stagger_after = 0;
  bool stagger_sent = false;
  std::vector<pe::Param> params;
  std::vector<pe::Buf> bufs;
  size_t weight_bytes = 0, ws_bytes = 0;
  int fc;         // channels entering RPN / ROI heads (256, or 512 for middle fusion)
  int stem_c;     // channels per backbone pass
  int stem_kp;    // padded K of the stem GEMM
  int npad;       // padded predictor width
  int H[6], W[6]; // canvas sizes at strides 2,4,8,16,32,64

  int add_param(const std::string& name, int kind, int Cout, int KH, int KW, int Cin, int welem) {
    pe::Param p;
    p.name = name; p.kind = kind; p.Cout = Cout; p.KH = KH; p.KW = KW; p.Cin = Cin;
    p.w_off = weight_bytes;
    weight_bytes = pe::align_up(weight_bytes + (size_t)Cout * KH * KW * Cin * welem, 256);
    p.b_off = weight_bytes;
    weight_bytes = pe::align_up(weight_bytes + (size_t)Cout * 4, 256);
    params.push_back(p);
    return (int)params.size() - 1;
  }
  int find_param(const std::string& name) const {
    for (size_t i = 0; i < params.size(); ++i)
      if (params[i].name == name) return (int)i;
    return -1;
  }
  size_t add_buf(const std::string& name, int n, int h, int w, int c, int elem) {
    pe::Buf b;
    b.name = name; b.off = ws_bytes; b.bytes = (size_t)n * h * w * c * elem; b.d[0] = n; b.d[1] = h; b.d[2] = w; b.d[3] = c; b.elem = elem;
    ws_bytes = pe::align_up(ws_bytes + b.bytes, 1024);
    bufs.push_back(b);
    return b.off;
  }
  const pe::Buf* find_buf(const std::string& name) const {
    for (const auto& b : bufs)
      if (b.name == name) return &b;
    return nullptr;
  }
};

namespace pe {
namespace {

const int kStageBlocks50[4] = {3, 4, 6, 3};
const int kStageBlocks101[4] = {3, 4, 23, 3};

void build_plan(pe_detector* d) {
  const pe_detector_config& c = d->cfg;
  const int* blocks = c.depth == 101 ? kStageBlocks101 : kStageBlocks50;
  d->stem_c = c.middle_fusion ? 3 : c.in_channels;
  d->stem_kp = kStemK;
  d->fc = c.middle_fusion ? 512 : 256;
  d->npad = (int)align_up((size_t)(c.num_classes + 1 + 4 * c.num_classes + 1), 16);
  for (int i = 0; i < 6; ++i) {
    d->H[i] = c.canvas_h >> (i + 1);
    d->W[i] = c.canvas_w >> (i + 1);
  }
  d->H[5] = (d->H[4] - 1) / 2 + 1;
  d->W[5] = (d->W[4] - 1) / 2 + 1;
  // ---- parameters (manifest order == blob order)
  const std::string bu = "backbone.bottom_up";
  d->add_param(bu + ".stem.conv1", kStem, 64, 1, 1, d->stem_kp, 2);
  int cin = 64;
  for (int s = 0; s < 4; ++s) {
    const int mid = 64 << s, cout = 256 << s;
    for (int b = 0; b < blocks[s]; ++b) {
      char q[96];
      snprintf(q, sizeof(q), "%s.res%d.%d", bu.c_str(), s + 2, b);
      d->add_param(std::string(q) + ".conv1", kConvBN, mid, 1, 1, cin, 2);
      d->add_param(std::string(q) + ".conv2", kConvBN, mid, 3, 3, mid, 2);
      if (b == 0) d->add_param(std::string(q) + ".conv3", kConv3Shortcut, cout, 1, 1, mid + cin, 2);
      else d->add_param(std::string(q) + ".conv3", kConvBN, cout, 1, 1, mid, 2);
      cin = cout;
    }
  }
  for (int l = 5; l >= 2; --l) {
    char q[64];
    snprintf(q, sizeof(q), "backbone.fpn_lateral%d", l);
    d->add_param(q, kConvBias, 256, 1, 1, 256 << (l - 2), 2);
    snprintf(q, sizeof(q), "backbone.fpn_output%d", l);
    d->add_param(q, kConvBias, 256, 3, 3, 256, 2);
  }
  d->add_param("proposal_generator.rpn_head.conv", kConvBias, d->fc, 3, 3, d->fc, 2);
  d->add_param("proposal_generator.rpn_head", kRpnPred, kRpnOutC, 1, 1, d->fc, 2);
  d->add_param("roi_heads.box_head.fc1", kFc1, 1024, 1, 1, 49 * d->fc, 2);
  d->add_param("roi_heads.box_head.fc2", kLinear, 1024, 1, 1, 1024, 2);
  d->add_param("roi_heads.box_predictor", kPredictor, d->npad, 1, 1, 1024, 2);

  // ---- workspace
  const int B = c.max_batch;
  const int passes = c.middle_fusion ? 2 : 1;
  d->add_buf("stem_canvas", B, c.canvas_h + 6, c.canvas_w + 8, 4, 2);
  d->add_buf("pil_taps", 1, 1, c.canvas_h + c.canvas_w + 128, 4, 4);  // Pillow resize tap tables (rows, then columns) + 2 KB normalisation table
  d->add_buf("stem_out", B, d->H[0], d->W[0], 64, 2);
  d->add_buf("pool_out", B, d->H[1], d->W[1], 64, 2);
  d->add_buf("x0", B, d->H[1], d->W[1], 256, 2);
  d->add_buf("x1", B, d->H[1], d->W[1], 256, 2);
  d->add_buf("t1", B, d->H[1], d->W[1], 64, 2);
  d->add_buf("t2", B, d->H[1], d->W[1], 64, 2);
  for (int s = 0; s < 4; ++s) {
    char q[16];
    snprintf(q, sizeof(q), "res%d", s + 2);
    d->add_buf(q, B, d->H[s + 1], d->W[s + 1], 256 << s, 2);
  }
  for (int p = 0; p < passes; ++p)
    for (int l = 5; l >= 2; --l) {
      char q[24];
      snprintf(q, sizeof(q), "inner%d_%d", l, p);
      d->add_buf(q, B, d->H[l - 1], d->W[l - 1], 256, 2);
      snprintf(q, sizeof(q), "pout%d_%d", l, p);
      d->add_buf(q, B, d->H[l - 1], d->W[l - 1], 256, 2);
    }
  for (int l = 2; l <= 6; ++l) {
    char q[16];
    snprintf(q, sizeof(q), "p%d", l);
    if (c.middle_fusion || l == 6) d->add_buf(q, B, d->H[l - 1], d->W[l - 1], d->fc, 2);
    snprintf(q, sizeof(q), "rpn_out%d", l);
    d->add_buf(q, B, d->H[l - 1], d->W[l - 1], kRpnOutC, 4);
    snprintf(q, sizeof(q), "rpn_logit%d", l);  // dense copy of the objectness logits (3 | pad) for the top-k passes
    d->add_buf(q, B, d->H[l - 1], d->W[l - 1], 4, 4);
  }
  d->add_buf("rpn_t", B, d->H[1], d->W[1], d->fc, 2);
  d->add_buf("cand_box", B, kRpnLevels, kTopkSlots, 4, 4);
  d->add_buf("cand_score", B, kRpnLevels, kTopkSlots, 1, 4);
  d->add_buf("cand_valid", B, kRpnLevels, kTopkSlots, 1, 1);
  d->add_buf("cand_count", B, kRpnLevels, 1, 1, 4);
  d->add_buf("keep_idx", B, kRpnLevels, kTopkSlots, 1, 4);
  d->add_buf("keep_count", B, kRpnLevels, 1, 1, 4);
  d->add_buf("nms_mask", B, kRpnLevels, 1024, 32, 4);
  d->add_buf("proposals", B, kMaxProps, 1, 4, 4);
  d->add_buf("prop_count", B, 1, 1, 1, 4);
  d->add_buf("roi_feats", B * kMaxProps, 1, 1, 49 * d->fc, 2);
  d->add_buf("fc1_out", B * kMaxProps, 1, 1, 1024, 2);
  d->add_buf("fc2_out", B * kMaxProps, 1, 1, 1024, 2);
  d->add_buf("head_out", B * kMaxProps, 1, 1, d->npad, 4);
}

struct Runner {
  pe_detector* d;
  const unsigned char* wts;
  unsigned char* ws;
  cudaStream_t st;
  int B;
  int status = PE_OK;
  int rev = 0;  // tile direction of the last GEMM launch

  void* buf(const char* name) const { return ws + d->find_buf(name)->off; }

  int gemm(const pe_conv_desc& cd, const void* x, const void* w, const float* bias, const void* res, void* y,
           const ConvSecondInput* x2 = nullptr, const ConvChain* chain = nullptr) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (d->profiling) {
      while ((int)d->ev.size() < d->ev_used + 2) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return PE_ERR_CUDA;
        d->ev.push_back(e);
      }
      if (!d->profile_runs || !d->run_open) {
        e0 = d->ev[d->ev_used++];
        e1 = d->ev[d->ev_used++];  // recorded by end_run() in run mode
      }
      const int pad = (cd.KH - 1) / 2;
      const double Ho = (cd.H + 2 * pad - cd.KH) / cd.stride + 1, Wo = (cd.W + 2 * pad - cd.KW) / cd.stride + 1;
      const double taps = (double)cd.KH * cd.KW * cd.Cin + (x2 ? x2->Cin : 0), opix = (double)cd.N * Ho * Wo;
      double bytes = (double)cd.N * cd.H * cd.W * cd.Cin * 2 / ((cd.KH == 1 && cd.stride == 2) ? 4.0 : 1.0) +
                     taps * cd.Cout * 2 + cd.Cout * 4.0 + opix * cd.Cout * (cd.out_fp32 ? 4 : 2);
      if (x2) bytes += opix * x2->Cin * 2;  // the block input at the positions the (strided) shortcut reads
      if (cd.residual_mode == 1) bytes += opix * cd.Cout * 2;
      if (cd.residual_mode == 2) bytes += opix * cd.Cout * 2 / 4.0;
      double flops = 2.0 * opix * cd.Cout * taps;
      if (chain) {  // + the chained 1x1 (the next block's conv1): its weights and its output; its input never leaves the SM
        flops += 2.0 * opix * cd.Cout * chain->N;
        bytes += (double)chain->N * cd.Cout * 2 + chain->N * 4.0 + opix * chain->N * (chain->out_fp32 ? 4 : 2);
        if (chain->out_fp32) bytes -= opix * cd.Cout * 2;  // the RPN head's hidden tensor is not stored at all
        if (chain->out_fp32 && chain->y2) bytes += opix * 16;  // dense logit plane
      }
      if (e0) {
        d->prof_flops.push_back(flops);
        d->prof_bytes.push_back(bytes);
        cudaEventRecord(e0, st);
      } else {  // the run's totals accumulate in its slot
        d->prof_flops.back() += flops;
        d->prof_bytes.back() += bytes;
      }
      if (d->profile_runs) { d->run_open = true; e1 = nullptr; }
    }
    // consecutive layers walk their tiles in opposite directions (conv_gemm.cu ConvArgs::reverse); PE_CONV_REVERSE=0 disables
    static const int alternate = [] { const char* e = getenv("PE_CONV_REVERSE"); return e ? atoi(e) : 1; }();
    rev ^= alternate;
    const int s = conv2d_launch(cd, x, w, bias, res, y, st, x2, rev, chain);
    if (d->profiling && e1) cudaEventRecord(e1, st);
    d->last_launches++;
    d->last_gemm_launches++;
    signal_stagger();
    return s;
  }

  static int stem_reverse() {
    static const int alternate = [] { const char* e = getenv("PE_CONV_REVERSE"); return e ? atoi(e) : 1; }();
    return alternate;
  }

  // run mode: closes the open run of GEMM launches (called before every non-GEMM launch and at the end of the forward)
  void end_run() {
    if (d->profiling && d->profile_runs && d->run_open) {
      cudaEventRecord(d->ev[d->ev_used - 1], st);
      d->run_open = false;
    }
  }

  void conv(const std::string& pname, const void* x, int H, int W, int stride, bool relu, int rmode, const void* res, void* y,
            bool out_fp32 = false, bool in_fp16 = false) {
    if (status != PE_OK) return;
    const int pi = d->find_param(pname);
    if (pi < 0) { status = PE_ERR_INVALID_ARGUMENT; return; }
    const Param& p = d->params[pi];
    pe_conv_desc cd;
    cd.N = B; cd.H = H; cd.W = W; cd.Cin = p.Cin; cd.Cout = p.Cout; cd.KH = p.KH; cd.KW = p.KW; cd.stride = stride;
    cd.relu = relu; cd.residual_mode = rmode; cd.out_fp32 = out_fp32; cd.in_fp16 = in_fp16;
    status = gemm(cd, x, wts + p.w_off, reinterpret_cast<const float*>(wts + p.b_off), res, y);
  }
  void linear(const std::string& pname, const void* x, int rows, bool relu, void* y, bool out_fp32) {
    if (status != PE_OK) return;
    const Param& p = d->params[d->find_param(pname)];
    pe_conv_desc cd;
    cd.N = 1; cd.H = 1; cd.W = rows; cd.Cin = p.Cin; cd.Cout = p.Cout; cd.KH = 1; cd.KW = 1; cd.stride = 1;
    cd.relu = relu; cd.residual_mode = 0; cd.out_fp32 = out_fp32; cd.in_fp16 = 0;
    status = gemm(cd, x, wts + p.w_off, reinterpret_cast<const float*>(wts + p.b_off), nullptr, y);
  }
  void check(int s, int launches = 1) { if (status == PE_OK) status = s; d->last_launches += launches; signal_stagger(); }
  void signal_stagger(bool force = false) {
    if (!d->stagger_ev || d->stagger_sent || (!force && d->last_launches < d->stagger_after)) return;
    cudaEventRecord(d->stagger_ev, st);
    d->stagger_sent = true;
  }
  // mode 1: event pair around a non-GEMM launch group, labelled with the launcher's name (pe_detector_profile_kernels)
  void aux_begin(const char* call) {
    if (!d->profiling || d->profile_runs || status != PE_OK) return;
    while ((int)d->aux_ev.size() < d->aux_used + 2) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) { status = PE_ERR_CUDA; return; }
      d->aux_ev.push_back(e);
    }
    const char* par = strchr(call, '(');
    d->aux_names.emplace_back(call, par ? (size_t)(par - call) : strlen(call));
    cudaEventRecord(d->aux_ev[d->aux_used], st);
    aux_open = true;
  }
  void aux_end() {
    if (!aux_open) return;
    cudaEventRecord(d->aux_ev[d->aux_used + 1], st);
    d->aux_used += 2;
    aux_open = false;
  }
  bool aux_open = false;
#define PE_NONGEMM(call) do { end_run(); aux_begin(#call); check(call); aux_end(); } while (0)

  // raw uint8 frames (resized on the fly) instead of a float32 tensor when frames != nullptr
  const unsigned char* frames = nullptr;
  int src_h = 0, src_w = 0, round_u8 = 0;
  // module seams (pe_detector_forward_stages): which of backbone / proposal generator / ROI heads run, and whether the images
  // are already normalised (Backbone.forward receives ImageList.tensor, rcnn.py:269-286 has been applied by the caller)
  int stages = PE_STAGE_ALL;
  bool prenormalized = false;

  // ResNet bottom-up + FPN for one backbone pass (input channels [c0, c0 + stem_c) of the image tensor)
  void backbone(const float* images, int Ctot, int c0, int img_h, int img_w, int pass) {
    const pe_detector_config& c = d->cfg;
    StemNorm nrm;
    for (int i = 0; i < 8; ++i) { nrm.mean[i] = 0.f; nrm.std[i] = 1.f; }
    if (!prenormalized)
      for (int i = 0; i < d->stem_c; ++i) { nrm.mean[i] = c.pixel_mean[c0 + i]; nrm.std[i] = c.pixel_std[c0 + i]; }
    if (frames)
      PE_NONGEMM(launch_stem_im2col_u8(frames, buf("stem_canvas"), nullptr, B, Ctot, c0, d->stem_c, src_h, src_w, img_h, img_w, c.canvas_h,
                                  c.canvas_w, round_u8, nrm, st, buf("pil_taps")));
    else
      PE_NONGEMM(launch_stem_im2col(images, buf("stem_canvas"), nullptr, B, Ctot, c0, d->stem_c, img_h, img_w, c.canvas_h, c.canvas_w, nrm, st));
    if (status == PE_OK) {  // 7x7/2 conv: tcgen05 GEMM whose A operand is TMA-read straight from the canvas (fp16 operands)
      const Param& p = d->params[d->find_param("backbone.bottom_up.stem.conv1")];
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (d->profiling) {
        while ((int)d->ev.size() < d->ev_used + 2) {
          cudaEvent_t e;
          if (cudaEventCreate(&e) != cudaSuccess) { status = PE_ERR_CUDA; break; }
          d->ev.push_back(e);
        }
        if (status == PE_OK) {
          e0 = d->ev[d->ev_used++];
          e1 = d->ev[d->ev_used++];
          const double opix = (double)B * (c.canvas_h / 2) * (c.canvas_w / 2);
          d->prof_flops.push_back(2.0 * opix * 64 * 49 * d->stem_c);
          d->prof_bytes.push_back((double)B * c.canvas_h * c.canvas_w * 8 + opix * 64 * 2 + 64.0 * kStemK * 2);
          cudaEventRecord(e0, st);
        }
      }
      if (status == PE_OK)
        status = conv_stem_launch(buf("stem_canvas"), wts + p.w_off, reinterpret_cast<const float*>(wts + p.b_off), buf("stem_out"), B,
                                  c.canvas_h, c.canvas_w, st, rev = stem_reverse());
      if (e1) cudaEventRecord(e1, st);
      d->last_launches++;
      d->last_gemm_launches++;
      signal_stagger();
    }
    // direction chain (L2 reuse, see ConvArgs::reverse): canvas staging walks forward -> stem conv backwards -> max-pool forward ->
    // res2.0.conv1 backwards -> ...
    PE_NONGEMM(launch_maxpool(buf("stem_out"), buf("pool_out"), B, d->H[0], d->W[0], 64, st, 0));
    rev = 0;  // the next GEMM toggles to 1 = backwards
    const int* blocks = c.depth == 101 ? kStageBlocks101 : kStageBlocks50;
    const void* x = buf("pool_out");
    int H = d->H[1], W = d->W[1];
    bool chained_c1 = false;
    for (int s = 0; s < 4; ++s) {
      for (int b = 0; b < blocks[s]; ++b) {
        char q[96];
        snprintf(q, sizeof(q), "backbone.bottom_up.res%d.%d", s + 2, b);
        const std::string base(q);
        const int stride = (b == 0 && s > 0) ? 2 : 1;
        const int Hin = H, Win = W;
        // conv1 of every block but the first was already computed by the previous block's conv3 kernel (chained 1x1)
        if (!chained_c1) conv(base + ".conv1", x, H, W, stride, true, 0, nullptr, buf("t1"));
        if (stride == 2) { H = (H - 1) / 2 + 1; W = (W - 1) / 2 + 1; }
        conv(base + ".conv2", buf("t1"), H, W, 1, true, 0, nullptr, buf("t2"));
        char rq[16];
        snprintf(rq, sizeof(rq), "res%d", s + 2);
        void* y = (b == blocks[s] - 1) ? buf(rq) : (x == buf("x0") ? buf("x1") : buf("x0"));
        // chain the NEXT block's conv1 (cout -> mid, ReLU) onto this conv3: its A operand is this kernel's own output tile
        static const int chain_mask = [] { const char* e = getenv("PE_CONV_CHAIN"); return e ? atoi(e) : 3; }();  // bit s: stage res(2+s); measured:
        // res2 / res3 gain (conv1's 419 / 210 MB re-read disappears), res4 loses (its single main accumulator serialises the longer K loop)
        ConvChain ch = {nullptr, nullptr, nullptr, 0, 0, 0};
        chained_c1 = false;
        if (b + 1 < blocks[s] && ((chain_mask >> s) & 1) && (64 << s) <= 256 && status == PE_OK) {
          char nq[96];
          snprintf(nq, sizeof(nq), "backbone.bottom_up.res%d.%d.conv1", s + 2, b + 1);
          const Param& p1 = d->params[d->find_param(nq)];
          ch.w = wts + p1.w_off; ch.bias = reinterpret_cast<const float*>(wts + p1.b_off); ch.y = buf("t1"); ch.N = p1.Cout; ch.relu = 1;
          chained_c1 = true;
        }
        const ConvChain* chp = chained_c1 ? &ch : nullptr;
        if (b == 0) {  // conv3 + projection shortcut as one GEMM over K = [t2 | block input] (the shortcut tensor never exists)
          if (status == PE_OK) {
            const Param& p = d->params[d->find_param(base + ".conv3")];
            const int mid = 64 << s;
            pe_conv_desc cd;
            cd.N = B; cd.H = H; cd.W = W; cd.Cin = mid; cd.Cout = p.Cout; cd.KH = 1; cd.KW = 1; cd.stride = 1;
            cd.relu = 1; cd.residual_mode = 0; cd.out_fp32 = 0; cd.in_fp16 = 0;
            ConvSecondInput x2 = {x, p.Cin - mid, Hin, Win, stride};
            status = gemm(cd, buf("t2"), wts + p.w_off, reinterpret_cast<const float*>(wts + p.b_off), nullptr, y, &x2, chp);
          }
        } else if (status == PE_OK) {
          const Param& p = d->params[d->find_param(base + ".conv3")];
          pe_conv_desc cd;
          cd.N = B; cd.H = H; cd.W = W; cd.Cin = p.Cin; cd.Cout = p.Cout; cd.KH = 1; cd.KW = 1; cd.stride = 1;
          cd.relu = 1; cd.residual_mode = 1; cd.out_fp32 = 0; cd.in_fp16 = 0;
          status = gemm(cd, buf("t2"), wts + p.w_off, reinterpret_cast<const float*>(wts + p.b_off), x, y, nullptr, chp);
        }
        x = y;
      }
    }
    // FPN top-down (fpn.py:110-145)
    char qi[24], qo[24], qp[24], rq[16];
    for (int l = 5; l >= 2; --l) {
      snprintf(qi, sizeof(qi), "inner%d_%d", l, pass);
      snprintf(qo, sizeof(qo), "pout%d_%d", l, pass);
      snprintf(rq, sizeof(rq), "res%d", l);
      char ln[40], on[40];
      snprintf(ln, sizeof(ln), "backbone.fpn_lateral%d", l);
      snprintf(on, sizeof(on), "backbone.fpn_output%d", l);
      if (l == 5) conv(ln, buf(rq), d->H[l - 1], d->W[l - 1], 1, false, 0, nullptr, buf(qi));
      else {
        snprintf(qp, sizeof(qp), "inner%d_%d", l + 1, pass);
        conv(ln, buf(rq), d->H[l - 1], d->W[l - 1], 1, false, 2, buf(qp), buf(qi));
      }
      conv(on, buf(qi), d->H[l - 1], d->W[l - 1], 1, false, 0, nullptr, buf(qo));
    }
  }

  const void* level_feat(int l) const {  // p2..p6 as seen by RPN / ROI heads
    char q[24];
    if (d->cfg.middle_fusion || l == 6) { snprintf(q, sizeof(q), "p%d", l); return buf(q); }
    snprintf(q, sizeof(q), "pout%d_0", l);
    return buf(q);
  }

  int run(const float* images, int img_h, int img_w, float out_h, float out_w, const pe_detections& o) {
    const pe_detector_config& c = d->cfg;
    const int Ctot = c.in_channels;
    if (stages & PE_STAGE_BACKBONE) {
      if (c.middle_fusion) {  // shared backbone on both halves, channel concat (rcnn.py:240-248)
        backbone(images, Ctot, 0, img_h, img_w, 0);
        backbone(images, Ctot, 3, img_h, img_w, 1);
        for (int l = 2; l <= 5 && status == PE_OK; ++l) {
          char qa[24], qb[24], qp[16];
          snprintf(qa, sizeof(qa), "pout%d_0", l);
          snprintf(qb, sizeof(qb), "pout%d_1", l);
          snprintf(qp, sizeof(qp), "p%d", l);
          PE_NONGEMM(launch_concat_channels(buf(qa), buf(qb), buf(qp), (long long)B * d->H[l - 1] * d->W[l - 1], 256, st));
        }
      } else {
        backbone(images, Ctot, 0, img_h, img_w, 0);
      }
      PE_NONGEMM(launch_subsample2(level_feat(5), buf("p6"), B, d->H[4], d->W[4], d->fc, st));
    }
    float4* props = reinterpret_cast<float4*>(buf("proposals"));
    int* prop_count = reinterpret_cast<int*>(buf("prop_count"));
    if (stages & PE_STAGE_RPN) {
      // RPN head on p2..p6 (rpn.py:74-85): 3x3+ReLU, then objectness + deltas as one 16-wide fp32 GEMM
      RpnLevels lv = {};
      for (int l = 2; l <= 6; ++l) {
        char q[16];
        snprintf(q, sizeof(q), "rpn_out%d", l);
        // 3x3 + ReLU and objectness | deltas as ONE kernel: the 256-channel hidden tensor stays in shared memory as the A operand of
        // the chained 16-wide GEMM (conv_gemm.cu ConvArgs::chain_fp32); PE_RPN_CHAIN=0 runs the two layers separately (A/B switch)
        static const int rpn_chain = [] { const char* e = getenv("PE_RPN_CHAIN"); return e ? atoi(e) : 1; }();
        if (rpn_chain && d->fc == 256) {
          if (status == PE_OK) {
            const Param& p3 = d->params[d->find_param("proposal_generator.rpn_head.conv")];
            const Param& p1 = d->params[d->find_param("proposal_generator.rpn_head")];
            pe_conv_desc cd;
            cd.N = B; cd.H = d->H[l - 1]; cd.W = d->W[l - 1]; cd.Cin = p3.Cin; cd.Cout = p3.Cout; cd.KH = 3; cd.KW = 3; cd.stride = 1;
            cd.relu = 1; cd.residual_mode = 0; cd.out_fp32 = 0; cd.in_fp16 = 0;
            char ql[20];
            snprintf(ql, sizeof(ql), "rpn_logit%d", l);
            ConvChain ch = {wts + p1.w_off, reinterpret_cast<const float*>(wts + p1.b_off), buf(q), 16, 0, 0, 1, buf(ql)};
            lv.logit[l - 2] = reinterpret_cast<const float*>(buf(ql));
            status = gemm(cd, level_feat(l), wts + p3.w_off, reinterpret_cast<const float*>(wts + p3.b_off), nullptr, nullptr, nullptr, &ch);
          }
        } else {
          conv("proposal_generator.rpn_head.conv", level_feat(l), d->H[l - 1], d->W[l - 1], 1, true, 0, nullptr, buf("rpn_t"));
          conv("proposal_generator.rpn_head", buf("rpn_t"), d->H[l - 1], d->W[l - 1], 1, false, 0, nullptr, buf(q), true);
        }
        lv.out[l - 2] = reinterpret_cast<const float*>(buf(q));
        if (!(rpn_chain && d->fc == 256)) lv.logit[l - 2] = nullptr;
        lv.H[l - 2] = d->H[l - 1];
        lv.W[l - 2] = d->W[l - 1];
        lv.stride[l - 2] = 2 << (l - 1);
        const double size = 32.0 * (1 << (l - 2));
        const double ratios[3] = {0.5, 1.0, 2.0};
        for (int a = 0; a < 3; ++a) {  // anchor_generator.py:151-187
          const double w = sqrt(size * size / ratios[a]), h = ratios[a] * w;
          lv.anchor[l - 2][a][0] = (float)(-w / 2.0);
          lv.anchor[l - 2][a][1] = (float)(-h / 2.0);
          lv.anchor[l - 2][a][2] = (float)(w / 2.0);
          lv.anchor[l - 2][a][3] = (float)(h / 2.0);
        }
      }
      RpnScratch rs;
      rs.cand_box = reinterpret_cast<float4*>(buf("cand_box"));
      rs.cand_score = reinterpret_cast<float*>(buf("cand_score"));
      rs.cand_valid = reinterpret_cast<unsigned char*>(buf("cand_valid"));
      rs.cand_count = reinterpret_cast<int*>(buf("cand_count"));
      rs.keep_idx = reinterpret_cast<int*>(buf("keep_idx"));
      rs.keep_count = reinterpret_cast<int*>(buf("keep_count"));
      rs.nms_mask = reinterpret_cast<unsigned*>(buf("nms_mask"));
      if (status == PE_OK) {
        end_run();
        aux_begin("launch_rpn_proposals(");
        check(launch_rpn_proposals(lv, B, c.pre_nms_topk, c.post_nms_topk, c.rpn_nms_thresh, (float)img_h, (float)img_w, rs, kMaxProps,
                                   props, prop_count, st), 4);
        aux_end();
      }
    }
    if (!(stages & PE_STAGE_ROI_HEADS)) { end_run(); signal_stagger(true); return status; }
    // ROI heads (roi_heads.py:595-631)
    RoiLevels fl;
    for (int l = 2; l <= 5; ++l) {
      fl.feat[l - 2] = reinterpret_cast<const __nv_bfloat16*>(level_feat(l));
      fl.H[l - 2] = d->H[l - 1];
      fl.W[l - 2] = d->W[l - 1];
      fl.scale[l - 2] = 1.0f / (float)(2 << (l - 1));
    }
    if (status == PE_OK) PE_NONGEMM(launch_roi_align(fl, props, prop_count, B, kMaxProps, d->fc, buf("roi_feats"), st));
    linear("roi_heads.box_head.fc1", buf("roi_feats"), B * kMaxProps, true, buf("fc1_out"), false);
    linear("roi_heads.box_head.fc2", buf("fc1_out"), B * kMaxProps, true, buf("fc2_out"), false);
    linear("roi_heads.box_predictor", buf("fc2_out"), B * kMaxProps, false, buf("head_out"), true);
    HeadParams hp;
    hp.img_h = (float)img_h; hp.img_w = (float)img_w; hp.out_h = out_h; hp.out_w = out_w;
    hp.scale_x = (float)((double)out_w / (double)img_w);
    hp.scale_y = (float)((double)out_h / (double)img_h);
    hp.score_thresh = c.score_thresh; hp.nms_thresh = c.nms_thresh; hp.max_det = c.detections_per_image;
    DetOut out;
    out.boxes = reinterpret_cast<float4*>(o.boxes); out.scores = o.scores; out.classes = o.classes; out.logits = o.class_logits;
    out.probs = o.probs; out.vars = o.vars; out.roi_index = o.roi_index; out.count = o.counts;
    if (status == PE_OK)
      PE_NONGEMM(launch_head_post(reinterpret_cast<const float*>(buf("head_out")), d->npad, props, prop_count, B, kMaxProps, c.num_classes, hp, out, st));
    end_run();
    signal_stagger(true);  // a waiter must never be left without its event (short stage masks, errors)
    return status;
  }
};

}  // namespace
}  // namespace pe

extern "C" PE_API int pe_detector_create(const pe_detector_config* cfg, pe_detector** out) {
  if (!cfg || !out) return PE_ERR_INVALID_ARGUMENT;
  if (cfg->depth != 50 && cfg->depth != 101) return PE_ERR_UNSUPPORTED;
  // K = 1 (KAIST) and K = 3 (FLIR) have compile-time head kernels; any other class count up to 1000 (the 80 COCO classes
  // of the reference's rgb_only zoo model) takes the run-time variant, which needs score_thresh >= 0.5 (one class per ROI)
  if (cfg->num_classes < 1 || cfg->num_classes > 1000) return PE_ERR_UNSUPPORTED;
  if (cfg->num_classes != 1 && cfg->num_classes != 3 && !(cfg->score_thresh >= 0.5f)) return PE_ERR_UNSUPPORTED;
  if (cfg->max_batch < 1 || cfg->canvas_h % 32 || cfg->canvas_w % 32 || cfg->canvas_h < 64 || cfg->canvas_w < 64) return PE_ERR_INVALID_ARGUMENT;
  if (cfg->middle_fusion ? cfg->in_channels != 6 : (cfg->in_channels < 1 || cfg->in_channels > 4)) return PE_ERR_INVALID_ARGUMENT;
  if (cfg->pre_nms_topk < 1 || cfg->pre_nms_topk > pe::kTopkSlots || cfg->post_nms_topk < 1 || cfg->post_nms_topk > pe::kMaxProps)
    return PE_ERR_UNSUPPORTED;
  if (cfg->detections_per_image < 1 || cfg->detections_per_image > pe::kMaxDet) return PE_ERR_UNSUPPORTED;
  pe_detector* d = new pe_detector();
  d->cfg = *cfg;
  pe::build_plan(d);
  *out = d;
  return PE_OK;
}

extern "C" PE_API void pe_detector_destroy(pe_detector* d) {
  if (d) for (cudaEvent_t e : d->ev) cudaEventDestroy(e);
  if (d) for (cudaEvent_t e : d->aux_ev) cudaEventDestroy(e);
  delete d;
}

extern "C" PE_API int pe_detector_set_stagger_event(pe_detector* d, void* event, int after_launches) {
  if (!d || after_launches < 0) return PE_ERR_INVALID_ARGUMENT;
  d->stagger_ev = reinterpret_cast<cudaEvent_t>(event);
  d->stagger_after = after_launches;
  return PE_OK;
}

extern "C" PE_API int pe_detector_set_profiling(pe_detector* d, int enabled) {
  if (!d) return PE_ERR_INVALID_ARGUMENT;
  d->profiling = enabled != 0;
  d->profile_runs = enabled == 2;  // 2: one event pair per run of back-to-back GEMM launches
  d->run_open = false;
  return PE_OK;
}

// Synchronises the stream of the last forward through the recorded events and reports the summed device time of
// the tensor-core GEMM launches, the span from the first to the last of them, and the launch counts.
extern "C" PE_API int pe_detector_last_profile(pe_detector* d, float* gemm_ms, float* span_ms, int* launches, int* gemm_launches) {
  if (!d) return PE_ERR_INVALID_ARGUMENT;
  if (launches) *launches = d->last_launches;
  if (gemm_launches) *gemm_launches = d->last_gemm_launches;
  float sum = 0.f, span = 0.f;
  if (d->profiling && d->ev_used >= 2) {
    PE_CUDA_CHECK(cudaEventSynchronize(d->ev[d->ev_used - 1]));
    for (int i = 0; i + 1 < d->ev_used; i += 2) {
      float ms = 0.f;
      PE_CUDA_CHECK(cudaEventElapsedTime(&ms, d->ev[i], d->ev[i + 1]));
      sum += ms;
    }
    PE_CUDA_CHECK(cudaEventElapsedTime(&span, d->ev[0], d->ev[d->ev_used - 1]));
  }
  if (gemm_ms) *gemm_ms = sum;
  if (span_ms) *span_ms = span;
  return PE_OK;
}

// Per-launch view of the same instrumentation: device time, algorithmic FLOPs and algorithmic bytes (each operand
// once: input, weights, bias, residual, output) of GEMM launch i of the last forward.  Returns the launch count.
extern "C" PE_API int pe_detector_profile_launches(pe_detector* d, float* ms, double* flops, double* bytes, int capacity) {
  if (!d) return 0;
  const int n = d->ev_used / 2;
  if (!d->profiling || n == 0) return 0;
  if (cudaEventSynchronize(d->ev[d->ev_used - 1]) != cudaSuccess) return 0;
  for (int i = 0; i < n && i < capacity; ++i) {
    float t = 0.f;
    cudaEventElapsedTime(&t, d->ev[2 * i], d->ev[2 * i + 1]);
    if (ms) ms[i] = t;
    if (flops) flops[i] = i < (int)d->prof_flops.size() ? d->prof_flops[i] : 0.0;
    if (bytes) bytes[i] = i < (int)d->prof_bytes.size() ? d->prof_bytes[i] : 0.0;
  }
  return n;
}

// Non-GEMM launch groups of the last mode-1 forward: launcher name (31 chars + NUL per entry) and device time.
extern "C" PE_API int pe_detector_profile_kernels(pe_detector* d, char* names, float* ms, int capacity) {
  if (!d) return 0;
  const int n = d->aux_used / 2;
  if (!d->profiling || n == 0) return 0;
  if (cudaEventSynchronize(d->aux_ev[d->aux_used - 1]) != cudaSuccess) return 0;
  for (int i = 0; i < n && i < capacity; ++i) {
    float t = 0.f;
    cudaEventElapsedTime(&t, d->aux_ev[2 * i], d->aux_ev[2 * i + 1]);
    if (ms) ms[i] = t;
    if (names) snprintf(names + (size_t)i * 32, 32, "%s", i < (int)d->aux_names.size() ? d->aux_names[i].c_str() : "");
  }
  return n;
}

extern "C" PE_API int pe_resize_frames(const uint8_t* frames, float* out, int B, int C, int src_h, int src_w, int dst_h, int dst_w,
                                       int round_u8, void* stream) {
  if (!frames || !out || B < 1 || C < 1 || src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1) return PE_ERR_INVALID_ARGUMENT;
  return pe::launch_resize_frames(frames, out, B, C, src_h, src_w, dst_h, dst_w, round_u8, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" PE_API int pe_detector_num_params(const pe_detector* d) { return d ? (int)d->params.size() : 0; }

extern "C" PE_API int pe_detector_param_info(const pe_detector* d, int i, pe_param_info* info) {
  if (!d || !info || i < 0 || i >= (int)d->params.size()) return PE_ERR_INVALID_ARGUMENT;
  const pe::Param& p = d->params[i];
  memset(info, 0, sizeof(*info));
  snprintf(info->name, sizeof(info->name), "%s", p.name.c_str());
  info->kind = p.kind; info->cout = p.Cout; info->kh = p.KH; info->kw = p.KW; info->cin = p.Cin;
  info->weight_offset = p.w_off; info->bias_offset = p.b_off;
  return PE_OK;
}

extern "C" PE_API size_t pe_detector_weight_bytes(const pe_detector* d) { return d ? d->weight_bytes : 0; }
extern "C" PE_API size_t pe_detector_workspace_bytes(const pe_detector* d) { return d ? d->ws_bytes : 0; }

extern "C" PE_API int pe_detector_buffer_info(const pe_detector* d, const char* name, size_t* offset, int* dims4, int* elem_bytes) {
  if (!d || !name) return PE_ERR_INVALID_ARGUMENT;
  const pe::Buf* b = d->find_buf(name);
  if (!b) return PE_ERR_INVALID_ARGUMENT;
  if (offset) *offset = b->off;
  if (dims4) for (int i = 0; i < 4; ++i) dims4[i] = b->d[i];
  if (elem_bytes) *elem_bytes = b->elem;
  return PE_OK;
}

extern "C" PE_API int pe_detector_forward(pe_detector* d, const void* weights, const float* images, int B, int img_h, int img_w,
                                          float out_h, float out_w, const pe_detections* out, void* workspace, size_t workspace_bytes,
                                          void* stream) {
  return pe_detector_forward_stages(d, weights, images, B, img_h, img_w, out_h, out_w, out, workspace, workspace_bytes, PE_STAGE_ALL, 0,
                                    stream);
}

extern "C" PE_API int pe_detector_forward_stages(pe_detector* d, const void* weights, const float* images, int B, int img_h, int img_w,
                                                 float out_h, float out_w, const pe_detections* out, void* workspace,
                                                 size_t workspace_bytes, int stages, int prenormalized, void* stream) {
  if (!d || !weights || !workspace || !(stages & PE_STAGE_ALL)) return PE_ERR_INVALID_ARGUMENT;
  if ((stages & PE_STAGE_BACKBONE) && !images) return PE_ERR_INVALID_ARGUMENT;
  if ((stages & PE_STAGE_ROI_HEADS) && !out) return PE_ERR_INVALID_ARGUMENT;
  if (B < 1 || B > d->cfg.max_batch) return PE_ERR_INVALID_ARGUMENT;
  if (img_h < 1 || img_w < 1 || img_h > d->cfg.canvas_h || img_w > d->cfg.canvas_w) return PE_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < d->ws_bytes) return PE_ERR_WORKSPACE_TOO_SMALL;
  d->ev_used = 0;
  d->aux_used = 0;
  d->aux_names.clear();
  d->run_open = false;
  d->prof_flops.clear();
  d->prof_bytes.clear();
  d->last_launches = 0;
  d->stagger_sent = false;
  d->last_gemm_launches = 0;
  pe::Runner r;
  r.d = d;
  r.wts = reinterpret_cast<const unsigned char*>(weights);
  r.ws = reinterpret_cast<unsigned char*>(workspace);
  r.st = reinterpret_cast<cudaStream_t>(stream);
  r.B = B;
  r.stages = stages;
  r.prenormalized = prenormalized != 0;
  static const pe_detections none = {};
  return r.run(images, img_h, img_w, out_h, out_w, out ? *out : none);
}

extern "C" PE_API int pe_detector_forward_frames(pe_detector* d, const void* weights, const uint8_t* frames, int B, int src_h, int src_w,
                                                 int img_h, int img_w, int round_u8, float out_h, float out_w, const pe_detections* out,
                                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!d || !weights || !frames || !out || !workspace) return PE_ERR_INVALID_ARGUMENT;
  if (B < 1 || B > d->cfg.max_batch || src_h < 1 || src_w < 1) return PE_ERR_INVALID_ARGUMENT;
  if (img_h < 1 || img_w < 1 || img_h > d->cfg.canvas_h || img_w > d->cfg.canvas_w) return PE_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < d->ws_bytes) return PE_ERR_WORKSPACE_TOO_SMALL;
  d->ev_used = 0;
  d->aux_used = 0;
  d->aux_names.clear();
  d->run_open = false;
  d->prof_flops.clear();
  d->prof_bytes.clear();
  d->last_launches = 0;
  d->stagger_sent = false;
  d->last_gemm_launches = 0;
  pe::Runner r;
  r.d = d;
  r.wts = reinterpret_cast<const unsigned char*>(weights);
  r.ws = reinterpret_cast<unsigned char*>(workspace);
  r.st = reinterpret_cast<cudaStream_t>(stream);
  r.B = B;
  r.frames = frames;
  r.src_h = src_h;
  r.src_w = src_w;
  r.round_u8 = round_u8;
  return r.run(nullptr, img_h, img_w, out_h, out_w, *out);
}

extern "C" PE_API int pe_pack_detections(const pe_detections* models, int M, int B, int K, int32_t* det_offsets, float* boxes,
                                         float* scores, int32_t* classes, float* probs, float* vars, void* stream) {
  if (!models || M < 1 || M > 4 || B < 1 || !det_offsets) return PE_ERR_INVALID_ARGUMENT;
  pe::PackIn in;
  for (int m = 0; m < M; ++m) {
    in.boxes[m] = reinterpret_cast<const float4*>(models[m].boxes);
    in.scores[m] = models[m].scores; in.classes[m] = models[m].classes; in.probs[m] = models[m].probs;
    in.vars[m] = models[m].vars; in.count[m] = models[m].counts;
  }
  return pe::launch_pack(in, B, M, K, det_offsets, reinterpret_cast<float4*>(boxes), scores, classes, probs, vars,
                         reinterpret_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------- op-level entry points
extern "C" PE_API int pe_rpn_proposals(const float* const* rpn_out, const int* H, const int* W, int B, int pre_nms_topk,
                                       int post_nms_topk, float nms_thresh, float img_h, float img_w, float* proposals,
                                       int32_t* proposal_counts, void* workspace, size_t workspace_bytes, void* stream) {
  if (!rpn_out || !H || !W || !proposals || !proposal_counts || !workspace || B < 1) return PE_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < pe_rpn_proposals_workspace_bytes(B)) return PE_ERR_WORKSPACE_TOO_SMALL;
  pe::RpnLevels lv;
  for (int l = 0; l < pe::kRpnLevels; ++l) {
    lv.out[l] = rpn_out[l]; lv.logit[l] = nullptr; lv.H[l] = H[l]; lv.W[l] = W[l]; lv.stride[l] = 4 << l;
    const double size = 32.0 * (1 << l);
    const double ratios[3] = {0.5, 1.0, 2.0};
    for (int a = 0; a < 3; ++a) {
      const double w = sqrt(size * size / ratios[a]), h = ratios[a] * w;
      lv.anchor[l][a][0] = (float)(-w / 2.0); lv.anchor[l][a][1] = (float)(-h / 2.0);
      lv.anchor[l][a][2] = (float)(w / 2.0); lv.anchor[l][a][3] = (float)(h / 2.0);
    }
  }
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  const size_t n = (size_t)B * pe::kRpnLevels * pe::kTopkSlots;
  pe::RpnScratch rs;
  rs.cand_box = reinterpret_cast<float4*>(ws); ws += n * 16;
  rs.nms_mask = reinterpret_cast<unsigned*>(ws); ws += n * 32 * 4;
  rs.cand_score = reinterpret_cast<float*>(ws); ws += n * 4;
  rs.keep_idx = reinterpret_cast<int*>(ws); ws += n * 4;
  rs.cand_count = reinterpret_cast<int*>(ws); ws += (size_t)B * pe::kRpnLevels * 4;
  rs.keep_count = reinterpret_cast<int*>(ws); ws += (size_t)B * pe::kRpnLevels * 4;
  rs.cand_valid = ws;
  return pe::launch_rpn_proposals(lv, B, pre_nms_topk, post_nms_topk, nms_thresh, img_h, img_w, rs, pe::kMaxProps,
                                  reinterpret_cast<float4*>(proposals), proposal_counts, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" PE_API size_t pe_rpn_proposals_workspace_bytes(int B) {
  const size_t n = (size_t)(B > 0 ? B : 0) * pe::kRpnLevels * pe::kTopkSlots;
  return n * (16 + 4 + 4 + 1 + 128) + (size_t)(B > 0 ? B : 0) * pe::kRpnLevels * 8 + 256;
}

extern "C" PE_API int pe_roi_align_fwd(const void* const* features, const int* H, const int* W, int C, const float* proposals,
                                       const int32_t* proposal_counts, int B, int max_props, void* out, void* stream) {
  if (!features || !H || !W || !proposals || !proposal_counts || !out || B < 1 || max_props < 1) return PE_ERR_INVALID_ARGUMENT;
  pe::RoiLevels fl;
  for (int l = 0; l < 4; ++l) {
    fl.feat[l] = reinterpret_cast<const __nv_bfloat16*>(features[l]);
    fl.H[l] = H[l]; fl.W[l] = W[l]; fl.scale[l] = 1.0f / (float)(4 << l);
  }
  return pe::launch_roi_align(fl, reinterpret_cast<const float4*>(proposals), proposal_counts, B, max_props, C, out,
                              reinterpret_cast<cudaStream_t>(stream));
}

extern "C" PE_API int pe_head_postprocess(const float* head_out, int npad, const float* proposals, const int32_t* proposal_counts,
                                          int B, int max_props, int K, float img_h, float img_w, float out_h, float out_w,
                                          float score_thresh, float nms_thresh, int detections_per_image,
                                          const pe_detections* out, void* stream) {
  if (!head_out || !proposals || !proposal_counts || !out || B < 1) return PE_ERR_INVALID_ARGUMENT;
  if (npad < (K + 1) + 4 * K + 1) return PE_ERR_INVALID_ARGUMENT;
  pe::HeadParams hp;
  hp.img_h = img_h; hp.img_w = img_w; hp.out_h = out_h; hp.out_w = out_w;
  hp.scale_x = (float)((double)out_w / (double)img_w);
  hp.scale_y = (float)((double)out_h / (double)img_h);
  hp.score_thresh = score_thresh; hp.nms_thresh = nms_thresh; hp.max_det = detections_per_image;
  pe::DetOut o;
  o.boxes = reinterpret_cast<float4*>(out->boxes); o.scores = out->scores; o.classes = out->classes; o.logits = out->class_logits;
  o.probs = out->probs; o.vars = out->vars; o.roi_index = out->roi_index; o.count = out->counts;
  return pe::launch_head_post(head_out, npad, reinterpret_cast<const float4*>(proposals), proposal_counts, B, max_props, K, hp, o,
                              reinterpret_cast<cudaStream_t>(stream));
}
