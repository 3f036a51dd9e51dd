// ProbEn late fusion on sm_100a: cross-model IoU match + score fusion + box fusion + per-class NMS,
// one launch for a whole batch of images.
//
// Replaces the Python loop of the reference's demo/FLIR/demo_probEn.py (:189-196 fusion, :92-187
// nms_bayesian, :32-42 bayesian_fusion_multiclass, :73-77 weighted_box_fusion, :44-71 nms_1, :236-267
// per-image dispatch).  See include/probenb200.h for the data layout.
//
// Mapping: one WARP per image when the image has <= 32 detections over all models (the realistic
// regime: score>0.5 leaves tens), lane j <-> detection j; each lane loads its record with coalesced
// loads (float4 box), the sort is rank-by-counting over warp shuffles, the greedy clustering walks
// cluster heads with redux.sync min + ballot, members are folded into their head with shuffles.
// Images with 33..256 detections (the detector pipeline: 100 per model) go to a work list handled by
// fuse_mid_kernel (thread per detection, symmetric pair circle, bit fixed-point clustering, several
// blocks per SM); 257..1024 detections take the block-per-image kernel that materialises the full
// suppression bitmask in shared memory and walks the heads serially.
//
// Numerics: the reference works in float64 on float32-exact inputs.  Membership (iou > thr) is decided
// in float32 and re-evaluated with the reference's exact float64 expression whenever the float32 margin
// is within 1e-5 of the threshold, so cluster membership is decision-exact; box fusion accumulates in
// float64; probEn log/exp run in float32 with the background mass 1-sum(p) formed in float64 (its sign
// drives the reference's NaN behaviour).
#include <math.h>
#include <stdlib.h>
#include "common.cuh"

namespace pe {
namespace {

constexpr int kMaxBlockDets = 1024;  // cap for the block-per-image path
constexpr int kBlockThreads = 256;
constexpr int kMidDets = 256;        // 33..256 detections: thread-per-detection block kernel (fuse_mid_kernel)

struct FuseArgs {
  const float4* boxes;
  const float* scores;
  const int* classes;
  const float* probs;
  const float* vars;
  const int* offs;
  int B, M;
  float thr;
  int score_mode, box_mode;
  float img_w, img_h;
  float4* out_boxes;
  float* out_scores;
  int* out_classes;
  int* out_counts;
  int bucket;      // fuse_mid_kernel: walk same-class partners only (PE_FUSE_BUCKET, default 1)
  int* big_count;  // workspace[0]: images with 33..kMidDets detections, workspace[1]: larger ones
  int* big_list;   // workspace[4..4+B): medium images from the front, large images from the back
};

// ---- IoU decisions ---------------------------------------------------------------------------------

// Reference expression, float64, legacy +1 areas and class offsets (demo_probEn.py:100-105,115-123).
__device__ __noinline__ bool match_exact_f64(float4 a, int ca, float4 b, int cb, float img_w, float img_h,
                                             float thr) {
  const double W = (double)img_w, H = (double)img_h;
  const double ax1 = __dadd_rn((double)a.x, __dmul_rn((double)ca, W));
  const double ay1 = __dadd_rn((double)a.y, __dmul_rn((double)ca, H));
  const double ax2 = __dadd_rn((double)a.z, __dmul_rn((double)ca, W));
  const double ay2 = __dadd_rn((double)a.w, __dmul_rn((double)ca, H));
  const double bx1 = __dadd_rn((double)b.x, __dmul_rn((double)cb, W));
  const double by1 = __dadd_rn((double)b.y, __dmul_rn((double)cb, H));
  const double bx2 = __dadd_rn((double)b.z, __dmul_rn((double)cb, W));
  const double by2 = __dadd_rn((double)b.w, __dmul_rn((double)cb, H));
  const double aa = __dmul_rn(__dadd_rn(__dsub_rn(ax2, ax1), 1.0), __dadd_rn(__dsub_rn(ay2, ay1), 1.0));
  const double ab = __dmul_rn(__dadd_rn(__dsub_rn(bx2, bx1), 1.0), __dadd_rn(__dsub_rn(by2, by1), 1.0));
  const double w = fmax(0.0, __dadd_rn(__dsub_rn(fmin(ax2, bx2), fmax(ax1, bx1)), 1.0));
  const double h = fmax(0.0, __dadd_rn(__dsub_rn(fmin(ay2, by2), fmax(ay1, by1)), 1.0));
  const double inter = __dmul_rn(w, h);
  const double ovr = __ddiv_rn(inter, __dsub_rn(__dadd_rn(aa, ab), inter));
  return ovr > (double)thr;
}

// Decision-exact "iou(head, cand) > thr" for the bayesian path.  area_* are float32 (+1) areas.
__device__ __forceinline__ bool match_bayes(float4 h, int hc, float harea, float4 c, int cc, float carea,
                                            float img_w, float img_h, float thr) {
  if (hc == cc) {
    const float w = fmaxf(0.f, fminf(h.z, c.z) - fmaxf(h.x, c.x) + 1.f);
    const float hh = fmaxf(0.f, fminf(h.w, c.w) - fmaxf(h.y, c.y) + 1.f);
    const float inter = w * hh;
    const float uni = harea + carea - inter;
    const float d = inter - thr * uni;
    if (fabsf(d) <= 1e-5f * fabsf(uni)) return match_exact_f64(h, hc, c, cc, img_w, img_h, thr);
    return d > 0.f && uni > 0.f ? true : (uni > 0.f ? false : match_exact_f64(h, hc, c, cc, img_w, img_h, thr));
  }
  // different classes live in different offset tiles; they can only touch through the +1 border.
  const float dx = (float)(cc - hc) * img_w, dy = (float)(cc - hc) * img_h;
  const float w = fminf(h.z, c.z + dx) - fmaxf(h.x, c.x + dx) + 1.f;
  const float hh = fminf(h.w, c.w + dy) - fmaxf(h.y, c.y + dy) + 1.f;
  if (w > -0.01f && hh > -0.01f) return match_exact_f64(h, hc, c, cc, img_w, img_h, thr);
  return false;
}

// torchvision nms float32 expression on coordinate-offset boxes (no FMA contraction).
__device__ __forceinline__ float nms_area(float4 b) {
  return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}
__device__ __forceinline__ bool match_nms(float4 h, float harea, float4 c, float carea, float thr) {
  const float w = fmaxf(0.f, __fsub_rn(fminf(h.z, c.z), fmaxf(h.x, c.x)));
  const float hh = fmaxf(0.f, __fsub_rn(fminf(h.w, c.w), fmaxf(h.y, c.y)));
  const float inter = __fmul_rn(w, hh);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(harea, carea), inter));
  return ovr > thr;
}
__device__ __forceinline__ float4 offset_box(float4 b, float off) {
  return make_float4(__fadd_rn(b.x, off), __fadd_rn(b.y, off), __fadd_rn(b.z, off), __fadd_rn(b.w, off));
}

// ---- per-detection precomputation for the fusion stage -------------------------------------------------

template <int K>
struct DetAux {
  float lg[K + 1];  // log p_k, log(1 - sum p)
  bool bad;         // reference would take log of a negative number -> NaN posterior
  float pmax;       // max_k p_k  (score_mode max reads probs, not scores)
  double wgt;       // box-fusion weight
};

template <int K>
__device__ __forceinline__ DetAux<K> make_aux(const float* p, float score, float var, int box_mode) {
  DetAux<K> a;
  double s = 0.0;
  a.bad = false;
  a.pmax = -INFINITY;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    s = __dadd_rn(s, (double)p[k]);
    a.lg[k] = logf(p[k]);
    a.bad |= (p[k] < 0.f);
    a.pmax = fmaxf(a.pmax, p[k]);
  }
  const double bg = __dsub_rn(1.0, s);
  a.bad |= (bg < 0.0);
  a.lg[K] = logf((float)bg);
  a.wgt = box_mode == PE_BOX_VAVG ? __ddiv_rn(1.0, (double)var) : (box_mode == PE_BOX_SAVG ? (double)score : 1.0);
  return a;
}

// Posterior of the product model (demo_probEn.py:32-42) from summed logs; max-subtracted softmax.
template <int K>
__device__ __forceinline__ void probEn_finish(const float* S, bool bad, float* score, int* cls) {
  if (bad) { *score = __int_as_float(0x7fc00000); *cls = 0; return; }
  float mx = S[0];
  int am = 0;
#pragma unroll
  for (int k = 1; k <= K; ++k)
    if (S[k] > mx) { mx = S[k]; am = k; }
  float den = 0.f;
#pragma unroll
  for (int k = 0; k <= K; ++k) den += expf(S[k] - mx);
  const float r = 1.f / den;  // exp(mx - mx) == 1
  if (r != r) { *score = r; *cls = 0; return; }
  *score = r;
  *cls = am;
}

// ---- packed-warp kernel ----------------------------------------------------------------------------------
// One warp walks a window of kWin consecutive images and packs as many of them as fit into its 32 lanes
// (lane <-> detection); every warp collective below runs per segment through member masks, so ~2 typical
// images (about 15 detections each) share every instruction.  No sort: the next cluster head of a segment is
// the maximum remaining score (redux.sync max + ballot), ties resolved by lane index.

__device__ __forceinline__ float4 shfl4(float4 v, int src) {
  return make_float4(__shfl_sync(kFullMask, v.x, src), __shfl_sync(kFullMask, v.y, src),
                     __shfl_sync(kFullMask, v.z, src), __shfl_sync(kFullMask, v.w, src));
}
__device__ __forceinline__ unsigned score_key(float s) {  // order-preserving float -> uint, never 0 for real scores
  const unsigned u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float key_score(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Pair loop of the packed kernel.  Every lane walks its segment's lanes in a circle (partner = next lane, wrapping
// at the segment end), so after n/2 rounds each unordered pair has been evaluated by one of its two lanes; visits
// beyond that (shorter segments keep turning while the longest one finishes) repeat consistent verdicts, and a
// lane meeting itself only sets its own bit, which is cleared at the end.  kExact = false: float32 decision with an
// error margin, returns true if some decision was inside the margin.  kExact = true: the reference's arithmetic.
struct PairCtx {
  float4 mbox;
  float area;
  unsigned skey;
  int lane, lo, hi, rounds;
  float nthr, thr, ebound;
  const float4* rec;
};

template <bool kNms, bool kExact>
__device__ __forceinline__ bool pair_rounds(const PairCtx& c, const FuseArgs& a, float4 box, int cls, int row0,
                                            unsigned& up_out, unsigned& higher_out) {
  const unsigned mybit = 1u << c.lane, lobit = 1u << c.lo, topbit = 1u << (c.hi - 1);
  const float4* const p_lo = c.rec + 2 * c.lo;
  const float4* pp = c.rec + 2 * c.lane;
  unsigned pbit = mybit, qbit = mybit, up = 0u, higher = 0u;
  bool unsure = false;
#pragma unroll 2
  for (int r = 0; r < c.rounds; ++r) {
    const bool pwrap = pbit == topbit, qwrap = qbit == lobit;
    pbit = pwrap ? lobit : pbit << 1;
    pp = pwrap ? p_lo : pp + 2;
    qbit = qwrap ? topbit : qbit >> 1;
    const float4 pb = pp[0];
    const float4 pa = pp[1];
    bool mt;
    if (kNms) {
      const float w = fmaxf(0.f, __fsub_rn(fminf(c.mbox.z, pb.z), fmaxf(c.mbox.x, pb.x)));
      const float hh = fmaxf(0.f, __fsub_rn(fminf(c.mbox.w, pb.w), fmaxf(c.mbox.y, pb.y)));
      const float inter = __fmul_rn(w, hh);
      const float den = __fsub_rn(__fadd_rn(c.area, pa.x), inter);
      if (kExact) {
        mt = __fdiv_rn(inter, den) > c.thr;
      } else {
        const float d = fmaf(c.nthr, den, inter);
        mt = d > 0.f;
        unsure |= !(fabsf(d) > 1e-5f * den) || !(den > 0.f);  // also when den is NaN
      }
    } else if (kExact) {
      const int prow = row0 + (__ffs(pbit) - 1);
      mt = match_exact_f64(box, cls, __ldg(a.boxes + prow), __ldg(a.classes + prow), a.img_w, a.img_h, c.thr);
    } else {
      // |error of w, hh| <= 2^-21 * cbound each (offset rounding + two float32 operations), so
      // |error of inter - thr * uni| < ebound * (w + hh + 1) + 1e-5 * uni
      const float w = fmaxf(fminf(c.mbox.z, pb.z) - fmaxf(c.mbox.x, pb.x), 0.f);
      const float hh = fmaxf(fminf(c.mbox.w, pb.w) - fmaxf(c.mbox.y, pb.y), 0.f);
      const float inter = w * hh;
      const float uni = (c.area + pa.x) - inter;
      const float d = fmaf(c.nthr, uni, inter);
      const float tol = fmaf(c.ebound, w + hh, fmaf(1e-5f, uni, c.ebound));
      mt = d > 0.f;
      unsure |= !(fabsf(d) > tol);  // also when an area is NaN (flagged detections)
    }
    const unsigned pk = __float_as_uint(pa.y);
    // score desc; ties: higher lane first for the bayesian order, lower lane first for torchvision's stable sort
    const bool before = pk > c.skey || (pk == c.skey && (kNms ? pbit < mybit : pbit > mybit));
    const unsigned all_mt = __ballot_sync(kFullMask, mt);
    const unsigned all_bf = __ballot_sync(kFullMask, before);
    if (before) {
      higher |= pbit;
      if (mt) up |= pbit;
    }
    // the lane whose partner I am ranks me after itself iff its 'before' bit is clear
    higher |= ~all_bf & qbit;
    up |= all_mt & ~all_bf & qbit;
  }
  up_out = up & ~mybit;
  higher_out = higher & ~mybit;
  return unsure;
}

constexpr int kWin = 4;

// SCORE: PE_SCORE_* ; kNms: the ('max','argmax') torchvision-NMS path
template <int K, int SCORE, bool kNms>
__global__ void __launch_bounds__(kBlockThreads, 4) fuse_packed_kernel(const FuseArgs a) {
  __shared__ float4 s_tile[kBlockThreads / 32][64];  // per warp and lane: matching box | (area, score key, -, -)
  const int lane = threadIdx.x & 31;
  float4* const s_rec = s_tile[threadIdx.x >> 5];
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int win = a.M <= 7 ? kWin : 1;
  const int nwin = (a.B + win - 1) / win;

  // fast-decision margin of the class-offset float32 boxes: |coordinate| <= cbound is checked per detection
  const float cbound = (float)(K + 2) * fmaxf(fmaxf(a.img_w, a.img_h), 1.f);
  const float ebound = 1.9073486328125e-6f * cbound;  // 2^-19 * cbound, see the pair loop
  const float nthr = -a.thr;

  for (int wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < nwin; wi += warps_total) {
    const int img0 = wi * win;
    const int nimg = min(win, a.B - img0);
    // offsets of the window: lane L holds offs[img0*M + min(L, nimg*M)]
    const int no = nimg * a.M;
    const int o = __ldg(a.offs + (size_t)img0 * a.M + min(lane, no));
    const int len = __shfl_down_sync(kFullMask, o, 1) - o;
    const unsigned nonempty = __ballot_sync(kFullMask, lane < no && len > 0);
    {  // images that never enter a pack: empty, too large for a warp (block kernel's work list), over the cap
      const int li = min(lane, nimg);
      const int ni = __shfl_sync(kFullMask, o, min(li + 1, nimg) * a.M) - __shfl_sync(kFullMask, o, li * a.M);
      if (lane < nimg) {
        if (ni <= 0) a.out_counts[img0 + lane] = 0;
        else if (ni > kMaxBlockDets) a.out_counts[img0 + lane] = -1;
        else if (ni > kMidDets) a.big_list[a.B - 1 - atomicAdd(a.big_count + 1, 1)] = img0 + lane;  // large: from the back
        else if (ni > 32) a.big_list[atomicAdd(a.big_count, 1)] = img0 + lane;                          // medium: from the front
      }
    }
    int cur = 0;
    while (cur < nimg) {
      // ---- greedy packing of images cur.. into the 32 lanes: cumulative detection counts t1..t4 (warp-uniform)
      const int c0 = __shfl_sync(kFullMask, o, cur * a.M);
      const int t1 = __shfl_sync(kFullMask, o, min(cur + 1, nimg) * a.M) - c0;
      const int t2 = __shfl_sync(kFullMask, o, min(cur + 2, nimg) * a.M) - c0;
      const int t3 = __shfl_sync(kFullMask, o, min(cur + 3, nimg) * a.M) - c0;
      const int t4 = __shfl_sync(kFullMask, o, min(cur + 4, nimg) * a.M) - c0;
      const int npk = min((t1 <= 32) + (t2 <= 32) + (t3 <= 32) + (t4 <= 32), nimg - cur);
      if (npk == 0) { ++cur; continue; }  // more than 32 detections: handled above
      const int used = npk == 1 ? t1 : (npk == 2 ? t2 : (npk == 3 ? t3 : t4));
      // ---- lane -> (segment, row); idle lanes form one-lane segments of their own
      const bool act = lane < used;
      const int sg = (lane >= t1) + (lane >= t2) + (lane >= t3);
      const int lo = !act ? lane : (sg == 0 ? 0 : (sg == 1 ? t1 : (sg == 2 ? t2 : t3)));
      const int hi = !act ? lane + 1 : (sg == 0 ? t1 : (sg == 1 ? t2 : (sg == 2 ? t3 : t4)));
      const int n = hi - lo;
      const int base = c0 + lo;
      const int my_img = img0 + cur + sg;
      const int live = act ? __popc((nonempty >> ((cur + sg) * a.M)) & ((1u << a.M) - 1u)) : 0;
      cur += npk;
      if (used == 0) continue;
      const unsigned segmask = (n == 32 ? kFullMask : ((1u << n) - 1u)) << lo;
      const int row = act ? c0 + lane : 0;
      const bool cluster = live >= 2;
      const float4 box = act ? __ldg(a.boxes + row) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float score = act ? __ldg(a.scores + row) : 0.f;
      const int cls = act ? __ldg(a.classes + row) : 0;
      // fusion-stage inputs are requested now so that their latency hides behind the pair loop
      float pr[K];
#pragma unroll
      for (int k = 0; k < K; ++k) pr[k] = (cluster && !kNms) ? __ldg(a.probs + (size_t)row * K + k) : 0.25f;
      const float var = (cluster && !kNms && a.box_mode == PE_BOX_VAVG) ? __ldg(a.vars + row) : 1.f;
      if (live == 1) {  // single contributing model: pass-through in input order (demo_probEn.py:240-252)
        a.out_boxes[row] = box;
        a.out_scores[row] = score;
        a.out_classes[row] = cls;
        if (lane == lo) a.out_counts[my_img] = n;
      }
      if (!__any_sync(kFullMask, cluster)) continue;

      // ---- stage the matching records of the pack in shared memory (one tile per warp)
      float4 mbox;
      float area;
      if (kNms) {  // torchvision batched_nms coordinate trick: boxes + class * (max coordinate + 1), float32
        float mc = cluster ? fmaxf(fmaxf(box.x, box.y), fmaxf(box.z, box.w)) : -INFINITY;
        {  // segment max through order-preserving keys, full-warp collectives
          const unsigned kk = score_key(mc);
          unsigned best = 0;
#pragma unroll
          for (int s2 = 0; s2 < kWin; ++s2)
            if (s2 < npk) {
              const unsigned ms = __reduce_max_sync(kFullMask, (act && sg == s2) ? kk : 0u);
              if (sg == s2) best = ms;
            }
          mc = key_score(best);
        }
        mbox = offset_box(box, __fmul_rn((float)cls, __fadd_rn(mc, 1.f)));
        area = nms_area(mbox);
      } else {
        // class-offset boxes (demo_probEn.py:100-105) in float32 with the legacy +1 folded into x2, y2: only used
        // for the fast decision, whose margin covers the rounding; borderline pairs re-evaluate the float64
        // expression.  Detections outside the margin's coordinate bound, with non-positive area or non-finite
        // fields get a NaN area, which sends every pair they take part in to the exact path.
        const float fc = (float)cls;
        mbox = make_float4(fmaf(fc, a.img_w, box.x), fmaf(fc, a.img_h, box.y), fmaf(fc, a.img_w, box.z) + 1.f,
                           fmaf(fc, a.img_h, box.w) + 1.f);
        area = (box.z - box.x + 1.f) * (box.w - box.y + 1.f);
        const float emax = fmaxf(fmaxf(fabsf(mbox.x), fabsf(mbox.y)), fmaxf(fabsf(mbox.z), fabsf(mbox.w)));
        if (!(emax <= cbound) || !(area > 0.f)) area = __int_as_float(0x7fc00000);
      }
      const unsigned skey = score_key(score);
      __syncwarp();  // the previous pack's readers are done with the tile
      s_rec[2 * lane] = mbox;
      s_rec[2 * lane + 1] = make_float4(area, __uint_as_float(skey), 0.f, 0.f);
      __syncwarp();

      // ---- all pairs of a segment: in round r lane idx meets idx + r (mod n) and the verdicts travel back
      //      through two ballots.  up = earlier-ranked lanes I match, higher = lanes ranked before me.
      //      Borderline float32 decisions only raise a flag; a pack that saw one is redone with exact decisions.
      unsigned up, higher;
      {
        const int rounds = __reduce_max_sync(kFullMask, cluster ? (n >> 1) : 0);
        PairCtx c;
        c.mbox = mbox; c.area = area; c.skey = skey; c.lane = lane; c.lo = lo; c.hi = hi; c.rounds = rounds;
        c.nthr = nthr; c.thr = a.thr; c.ebound = ebound; c.rec = s_rec;
        const bool unsure = pair_rounds<kNms, false>(c, a, box, cls, c0, up, higher);
        if (__any_sync(kFullMask, unsure)) pair_rounds<kNms, true>(c, a, box, cls, c0, up, higher);
        if (!cluster) up = 0u;
      }

      // ---- greedy clustering as a fixed point over the pair bits: a detection is a cluster head iff every
      //      earlier-ranked detection it matches was itself absorbed; absorbed iff it matches an earlier head.
      //      Each round settles at least the earliest undecided detection of every segment.
      const unsigned clmask = __ballot_sync(kFullMask, cluster);
      bool is_head = cluster && up == 0u, is_rem = false;
      unsigned heads = __ballot_sync(kFullMask, is_head);
      while (true) {
        is_rem = (up & heads) != 0u;
        const unsigned rem = __ballot_sync(kFullMask, is_rem);
        is_head = cluster && (up & ~rem) == 0u;
        heads = __ballot_sync(kFullMask, is_head);
        if (((heads | rem) & clmask) == clmask) break;
      }
      // owner of an absorbed detection: the earliest-ranked head it matches
      int owner = lane;
      {
        unsigned cand = is_rem ? (up & heads) : 0u;
        if (cand) owner = __ffs(cand) - 1;
        if (__any_sync(kFullMask, (cand & (cand - 1u)) != 0u)) {
          const int myrank = __popc(higher);
          int bestrank = 0x7fffffff;
          while (__any_sync(kFullMask, cand != 0u)) {
            const bool has = cand != 0u;
            const int c = has ? __ffs(cand) - 1 : lane;
            if (has) cand &= cand - 1u;
            const int rc = __shfl_sync(kFullMask, myrank, c);
            if (has && rc < bestrank) { bestrank = rc; owner = c; }
          }
        }
      }
      const unsigned same_owner = __match_any_sync(kFullMask, owner);
      const unsigned my_cluster = is_head ? (same_owner & ~(1u << lane)) : 0u;
      const int my_pos = is_head ? __popc(heads & higher) : -1;
      if (cluster && lane == lo) a.out_counts[my_img] = __popc(heads & segmask);

      if (kNms) {  // survivors keep their own record (demo_probEn.py:66-69)
        if (my_pos >= 0) {
          a.out_boxes[base + my_pos] = box;
          a.out_scores[base + my_pos] = score;
          a.out_classes[base + my_pos] = cls;
        }
        continue;
      }

      // ---- per-detection terms, then fold members into their head
      float lg[K + 1];
      float pmax = -INFINITY;
      {
        double sp = 0.0;
        bool bad = false;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float p = pr[k];
          sp = __dadd_rn(sp, (double)p);
          lg[k] = __logf(p);
          bad |= p < 0.f;
          pmax = fmaxf(pmax, p);
        }
        const double bg = __dsub_rn(1.0, sp);
        lg[K] = __logf((float)bg);
        if (bad || bg < 0.0) lg[0] = __int_as_float(0x7fc00000);  // the reference's log(negative) -> NaN posterior
      }
      float wgt = 1.f;
      if (a.box_mode == PE_BOX_VAVG) wgt = __frcp_rn(var);
      else if (a.box_mode == PE_BOX_SAVG) wgt = score;

      float S[K + 1];
#pragma unroll
      for (int k = 0; k <= K; ++k) S[k] = lg[k];
      float ssum = score, wsum = wgt, pm = pmax;
      float dx = 0.f, dy = 0.f, dz = 0.f, dw = 0.f;  // sum of w * (member box - head box): small, exact-ish in fp32
      int cnt = 1;
      int best_lane = -1;  // argmax box: first member in cluster order (= highest lane) whose score ties the head's
      unsigned rem = my_cluster;
      while (__any_sync(kFullMask, rem != 0u)) {
        const bool has = rem != 0u;
        const int src = has ? 31 - __clz(rem) : lane;
        if (has) rem &= ~(1u << src);
        const float4 ob = shfl4(box, src);
        const float ow = __shfl_sync(kFullMask, wgt, src);
        float ol[K + 1];
        float os = 0.f, opm = 0.f;
        if (SCORE == PE_SCORE_PROBEN) {
#pragma unroll
          for (int k = 0; k <= K; ++k) ol[k] = __shfl_sync(kFullMask, lg[k], src);
        } else if (SCORE == PE_SCORE_MAX) {
          opm = __shfl_sync(kFullMask, pmax, src);
        }
        if (SCORE == PE_SCORE_AVG || a.box_mode == PE_BOX_ARGMAX) os = __shfl_sync(kFullMask, score, src);
        if (has) {
          if (SCORE == PE_SCORE_PROBEN) {
#pragma unroll
            for (int k = 0; k <= K; ++k) S[k] += ol[k];
          } else if (SCORE == PE_SCORE_MAX) {
            pm = fmaxf(pm, opm);
          } else {
            ssum += os;
          }
          wsum += ow;
          dx += ow * (ob.x - box.x); dy += ow * (ob.y - box.y); dz += ow * (ob.z - box.z); dw += ow * (ob.w - box.w);
          ++cnt;
          if (os == score && best_lane < 0) best_lane = src;
        }
      }
      float4 abox = box;  // argmax box: the tied member's box if there is one, else the head's own
      if (a.box_mode == PE_BOX_ARGMAX) abox = shfl4(box, best_lane >= 0 ? best_lane : lane);
      if (my_pos >= 0) {
        float fs = score;
        int fc = cls;
        float4 fb = box;
        if (cnt > 1) {
          if (SCORE == PE_SCORE_PROBEN) {
            float mx = S[0];
            int am = 0;
#pragma unroll
            for (int k = 1; k <= K; ++k)
              if (S[k] > mx) { mx = S[k]; am = k; }
            float den = 0.f;
#pragma unroll
            for (int k = 0; k <= K; ++k) den += __expf(S[k] - mx);
            fs = __fdividef(1.f, den);
            fc = am;
            bool nan_any = false;
#pragma unroll
            for (int k = 0; k <= K; ++k) nan_any |= S[k] != S[k];
            if (nan_any || fs != fs) { fs = __int_as_float(0x7fc00000); fc = 0; }
          } else if (SCORE == PE_SCORE_AVG) {
            fs = ssum / (float)cnt;
          } else {
            fs = pm;
          }
          if (a.box_mode == PE_BOX_ARGMAX) fb = abox;
          else {
            const float inv = 1.f / wsum;
            fb = make_float4(box.x + dx * inv, box.y + dy * inv, box.z + dz * inv, box.w + dw * inv);
          }
        }
        a.out_boxes[base + my_pos] = fb;
        a.out_scores[base + my_pos] = fs;
        a.out_classes[base + my_pos] = fc;
      }
    }
  }
}

// ---- block-per-image kernel for 33..1024 detections ------------------------------------------------------

template <int K>
__global__ void __launch_bounds__(kBlockThreads) fuse_block_kernel(const FuseArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_box = reinterpret_cast<float4*>(smem_raw);                 // original boxes, rank order
  float4* s_mbox = s_box + kMaxBlockDets;                              // matching boxes (nms: offset)
  float* s_score = reinterpret_cast<float*>(s_mbox + kMaxBlockDets);
  float* s_area = s_score + kMaxBlockDets;
  int* s_cls = reinterpret_cast<int*>(s_area + kMaxBlockDets);
  int* s_src = s_cls + kMaxBlockDets;                                  // rank -> row offset in the image
  int* s_pos = s_src + kMaxBlockDets;                                  // rank -> output position or -1
  unsigned* s_mask = reinterpret_cast<unsigned*>(s_pos + kMaxBlockDets);  // [n][W]
  __shared__ float s_red[kBlockThreads / 32];
  __shared__ int s_nheads;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool nms_path = a.score_mode == PE_SCORE_MAX && a.box_mode == PE_BOX_ARGMAX;
  const int nbig = a.big_count[1];

  for (int bi = blockIdx.x; bi < nbig; bi += gridDim.x) {
    const int img = a.big_list[a.B - 1 - bi];
    const int base = a.offs[(size_t)img * a.M];
    const int n = a.offs[(size_t)img * a.M + a.M] - base;
    const int W = (n + 31) >> 5;
    int live = 0;
    for (int m = 0; m < a.M; ++m) live += a.offs[(size_t)img * a.M + m + 1] > a.offs[(size_t)img * a.M + m];
    __syncthreads();  // previous iteration done with smem
    if (live == 1) {
      for (int j = tid; j < n; j += blockDim.x) {
        a.out_boxes[base + j] = a.boxes[base + j];
        a.out_scores[base + j] = a.scores[base + j];
        a.out_classes[base + j] = a.classes[base + j];
      }
      if (tid == 0) a.out_counts[img] = n;
      continue;
    }
    // stage scores (input order) in s_area, compute ranks, scatter records to rank order
    for (int j = tid; j < n; j += blockDim.x) s_area[j] = a.scores[base + j];
    float mc = -INFINITY;
    __syncthreads();
    for (int j = tid; j < n; j += blockDim.x) {
      const float sj = s_area[j];
      int rank = 0;
      for (int k = 0; k < n; ++k) {
        const float sk = s_area[k];
        rank += (sk > sj) || (sk == sj && (nms_path ? k < j : k > j));
      }
      const float4 b = a.boxes[base + j];
      s_box[rank] = b;
      s_score[rank] = sj;
      s_cls[rank] = a.classes[base + j];
      s_src[rank] = j;
      s_pos[rank] = -1;
      mc = fmaxf(mc, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
    }
    if (nms_path) {  // block max of all coordinates
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) mc = fmaxf(mc, __shfl_xor_sync(kFullMask, mc, s));
      if (lane == 0) s_red[wid] = mc;
    }
    __syncthreads();
    if (nms_path) {
      mc = s_red[0];
      for (int w = 1; w < kBlockThreads / 32; ++w) mc = fmaxf(mc, s_red[w]);
    }
    for (int r = tid; r < n; r += blockDim.x) {
      const float4 b = s_box[r];
      if (nms_path) {
        const float4 ob = offset_box(b, __fmul_rn((float)s_cls[r], __fadd_rn(mc, 1.f)));
        s_mbox[r] = ob;
        s_area[r] = nms_area(ob);
      } else {
        s_mbox[r] = b;
        s_area[r] = (b.z - b.x + 1.f) * (b.w - b.y + 1.f);
      }
    }
    __syncthreads();
    // suppression bitmask: word (i, w) holds bits j = 32w.. with j > i and iou(i, j) > thr
    for (int item = tid; item < n * W; item += blockDim.x) {
      const int i = item / W, w = item - i * W;
      unsigned bits = 0;
      if (w >= (i >> 5)) {
        const float4 hb = s_mbox[i];
        const float ha = s_area[i];
        const int hc = s_cls[i];
        const int j0 = w << 5;
        for (int t = 0; t < 32; ++t) {
          const int j = j0 + t;
          if (j > i && j < n) {
            const bool m = nms_path ? match_nms(hb, ha, s_mbox[j], s_area[j], a.thr)
                                    : match_bayes(hb, hc, ha, s_mbox[j], s_cls[j], s_area[j], a.img_w, a.img_h, a.thr);
            bits |= (unsigned)m << t;
          }
        }
      }
      s_mask[item] = bits;
    }
    __syncthreads();
    // serial head scan by warp 0: lane w owns word w of the removed set (W <= 32)
    if (wid == 0) {
      unsigned removed = 0;
      int nheads = 0;
      for (int i = 0; i < n; ++i) {
        const unsigned rw = __shfl_sync(kFullMask, removed, i >> 5);
        if ((rw >> (i & 31)) & 1u) continue;
        unsigned rowbits = 0;
        if (lane < W) {
          rowbits = s_mask[i * W + lane] & ~removed;
          s_mask[i * W + lane] = rowbits;  // now: members of cluster i
          removed |= rowbits;
          if (lane == (i >> 5)) removed |= 1u << (i & 31);
        }
        if (lane == 0) s_pos[i] = nheads;
        ++nheads;
      }
      if (lane == 0) { s_nheads = nheads; a.out_counts[img] = nheads; }
    }
    __syncthreads();
    // one thread per head folds its members in rank order, head last (the reference's order)
    for (int i = tid; i < n; i += blockDim.x) {
      const int pos = s_pos[i];
      if (pos < 0) continue;
      const float4 hbox = s_box[i];
      const float hscore = s_score[i];
      float fs = hscore;
      int fc = s_cls[i];
      float4 fb = hbox;
      if (!nms_path) {
        float S[K + 1];
#pragma unroll
        for (int k = 0; k <= K; ++k) S[k] = 0.f;
        bool bad = false;
        float pmax = -INFINITY;
        double ssum = 0.0, wsum = 0.0, bx = 0.0, by = 0.0, bz = 0.0, bw = 0.0;
        int cnt = 0;
        bool have_best = false;
        float4 best_box = hbox;
        auto fold = [&](int r) {
          const int rowm = base + s_src[r];
          float pr[K];
#pragma unroll
          for (int k = 0; k < K; ++k) pr[k] = a.probs[(size_t)rowm * K + k];
          const DetAux<K> d = make_aux<K>(pr, s_score[r], a.vars[rowm], a.box_mode);
          const float4 ob = s_box[r];
#pragma unroll
          for (int k = 0; k <= K; ++k) S[k] += d.lg[k];
          bad |= d.bad;
          pmax = fmaxf(pmax, d.pmax);
          ssum += (double)s_score[r];
          wsum += d.wgt;
          bx += d.wgt * (double)ob.x; by += d.wgt * (double)ob.y; bz += d.wgt * (double)ob.z; bw += d.wgt * (double)ob.w;
          ++cnt;
          if (!have_best && s_score[r] == hscore) { have_best = true; best_box = ob; }
        };
        for (int w = i >> 5; w < W; ++w) {
          unsigned bits = s_mask[i * W + w];
          while (bits) {
            const int t = __ffs(bits) - 1;
            bits &= bits - 1u;
            fold((w << 5) + t);
          }
        }
        if (cnt > 0) {
          fold(i);
          if (a.score_mode == PE_SCORE_PROBEN) probEn_finish<K>(S, bad, &fs, &fc);
          else if (a.score_mode == PE_SCORE_AVG) fs = (float)(ssum / (double)cnt);
          else fs = pmax;
          if (a.box_mode == PE_BOX_ARGMAX) fb = best_box;
          else {
            const double inv = 1.0 / wsum;
            fb = make_float4((float)(bx * inv), (float)(by * inv), (float)(bz * inv), (float)(bw * inv));
          }
        }
      }
      a.out_boxes[base + pos] = fb;
      a.out_scores[base + pos] = fs;
      a.out_classes[base + pos] = fc;
    }
  }
}

// ---- block-per-image kernel for 33..256 detections: thread <-> detection -------------------------------------
// The regime of the detector pipeline (100 detections per model and image).  Same decisions and the same fold order as
// fuse_block_kernel, restructured so that nothing runs serially over the detections:
//   * records are read once, the per-detection terms (logs, weight) are computed by the detection's own thread and everything is
//     scattered to rank order in shared memory (28 KB static: several blocks per SM);
//   * pair matrix: thread i evaluates partners (i + d) mod n, d = 1..n/2, so every unordered pair is evaluated once (twice for
//     the diametrical pairs of an even n) and a match sets both symmetric bits with shared-memory atomics (matches are rare);
//   * greedy clustering without the serial head scan: "i is a head iff no EARLIER head matches i" is iterated as a bit fixed
//     point over all detections at once (after t rounds the first t ranks are final; real scenes settle in 2-4 rounds), the owner
//     of a non-head is the earliest head that matches it, output position = number of earlier heads;
//   * one thread per head folds its members (rank order, head last) from shared memory.
template <int K>
__global__ void __launch_bounds__(kBlockThreads, 4) fuse_mid_kernel(const FuseArgs a) {
  constexpr int CAP = kMidDets, WMAX = kMidDets / 32;
  __shared__ float4 s_box[CAP], s_mbox[CAP];
  __shared__ float s_score[CAP], s_area[CAP], s_pmax[CAP], s_lg[K + 1][CAP];
  __shared__ double s_wgt[CAP];
  __shared__ int s_cls[CAP], s_owner[CAP];
  __shared__ unsigned char s_bad[CAP];
  __shared__ unsigned s_mask[CAP][WMAX];
  __shared__ unsigned s_head[2][WMAX];
  __shared__ float s_red[kBlockThreads / 32];
  __shared__ int s_wc[K + 1][kBlockThreads / 32], s_cq[CAP], s_flag[CAP], s_nflag;
  __shared__ float s_ext[4][kBlockThreads / 32];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool nms_path = a.score_mode == PE_SCORE_MAX && a.box_mode == PE_BOX_ARGMAX;
  const int nmid = a.big_count[0];

  for (int bi = blockIdx.x; bi < nmid; bi += gridDim.x) {
    const int img = a.big_list[bi];
    const int base = a.offs[(size_t)img * a.M];
    const int n = a.offs[(size_t)img * a.M + a.M] - base;
    const int W = (n + 31) >> 5;
    int live = 0;
    for (int m = 0; m < a.M; ++m) live += a.offs[(size_t)img * a.M + m + 1] > a.offs[(size_t)img * a.M + m];
    __syncthreads();  // previous image done with shared memory
    if (live == 1) {
      if (tid < n) {
        a.out_boxes[base + tid] = a.boxes[base + tid];
        a.out_scores[base + tid] = a.scores[base + tid];
        a.out_classes[base + tid] = a.classes[base + tid];
      }
      if (tid == 0) a.out_counts[img] = n;
      continue;
    }
    // ---- this thread's detection (input order): record, per-detection fusion terms, rank by counting
    const bool act = tid < n;
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    float sc = -INFINITY;
    int cls = 0;
    DetAux<K> aux;
    if (act) {
      box = a.boxes[base + tid];
      sc = a.scores[base + tid];
      cls = a.classes[base + tid];
      float pr[K];
#pragma unroll
      for (int k = 0; k < K; ++k) pr[k] = a.probs[(size_t)(base + tid) * K + k];
      aux = make_aux<K>(pr, sc, a.vars[base + tid], a.box_mode);
      s_area[tid] = sc;  // scores in input order (s_area is rewritten below)
    }
#pragma unroll
    for (int w = 0; w < WMAX; ++w) s_mask[tid][w] = 0u;
    __syncthreads();
    int rank = 0;
    if (act) {
      for (int k = 0; k < n; ++k) {
        const float sk = s_area[k];
        rank += (sk > sc) || (sk == sc && (nms_path ? k < tid : k > tid));
      }
    }
    float mc = act ? fmaxf(fmaxf(box.x, box.y), fmaxf(box.z, box.w)) : -INFINITY;
    if (nms_path) {  // block max of all coordinates (torchvision batched_nms offsets)
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) mc = fmaxf(mc, __shfl_xor_sync(kFullMask, mc, s));
      if (lane == 0) s_red[wid] = mc;
    }
    __syncthreads();  // every thread has read the input-order scores
    if (nms_path) {
      mc = s_red[0];
      for (int w = 1; w < kBlockThreads / 32; ++w) mc = fmaxf(mc, s_red[w]);
    }
    if (act) {
      s_box[rank] = box;
      s_score[rank] = sc;
      s_cls[rank] = cls;
#pragma unroll
      for (int k = 0; k <= K; ++k) s_lg[k][rank] = aux.lg[k];
      s_wgt[rank] = aux.wgt;
      s_pmax[rank] = aux.pmax;
      s_bad[rank] = aux.bad ? 1 : 0;
      if (nms_path) {
        const float4 ob = offset_box(box, __fmul_rn((float)cls, __fadd_rn(mc, 1.f)));
        s_mbox[rank] = ob;
        s_area[rank] = nms_area(ob);
      } else {
        s_mbox[rank] = box;
        s_area[rank] = (box.z - box.x + 1.f) * (box.w - box.y + 1.f);
      }
    }
    __syncthreads();
    // ---- from here on thread i <-> rank i.
    const int i = tid;
    // Class buckets: a match needs equal classes - different classes live in different offset tiles (demo_probEn.py:100-105) and
    // can only meet through the +1 border, which takes a box reaching the far corner of its tile and one starting at the near
    // corner of the next (the reference's legacy areas give such degenerate pairs a non-zero overlap).  So every thread walks the
    // circle of ITS CLASS only (K = 3: a third of the pair evaluations), and the rare corner-reaching / corner-starting boxes are
    // collected in a list and tested against each other with the reference's cross-class expression: the set of pairs that can
    // pass `w > -0.01 && h > -0.01` is a subset of (flagged x flagged), every other cross-class pair has an empty intersection.
    bool bucketed = false;
    if (a.bucket) {  // block-uniform
      const int mycls = act ? s_cls[i] : -1;
      const int cb = (mycls >= 0 && mycls < K) ? mycls : K;  // bucket K: class ids outside [0, K) -> no bucketing for this image
      int mypos = 0;
#pragma unroll
      for (int c = 0; c <= K; ++c) {
        const unsigned bal = __ballot_sync(kFullMask, act && cb == c);
        if (cb == c) mypos = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) s_wc[c][wid] = __popc(bal);
      }
      const float4 mb = act ? s_mbox[i] : make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
      float mnx = mb.x, mny = mb.y, mxz = mb.z, mxw = mb.w;
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(kFullMask, mnx, sft)); mny = fminf(mny, __shfl_xor_sync(kFullMask, mny, sft));
        mxz = fmaxf(mxz, __shfl_xor_sync(kFullMask, mxz, sft)); mxw = fmaxf(mxw, __shfl_xor_sync(kFullMask, mxw, sft));
      }
      if (lane == 0) { s_ext[0][wid] = mnx; s_ext[1][wid] = mny; s_ext[2][wid] = mxz; s_ext[3][wid] = mxw; }
      if (tid == 0) s_nflag = 0;
      __syncthreads();
      int cstart = 0, ccount = 0, other = 0;
#pragma unroll
      for (int c = 0; c <= K; ++c) {
        int tot = 0;
        for (int w = 0; w < kBlockThreads / 32; ++w) { const int v = s_wc[c][w]; if (c == cb && w < wid) mypos += v; tot += v; }
        if (c < cb) cstart += tot;
        if (c == cb) ccount = tot;
        if (c == K) other = tot;
      }
      bucketed = other == 0;  // block-uniform
      if (bucketed) {
        bool flagged = false;
        if (act) {
          s_cq[cstart + mypos] = i;
          if (!nms_path) {
            for (int w = 0; w < kBlockThreads / 32; ++w) {
              mnx = fminf(mnx, s_ext[0][w]); mny = fminf(mny, s_ext[1][w]); mxz = fmaxf(mxz, s_ext[2][w]); mxw = fmaxf(mxw, s_ext[3][w]);
            }
            const float tw = a.img_w - 1.02f, th = a.img_h - 1.02f;  // the predicate's -0.01 plus float32 round-off slack
            const bool fa = (mb.z - mnx > tw) && (mb.w - mny > th);   // can be the lower-class side of a border contact
            const bool fb = (mxz - mb.x > tw) && (mxw - mb.y > th);   // ... the higher-class side
            flagged = fa || fb;
            if (flagged) s_flag[atomicAdd(&s_nflag, 1)] = i;
          }
        }
        __syncthreads();
        if (act) {
          const float4 hb = s_mbox[i];
          const float ha = s_area[i];
          const int hc = s_cls[i];
          const int half_c = ccount >> 1, cend = cstart + ccount;
          int pos = cstart + mypos;
          for (int d = 1; d <= half_c; ++d) {
            pos = pos + 1 == cend ? cstart : pos + 1;
            const int j = s_cq[pos];
            const bool m = nms_path ? match_nms(hb, ha, s_mbox[j], s_area[j], a.thr)
                                    : match_bayes(hb, hc, ha, s_mbox[j], hc, s_area[j], a.img_w, a.img_h, a.thr);
            if (m) {
              atomicOr(&s_mask[i][j >> 5], 1u << (j & 31));
              atomicOr(&s_mask[j][i >> 5], 1u << (i & 31));
            }
          }
          if (flagged) {
            const int nf = s_nflag;
            for (int t = 0; t < nf; ++t) {
              const int j = s_flag[t];
              const int cc = s_cls[j];
              if (cc != hc && match_bayes(hb, hc, ha, s_mbox[j], cc, s_area[j], a.img_w, a.img_h, a.thr)) {
                atomicOr(&s_mask[i][j >> 5], 1u << (j & 31));
                atomicOr(&s_mask[j][i >> 5], 1u << (i & 31));
              }
            }
          }
        }
      }
    }
    // Pair matrix over the circle of all detections (no bucketing: class ids outside [0, K), or PE_FUSE_BUCKET=0).
    if (!bucketed && act) {
      const float4 hb = s_mbox[i];
      const float ha = s_area[i];
      const int hc = s_cls[i];
      const int half_n = n >> 1;
      int j = i;
      for (int d = 1; d <= half_n; ++d) {
        j = j + 1 == n ? 0 : j + 1;
        const bool m = nms_path ? match_nms(hb, ha, s_mbox[j], s_area[j], a.thr)
                                : match_bayes(hb, hc, ha, s_mbox[j], s_cls[j], s_area[j], a.img_w, a.img_h, a.thr);
        if (m) {
          atomicOr(&s_mask[i][j >> 5], 1u << (j & 31));
          atomicOr(&s_mask[j][i >> 5], 1u << (i & 31));
        }
      }
    }
    if (tid < WMAX) s_head[0][tid] = tid < W ? (tid == W - 1 && (n & 31) ? (1u << (n & 31)) - 1u : 0xffffffffu) : 0u;
    __syncthreads();
    // ---- heads: bit fixed point.  earlier[w] = bits of my row at ranks below mine
    unsigned row[WMAX];
#pragma unroll
    for (int w = 0; w < WMAX; ++w) {
      const unsigned m = s_mask[i][w];
      row[w] = w < (i >> 5) ? m : (w == (i >> 5) ? m & ((1u << (i & 31)) - 1u) : 0u);
    }
    int cur = 0;
    for (int round = 0; round <= n; ++round) {
      unsigned hit = 0u;
#pragma unroll
      for (int w = 0; w < WMAX; ++w) hit |= row[w] & s_head[cur][w];
      const bool is_head = act && hit == 0u;
      const unsigned bal = __ballot_sync(kFullMask, is_head);
      const bool changed = lane == 0 && bal != s_head[cur][wid];
      if (lane == 0) s_head[cur ^ 1][wid] = bal;
      cur ^= 1;
      if (!__syncthreads_or(changed)) break;
    }
    // ---- owner of every detection, output position of every head
    unsigned hw[WMAX];
#pragma unroll
    for (int w = 0; w < WMAX; ++w) hw[w] = s_head[cur][w];
    const bool is_head = act && ((s_head[cur][i >> 5] >> (i & 31)) & 1u);
    int pos = 0, owner = i, nheads = 0;
#pragma unroll
    for (int w = 0; w < WMAX; ++w) {
      nheads += __popc(hw[w]);
      pos += w < (i >> 5) ? __popc(hw[w]) : (w == (i >> 5) ? __popc(hw[w] & ((1u << (i & 31)) - 1u)) : 0);
    }
    if (act && !is_head) {
#pragma unroll
      for (int w = WMAX - 1; w >= 0; --w) {
        const unsigned h = row[w] & hw[w];
        if (h) owner = (w << 5) + __ffs(h) - 1;  // descending w: the earliest matching head wins
      }
    }
    if (act) s_owner[i] = owner;
    if (tid == 0) a.out_counts[img] = nheads;
    __syncthreads();
    // ---- fold: one thread per head, members in rank order, head last (the reference's order)
    if (is_head) {
      const float4 hbox = s_box[i];
      const float hscore = s_score[i];
      float fs = hscore;
      int fc = s_cls[i];
      float4 fb = hbox;
      if (!nms_path) {
        float S[K + 1];
#pragma unroll
        for (int k = 0; k <= K; ++k) S[k] = 0.f;
        bool bad = false;
        float pmax = -INFINITY;
        double ssum = 0.0, wsum = 0.0, bx = 0.0, by = 0.0, bz = 0.0, bw = 0.0;
        int cnt = 0;
        bool have_best = false;
        float4 best_box = hbox;
        auto fold = [&](int r) {
          const float4 ob = s_box[r];
          const double wg = s_wgt[r];
#pragma unroll
          for (int k = 0; k <= K; ++k) S[k] += s_lg[k][r];
          bad |= s_bad[r] != 0;
          pmax = fmaxf(pmax, s_pmax[r]);
          ssum += (double)s_score[r];
          wsum += wg;
          bx += wg * (double)ob.x; by += wg * (double)ob.y; bz += wg * (double)ob.z; bw += wg * (double)ob.w;
          ++cnt;
          if (!have_best && s_score[r] == hscore) { have_best = true; best_box = ob; }
        };
        for (int w = i >> 5; w < W; ++w) {
          unsigned bits = s_mask[i][w] & ~s_head[cur][w];  // later non-heads that match me ...
          if (w == (i >> 5)) bits &= ~((2u << (i & 31)) - 1u);
          while (bits) {
            const int t = __ffs(bits) - 1;
            bits &= bits - 1u;
            const int r = (w << 5) + t;
            if (s_owner[r] == i) fold(r);          // ... and were not claimed by an earlier head
          }
        }
        if (cnt > 0) {
          fold(i);
          if (a.score_mode == PE_SCORE_PROBEN) probEn_finish<K>(S, bad, &fs, &fc);
          else if (a.score_mode == PE_SCORE_AVG) fs = (float)(ssum / (double)cnt);
          else fs = pmax;
          if (a.box_mode == PE_BOX_ARGMAX) fb = best_box;
          else {
            const double inv = 1.0 / wsum;
            fb = make_float4((float)(bx * inv), (float)(by * inv), (float)(bz * inv), (float)(bw * inv));
          }
        }
      }
      a.out_boxes[base + pos] = fb;
      a.out_scores[base + pos] = fs;
      a.out_classes[base + pos] = fc;
    }
  }
}

constexpr size_t block_smem_bytes() {
  return (size_t)kMaxBlockDets * (2 * sizeof(float4) + 2 * sizeof(float) + 3 * sizeof(int)) +
         (size_t)kMaxBlockDets * (kMaxBlockDets / 32) * sizeof(unsigned);
}

template <int K>
int launch_fuse(const FuseArgs& a, cudaStream_t st) {
  PE_CUDA_CHECK(cudaMemsetAsync(a.big_count, 0, 2 * sizeof(int), st));
  const int warps_per_block = kBlockThreads / 32;
  const int sms = sm_count();
  // persistent grid: one wave of resident blocks (4 per SM at 64 registers), warps stride over the windows
  long long want = ceil_div<long long>(ceil_div<long long>(a.B, kWin), warps_per_block);
  long long cap = (long long)sms * 4;
  int grid = (int)(want < cap ? (want > 0 ? want : 1) : cap);
  const bool nms = a.score_mode == PE_SCORE_MAX && a.box_mode == PE_BOX_ARGMAX;
  if (nms) fuse_packed_kernel<K, PE_SCORE_MAX, true><<<grid, kBlockThreads, 0, st>>>(a);
  else if (a.score_mode == PE_SCORE_PROBEN) fuse_packed_kernel<K, PE_SCORE_PROBEN, false><<<grid, kBlockThreads, 0, st>>>(a);
  else if (a.score_mode == PE_SCORE_AVG) fuse_packed_kernel<K, PE_SCORE_AVG, false><<<grid, kBlockThreads, 0, st>>>(a);
  else fuse_packed_kernel<K, PE_SCORE_MAX, false><<<grid, kBlockThreads, 0, st>>>(a);
  PE_LAUNCH_CHECK();
  static DeviceOnce attr_once;  // one per K instantiation
  if (attr_once.needed()) {
    PE_CUDA_CHECK(cudaFuncSetAttribute(fuse_block_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)block_smem_bytes()));
    attr_once.mark();
  }
  // 33..256 detections (the detector pipeline's regime): thread-per-detection kernel, several blocks per SM
  const int grid_m = a.B < sms * 4 ? a.B : sms * 4;
  fuse_mid_kernel<K><<<grid_m, kBlockThreads, 0, st>>>(a);
  PE_LAUNCH_CHECK();
  const int grid_b = a.B < sms ? a.B : sms;
  fuse_block_kernel<K><<<grid_b, kBlockThreads, block_smem_bytes(), st>>>(a);
  PE_LAUNCH_CHECK();
  return PE_OK;
}

}  // namespace
}  // namespace pe

extern "C" PE_API size_t pe_fuse_workspace_bytes(int B) { return sizeof(int) * ((size_t)(B > 0 ? B : 0) + 4); }

extern "C" PE_API int pe_fuse_max_dets_per_image(void) { return pe::kMaxBlockDets; }

extern "C" PE_API int pe_fuse_batch(const float* boxes, const float* scores, const int32_t* classes, const float* probs,
                             const float* vars, const int32_t* det_offsets, int B, int M, int K, float iou_thr,
                             int score_mode, int box_mode, float img_w, float img_h, float* out_boxes,
                             float* out_scores, int32_t* out_classes, int32_t* out_counts, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (B < 0 || M < 1 || M > 31) return PE_ERR_INVALID_ARGUMENT;
  if (score_mode < PE_SCORE_PROBEN || score_mode > PE_SCORE_MAX) return PE_ERR_INVALID_ARGUMENT;
  if (box_mode < PE_BOX_VAVG || box_mode > PE_BOX_ARGMAX) return PE_ERR_INVALID_ARGUMENT;
  if (B == 0) return PE_OK;
  if (!det_offsets || !out_counts || !workspace) return PE_ERR_INVALID_ARGUMENT;
  if (!boxes || !scores || !classes || !probs || !vars || !out_boxes || !out_scores || !out_classes)
    return PE_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(boxes) & 15) || (reinterpret_cast<uintptr_t>(out_boxes) & 15))
    return PE_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < pe_fuse_workspace_bytes(B)) return PE_ERR_WORKSPACE_TOO_SMALL;
  pe::FuseArgs a;
  a.boxes = reinterpret_cast<const float4*>(boxes);
  a.scores = scores;
  a.classes = classes;
  a.probs = probs;
  a.vars = vars;
  a.offs = det_offsets;
  a.B = B;
  a.M = M;
  a.thr = iou_thr;
  a.score_mode = score_mode;
  a.box_mode = box_mode;
  a.img_w = img_w;
  a.img_h = img_h;
  a.out_boxes = reinterpret_cast<float4*>(out_boxes);
  a.out_scores = out_scores;
  a.out_classes = out_classes;
  a.out_counts = out_counts;
  static const int bucket_env = [] { const char* e = getenv("PE_FUSE_BUCKET"); return e ? atoi(e) : 1; }();
  a.bucket = bucket_env;
  a.big_count = reinterpret_cast<int*>(workspace);
  a.big_list = a.big_count + 4;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (K == 1) return pe::launch_fuse<1>(a, st);
  if (K == 3) return pe::launch_fuse<3>(a, st);
  return PE_ERR_UNSUPPORTED;
}
