// C ABI (include/vitsdec.h) and the decode schedule: which fused convolution runs on which buffer.
//
// Reference being replaced: Generator.__init__/forward (/root/reference/models.py:245-289) with
// ResBlock1/ResBlock2 (/root/reference/modules.py:187-256).  The schedule below is a restatement of that
// forward in terms of ONE primitive (conv + fused epilogue on channels-last bf16 a-form tensors):
//
//   pack_z           z fp32 NCL -> bf16 [B][T][C0]
//   cond             cb[b] = cond(g[b])                                           models.py:272-273
//   conv_pre         X = lrelu(conv7(z) + bias + cb[b], 0.1)                      models.py:271-276
//   per stage i:     U = lrelu(convT_i(X) + bias, 0.1)   (polyphase, 2-3 taps)    models.py:276-277
//     per branch j:  cur = U
//        per pair m: H   = lrelu(c1(cur) + b, 0.1)                                modules.py:212-216
//                    nxt = lrelu(c2(H) + b + x(cur), 0.1)                         modules.py:217-221
//        the last c2 of a branch feeds the MRF accumulator instead:               models.py:279-284
//                    first branch S = v; middle S += v; last X = lrelu((S+v)/nk, 0.1 | 0.01)
//   conv_post        out = tanh(conv7(X))  (X already holds lrelu(., 0.01))       models.py:285-287
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/vitsdec.h"
#include "common.cuh"
#include "conv_mrf128.h"
#include "conv_mrfp.h"
#include "conv_pair.h"
#include "conv_pairf.h"
#include "conv_tc.h"
#include "pack.h"

namespace vd {

static thread_local std::string g_err;
static std::atomic<unsigned long long*> g_trace_buffer{nullptr};  // vitsdec_debug_set_trace: device buffer of clock64 stamps
void set_error(const std::string& msg) { g_err = msg; }

typedef __nv_bfloat16 bf16;
constexpr float kSlope = 0.1f;       // modules.py:17 LRELU_SLOPE
constexpr float kPostSlope = 0.01f;  // F.leaky_relu default, models.py:285

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static int ceildiv(int a, int b) { return -floordiv(-a, b); }

enum LayerKind { kConv = 0, kConvT = 1, kPost = 2, kCond = 3, kMrf = 4, kPair = 5, kMrfPair = 6 };

struct Layer {
  std::string name;
  LayerKind kind;
  int c_in = 0, c_out = 0, k = 0, dil = 1, stride = 1;
  int cin_src = 0, cout_src = 0;   // the caller's weight dimensions when the packed operand is zero-padded (0 = same)
  ConvGeom geom{};        // B, L filled per decode
  bf16* w = nullptr;      // packed operand [ntaps][n_total][c_in]
  float* bias = nullptr;  // [n_total] (zero when the layer has no bias)
  float* wf32 = nullptr;  // conv_post / cond keep fp32 weights
  bool loaded = false;
  // kMrf (virtual layer): the last convs of one stage's MRF branches, accumulated in one launch
  std::vector<int> members;      // real layer ids, one per branch
  int mrf_group = -1;            // real layers: id of the virtual layer they also feed, and their tap base in it
  int mrf_tap_base = 0;
  int mrfp_group = -1;           // real layers: id of the stage's fused last-pairs layer (kMrfPair, conv_mrfp.cu) and
  int mrfp_tap_base = 0;         // their first tap inside its packed weights
  int pair_group = -1;           // real layers: id of the fused-pair virtual layer (kPair) and their tap base in it
  int pair_tap_base = 0;
  // time-folded form (fold_geom): r time samples per row, r*C virtual channels; 0 = layer has no folded form
  int fold_r = 0;
  ConvGeom fgeom{};
  bf16* wfold = nullptr;         // [folded taps][r*C][r*C]
  float* bias_fold = nullptr;    // [r*C]
  int mrf_ftap_base = 0;         // real layers: first folded tap inside the virtual MRF layer's folded weights
  int pair_ftap_base = 0;        // real layers: first folded tap inside the pair layer's folded weights (0 or nt)
  bool pair_plain = false;       // kPair: conv_pair.cu (resident unfolded weights) can run this pair
};

static void conv_geom(Layer& l) {
  ConvGeom& g = l.geom;
  g.c_in = l.c_in;
  g.n_total = l.c_out;
  g.ntaps = l.k;
  for (int j = 0; j < l.k; ++j) {
    g.tap_off[j] = (j - (l.k - 1) / 2) * l.dil;  // get_padding(k, d) = d(k-1)/2, commons.py:14-15
    g.tap_nlo[j] = 0;
    g.tap_nhi[j] = l.c_out;
    g.tap_kmask[j] = ~0u;
  }
  g.nseg = 1;
  g.seg_tap_end[0] = g.ntaps;
}

// Time folding (narrow layers).  A C-channel tensor [B][L][C] is bit-for-bit a (r*C)-channel tensor [B][L/r][r*C]:
// row n of the folded view holds time samples r*n .. r*n+r-1.  A k-tap dilation-1 convolution over time is then a
// convolution over folded rows with block-Toeplitz weights
//     W'[s][phi*C + co][psi*C + ci] = W[j][co][ci],   j - (k-1)/2 = r*s + psi - phi   (zero when no such tap exists)
// (dilated convs: see fold_geom) which turns the N = 32/64 tensor-core tiles of stages 2-3 (bound by shared-memory operand reads: 4 KB of activations
// per 16/32-cycle MMA) into 128-channel channels-as-M tiles (N = 256 time rows per instruction) at the price of
// (k + r - 1)/k more MACs.  K-chunks of a folded tap that are structurally zero are skipped (tap_kmask).
static int fold_factor(const Layer& l) {
  if (l.kind != kConv || l.c_in != l.c_out) return 0;
  if (l.c_in == 32) return 4;
  if (l.c_in == 64) return 2;
  return 0;
}

// folded taps of one k-tap conv appended to g (tap table only); returns the number of folded taps
static int fold_taps(int c, int k, int r, ConvGeom& g, int tap_base, int kc = 64) {
  const int hk = (k - 1) / 2;
  const int s_min = floordiv(-hk, r), s_max = floordiv(r - 1 + hk, r);
  const int per = kc / c;  // folded K-chunk = `per` consecutive time phases
  for (int s = s_min; s <= s_max; ++s) {
    const int i = tap_base + (s - s_min);
    g.tap_off[i] = s;
    g.tap_nlo[i] = 0;
    g.tap_nhi[i] = r * c;
    uint32_t mask = 0;
    for (int psi = 0; psi < r; ++psi)
      for (int phi = 0; phi < r; ++phi)
        if (std::abs(r * s + psi - phi) <= hk) mask |= 1u << (psi / per);
    g.tap_kmask[i] = mask;
  }
  return s_max - s_min + 1;
}

static void fold_geom(Layer& l) {
  l.fold_r = fold_factor(l);
  if (!l.fold_r) return;
  ConvGeom& g = l.fgeom;
  g = ConvGeom{};
  g.c_in = g.n_total = l.fold_r * l.c_in;
  // A dilation-d conv is a dilation-1 conv on each of the d sub-sequences t = d*q + rho: same Toeplitz weights, the
  // folded rows are taken from one sub-sequence (ConvGeom::rho_d; K-chunk = one time phase, phases are d*C apart).
  g.ntaps = fold_taps(l.c_in, l.k, l.fold_r, g, 0, l.dil > 1 ? l.c_in : 64);
  g.nseg = 1;
  g.seg_tap_end[0] = g.ntaps;
  if (l.dil > 1) {
    g.rho_d = l.dil;
    g.c_real = l.c_in;
  }
}

// ConvTranspose1d(k, s, p=(k-s)/2): output sample s*i + r reads input rows i + off, kernel index j = r + p - s*off
static void convT_geom(Layer& l) {
  ConvGeom& g = l.geom;
  const int s = l.stride, k = l.k, p = (k - s) / 2;
  g.c_in = l.c_in;
  g.n_total = s * l.c_out;
  const int off_min = ceildiv(p - k + 1, s), off_max = floordiv(s - 1 + p, s);
  g.ntaps = off_max - off_min + 1;
  for (int i = 0; i < g.ntaps; ++i) {
    const int off = off_min + i;
    g.tap_off[i] = off;
    const int rlo = std::max(0, s * off - p), rhi = std::min(s, s * off - p + k);
    g.tap_nlo[i] = rlo * l.c_out;
    g.tap_nhi[i] = rhi * l.c_out;
    g.tap_kmask[i] = ~0u;
  }
  g.nseg = 1;
  g.seg_tap_end[0] = g.ntaps;
}

struct PlanKey {
  int B, T, impl, desc_mode;
  const void* ws;
  bool operator<(const PlanKey& o) const {
    return std::tie(B, T, impl, desc_mode, ws) < std::tie(o.B, o.T, o.impl, o.desc_mode, o.ws);
  }
};

struct Step {           // one launch of the conv primitive
  int layer;
  const bf16* xs[kMaxSeg];
  ConvEpilogue ep;
  int L;
  ConvTcPlan tc;
  bool is_pair = false;     // fused ResBlock1 pair (conv_pair.cu): layer = the kPair virtual layer
  PairPlan pair;
  bool is_pairf = false;    // time-folded fused pair (conv_pairf.cu)
  PairFPlan pairf;
  bool is_mrfp = false;     // last pairs of all MRF branches + branch average in one launch (conv_mrfp.cu): layer = kMrfPair
  MrfpPlan mrfp;
  float mrfp_out_slope = 0.f;
  bool is_mrf128 = false;   // the same fusion for a 128-channel stage (conv_mrf128.cu): layer = kMrfPair
  Mrf128Plan mrf128;
  int branch = -1;          // MRF branch this launch belongs to (-1: trunk), for concurrent branches under the graph
  bf16* dbg_dst = nullptr;  // debug_keep: copy ep.out here after the launch
  size_t dbg_bytes = 0;
};

struct Plan {
  ~Plan() {
    for (cudaGraphExec_t e : graph_exec)
      if (e) cudaGraphExecDestroy(e);
  }
  // CUDA graph of the conv steps (they only touch plan-owned workspace pointers, so one capture serves every call
  // with this plan); index 0: no speaker conditioning, 1: conv_pre adds the per-utterance cond bias
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  bool graph_failed = false;
  // Capturing + instantiating the graph costs ~12 ms (a batch-1 decode takes 0.5 ms): a plan is replayed as a graph
  // only from its third use on, so a serving loop whose shape changes every call (tools/shape_churn.py: 12.1 ms per
  // new shape with eager capture, 1.3 ms with plain launches) never pays for graphs it will not reuse.
  std::atomic<int> uses{0};
  // Concurrent MRF branches: small decodes (a 2 s utterance has 12-170 tiles per launch for 148 SMs) are bound by
  // launch latency and under-filled kernels; the three branches of a stage are independent, so under the CUDA graph
  // they run on forked streams: 0.84 -> 0.53 ms at 173 frames, 1.19 -> 0.96 ms at 862, neutral to -1.5 % at 16 x 862
  // (tools/par_sweep.py), so it is always on (option par = 0 serialises).
  bool par = false;
  std::vector<Step> steps;
  bf16* a0;          // packed latent
  float* cb;         // cond bias [B][C]
  bf16* x_final;     // input of conv_post
  int L_final, C_final;
  bool post_tc = false;  // conv_post runs as a time-folded tensor-core launch (its output pointer is per call, so it
  Step post;             // stays outside the graph)
  std::vector<std::pair<std::string, std::tuple<const bf16*, int, int, float>>> debug;  // name -> (ptr, C, L, gain)
};

}  // namespace vd

using namespace vd;

struct vitsdec_decoder {
  vitsdec_hparams hp;
  int device = 0;
  int num_sms = 148;
  std::vector<Layer> layers;
  std::map<std::string, int> by_name;
  int l_pre = -1, l_post = -1, l_cond = -1;
  std::vector<int> l_ups;
  std::vector<int> l_mrf;              // per stage: virtual fused-MRF layer id, or -1 (fp32 accumulator path)
  std::vector<int> l_mrfp;             // per stage: kMrfPair layer id (last pairs + MRF in one launch), or -1
  int num_real_layers = 0;
  std::vector<std::vector<int>> l_rb;  // per resblock: conv layer ids in forward order
  std::vector<int> stage_ch;
  int hop = 1;
  float* scale_scratch = nullptr;
  int impl = 0, desc_mode = 0, debug_keep = 0, profile = 0, fuse_pairs = 1, use_graph = 1, fold = 1, pairf = 1, par = 1, mrfp = 5;
  int c_z = 0;   // initial_channel rounded up to a multiple of 32: the packed latent and conv_pre's K are zero-padded
  int pdl = 1;   // option "pdl": programmatic dependent launch for launches that leave SMs idle (0 off, 2 every launch)
  int fp16 = 0;  // option "fp16": weights and stored activations are IEEE fp16 instead of bf16 (ConvEpilogue::f16)
  cudaStream_t cstream = nullptr;  // capture-only stream (the caller's may be the legacy default stream)
  cudaStream_t bstream[VITSDEC_MAX_KERNELS] = {};  // capture-only streams of MRF branches 1.. (Plan::par)
  cudaEvent_t ev_fork = nullptr, ev_join[VITSDEC_MAX_KERNELS] = {};
  std::map<std::pair<int, int>, int> l_pair;  // (resblock index, pair index) -> kPair virtual layer id
  std::atomic<int> last_launches{0};
  std::atomic<int> graph_failures{0};   // plans whose CUDA-graph capture / instantiation failed (they run as plain launches)
  int max_dil = 1;                      // largest ResBlock dilation: sizes the slack behind the last workspace slot
  // profile=1: CUDA events around the convolution launches of every decode, accumulated on read
  cudaEvent_t ev_conv0 = nullptr, ev_conv1 = nullptr;
  bool ev_pending = false;
  double prof_conv_ms = 0.0;
  long prof_conv_launches = 0;
  std::mutex prof_mu;                   // guards the four profile fields above (option "profile" is a single-stream bench hook)
  std::mutex mu;
  std::list<std::pair<PlanKey, std::shared_ptr<Plan>>> plans;  // small LRU
  std::shared_ptr<Plan> last_plan;
  // decode_host resources
  cudaStream_t hstream = nullptr;
  void* hbuf = nullptr;
  size_t hbuf_bytes = 0;
};

namespace vd {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int add_layer(vitsdec_decoder* d, const std::string& name, LayerKind kind, int c_in, int c_out, int k, int dil,
                     int stride) {
  Layer l;
  l.name = name; l.kind = kind; l.c_in = c_in; l.c_out = c_out; l.k = k; l.dil = dil; l.stride = stride;
  if (kind == kConv) { conv_geom(l); fold_geom(l); }
  if (kind == kConvT) convT_geom(l);
  d->layers.push_back(l);
  d->by_name[name] = (int)d->layers.size() - 1;
  return (int)d->layers.size() - 1;
}

static int alloc_layer(Layer& l) {
  if (l.kind == kMrfPair) {
    VD_CUDA(cudaMalloc(&l.w, (size_t)l.k * l.c_out * l.c_in * sizeof(bf16)));   // l.k = total taps of the packed set
    return 0;
  }
  if (l.kind == kPair) {
    VD_CUDA(cudaMalloc(&l.w, (size_t)2 * l.k * l.c_out * l.c_in * sizeof(bf16)));
    if (l.fold_r) VD_CUDA(cudaMalloc(&l.wfold, (size_t)l.fgeom.ntaps * 128 * 128 * sizeof(bf16)));
    return 0;
  }
  if (l.kind == kConv || l.kind == kConvT || l.kind == kMrf) {
    const size_t wn = (size_t)l.geom.ntaps * l.geom.n_total * l.geom.c_in;
    VD_CUDA(cudaMalloc(&l.w, wn * sizeof(bf16)));
    VD_CUDA(cudaMalloc(&l.bias, (size_t)l.geom.n_total * sizeof(float)));
    VD_CUDA(cudaMemset(l.bias, 0, (size_t)l.geom.n_total * sizeof(float)));
  } else {
    VD_CUDA(cudaMalloc(&l.wf32, (size_t)l.c_out * l.c_in * l.k * sizeof(float)));
    if (l.kind == kCond) VD_CUDA(cudaMalloc(&l.bias, (size_t)l.c_out * sizeof(float)));
  }
  if (l.fold_r) {
    const size_t fn = (size_t)l.fgeom.ntaps * l.fgeom.n_total * l.fgeom.c_in;
    VD_CUDA(cudaMalloc(&l.wfold, fn * sizeof(bf16)));
    VD_CUDA(cudaMalloc(&l.bias_fold, (size_t)l.fgeom.n_total * sizeof(float)));
    VD_CUDA(cudaMemset(l.bias_fold, 0, (size_t)l.fgeom.n_total * sizeof(float)));
  }
  return 0;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct WsLayout {
  size_t slot;      // bytes of one bf16 activation slot
  size_t off_a0, off_cb, off_slots, off_dbg, total;
  int nslots;
};

// Slot map.  0: X (stage in/out)  1: U (upsampled)  2: T1 (c1 output of non-final pairs)  3: T2 (pair ping-pong)
//   fused MRF:  4+j: P_j (input + residual of branch j's last pair)   4+nk+j: H_j (c1 output of that last pair)
//               4+2nk+2(j-1), +1: T1_j, T2_j of branches j >= 1 (branches may run concurrently, see Plan::par)
//   fp32-accumulator MRF (fallback):  4: P   5,6: S (one fp32 tensor)
static bool all_fused(const vitsdec_decoder* d) {
  for (int id : d->l_mrf)
    if (id < 0) return false;
  return true;
}

static WsLayout ws_layout(const vitsdec_decoder* d, int B, int T) {
  WsLayout w{};
  size_t slot = (size_t)B * T * d->hp.upsample_initial_channel * 2;  // conv_pre output
  long L = T;
  size_t dbg = align_up(slot, 1024);
  for (size_t i = 0; i < d->stage_ch.size(); ++i) {
    L *= d->hp.upsample_rates[i];
    const size_t sz = (size_t)B * L * d->stage_ch[i] * 2;
    slot = std::max(slot, sz);
    dbg += 2 * align_up(sz, 1024);
  }
  slot = align_up(slot, 1024);
  size_t o = 0;
  w.slot = slot;
  w.off_a0 = o; o += align_up((size_t)B * T * d->c_z * 2, 1024);
  w.off_cb = o; o += align_up((size_t)B * d->hp.upsample_initial_channel * 4, 1024);
  w.nslots = all_fused(d) ? 4 + 2 * d->hp.num_kernels + 2 * (d->hp.num_kernels - 1)
                          : std::max(7, 4 + 2 * d->hp.num_kernels);
  w.off_slots = o; o += (size_t)w.nslots * slot;
  w.off_dbg = o;
  if (d->debug_keep) o += align_up(dbg, 1024);
  // dilated folded views read (and mask) up to rho_d*r rows (= 256 * dilation bytes) past the last utterance of a slot
  o += align_up((size_t)256 * d->max_dil + 1024, 4096);
  w.total = o;
  return w;
}

static int build_plan(vitsdec_decoder* d, Plan& pl, int B, int T, uint8_t* ws) {
  const WsLayout w = ws_layout(d, B, T);
  auto slot = [&](int i) { return reinterpret_cast<bf16*>(ws + w.off_slots + (size_t)i * w.slot); };
  bf16* X = slot(0);
  bf16* U = slot(1);
  bf16* T1 = slot(2);
  bf16* T2 = slot(3);
  const int nk = d->hp.num_kernels;
  pl.a0 = reinterpret_cast<bf16*>(ws + w.off_a0);
  pl.cb = reinterpret_cast<float*>(ws + w.off_cb);
  uint8_t* dbg = ws + w.off_dbg;

  auto push = [&](int layer, const bf16* const* xs, int L, const ConvEpilogue& ep) -> int {
    Step s{};
    s.layer = layer; s.ep = ep; s.L = L;
    Layer& ly = d->layers[layer];
    ConvGeom g = ly.geom;
    g.B = B; g.L = L;
    const bf16* wts = ly.w;
    if (d->impl == 0 && d->fold && ly.fold_r && (L % ly.fold_r == 0 || ly.fgeom.rho_d > 1) && ep.mrf == nullptr &&
        ep.bias_b == nullptr) {
      // time-folded launch: same bytes viewed as [B][L/r][r*C] (see fold_geom)
      g = ly.fgeom;
      g.B = B;
      if (g.rho_d > 1) {
        g.L_real = L;
        g.L = ceildiv(L, g.rho_d * ly.fold_r);
      } else {
        g.L = L / ly.fold_r;
      }
      wts = ly.wfold;
      s.ep.bias = ly.bias_fold;
    }
    for (int i = 0; i < g.nseg; ++i) s.xs[i] = xs[i];
    s.tc.p.g = g;
    if (d->impl == 0) {
      // conv_pre's per-utterance bias arrives at decode time (generic epilogue): never on paired tiles (desc_mode bit 12)
      const int dm = d->desc_mode | (layer == d->l_pre ? 4096 : 0);
      if (plan_conv_tc(&s.tc, g, s.xs, wts, d->num_sms, dm, ep.mrf == nullptr)) return 1;
      if (bind_residual_tc(s.tc, ep)) return 1;
    }
    pl.steps.push_back(s);
    return 0;
  };
  auto push1 = [&](int layer, const bf16* x, int L, const ConvEpilogue& ep) -> int { return push(layer, &x, L, ep); };
  auto keep = [&](const std::string& name, int C, int Lr, float gain) {
    if (!d->debug_keep) return;
    Step& s = pl.steps.back();
    s.dbg_dst = reinterpret_cast<bf16*>(dbg);
    s.dbg_bytes = (size_t)B * Lr * C * 2;
    pl.debug.push_back({name, {s.dbg_dst, C, Lr, gain}});
    dbg += align_up(s.dbg_bytes, 1024);
  };
  auto ep0 = [&](int layer) {
    ConvEpilogue e{};
    e.bias = d->layers[layer].bias;
    e.res_gain = 1.f / kSlope;
    e.out_slope = kSlope;
    e.mrf_scale = 1.f;
    e.f16 = d->fp16;
    return e;
  };

  // conv_pre (+cond): bias_b filled at decode time when g is given
  {
    ConvEpilogue e = ep0(d->l_pre);
    e.out = X;
    if (push1(d->l_pre, pl.a0, T, e)) return 1;
    keep("conv_pre", d->hp.upsample_initial_channel, T, 1.f / kSlope);
  }
  const int nstage = d->hp.num_upsamples;
  int L = T;
  for (int i = 0; i < nstage; ++i) {
    const int s = d->hp.upsample_rates[i];
    const int ch = d->stage_ch[i];
    {  // ups[i]: rows = input rows, columns = s * ch  ==  [B][L*s][ch]
      ConvEpilogue e = ep0(d->l_ups[i]);
      e.out = U;
      if (push1(d->l_ups[i], X, L, e)) return 1;
    }
    L *= s;
    keep("ups." + std::to_string(i), ch, L, 1.f / kSlope);
    const float next_slope = (i == nstage - 1) ? kPostSlope : kSlope;
    const bool fused = d->l_mrf[i] >= 0;
    // last pairs of all branches + MRF average as one launch (conv_mrfp.cu)
    const bool use_mrfp = fused && d->impl == 0 && (d->mrfp & (ch == 128 ? 4 : 1)) && d->fuse_pairs && d->l_mrfp[i] >= 0;
    float* S = reinterpret_cast<float*>(slot(5));
    const bf16* seg_in[kMaxSeg] = {nullptr, nullptr, nullptr, nullptr};
    const bf16* seg_res[kMaxSeg] = {nullptr, nullptr, nullptr, nullptr};
    for (int j = 0; j < nk; ++j) {
      const std::vector<int>& convs = d->l_rb[i * nk + j];
      const int nconv = (int)convs.size();
      const int npairs = d->hp.resblock == 1 ? nconv / 2 : nconv;
      bf16* Pj = fused ? slot(4 + j) : slot(4);
      // private temporaries per branch (fused schedule): the branches of a stage are independent between `ups` and
      // the MRF launch and may run concurrently
      bf16* T1j = (fused && j > 0) ? slot(4 + 2 * nk + 2 * (j - 1)) : T1;
      bf16* T2j = (fused && j > 0) ? slot(4 + 2 * nk + 2 * (j - 1) + 1) : T2;
      bf16* Hj = fused ? slot(4 + nk + j) : T1;
      const size_t first_step = pl.steps.size();
      const bf16* cur = U;
      for (int m = 0; m < npairs; ++m) {
        const bool last = m == npairs - 1;
        const bf16* conv_in = cur;
        int lid;
        auto pit = d->l_pair.find({i * nk + j, m});
        // pairf: 1 = where it is the faster kernel (pairf_preferred), 2 = wherever it exists (tests)
        const bool pair_f = pit != d->l_pair.end() && d->fold && d->pairf && d->layers[pit->second].fold_r &&
                            L % d->layers[pit->second].fold_r == 0 &&
                            (d->pairf == 2 || (pairf_preferred(d->layers[pit->second].c_out, d->layers[pit->second].k,
                                                               d->layers[pit->second].dil) &&
                                               (d->pairf != 3 || (d->layers[pit->second].c_out == 128 &&
                                                                  d->layers[pit->second].k <= 5))));
        const bool pair_p = pit != d->l_pair.end() && d->layers[pit->second].pair_plain;
        // C = 32: the same pair on the 2-sample folded view (conv_mrfp.cu with one branch): N = 64 MMAs, 512-sample tiles
        // (for k < 9 conv_pair.cu's 256-sample tiles are the faster pair kernel: 124 / 178 us against 170 / 189 us for
        // k = 3 / 7 in the 16 x 10 s decode; mrfp bit 3 runs every C = 32 pair here, the first round-2 rule)
        const bool pair_m = pit != d->l_pair.end() && (d->mrfp & 1) && L % 2 == 0 && d->layers[pit->second].pair_plain &&
                            ((d->layers[pit->second].c_out == 32 && (d->layers[pit->second].k >= 9 || (d->mrfp & 8))) ||
                             (d->mrfp & 2)) &&
                            mrfp_supported(d->layers[pit->second].c_out, 1, &d->layers[pit->second].k,
                                           &d->layers[pit->second].dil);
        if (!last && d->impl == 0 && d->fuse_pairs && pair_m && !(pair_f && d->pairf == 2)) {
          bf16* dst = ((npairs - 2 - m) % 2 == 0) ? Pj : T2j;
          const Layer& pv = d->layers[pit->second];
          Step s{};
          s.layer = pit->second;
          s.L = L;
          s.xs[0] = cur;
          s.ep = ep0(convs[2 * m]);
          s.ep.out = dst;
          s.is_mrfp = true;
          s.mrfp_out_slope = kSlope;
          s.tc.p.g.B = B;
          if (plan_conv_mrfp(&s.mrfp, B, L, pv.c_out, 1, &pv.k, &pv.dil, &cur, pv.w, d->num_sms)) return 1;
          pl.steps.push_back(s);
          cur = dst;
          continue;
        }
        if (!last && d->impl == 0 && d->fuse_pairs && (pair_f || pair_p)) {
          // whole pair in one launch: h stays in shared memory
          bf16* dst = ((npairs - 2 - m) % 2 == 0) ? Pj : T2j;
          Step s{};
          s.layer = pit->second;
          s.L = L;
          s.xs[0] = cur;
          s.ep = ep0(convs[2 * m]);
          s.ep.out = dst;
          const Layer& pv = d->layers[pit->second];
          if (pair_f) {
            s.is_pairf = true;
            if (plan_conv_pairf(&s.pairf, B, L, pv.c_out, pv.k, pv.dil, cur, pv.wfold, d->num_sms)) return 1;
          } else {
            s.is_pair = true;
            if (plan_conv_pair(&s.pair, B, L, pv.c_out, pv.k, pv.dil, cur, pv.w, d->num_sms)) return 1;
          }
          s.tc.p.g.B = B;
          pl.steps.push_back(s);
          cur = dst;
          continue;
        }
        if (last && use_mrfp) {   // the whole last pair runs inside the stage's conv_mrfp launch
          seg_res[j] = cur;
          break;
        }
        if (d->hp.resblock == 1) {
          ConvEpilogue e1 = ep0(convs[2 * m]);
          e1.out = last ? Hj : T1j;
          if (push1(convs[2 * m], cur, L, e1)) return 1;
          conv_in = e1.out;
          lid = convs[2 * m + 1];
        } else {
          lid = convs[m];
        }
        if (last && fused) {  // deferred: runs as segment j of the stage's fused MRF launch
          seg_in[j] = conv_in;
          seg_res[j] = cur;
          break;
        }
        ConvEpilogue e = ep0(lid);
        e.res[0] = cur;
        e.nres = 1;
        // ping-pong so that the input of the last pair lands in P_j
        bf16* dst = ((npairs - 2 - m) % 2 == 0) ? Pj : T2j;
        if (!last) {
          e.out = dst;
        } else {
          e.mrf = S;
          e.mrf_scale = 1.f / nk;
          if (nk == 1) { e.mrf_mode = 3; e.mrf = nullptr; }
          else if (j == 0) e.mrf_mode = 1;
          else if (j < nk - 1) e.mrf_mode = 2;
          else e.mrf_mode = 3;
          e.out = X;
          e.out_slope = next_slope;
        }
        if (push1(lid, conv_in, L, e)) return 1;
        cur = dst;
      }
      if (fused)
        for (size_t si = first_step; si < pl.steps.size(); ++si) pl.steps[si].branch = j;
    }
    if (use_mrfp) {
      const Layer& v = d->layers[d->l_mrfp[i]];
      Step s{};
      s.layer = d->l_mrfp[i];
      s.L = L;
      s.is_mrfp = true;
      s.ep = ep0(v.members[0]);
      s.ep.out = X;
      s.mrfp_out_slope = next_slope;
      s.tc.p.g.B = B;
      int ks[kMpMaxBr] = {0, 0, 0}, dl[kMpMaxBr] = {1, 1, 1};
      for (int j = 0; j < nk; ++j) { ks[j] = d->layers[v.members[j]].k; dl[j] = d->layers[v.members[j]].dil; }
      if (ch == 128) {   // streamed weights, channels-as-M tiles (conv_mrf128.cu)
        s.is_mrfp = false;
        s.is_mrf128 = true;
        if (plan_conv_mrf128(&s.mrf128, B, L, nk, ks, dl, seg_res, v.w, d->num_sms)) return 1;
      } else if (plan_conv_mrfp(&s.mrfp, B, L, ch, nk, ks, dl, seg_res, v.w, d->num_sms)) {
        return 1;
      }
      pl.steps.push_back(s);
    } else if (fused) {
      // models.py:279-284: x = (rb0(x) + rb1(x) + rb2(x)) / nk -- the three last convs accumulate in one TMEM tile
      ConvEpilogue e = ep0(d->l_mrf[i]);
      for (int j = 0; j < nk; ++j) e.res[j] = seg_res[j];
      e.nres = nk;
      e.mrf_mode = 3;
      e.mrf = nullptr;
      e.mrf_scale = 1.f / nk;
      e.out = X;
      e.out_slope = next_slope;
      if (push(d->l_mrf[i], seg_in, L, e)) return 1;
    }
    keep("mrf." + std::to_string(i), ch, L, 1.f / next_slope);
  }
  pl.x_final = X;
  pl.L_final = L;
  pl.C_final = d->stage_ch.back();
  pl.par = all_fused(d) && nk > 1 && d->par != 0;
  // Programmatic dependent launch: a short launch (at most two tile rounds) lets the next launch of its stream run its
  // prologue early (barrier init, TMEM allocation, descriptor prefetch, resident-weight loads: a few of the ~15 us a small
  // launch takes).  Long launches gain nothing (one CTA per SM, no room for the dependent's CTAs) and measured slower.
  for (Step& s : pl.steps) {
    const int tiles = s.is_mrfp ? s.mrfp.p.total_tiles : (s.is_pair ? s.pair.p.total_tiles : s.tc.p.total_tiles);
    const bool on = d->impl == 0 && !s.is_pairf && !s.is_mrf128 && (d->pdl == 2 || (d->pdl == 1 && tiles <= 2 * d->num_sms));   // a short launch
    s.tc.pdl = on && !s.is_pair && !s.is_mrfp;
    s.pair.pdl = on && s.is_pair;
    s.mrfp.pdl = on && s.is_mrfp;
  }
  pl.post_tc = false;
  Layer& lp = d->layers[d->l_post];
  if (d->impl == 0 && d->fold && lp.fold_r && L % lp.fold_r == 0) {
    Step s{};
    s.layer = d->l_post;
    s.L = L;
    s.xs[0] = X;
    ConvGeom g = lp.fgeom;
    g.B = B; g.L = L / lp.fold_r;
    s.ep.bias = lp.bias_fold;  // zeros: conv_post has no bias (models.py:264)
    s.ep.mrf_mode = 4;
    s.ep.post_c = lp.c_in;
    s.ep.res_gain = s.ep.out_slope = s.ep.mrf_scale = 1.f;
    s.ep.f16 = d->fp16;
    s.tc.p.g = g;
    if (plan_conv_tc(&s.tc, g, s.xs, lp.wfold, d->num_sms, d->desc_mode, true)) return 1;
    VD_CHECK(s.tc.swap, "conv_post: folded launch must be channels-as-M");
    pl.post = s;
    pl.post_tc = true;
  }
  return 0;
}

static int run_conv(vitsdec_decoder* d, Step& s, cudaStream_t st) {
  Layer& ly = d->layers[s.layer];
  if (s.is_mrf128) {
    const int nk = (int)ly.members.size() / 2;   // kMrfPair: c1 of every branch, then c2 of every branch
    const float* b1[kM8MaxBr] = {nullptr, nullptr, nullptr};
    for (int j = 0; j < nk; ++j) b1[j] = d->layers[ly.members[j]].bias;
    const float* b2 = d->layers[d->layers[ly.members[nk]].mrf_group].bias;   // sum of the c2 biases (fused-MRF virtual layer)
    return launch_conv_mrf128(s.mrf128, b1, b2, kSlope, s.mrfp_out_slope, s.ep.out, st, d->fp16);
  }
  if (s.is_mrfp) {
    const int nk = (int)ly.members.size() / 2;   // kPair: {c1, c2}; kMrfPair: c1 of every branch, then c2 of every branch
    const float* b1[kMpMaxBr] = {nullptr, nullptr, nullptr};
    for (int j = 0; j < nk; ++j) b1[j] = d->layers[ly.members[j]].bias;
    // MRF: sum of the c2 biases, kept by the fused-MRF virtual layer of the same stage (rebuilt at every load)
    const float* b2 = ly.kind == kMrfPair ? d->layers[d->layers[ly.members[nk]].mrf_group].bias
                                          : d->layers[ly.members[1]].bias;
    return launch_conv_mrfp(s.mrfp, b1, b2, kSlope, s.mrfp_out_slope, s.ep.out, st, d->fp16);
  }
  if (s.is_pairf)
    return launch_conv_pairf(s.pairf, d->layers[ly.members[0]].bias, d->layers[ly.members[1]].bias, kSlope, s.ep.out,
                             st, d->fp16);
  if (s.is_pair)
    return launch_conv_pair(s.pair, d->layers[ly.members[0]].bias, d->layers[ly.members[1]].bias, kSlope, s.ep.out, st,
                            d->fp16);
  if (d->impl == 0) return launch_conv_tc(s.tc, s.ep, st);
#ifdef VITSDEC_TESTING
  return launch_conv_simt(s.tc.p.g, s.ep, s.xs, ly.w, st);   // CUDA-core cross-check: test build only (build.py)
#else
  set_error("impl=1 (CUDA-core cross-check kernels) exists only in the test build libvitsdec_test.so");
  return 1;
#endif
}

}  // namespace vd

// =========================================================================================== C ABI
extern "C" {

int vitsdec_abi_version(void) { return VITSDEC_ABI_VERSION; }
const char* vitsdec_last_error(void) { return g_err.c_str(); }

int vitsdec_create(const vitsdec_hparams* hp, int device, vitsdec_decoder** out) {
  VD_CHECK(hp && out, "vitsdec_create: null argument");
  VD_CHECK(hp->resblock == 1 || hp->resblock == 2, "resblock must be 1 or 2");
  VD_CHECK(hp->num_upsamples >= 1 && hp->num_upsamples <= VITSDEC_MAX_UPSAMPLES, "bad num_upsamples");
  VD_CHECK(hp->num_kernels >= 1 && hp->num_kernels <= VITSDEC_MAX_KERNELS, "bad num_kernels");
  VD_CHECK(hp->initial_channel > 0 && hp->initial_channel <= 4096, "initial_channel out of range");
  VD_CHECK(hp->upsample_initial_channel % 32 == 0, "upsample_initial_channel must be a multiple of 32");
  int ndev = 0;
  VD_CUDA(cudaGetDeviceCount(&ndev));
  VD_CHECK(device >= 0 && device < ndev, "vitsdec_create: no such CUDA device (there is no CPU fallback)");
  cudaDeviceProp prop;
  VD_CUDA(cudaGetDeviceProperties(&prop, device));
  VD_CHECK(prop.major == 10, "vitsdec needs an sm_100 (B200) device: kernels are tcgen05/TMA only");
  DeviceGuard guard(device);
  VD_CHECK(guard.ok, "cudaSetDevice failed");

  // every error return below frees what was allocated so far (vitsdec_destroy tolerates a half-built decoder)
  std::unique_ptr<vitsdec_decoder, void (*)(vitsdec_decoder*)> d(new vitsdec_decoder(), vitsdec_destroy);
  d->hp = *hp;
  d->device = device;
  d->num_sms = prop.multiProcessorCount;
  const int c0 = hp->upsample_initial_channel;
  // any latent width (HiFi-GAN mel inputs have 80 channels): K is padded with zero channels to the 32-channel K-chunk
  d->c_z = (hp->initial_channel + 31) / 32 * 32;
  d->l_pre = add_layer(d.get(), "conv_pre", kConv, d->c_z, c0, 7, 1, 1);
  d->layers[d->l_pre].cin_src = hp->initial_channel;
  // Stage widths are the reference's c0 // 2^(i+1) (models.py:254,260); tensors narrower than 32 channels (HiFi-GAN V2:
  // 16 and 8) are carried with zero channels up to 32 -- zero weights and zero biases keep them zero through every
  // leaky-relu, so the padded decoder computes exactly the narrow one.
  auto pad32 = [](int c) { return (c + 31) / 32 * 32; };
  std::vector<int> real_ch;
  for (int i = 0; i < hp->num_upsamples; ++i) {
    const int ci_real = c0 >> i, co_real = c0 >> (i + 1);
    const int ci = pad32(ci_real), co = pad32(co_real);
    const int k = hp->upsample_kernel_sizes[i], s = hp->upsample_rates[i];
    VD_CHECK(co_real > 0, "upsample_initial_channel is too small for this many stages");
    VD_CHECK(k >= s && (k - s) % 2 == 0, "upsample kernel must satisfy k >= stride and (k - stride) even");
    d->stage_ch.push_back(co);
    real_ch.push_back(co_real);
    d->hop *= s;
    d->l_ups.push_back(add_layer(d.get(), "ups." + std::to_string(i), kConvT, ci, co, k, 1, s));
    d->layers.back().cin_src = ci_real;
    d->layers.back().cout_src = co_real;
    VD_CHECK(d->layers.back().geom.ntaps <= kMaxTaps, "upsample kernel too large");
  }
  int n = 0;
  for (int i = 0; i < hp->num_upsamples; ++i) {
    const int ch = d->stage_ch[i];
    for (int j = 0; j < hp->num_kernels; ++j, ++n) {
      const int k = hp->resblock_kernel_sizes[j];
      const int nd = hp->num_dilations[j];
      VD_CHECK(k % 2 == 1 && k <= kMaxTaps, "resblock kernel sizes must be odd and <= 31");
      VD_CHECK(nd >= 1 && nd <= VITSDEC_MAX_DILATIONS, "bad dilation count");
      for (int m = 0; m < nd; ++m) {
        VD_CHECK(hp->resblock_dilation_sizes[j][m] >= 1 && hp->resblock_dilation_sizes[j][m] <= 64,
                 "resblock dilations must be in 1..64");
        d->max_dil = std::max(d->max_dil, hp->resblock_dilation_sizes[j][m]);
      }
      std::vector<int> ids;
      const std::string base = "resblocks." + std::to_string(n) + ".";
      if (hp->resblock == 1) {
        VD_CHECK(nd >= 3, "ResBlock1 needs 3 dilations (modules.py:188-196 indexes dilation[0..2])");   // extras are ignored, like the reference
        std::vector<int> c1, c2;
        for (int m = 0; m < 3; ++m)
          c1.push_back(add_layer(d.get(), base + "convs1." + std::to_string(m), kConv, ch, ch, k,
                                 hp->resblock_dilation_sizes[j][m], 1));
        for (int m = 0; m < 3; ++m)
          c2.push_back(add_layer(d.get(), base + "convs2." + std::to_string(m), kConv, ch, ch, k, 1, 1));
        for (int m = 0; m < 3; ++m) { ids.push_back(c1[m]); ids.push_back(c2[m]); }
      } else {
        for (int m = 0; m < nd; ++m)
          ids.push_back(add_layer(d.get(), base + "convs." + std::to_string(m), kConv, ch, ch, k,
                                  hp->resblock_dilation_sizes[j][m], 1));
      }
      for (int id : ids) d->layers[id].cin_src = d->layers[id].cout_src = real_ch[i];
      d->l_rb.push_back(ids);
    }
  }
  d->l_post = add_layer(d.get(), "conv_post", kPost, d->stage_ch.back(), 1, 7, 1, 1);
  d->layers[d->l_post].cin_src = real_ch.back();
  {
    // conv_post on the tensor cores: time-folded like the narrow ResBlock convs, the single output channel padded to C
    // (zero weight rows), tanh + fp32 store in the epilogue (conv_tc.cu EPI 4)
    Layer& lp = d->layers[d->l_post];
    const int r = lp.c_in == 32 ? 4 : (lp.c_in == 64 ? 2 : 0);
    if (r) {
      lp.fold_r = r;
      ConvGeom& fg = lp.fgeom;
      fg = ConvGeom{};
      fg.c_in = fg.n_total = r * lp.c_in;
      // Two-term weights: conv_post sums 7*C products that largely cancel (the waveform is small next to the last
      // stage's activations), so the bf16 rounding of its weights alone cost 0.5-1.7 dB of waveform SNR.  Only channel
      // 0 of each phase is a real output, so channel 1 carries bf16(w - bf16(w)) and the epilogue adds the two rows:
      // no extra MMAs (doubling the taps instead made the launch 157 us instead of 89).
      fg.ntaps = fold_taps(lp.c_in, lp.k, r, fg, 0);
      fg.nseg = 1;
      fg.seg_tap_end[0] = fg.ntaps;
    }
  }
  if (hp->gin_channels > 0) d->l_cond = add_layer(d.get(), "cond", kCond, hp->gin_channels, c0, 1, 1, 1);
  d->num_real_layers = (int)d->layers.size();
  // Fused MRF: one virtual layer per stage whose segments are the LAST conv of every branch (weights concatenated
  // along taps, biases summed); possible when the residual count fits the epilogue (<= 3) and the taps fit the table.
  {
    int ksum = 0;
    for (int j = 0; j < hp->num_kernels; ++j) ksum += hp->resblock_kernel_sizes[j];
    const bool fusable = hp->num_kernels <= kMaxSeg - 1 && ksum <= kMaxTaps;
    for (int i = 0; i < hp->num_upsamples; ++i) {
      if (!fusable) { d->l_mrf.push_back(-1); continue; }
      Layer v;
      v.name = "mrf." + std::to_string(i);
      v.kind = kMrf;
      v.c_in = v.c_out = d->stage_ch[i];
      ConvGeom& g = v.geom;
      g.c_in = g.n_total = v.c_in;
      g.ntaps = 0;
      g.nseg = hp->num_kernels;
      const int vid = (int)d->layers.size();
      for (int j = 0; j < hp->num_kernels; ++j) {
        const int lid = d->l_rb[i * hp->num_kernels + j].back();
        Layer& m = d->layers[lid];
        m.mrf_group = vid;
        m.mrf_tap_base = g.ntaps;
        for (int t = 0; t < m.geom.ntaps; ++t, ++g.ntaps) {
          g.tap_off[g.ntaps] = m.geom.tap_off[t];
          g.tap_nlo[g.ntaps] = 0;
          g.tap_nhi[g.ntaps] = v.c_out;
          g.tap_kmask[g.ntaps] = ~0u;
        }
        g.seg_tap_end[j] = g.ntaps;
        v.members.push_back(lid);
      }
      {  // folded form of the same launch, when every member folds (they share C, dilation 1)
        int r = fold_factor(d->layers[v.members[0]]);
        int nft = 0;
        for (int lid : v.members) {
          const Layer& m = d->layers[lid];
          if (fold_factor(m) != r || m.dil != 1) r = 0;  // one launch = one view: dilation-1 members only
          if (r) nft += floordiv(r - 1 + (m.k - 1) / 2, r) - floordiv(-((m.k - 1) / 2), r) + 1;
        }
        if (r && nft <= kMaxTaps) {
          v.fold_r = r;
          ConvGeom& fg = v.fgeom;
          fg = ConvGeom{};
          fg.c_in = fg.n_total = r * v.c_in;
          fg.nseg = hp->num_kernels;
          for (int j = 0; j < hp->num_kernels; ++j) {
            Layer& m = d->layers[v.members[j]];
            m.mrf_ftap_base = fg.ntaps;
            fg.ntaps += fold_taps(m.c_in, m.k, r, fg, fg.ntaps);
            fg.seg_tap_end[j] = fg.ntaps;
          }
        }
      }
      d->layers.push_back(v);
      d->l_mrf.push_back(vid);
    }
  }
  // Fused pairs: every non-final (c1, c2) pair of a ResBlock1 whose two weight sets, activation stages and the
  // intermediate tile fit in shared memory (C <= 64) runs as one launch (conv_pair.cu).
  if (hp->resblock == 1) {
    for (size_t rb = 0; rb < d->l_rb.size(); ++rb) {
      const std::vector<int> convs = d->l_rb[rb];
      const int npairs = (int)convs.size() / 2;
      for (int m = 0; m + 1 < npairs; ++m) {
        const Layer c1 = d->layers[convs[2 * m]];
        const bool plain = pair_supported(c1.c_out, c1.k, c1.dil), folded = pairf_supported(c1.c_out, c1.k, c1.dil);
        if (!plain && !folded) continue;
        Layer v;
        v.name = "pair." + c1.name;
        v.kind = kPair;
        v.c_in = v.c_out = c1.c_out;
        v.k = c1.k;
        v.dil = c1.dil;
        v.pair_plain = plain;
        v.members = {convs[2 * m], convs[2 * m + 1]};
        if (folded) {  // conv_pairf.cu: block-Toeplitz taps of c1, then of c2
          v.fold_r = 128 / c1.c_out;
          v.fgeom = ConvGeom{};
          v.fgeom.c_in = v.fgeom.n_total = 128;
          v.fgeom.ntaps = 2 * pairf_taps(c1.c_out, c1.k);
          d->layers[convs[2 * m + 1]].pair_ftap_base = v.fgeom.ntaps / 2;
        }
        const int vid = (int)d->layers.size();
        d->layers[convs[2 * m]].pair_group = vid;
        d->layers[convs[2 * m]].pair_tap_base = 0;
        d->layers[convs[2 * m + 1]].pair_group = vid;
        d->layers[convs[2 * m + 1]].pair_tap_base = c1.k;
        d->layers.push_back(v);
        d->l_pair[{(int)rb, m}] = vid;
      }
    }
  }
  // Fused last pairs: the LAST (c1, c2) pair of every MRF branch of a stage, the branch sum and the average in ONE launch
  // (conv_mrfp.cu) where both weight sets of all branches fit in shared memory next to the tiles (C = 32).
  d->l_mrfp.assign(hp->num_upsamples, -1);
  if (hp->resblock == 1) {
    for (int i = 0; i < hp->num_upsamples; ++i) {
      if (d->l_mrf[i] < 0 || hp->num_kernels > kMpMaxBr) continue;
      int ks[kMpMaxBr] = {0, 0, 0}, dl[kMpMaxBr] = {1, 1, 1};
      for (int j = 0; j < hp->num_kernels; ++j) {
        const std::vector<int>& convs = d->l_rb[i * hp->num_kernels + j];
        ks[j] = d->layers[convs[convs.size() - 2]].k;
        dl[j] = d->layers[convs[convs.size() - 2]].dil;
      }
      if (!mrfp_supported(d->stage_ch[i], hp->num_kernels, ks, dl) &&
          !mrf128_supported(d->stage_ch[i], hp->num_kernels, ks, dl))
        continue;
      Layer v;
      v.name = "mrfp." + std::to_string(i);
      v.kind = kMrfPair;
      v.c_in = v.c_out = d->stage_ch[i];
      const int vid = (int)d->layers.size();
      int tap = 0;
      for (int pass = 0; pass < 2; ++pass)       // packed order: c1 of every branch, then c2 of every branch
        for (int j = 0; j < hp->num_kernels; ++j) {
          const std::vector<int>& convs = d->l_rb[i * hp->num_kernels + j];
          const int lid = convs[convs.size() - 2 + pass];
          d->layers[lid].mrfp_group = vid;
          d->layers[lid].mrfp_tap_base = tap;
          tap += d->layers[lid].k;
          v.members.push_back(lid);
        }
      v.k = tap;
      d->layers.push_back(v);
      d->l_mrfp[i] = vid;
    }
  }
  for (Layer& l : d->layers)
    if (alloc_layer(l)) return 1;
  VD_CUDA(cudaMalloc(&d->scale_scratch, 4096 * sizeof(float)));
  *out = d.release();
  return 0;
}

void vitsdec_destroy(vitsdec_decoder* d) {
  if (!d) return;
  DeviceGuard guard(d->device);
  cudaDeviceSynchronize();
  for (Layer& l : d->layers) {
    cudaFree(l.w); cudaFree(l.bias); cudaFree(l.wf32); cudaFree(l.wfold); cudaFree(l.bias_fold);
  }
  cudaFree(d->scale_scratch);
  if (d->hbuf) cudaFree(d->hbuf);
  if (d->hstream) cudaStreamDestroy(d->hstream);
  d->plans.clear();
  d->last_plan.reset();
  if (d->cstream) cudaStreamDestroy(d->cstream);
  for (int j = 0; j < VITSDEC_MAX_KERNELS; ++j) {
    if (d->bstream[j]) cudaStreamDestroy(d->bstream[j]);
    if (d->ev_join[j]) cudaEventDestroy(d->ev_join[j]);
  }
  if (d->ev_fork) cudaEventDestroy(d->ev_fork);
  if (d->ev_conv0) cudaEventDestroy(d->ev_conv0);
  if (d->ev_conv1) cudaEventDestroy(d->ev_conv1);
  delete d;
}

int vitsdec_num_layers(const vitsdec_decoder* d) { return d ? d->num_real_layers : 0; }
const char* vitsdec_layer_name(const vitsdec_decoder* d, int i) {
  if (!d || i < 0 || i >= d->num_real_layers) return nullptr;
  return d->layers[i].name.c_str();
}

int vitsdec_load_layer(vitsdec_decoder* d, const char* name, const float* w, const float* wg, const float* bias,
                       void* stream) {
  VD_CHECK(d && name && w, "vitsdec_load_layer: null argument");
  auto it = d->by_name.find(name);
  VD_CHECK(it != d->by_name.end(), std::string("vitsdec_load_layer: unknown layer ") + name);
  DeviceGuard guard(d->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::lock_guard<std::mutex> lock(d->mu);
  Layer& l = d->layers[it->second];
  // the caller's tensors have the reference's (source) dimensions; packed operands are zero-padded to the 32-channel
  // granularity of the tiles (conv_pre's latent width, stages narrower than 32 channels)
  const int ci_s = l.cin_src ? l.cin_src : l.c_in, co_s = l.cout_src ? l.cout_src : l.c_out;
  if (l.kind == kConv) {
    VD_CHECK(l.c_out <= 4096, "too many channels");
    if (launch_wn_scale(w, wg, d->scale_scratch, co_s, ci_s * l.k, st)) return 1;
    if (launch_pack_conv(w, d->scale_scratch, l.w, l.c_out, l.c_in, l.k, st, 0, d->fp16, ci_s, co_s)) return 1;
    if (launch_replicate_bias(bias, l.bias, l.c_out, 1, st, co_s)) return 1;
    if (l.fold_r) {
      if (launch_pack_conv_fold(w, d->scale_scratch, l.wfold, l.c_in, co_s, l.k, l.fold_r, st, 0, d->fp16, ci_s)) return 1;
      if (launch_replicate_bias(bias, l.bias_fold, l.c_out, l.fold_r, st, co_s)) return 1;
    }
    if (l.mrf_group >= 0) {
      Layer& v = d->layers[l.mrf_group];
      if (launch_pack_conv(w, d->scale_scratch, v.w + (size_t)l.mrf_tap_base * l.c_out * l.c_in, l.c_out, l.c_in, l.k,
                           st, 0, d->fp16, ci_s, co_s))
        return 1;
      if (v.fold_r &&
          launch_pack_conv_fold(w, d->scale_scratch,
                                v.wfold + (size_t)l.mrf_ftap_base * v.fgeom.n_total * v.fgeom.c_in, l.c_in, co_s,
                                l.k, v.fold_r, st, 0, d->fp16, ci_s))
        return 1;
      // combined bias of the fused launch = sum of the member biases (unloaded members still hold zeros), rebuilt HERE
      // on the load stream -- which the caller synchronises -- so that no load-time work is ever deferred into a decode
      // (a concurrent first decode on another stream could have read it before the deferred kernel ran)
      const float* bs[kMaxSeg] = {nullptr, nullptr, nullptr, nullptr};
      for (size_t j = 0; j < v.members.size(); ++j) bs[j] = d->layers[v.members[j]].bias;
      if (launch_sum_bias(bs[0], bs[1], bs[2], bs[3], v.bias, v.c_out, st)) return 1;
      if (v.fold_r && launch_replicate_bias(v.bias, v.bias_fold, v.c_out, v.fold_r, st)) return 1;
    }
    if (l.mrfp_group >= 0) {
      Layer& v = d->layers[l.mrfp_group];
      if (launch_pack_conv(w, d->scale_scratch, v.w + (size_t)l.mrfp_tap_base * l.c_out * l.c_in, l.c_out, l.c_in, l.k,
                           st, 0, d->fp16, ci_s, co_s))
        return 1;
    }
    if (l.pair_group >= 0) {
      Layer& v = d->layers[l.pair_group];
      if (launch_pack_conv(w, d->scale_scratch, v.w + (size_t)l.pair_tap_base * l.c_out * l.c_in, l.c_out, l.c_in, l.k,
                           st, 0, d->fp16, ci_s, co_s))
        return 1;
      if (v.fold_r && launch_pack_conv_fold(w, d->scale_scratch, v.wfold + (size_t)l.pair_ftap_base * 128 * 128, l.c_in,
                                            co_s, l.k, v.fold_r, st, 0, d->fp16, ci_s))
        return 1;
    }
  } else if (l.kind == kConvT) {
    VD_CHECK(l.c_in <= 4096, "too many channels");
    if (launch_wn_scale(w, wg, d->scale_scratch, ci_s, co_s * l.k, st)) return 1;  // dim 0 of [C_in,C_out,k]
    if (launch_pack_convT(w, d->scale_scratch, l.w, l.c_in, l.c_out, l.k, l.stride, (l.k - l.stride) / 2,
                          l.geom.ntaps, l.geom.tap_off[0], st, d->fp16, ci_s, co_s))
      return 1;
    if (launch_replicate_bias(bias, l.bias, l.c_out, l.stride, st, co_s)) return 1;
  } else {
    VD_CHECK(wg == nullptr, "conv_post / cond are not weight-normed in the reference (models.py:264,268)");
    // [c_out][ci_s][k] -> the first ci_s rows of the (zero-initialised) padded copy; c_out = 1 for conv_post and cond
    // has no padding, so one contiguous copy serves both
    VD_CUDA(cudaMemsetAsync(l.wf32, 0, (size_t)l.c_out * l.c_in * l.k * sizeof(float), st));
    VD_CHECK(l.kind == kCond || l.c_out == 1, "conv_post has one output channel");
    VD_CUDA(cudaMemcpyAsync(l.wf32, w, (size_t)l.c_out * ci_s * l.k * sizeof(float), cudaMemcpyDeviceToDevice, st));
    // two-term weights: virtual channel 0 = bf16(w), channel 1 = bf16(w - bf16(w))
    if (l.kind == kPost && l.fold_r && launch_pack_conv_fold(w, nullptr, l.wfold, l.c_in, 1, l.k, l.fold_r, st, 1,
                                                                     d->fp16, ci_s))
      return 1;
    if (l.kind == kCond) {
      VD_CHECK(bias != nullptr, "cond needs a bias");
      VD_CUDA(cudaMemcpyAsync(l.bias, bias, (size_t)l.c_out * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
  }
  l.loaded = true;
  return 0;
}

size_t vitsdec_workspace_bytes(const vitsdec_decoder* d, int batch, int frames) {
  if (!d || batch <= 0 || frames <= 0) return 0;
  return ws_layout(d, batch, frames).total;
}

int vitsdec_decode(vitsdec_decoder* d, const float* z, int64_t zsb, int64_t zsc, const float* g, float* out, int B,
                   int T, void* ws, size_t ws_bytes, void* stream) {
  VD_CHECK(d && z && out && ws, "vitsdec_decode: null argument");
  VD_CHECK(B > 0 && T > 0, "vitsdec_decode: empty batch or zero frames");
  VD_CHECK(B <= 65535, "vitsdec_decode: batch too large");
  VD_CHECK((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "workspace must be 256-byte aligned");
  VD_CHECK(g == nullptr || d->l_cond >= 0, "g given but the decoder was built with gin_channels=0 (models.py:267)");
  for (int i = 0; i < d->num_real_layers; ++i)
    VD_CHECK(d->layers[i].loaded, "vitsdec_decode: layer " + d->layers[i].name + " has no weights loaded");
  DeviceGuard guard(d->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::shared_ptr<Plan> plan;
  {
    std::lock_guard<std::mutex> lock(d->mu);
    VD_CHECK(ws_bytes >= ws_layout(d, B, T).total, "vitsdec_decode: workspace too small");
    const PlanKey key{B, T, d->impl, ((((((((d->desc_mode * 16 + d->mrfp) * 4 + d->pdl) * 2 + d->par) * 8 + d->pairf) * 2 + d->fold) * 2 + d->debug_keep) * 2) + d->fuse_pairs), ws};
    for (auto it = d->plans.begin(); it != d->plans.end(); ++it) {
      if (!(it->first < key) && !(key < it->first)) {
        plan = it->second;
        d->plans.splice(d->plans.begin(), d->plans, it);
        break;
      }
    }
    if (!plan) {
      plan = std::make_shared<Plan>();
      if (build_plan(d, *plan, B, T, static_cast<uint8_t*>(ws))) return 1;
      d->plans.emplace_front(key, plan);
      if (d->plans.size() > 64) d->plans.pop_back();
    }
    d->last_plan = plan;
  }

  // option "profile": one pair of events per decoder, so profiled decodes are serialised (a bench hook, not a serving mode)
  std::unique_lock<std::mutex> prof_lock(d->prof_mu, std::defer_lock);
  if (d->profile) prof_lock.lock();
  int launches = 0;
  if (launch_pack_z(z, zsb, zsc, plan->a0, B, d->hp.initial_channel, T, st, d->fp16, d->c_z)) return 1;
  ++launches;
  if (g) {
    const Layer& lc = d->layers[d->l_cond];
    if (launch_cond(lc.wf32, lc.bias, g, plan->cb, B, lc.c_out, lc.c_in, st)) return 1;
    ++launches;
  }
  if (d->profile) {
    if (!d->ev_conv0) {
      VD_CUDA(cudaEventCreate(&d->ev_conv0));
      VD_CUDA(cudaEventCreate(&d->ev_conv1));
    }
    if (d->ev_pending) {  // fold the previous decode's interval in before re-recording
      float ms = 0.f;
      VD_CUDA(cudaEventSynchronize(d->ev_conv1));
      VD_CUDA(cudaEventElapsedTime(&ms, d->ev_conv0, d->ev_conv1));
      d->prof_conv_ms += ms;
      d->ev_pending = false;
    }
    VD_CUDA(cudaEventRecord(d->ev_conv0, st));
  }
  auto enqueue_steps = [&](cudaStream_t qs, bool fork) -> int {
    int prev = -1;
    unsigned open_branches = 0;   // branches >= 1 that have launches since the last join
    for (size_t i = 0; i < plan->steps.size(); ++i) {
      Step s = plan->steps[i];  // copy: per-call epilogue fields, re-entrant across threads
      if (i == 0) s.ep.bias_b = g ? plan->cb : nullptr;
      cudaStream_t ls = qs;
      if (fork) {
        const int b = s.branch;
        if (b == 0 && prev < 0) VD_CUDA(cudaEventRecord(d->ev_fork, qs));   // the point right after `ups`
        if (b > 0) {
          if (prev != b) VD_CUDA(cudaStreamWaitEvent(d->bstream[b], d->ev_fork, 0));
          ls = d->bstream[b];
          open_branches |= 1u << b;
        }
        if (b < 0 && open_branches) {   // back on the trunk (the MRF launch): join
          for (int j = 1; j < VITSDEC_MAX_KERNELS; ++j)
            if ((open_branches >> j) & 1u) {
              VD_CUDA(cudaEventRecord(d->ev_join[j], d->bstream[j]));
              VD_CUDA(cudaStreamWaitEvent(qs, d->ev_join[j], 0));
            }
          open_branches = 0;
        }
        prev = b;
      }
      if (run_conv(d, s, ls)) return 1;
      if (s.dbg_dst) VD_CUDA(cudaMemcpyAsync(s.dbg_dst, s.ep.out, s.dbg_bytes, cudaMemcpyDeviceToDevice, ls));
    }
    return 0;
  };
  bool launched = false;
  if (d->use_graph && !d->debug_keep && !plan->graph_failed && (++plan->uses >= 3 || d->use_graph == 2)) {
    // one graph launch instead of ~60 kernel launches: what makes a 2 s / batch-1 decode launch-bound otherwise
    std::lock_guard<std::mutex> lock(d->mu);
    cudaGraphExec_t& exec = plan->graph_exec[g ? 1 : 0];
    if (!exec) {
      if (!d->cstream) VD_CUDA(cudaStreamCreateWithFlags(&d->cstream, cudaStreamNonBlocking));
      if (plan->par && !d->ev_fork) {
        VD_CUDA(cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming));
        for (int j = 1; j < d->hp.num_kernels; ++j) {
          VD_CUDA(cudaStreamCreateWithFlags(&d->bstream[j], cudaStreamNonBlocking));
          VD_CUDA(cudaEventCreateWithFlags(&d->ev_join[j], cudaEventDisableTiming));
        }
      }
      cudaGraph_t graph = nullptr;
      bool ok = cudaStreamBeginCapture(d->cstream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) {
        const int rc = enqueue_steps(d->cstream, plan->par);
        ok = cudaStreamEndCapture(d->cstream, &graph) == cudaSuccess && rc == 0 && graph != nullptr;
      }
      if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
      if (graph) cudaGraphDestroy(graph);
      if (!ok) {
        cudaGetLastError();
        exec = nullptr;
        plan->graph_failed = true;  // this plan runs as plain launches from now on; visible as option "graph_failed"
        d->graph_failures.fetch_add(1);
      }
    }
    if (exec) {
      VD_CUDA(cudaGraphLaunch(exec, st));
      launched = true;
    }
  }
  if (!launched && enqueue_steps(st, false)) return 1;
  launches += (int)plan->steps.size();
  if (plan->post_tc) {
    Step s = plan->post;
    s.ep.out_f32 = out;
    if (launch_conv_tc(s.tc, s.ep, st)) return 1;
  } else if (launch_conv_post(plan->x_final, d->layers[d->l_post].wf32, out, B, plan->L_final, plan->C_final, st,
                              d->fp16)) {
    return 1;
  }
  if (d->profile) {  // the bracket covers every tcgen05 launch of the decode (conv_post included when it is one)
    VD_CUDA(cudaEventRecord(d->ev_conv1, st));
    d->ev_pending = true;
    d->prof_conv_launches += (long)plan->steps.size() + (plan->post_tc ? 1 : 0);
  }
  ++launches;
  d->last_launches = launches;
  return 0;
}

int vitsdec_decode_host(vitsdec_decoder* d, const float* z, const float* g, float* out, int B, int T) {
  VD_CHECK(d && z && out, "vitsdec_decode_host: null argument");
  DeviceGuard guard(d->device);
  if (!d->hstream) VD_CUDA(cudaStreamCreateWithFlags(&d->hstream, cudaStreamNonBlocking));
  const size_t zb = align_up((size_t)B * d->hp.initial_channel * T * 4, 1024);
  const size_t gb = align_up((size_t)B * std::max(1, d->hp.gin_channels) * 4, 1024);
  const size_t ob = align_up((size_t)B * T * d->hop * 4, 1024);
  const size_t wb = vitsdec_workspace_bytes(d, B, T);
  const size_t need = zb + gb + ob + wb;
  if (need > d->hbuf_bytes) {
    VD_CUDA(cudaStreamSynchronize(d->hstream));
    if (d->hbuf) VD_CUDA(cudaFree(d->hbuf));
    d->hbuf = nullptr; d->hbuf_bytes = 0;
    {
      std::lock_guard<std::mutex> lock(d->mu);
      d->plans.clear();
    }
    VD_CUDA(cudaMalloc(&d->hbuf, need));
    d->hbuf_bytes = need;
  }
  uint8_t* base = static_cast<uint8_t*>(d->hbuf);
  float* zd = reinterpret_cast<float*>(base);
  float* gd = reinterpret_cast<float*>(base + zb);
  float* od = reinterpret_cast<float*>(base + zb + gb);
  void* ws = base + zb + gb + ob;
  VD_CUDA(cudaMemcpyAsync(zd, z, (size_t)B * d->hp.initial_channel * T * 4, cudaMemcpyHostToDevice, d->hstream));
  if (g) VD_CUDA(cudaMemcpyAsync(gd, g, (size_t)B * d->hp.gin_channels * 4, cudaMemcpyHostToDevice, d->hstream));
  if (vitsdec_decode(d, zd, (int64_t)d->hp.initial_channel * T, T, g ? gd : nullptr, od, B, T, ws, wb, d->hstream))
    return 1;
  VD_CUDA(cudaMemcpyAsync(out, od, (size_t)B * T * d->hop * 4, cudaMemcpyDeviceToHost, d->hstream));
  VD_CUDA(cudaStreamSynchronize(d->hstream));
  return 0;
}

int vitsdec_wav_pcm16(int device, const float* wav_dev, int16_t* pcm_dev, int64_t samples, void* stream) {
  VD_CHECK(samples >= 0 && (samples == 0 || (wav_dev && pcm_dev)), "vitsdec_wav_pcm16: null argument");
  DeviceGuard guard(device);
  VD_CHECK(guard.ok, "cudaSetDevice failed");
  return launch_pcm16(wav_dev, pcm_dev, (long)samples, static_cast<cudaStream_t>(stream));
}

int vitsdec_set_option(vitsdec_decoder* d, const char* key, int value) {
  VD_CHECK(d && key, "vitsdec_set_option: null argument");
  std::lock_guard<std::mutex> lock(d->mu);
  if (!strcmp(key, "impl")) {
#ifdef VITSDEC_TESTING
    VD_CHECK(value == 0 || value == 1, "impl: 0 (tcgen05) or 1 (simt)");
#else
    VD_CHECK(value == 0, "impl=1 (CUDA-core cross-check kernels) exists only in the test build libvitsdec_test.so");
#endif
    d->impl = value;
  }
  else if (!strcmp(key, "desc_mode")) d->desc_mode = value;
  else if (!strcmp(key, "debug_keep")) d->debug_keep = value ? 1 : 0;
  else if (!strcmp(key, "fuse_pairs")) d->fuse_pairs = value ? 1 : 0;
  else if (!strcmp(key, "graph")) d->use_graph = value < 0 ? 0 : (value > 2 ? 2 : value);  // 2: capture on first use
  else if (!strcmp(key, "fold")) d->fold = value ? 1 : 0;
  else if (!strcmp(key, "pairf")) d->pairf = value < 0 ? 0 : (value > 3 ? 3 : value);
  else if (!strcmp(key, "par")) d->par = value ? 1 : 0;
  else if (!strcmp(key, "mrfp")) d->mrfp = value < 0 ? 0 : (value > 15 ? 15 : value);   // bit 0: C = 32 stage, bit 1: C = 64 pairs, bit 2: C = 128 stage tail, bit 3: every C = 32 pair
  else if (!strcmp(key, "pdl")) d->pdl = value < 0 ? 0 : (value > 2 ? 2 : value);
  else if (!strcmp(key, "fp16")) {
    // the 16-bit storage format of weights AND activations: packed weights of the other format are useless, so every
    // layer must be loaded again (the Python Generator re-folds by itself) and cached plans are dropped
    const int v = value ? 1 : 0;
    if (v != d->fp16) {
      d->fp16 = v;
      for (int i = 0; i < d->num_real_layers; ++i) d->layers[i].loaded = false;
      d->plans.clear();
      d->last_plan.reset();
    }
  }
  else if (!strcmp(key, "profile")) {
    std::lock_guard<std::mutex> pl(d->prof_mu);
    d->profile = value ? 1 : 0;
    d->prof_conv_ms = 0.0;
    d->prof_conv_launches = 0;
    d->ev_pending = false;
  }
  else { set_error(std::string("unknown option ") + key); return 1; }
  return 0;
}

int vitsdec_profile_read(vitsdec_decoder* d, double* conv_ms, int64_t* conv_launches) {
  VD_CHECK(d && conv_ms && conv_launches, "vitsdec_profile_read: null argument");
  DeviceGuard guard(d->device);
  std::lock_guard<std::mutex> pl(d->prof_mu);
  if (d->ev_pending) {
    float ms = 0.f;
    VD_CUDA(cudaEventSynchronize(d->ev_conv1));
    VD_CUDA(cudaEventElapsedTime(&ms, d->ev_conv0, d->ev_conv1));
    d->prof_conv_ms += ms;
    d->ev_pending = false;
  }
  *conv_ms = d->prof_conv_ms;
  *conv_launches = d->prof_conv_launches;
  return 0;
}

int vitsdec_get_option(const vitsdec_decoder* d, const char* key, int* value) {
  VD_CHECK(d && key && value, "vitsdec_get_option: null argument");
  if (!strcmp(key, "impl")) *value = d->impl;
  else if (!strcmp(key, "desc_mode")) *value = d->desc_mode;
  else if (!strcmp(key, "debug_keep")) *value = d->debug_keep;
  else if (!strcmp(key, "fuse_pairs")) *value = d->fuse_pairs;
  else if (!strcmp(key, "graph")) *value = d->use_graph;
  else if (!strcmp(key, "fold")) *value = d->fold;
  else if (!strcmp(key, "pairf")) *value = d->pairf;
  else if (!strcmp(key, "par")) *value = d->par;
  else if (!strcmp(key, "mrfp")) *value = d->mrfp;
  else if (!strcmp(key, "pdl")) *value = d->pdl;
  else if (!strcmp(key, "fp16")) *value = d->fp16;
  else if (!strcmp(key, "graph_failed")) *value = d->graph_failures.load();
  else if (!strcmp(key, "testing_build")) {
#ifdef VITSDEC_TESTING
    *value = 1;
#else
    *value = 0;
#endif
  }
  else if (!strcmp(key, "hop")) *value = d->hop;
  else if (!strcmp(key, "num_sms")) *value = d->num_sms;
  else { set_error(std::string("unknown option ") + key); return 1; }
  return 0;
}

int vitsdec_last_launch_count(const vitsdec_decoder* d) { return d ? d->last_launches.load() : 0; }

int vitsdec_debug_read(vitsdec_decoder* d, const char* name, float* out, size_t out_elems, int* channels, int* length,
                       void* stream) {
  VD_CHECK(d && name && out, "vitsdec_debug_read: null argument");
  std::shared_ptr<Plan> plan;
  {
    std::lock_guard<std::mutex> lock(d->mu);
    plan = d->last_plan;
  }
  VD_CHECK(plan != nullptr, "vitsdec_debug_read: no decode has run");
  DeviceGuard guard(d->device);
  for (auto& e : plan->debug) {
    if (e.first == name) {
      const int C = std::get<1>(e.second), L = std::get<2>(e.second);
      const int B = plan->steps[0].tc.p.g.B;
      VD_CHECK(out_elems >= (size_t)B * C * L, "vitsdec_debug_read: output too small");
      if (channels) *channels = C;
      if (length) *length = L;
      return launch_unpack_debug(std::get<0>(e.second), std::get<3>(e.second), out, B, L, C,
                                 static_cast<cudaStream_t>(stream), d->fp16);
    }
  }
  set_error(std::string("vitsdec_debug_read: unknown or not-kept tensor ") + name + " (set option debug_keep=1)");
  return 1;
}

static int op_conv_common(int device, Layer& l, const void* x, const float* w, const float* bias, const void* res,
                          float res_gain, float out_slope, void* y, int B, int L, int impl, int desc_mode,
                          cudaStream_t st) {
  DeviceGuard guard(device);
  VD_CHECK(guard.ok, "cudaSetDevice failed");
  cudaDeviceProp prop;
  VD_CUDA(cudaGetDeviceProperties(&prop, device));
  VD_CHECK(prop.major == 10, "vitsdec needs an sm_100 (B200) device");
  const bool fold = (desc_mode & 16) != 0;  // test knob: run the time-folded form of a narrow dilation-1 layer
  const int f16 = (desc_mode & 1024) ? 1 : 0;  // x / res / y are IEEE fp16 instead of bf16 (option "fp16" of the decoder)
  if (fold) {
    fold_geom(l);
    VD_CHECK(impl == 0 && l.fold_r && (L % l.fold_r == 0 || l.dil > 1), "op_conv: layer has no time-folded form");
  }
  if (alloc_layer(l)) return 1;
  float* scale = nullptr;
  VD_CUDA(cudaMalloc(&scale, 4096 * sizeof(float)));
  int rc = 0;
  if (l.kind == kConv) {
    rc = launch_wn_scale(w, nullptr, scale, l.c_out, l.c_in * l.k, st) ||
         launch_pack_conv(w, scale, l.w, l.c_out, l.c_in, l.k, st, 0, f16) ||
         launch_replicate_bias(bias, l.bias, l.c_out, 1, st);
    if (!rc && fold)
      rc = launch_pack_conv_fold(w, scale, l.wfold, l.c_in, l.c_out, l.k, l.fold_r, st, 0, f16) ||
           launch_replicate_bias(bias, l.bias_fold, l.c_out, l.fold_r, st);
  } else {
    rc = launch_wn_scale(w, nullptr, scale, l.c_in, l.c_out * l.k, st) ||
         launch_pack_convT(w, scale, l.w, l.c_in, l.c_out, l.k, l.stride, (l.k - l.stride) / 2, l.geom.ntaps,
                           l.geom.tap_off[0], st, f16) ||
         launch_replicate_bias(bias, l.bias, l.c_out, l.stride, st);
  }
  if (!rc) {
    ConvGeom g = fold ? l.fgeom : l.geom;
    g.B = B; g.L = fold ? L / l.fold_r : L;
    if (fold && g.rho_d > 1) {  // NB reads up to rho_d*r rows past the end of x (masked to zero in the kernel)
      g.L_real = L;
      g.L = ceildiv(L, g.rho_d * l.fold_r);
    }
    ConvEpilogue e{};
    e.bias = fold ? l.bias_fold : l.bias;
    e.res[0] = static_cast<const bf16*>(res);
    e.nres = res ? 1 : 0;
    e.res_gain = res_gain;
    e.out_slope = out_slope;
    e.mrf_scale = 1.f;
    e.out = static_cast<bf16*>(y);
    e.f16 = f16;
    if (impl == 0) {
      ConvTcPlan pl{};
      const bf16* xs[kMaxSeg] = {static_cast<const bf16*>(x), nullptr, nullptr, nullptr};
      rc = plan_conv_tc(&pl, g, xs, fold ? l.wfold : l.w, prop.multiProcessorCount, desc_mode);
      pl.p.trace = g_trace_buffer.load();
      rc = rc || launch_conv_tc(pl, e, st);
    } else {
#ifdef VITSDEC_TESTING
      const bf16* xs[kMaxSeg] = {static_cast<const bf16*>(x), nullptr, nullptr, nullptr};
      rc = launch_conv_simt(g, e, xs, l.w, st);
#else
      set_error("impl=1 (CUDA-core cross-check kernels) exists only in the test build libvitsdec_test.so");
      rc = 1;
#endif
    }
  }
  cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(scale); cudaFree(l.w); cudaFree(l.bias); cudaFree(l.wfold); cudaFree(l.bias_fold);
  if (!rc && se != cudaSuccess) { set_error(std::string("op_conv: ") + cudaGetErrorString(se)); rc = 1; }
  return rc;
}

int vitsdec_debug_set_trace(void* trace_dev) {
  g_trace_buffer.store(static_cast<unsigned long long*>(trace_dev));
  return 0;
}

int vitsdec_op_conv1d(int device, const void* x, const float* w, const float* bias, const void* res, float res_gain,
                      float out_slope, void* y, int B, int L, int c_in, int c_out, int k, int dilation, int impl,
                      int desc_mode, void* stream) {
  VD_CHECK(x && w && y, "vitsdec_op_conv1d: null argument");
  VD_CHECK(k % 2 == 1 && k <= kMaxTaps && k >= 1, "k must be odd and <= 31");
  Layer l;
  l.kind = kConv; l.c_in = c_in; l.c_out = c_out; l.k = k; l.dil = dilation;
  conv_geom(l);
  return op_conv_common(device, l, x, w, bias, res, res_gain, out_slope, y, B, L, impl, desc_mode,
                        static_cast<cudaStream_t>(stream));
}

int vitsdec_op_resblock_pair(int device, const void* x, const float* w1, const float* b1, const float* w2,
                             const float* b2, void* y, int B, int L, int channels, int k, int dilation, float slope,
                             void* stream) {
  VD_CHECK(x && w1 && w2 && y, "vitsdec_op_resblock_pair: null argument");
  VD_CHECK(pair_supported(channels, k, dilation), "vitsdec_op_resblock_pair: shape not supported by the fused kernel");
  DeviceGuard guard(device);
  VD_CHECK(guard.ok, "cudaSetDevice failed");
  cudaDeviceProp prop;
  VD_CUDA(cudaGetDeviceProperties(&prop, device));
  VD_CHECK(prop.major == 10, "vitsdec needs an sm_100 (B200) device");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bf16* w = nullptr;
  float *scale = nullptr, *bias = nullptr;
  const size_t per = (size_t)k * channels * channels;
  VD_CUDA(cudaMalloc(&w, 2 * per * sizeof(bf16)));
  VD_CUDA(cudaMalloc(&scale, 4096 * sizeof(float)));
  VD_CUDA(cudaMalloc(&bias, 2 * channels * sizeof(float)));
  int rc = launch_wn_scale(w1, nullptr, scale, channels, channels * k, st) ||
           launch_pack_conv(w1, scale, w, channels, channels, k, st) ||
           launch_pack_conv(w2, scale, w + per, channels, channels, k, st) ||
           launch_replicate_bias(b1, bias, channels, 1, st) || launch_replicate_bias(b2, bias + channels, channels, 1, st);
  if (!rc) {
    PairPlan pl{};
    rc = plan_conv_pair(&pl, B, L, channels, k, dilation, static_cast<const bf16*>(x), w, prop.multiProcessorCount);
    pl.p.trace = g_trace_buffer.load();
    rc = rc || launch_conv_pair(pl, bias, bias + channels, slope, static_cast<bf16*>(y), st);
  }
  cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(w); cudaFree(scale); cudaFree(bias);
  if (!rc && se != cudaSuccess) { set_error(std::string("op_resblock_pair: ") + cudaGetErrorString(se)); rc = 1; }
  return rc;
}

int vitsdec_op_mrf_pairs(int device, int nbr, const void* const* xs, const float* const* w1, const float* const* b1,
                         const float* const* w2, const float* const* b2, void* y, int B, int L, int channels, const int* k,
                         const int* dilation, float slope, float out_slope, void* stream) {
  VD_CHECK(xs && w1 && b1 && w2 && b2 && y && k && dilation, "vitsdec_op_mrf_pairs: null argument");
  const bool c128 = channels == 128 && nbr >= 1 && nbr <= kM8MaxBr && mrf128_supported(channels, nbr, k, dilation);
  VD_CHECK(c128 || (nbr >= 1 && nbr <= kMpMaxBr && (channels == 32 || channels == 64) && L % (64 / channels) == 0 &&
               mrfp_supported(channels, nbr, k, dilation)),
           "vitsdec_op_mrf_pairs: shape not supported by the fused kernel (C = 32 with an even length or C = 64, <= 3 "
           "branches, all weights resident in shared memory)");
  DeviceGuard guard(device);
  VD_CHECK(guard.ok, "cudaSetDevice failed");
  cudaDeviceProp prop;
  VD_CUDA(cudaGetDeviceProperties(&prop, device));
  VD_CHECK(prop.major == 10, "vitsdec needs an sm_100 (B200) device");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int sumk = 0;
  for (int j = 0; j < nbr; ++j) sumk += k[j];
  bf16* w = nullptr;
  float *scale = nullptr, *bias = nullptr;
  const size_t tapsz = (size_t)channels * channels;
  VD_CUDA(cudaMalloc(&w, 2 * sumk * tapsz * sizeof(bf16)));
  VD_CUDA(cudaMalloc(&scale, 4096 * sizeof(float)));
  VD_CUDA(cudaMalloc(&bias, (size_t)(2 * kMpMaxBr + 1) * channels * sizeof(float)));
  int rc = 0, tap = 0;
  for (int pass = 0; pass < 2 && !rc; ++pass)        // packed order: c1 of every branch, then c2 of every branch
    for (int j = 0; j < nbr && !rc; ++j) {
      const float* src = pass == 0 ? w1[j] : w2[j];
      rc = launch_wn_scale(src, nullptr, scale, channels, channels * k[j], st) ||
           launch_pack_conv(src, scale, w + tap * tapsz, channels, channels, k[j], st) ||
           launch_replicate_bias(pass == 0 ? b1[j] : b2[j], bias + (pass * kMpMaxBr + j) * channels, channels, 1, st);
      tap += k[j];
    }
  float* b2sum = bias + 2 * kMpMaxBr * channels;
  if (!rc)
    rc = launch_sum_bias(bias + kMpMaxBr * channels, nbr > 1 ? bias + (kMpMaxBr + 1) * channels : nullptr,
                         nbr > 2 ? bias + (kMpMaxBr + 2) * channels : nullptr, nullptr, b2sum, channels, st);
  if (!rc && c128) {
    static Mrf128Plan pl128;   // (a plan is ~1.5 KB of tensor maps and tables; op entries are serial test hooks)
    const bf16* xin[kM8MaxBr] = {nullptr, nullptr, nullptr};
    const float* bb[kM8MaxBr] = {nullptr, nullptr, nullptr};
    for (int j = 0; j < nbr; ++j) { xin[j] = static_cast<const bf16*>(xs[j]); bb[j] = bias + j * channels; }
    rc = plan_conv_mrf128(&pl128, B, L, nbr, k, dilation, xin, w, prop.multiProcessorCount);
    rc = rc || launch_conv_mrf128(pl128, bb, b2sum, slope, out_slope, static_cast<bf16*>(y), st);
  } else if (!rc) {
    MrfpPlan pl{};
    const bf16* xin[kMpMaxBr] = {nullptr, nullptr, nullptr};
    const float* bb[kMpMaxBr] = {nullptr, nullptr, nullptr};
    for (int j = 0; j < nbr; ++j) { xin[j] = static_cast<const bf16*>(xs[j]); bb[j] = bias + j * channels; }
    rc = plan_conv_mrfp(&pl, B, L, channels, nbr, k, dilation, xin, w, prop.multiProcessorCount);
    pl.p.trace = g_trace_buffer.load();
    rc = rc || launch_conv_mrfp(pl, bb, b2sum, slope, out_slope, static_cast<bf16*>(y), st);
  }
  cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(w); cudaFree(scale); cudaFree(bias);
  if (!rc && se != cudaSuccess) { set_error(std::string("op_mrf_pairs: ") + cudaGetErrorString(se)); rc = 1; }
  return rc;
}

int vitsdec_op_resblock_pair_folded(int device, const void* x, const float* w1, const float* b1, const float* w2,
                                    const float* b2, void* y, int B, int L, int channels, int k, int dilation,
                                    float slope, void* stream) {
  VD_CHECK(x && w1 && w2 && y, "vitsdec_op_resblock_pair_folded: null argument");
  VD_CHECK(pairf_supported(channels, k, dilation) && L % (128 / channels) == 0,
           "vitsdec_op_resblock_pair_folded: shape not supported by the folded fused kernel");
  DeviceGuard guard(device);
  VD_CHECK(guard.ok, "cudaSetDevice failed");
  cudaDeviceProp prop;
  VD_CUDA(cudaGetDeviceProperties(&prop, device));
  VD_CHECK(prop.major == 10, "vitsdec needs an sm_100 (B200) device");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nt = pairf_taps(channels, k), r = 128 / channels;
  bf16* w = nullptr;
  float *scale = nullptr, *bias = nullptr;
  VD_CUDA(cudaMalloc(&w, (size_t)2 * nt * 128 * 128 * sizeof(bf16)));
  VD_CUDA(cudaMalloc(&scale, 4096 * sizeof(float)));
  VD_CUDA(cudaMalloc(&bias, 2 * channels * sizeof(float)));
  int rc = launch_wn_scale(w1, nullptr, scale, channels, channels * k, st) ||
           launch_pack_conv_fold(w1, scale, w, channels, channels, k, r, st) ||
           launch_pack_conv_fold(w2, scale, w + (size_t)nt * 128 * 128, channels, channels, k, r, st) ||
           launch_replicate_bias(b1, bias, channels, 1, st) || launch_replicate_bias(b2, bias + channels, channels, 1, st);
  if (!rc) {
    PairFPlan pl{};
    rc = plan_conv_pairf(&pl, B, L, channels, k, dilation, static_cast<const bf16*>(x), w, prop.multiProcessorCount);
    pl.p.trace = g_trace_buffer.load();
    rc = rc || launch_conv_pairf(pl, bias, bias + channels, slope, static_cast<bf16*>(y), st);
  }
  cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(w); cudaFree(scale); cudaFree(bias);
  if (!rc && se != cudaSuccess) { set_error(std::string("op_resblock_pair_folded: ") + cudaGetErrorString(se)); rc = 1; }
  return rc;
}

int vitsdec_op_conv_transpose1d(int device, const void* x, const float* w, const float* bias, float out_slope, void* y,
                                int B, int L, int c_in, int c_out, int k, int stride, int impl, void* stream) {
  VD_CHECK(x && w && y, "vitsdec_op_conv_transpose1d: null argument");
  VD_CHECK(k >= stride && (k - stride) % 2 == 0, "need k >= stride and (k - stride) even");
  Layer l;
  l.kind = kConvT; l.c_in = c_in; l.c_out = c_out; l.k = k; l.stride = stride;
  convT_geom(l);
  VD_CHECK(l.geom.ntaps <= kMaxTaps, "kernel too large");
  return op_conv_common(device, l, x, w, bias, nullptr, 1.f, out_slope, y, B, L, impl, 0,
                        static_cast<cudaStream_t>(stream));
}

}  // extern "C"
