// The normalising flow between the prior and the waveform decoder, on the same tcgen05 convolution primitive.
//
// Reference being replaced: ResidualCouplingBlock.forward(x, x_mask, g, reverse) (/root/reference/models.py:179-209),
// i.e. n_flows x [ResidualCouplingLayer (modules.py:298-343, mean_only=True), Flip (modules.py:270-277)] with the
// WaveNet-style WN (modules.py:111-184) and its gate fused_add_tanh_sigmoid_multiply (commons.py:103-110).
// SynthesizerTrn.infer runs it in reverse right before the decoder (models.py:521), voice_conversion in both directions
// (models.py:530-531).  SURVEY.md section 8f ranks it first among the callers of the decoder path.
//
// Data layout: the latent stays fp32 channels-last X[B][T][C] for the whole block (the coupling x1 -/+= m is exact in
// fp32); convolution operands are bf16 channels-last like the decoder's.  `* x_mask` after every layer is a per-row
// multiply in the conv epilogues (ConvEpilogue::rowmask; the mask is binary in the reference, commons.sequence_mask).
// One coupling layer:
//   flip_split   X <- flip(X) (reverse) ; X0 = bf16(X[:, :C/2])
//   pre          H = (W_pre X0 + b) * mask                                               modules.py:326
//   cond         CB[l][b] = cond_layer(g)[l]            (weight-normed 1x1 conv = GEMV)   modules.py:153-154
//   per layer l  ACT = tanh(a) * sigmoid(b),  (a | b) = in_l(H) + bias + CB[l][b]        modules.py:157, commons.py:105-110
//                (conv k, dilation rate^l; the gate runs in the conv's epilogue: in_l's output channels are packed
//                interleaved, (a_j, b_j) in adjacent columns, so one thread holds both halves of a pair)
//                RS  = res_skip_l(ACT): ONE launch with a split epilogue (ConvEpilogue::split_col) --
//                H   = (RS[:, :Hc] + H) * mask               (bf16, not for the last layer)   modules.py:171-172
//                S  += RS[:, Hc:]                            (fp32 accumulator)               modules.py:173-175
//   post         M = (W_post bf16(S * mask) + b) * mask                (fp32)            modules.py:328
//   couple       X[:, C/2:] = (X[:, C/2:] - M) * mask  (reverse)  |  M + X[:, C/2:] * mask (forward)   modules.py:335-343
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/vitsdec.h"
#include "common.cuh"
#include "conv_tc.h"
#include "pack.h"

namespace vd {

typedef __nv_bfloat16 bf16;

// z fp32 [B][C][T] (strided) -> fp32 [B][T][C], with the first coupling's Flip (reverse pass) and its operand
// X0 = 16-bit copy of the first C/2 channels in the same pass (was a launch of its own)
__global__ void flow_in_kernel(const float* __restrict__ z, long sb, long sc, float* __restrict__ out,
                               bf16* __restrict__ x0, int C, int T, int flip, int f16) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < T) ? z[b * sb + c * sc + t] : 0.f;
  }
  __syncthreads();
  const int half = C / 2;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T && c < C) {
      const int cc = flip ? C - 1 - c : c;
      const float v = tile[threadIdx.x][i];
      const long row = (long)b * T + t;
      out[row * C + cc] = v;
      if (cc < half) x0[row * half + cc] = pack_act_rt(v, f16);
    }
  }
}

// fp32 [B][T][C] -> fp32 [B][C][T], optionally with the channel flip of a trailing Flip module; the LAST coupling's
// x1 = (x1 - m) * mask (reverse) / m + x1 * mask (forward) is applied on the way (m: [rows][C/2], was a launch of its own)
__global__ void flow_out_kernel(const float* __restrict__ x, const float* __restrict__ m, const float* __restrict__ mask,
                                float* __restrict__ out, int C, int T, int flip, int reverse) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int half = C / 2;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (t < T && c < C) {
      const int cs = flip ? C - 1 - c : c;
      const long row = (long)b * T + t;
      v = x[row * C + cs];
      if (cs >= half) {
        const float mv = m[row * half + cs - half], mk = mask[row];
        v = reverse ? (v - mv) * mk : mv + v * mk;
      }
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    if (c < C && t < T) out[((long)b * C + c) * T + t] = tile[threadIdx.x][i];
  }
}

// One coupling's x1 = (x1 - m) * mask (reverse) or m + x1 * mask (forward) -- m is already masked (modules.py:328,
// 335-343, logs = 0) -- THEN the Flip in front of the next coupling (modules.py:272) and that coupling's operand
// X0 = 16-bit copy of the new first half, in one pass over X (were two launches between every pair of couplings)
__global__ void flow_couple_flip_split_kernel(float* __restrict__ x, const float* __restrict__ m,
                                              const float* __restrict__ mask, bf16* __restrict__ x0, long rows, int C,
                                              int reverse, int f16) {
  const int half = C / 2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < rows * half; i += (long)gridDim.x * blockDim.x) {
    const long r = i / half;
    const int c = i % half;
    float* row = x + r * C;
    const float lo = row[c];                               // x0 of this coupling: unchanged
    const float mv = m[r * half + (half - 1 - c)], mk = mask[r];
    float hi = row[C - 1 - c];                             // x1 element that the flip brings to position c
    hi = reverse ? (hi - mv) * mk : mv + hi * mk;
    row[c] = hi;
    row[C - 1 - c] = lo;
    x0[i] = pack_act_rt(hi, f16);
  }
}

__global__ void flow_fill_mask_kernel(float* mask, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) mask[i] = 1.f;
}

// cb[coupling][l][b][n] = cond.bias[l*N + n] + sum_ci w[l*N + n][ci] * g[b][ci]   (one warp per output; N = 2*hidden).
// All couplings in ONE launch (blockIdx.z): the conditioning depends on g only, so it leaves the per-coupling chain.
struct FlowCondPtrs {
  const float* w[16];
  const float* b[16];
};
__global__ void flow_cond_kernel(const FlowCondPtrs ptrs, const float* __restrict__ g, float* __restrict__ cb_all, int B,
                                 int N, int nl, int gin) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  if (warp >= N * nl) return;
  const float* __restrict__ wc = ptrs.w[blockIdx.z];
  const float* __restrict__ bc = ptrs.b[blockIdx.z];
  float* __restrict__ cb = cb_all + (size_t)blockIdx.z * nl * B * N;
  float s = 0.f;
  for (int i = lane; i < gin; i += 32) s = fmaf(wc[(long)warp * gin + i], g[(long)b * gin + i], s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  // column order of the interleaved in_layer: output o < N/2 (tanh half) -> 2o, o >= N/2 (sigmoid half) -> 2(o-N/2)+1
  const int o = warp % N, col = o < N / 2 ? 2 * o : 2 * (o - N / 2) + 1;
  if (lane == 0) cb[((long)(warp / N) * B + b) * N + col] = s + bc[warp];
}

// w_eff = v * g / ||v|| kept in fp32 (cond_layer: used by the GEMV above, not by the tensor cores)
__global__ void flow_fold_f32_kernel(const float* __restrict__ v, const float* __restrict__ scale, float* __restrict__ w,
                                     long rows, int inner) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < rows * inner; i += (long)gridDim.x * blockDim.x)
    w[i] = v[i] * scale[i / inner];
}

struct FlowConv {             // one tensor-core convolution of the block
  int c_in = 0, c_out = 0, k = 1, dil = 1;
  bf16* w = nullptr;          // packed [k][c_out][c_in]
  float* bias = nullptr;      // [c_out]
  ConvGeom geom{};
};

struct FlowLayer {            // one WN layer
  FlowConv in, rs;            // rs: res_skip_layers[l], 2*Hc outputs (Hc for the last layer: skip only)
  bool has_res = false;
};

struct FlowCoupling {
  FlowConv pre, post;
  // skip sum of a WN as ONE launch (n_layers <= kMaxSeg): segment l = the skip rows of res_skip_layers[l] on acts_l, all
  // accumulated in TMEM; the per-layer res_skip launch then only updates the residual stream (rows [0, H)) -- no fp32
  // skip tensor read-modify-written by every layer (it made a 2 GFLOP 1 x 1 conv cost 23-28 us at 16 x 862)
  FlowConv skip;
  float* skip_b = nullptr;    // [n_layers][H]: the layers' skip biases; skip.bias is their sum
  std::vector<FlowLayer> layers;
  float* cond_w = nullptr;    // folded fp32 [nl*2H][gin]
  float* cond_b = nullptr;
  std::map<std::string, bool> loaded;
};

static void flow_conv_geom(FlowConv& c) {
  ConvGeom& g = c.geom;
  g = ConvGeom{};
  g.c_in = c.c_in;
  g.n_total = c.c_out;
  g.ntaps = c.k;
  for (int j = 0; j < c.k; ++j) {
    g.tap_off[j] = (j - (c.k - 1) / 2) * c.dil;   // padding = (k*d - d)/2, modules.py:133
    g.tap_nlo[j] = 0;
    g.tap_nhi[j] = c.c_out;
    g.tap_kmask[j] = ~0u;
  }
  g.nseg = 1;
  g.seg_tap_end[0] = c.k;
}

static int flow_conv_alloc(FlowConv& c, int c_in, int c_out, int k, int dil) {
  c.c_in = c_in; c.c_out = c_out; c.k = k; c.dil = dil;
  flow_conv_geom(c);
  VD_CUDA(cudaMalloc(&c.w, (size_t)k * c_out * c_in * sizeof(bf16)));
  VD_CUDA(cudaMalloc(&c.bias, (size_t)c_out * sizeof(float)));
  VD_CUDA(cudaMemset(c.bias, 0, (size_t)c_out * sizeof(float)));
  return 0;
}

struct FlowStep {
  ConvTcPlan tc;
  ConvEpilogue ep;
};

struct FlowPlan {            // per (B, T, workspace): tensor maps of every conv launch, in execution order per coupling
  ~FlowPlan() {
    for (cudaGraphExec_t e : graph_exec)
      if (e) cudaGraphExecDestroy(e);
  }
  std::vector<std::vector<FlowStep>> steps;   // [coupling][launch]
  // the whole block between the input and output transposes replays as one CUDA graph (it only touches workspace
  // pointers); index = reverse * 2 + has_g
  cudaGraphExec_t graph_exec[4] = {nullptr, nullptr, nullptr, nullptr};
  bool graph_failed = false;
  int uses = 0;   // the graph is captured at the third use of a plan (capture + instantiate ~10 ms, see decoder.cu Plan)
};

static size_t fl_align(size_t v) { return (v + 1023) / 1024 * 1024; }

}  // namespace vd

using namespace vd;

struct vitsdec_flow {
  vitsdec_flow_hparams hp;
  int device = 0, num_sms = 148;
  std::vector<FlowCoupling> cpl;
  std::vector<std::string> names;
  float* scale_scratch = nullptr;
  std::mutex mu;
  std::list<std::pair<std::tuple<int, int, const void*>, std::shared_ptr<FlowPlan>>> plans;
  cudaStream_t cstream = nullptr;   // capture-only stream
  bool skipsum = false;             // the WN skip sum runs as one multi-segment launch (FlowCoupling::skip)
  int fp16 = 0;                     // vitsdec_flow_set_option("fp16"): conv operands / stored activations are fp16
  int pdl = 1;                      // vitsdec_flow_set_option("pdl"): 0 = no programmatic dependent launch
};

namespace vd {

struct FlowWs {
  size_t x, x0, h0, h1, act, s, outb, m, cb, mask, g, total;
};

static FlowWs flow_ws(const vitsdec_flow* f, int B, int T) {
  const size_t rows = (size_t)B * T, C = f->hp.channels, H = f->hp.hidden_channels;
  FlowWs w{};
  size_t o = 0;
  w.x = o; o += fl_align(rows * C * 4);
  w.x0 = o; o += fl_align(rows * (C / 2) * 2);
  w.h0 = o; o += fl_align(rows * H * 2);
  w.h1 = o; o += fl_align(rows * H * 2);
  w.act = o; o += (size_t)(f->skipsum ? f->hp.n_layers : 1) * fl_align(rows * H * 2);
  w.s = o; o += fl_align(rows * H * 4);
  w.outb = o; o += fl_align(rows * H * 2);
  w.m = o; o += fl_align(rows * (C / 2) * 4);
  w.cb = o; o += fl_align((size_t)f->hp.n_flows * f->hp.n_layers * B * 2 * H * 4);
  w.mask = o; o += fl_align(rows * 4);
  w.g = o; o += fl_align((size_t)B * (f->hp.gin_channels > 0 ? f->hp.gin_channels : 1) * 4);
  w.total = o + 4096;
  return w;
}

struct FlowDeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit FlowDeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~FlowDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int flow_build_plan(vitsdec_flow* f, FlowPlan& pl, int B, int T, uint8_t* ws) {
  const FlowWs w = flow_ws(f, B, T);
  const int H = f->hp.hidden_channels, nl = f->hp.n_layers;
  bf16* X0 = reinterpret_cast<bf16*>(ws + w.x0);
  bf16* Hb[2] = {reinterpret_cast<bf16*>(ws + w.h0), reinterpret_cast<bf16*>(ws + w.h1)};
  bf16* ACT = reinterpret_cast<bf16*>(ws + w.act);
  float* S = reinterpret_cast<float*>(ws + w.s);
  bf16* OUTB = reinterpret_cast<bf16*>(ws + w.outb);
  float* M = reinterpret_cast<float*>(ws + w.m);
  float* CB = reinterpret_cast<float*>(ws + w.cb);
  const float* mask = reinterpret_cast<const float*>(ws + w.mask);
  pl.steps.resize(f->cpl.size());
  for (size_t ci = 0; ci < f->cpl.size(); ++ci) {
    FlowCoupling& c = f->cpl[ci];
    auto push_n = [&](const FlowConv& cv, const bf16* const* xin, ConvEpilogue e) -> int {
      FlowStep s{};
      ConvGeom g = cv.geom;
      g.B = B; g.L = T;
      e.bias = cv.bias;
      e.rowmask = mask;
      e.res_gain = 1.f;
      e.f16 = f->fp16;
      if (e.out_slope == 0.f) e.out_slope = 1.f;
      if (e.mrf_scale == 0.f) e.mrf_scale = 1.f;
      const bf16* xs[kMaxSeg] = {nullptr, nullptr, nullptr, nullptr};
      for (int i = 0; i < g.nseg; ++i) xs[i] = xin[i];
      if (plan_conv_tc(&s.tc, g, xs, cv.w, f->num_sms, 0, /*allow_swap=*/false)) return 1;
      s.ep = e;
      // programmatic dependent launch (see decoder.cu): the flow's launches are all short; the small kernels between
      // them (flip / cond / couple) neither trigger nor wait, which degrades to ordinary stream order around them
      s.tc.pdl = f->pdl && s.tc.p.total_tiles <= 2 * f->num_sms;
      if (bind_residual_tc(s.tc, e)) return 1;
      pl.steps[ci].push_back(s);
      return 0;
    };
    auto push = [&](const FlowConv& cv, const bf16* x, ConvEpilogue e) -> int { return push_n(cv, &x, e); };
    const size_t act_stride = fl_align((size_t)B * T * H * 2) / 2;   // elements between the layers' gate outputs (skipsum)
    const bf16* acts[kMaxSeg] = {nullptr, nullptr, nullptr, nullptr};
    int cur = 0;
    {  // pre
      ConvEpilogue e{};
      e.out = Hb[cur];
      if (push(c.pre, X0, e)) return 1;
    }
    for (int l = 0; l < nl; ++l) {
      FlowLayer& ly = c.layers[l];
      bf16* act_l = f->skipsum ? ACT + (size_t)l * act_stride : ACT;
      {
        ConvEpilogue e{};
        e.out = act_l;
        e.gate = 1;
        if (f->hp.gin_channels) e.bias_b = CB + ((size_t)ci * nl + l) * B * 2 * H;  // only used when g is given
        if (push(ly.in, Hb[cur], e)) return 1;
      }
      if (f->skipsum) {
        acts[l] = act_l;
        if (ly.has_res) {           // residual stream only: x = (x + res_acts) * mask, modules.py:170-171
          ConvEpilogue e{};
          e.res[0] = Hb[cur];
          e.nres = 1;
          e.out = Hb[cur ^ 1];
          if (push(ly.rs, act_l, e)) return 1;
          cur ^= 1;
        }
      } else {
        ConvEpilogue e{};
        e.mrf = S;
        if (ly.has_res) {           // residual half -> next H (bf16), skip half -> S (fp32), one launch
          e.split_col = H;
          e.res[0] = Hb[cur];
          e.nres = 1;
          e.out = Hb[cur ^ 1];
          e.mrf_mode = l == 0 ? 1 : 2;
        } else {                    // last layer: bf16(S + v) is post's operand
          e.mrf_mode = 3;
          if (nl == 1) e.mrf = nullptr;
          e.out = OUTB;
        }
        if (push(ly.rs, ACT, e)) return 1;
        if (ly.has_res) cur ^= 1;
      }
    }
    if (f->skipsum) {   // output = sum_l skip_l(acts_l), * mask (modules.py:173-176): one launch, the sum lives in TMEM
      ConvEpilogue e{};
      e.out = OUTB;
      if (push_n(c.skip, acts, e)) return 1;
    }
    {  // post: fp32 store of m
      ConvEpilogue e{};
      e.mrf = M;
      e.mrf_mode = 1;
      if (push(c.post, OUTB, e)) return 1;
    }
  }
  return 0;
}

static int grid1d(long n) { return (int)std::min<long>((n + 255) / 256, 148 * 8); }

}  // namespace vd

extern "C" {

int vitsdec_flow_create(const vitsdec_flow_hparams* hp, int device, vitsdec_flow** out) {
  VD_CHECK(hp && out, "vitsdec_flow_create: null argument");
  VD_CHECK(hp->channels % 64 == 0 && hp->channels > 0, "flow: channels must be a multiple of 64 (32-channel halves)");
  VD_CHECK(hp->hidden_channels % 32 == 0 && hp->hidden_channels > 0, "flow: hidden_channels must be a multiple of 32");
  VD_CHECK(hp->kernel_size % 2 == 1 && hp->kernel_size >= 1 && hp->kernel_size <= kMaxTaps, "flow: bad kernel_size");
  VD_CHECK(hp->n_layers >= 1 && hp->n_layers <= 16 && hp->n_flows >= 1 && hp->n_flows <= 16, "flow: bad layer counts");
  VD_CHECK(hp->dilation_rate >= 1 && hp->gin_channels >= 0, "flow: bad dilation_rate / gin_channels");
  int ndev = 0;
  VD_CUDA(cudaGetDeviceCount(&ndev));
  VD_CHECK(device >= 0 && device < ndev, "vitsdec_flow_create: no such CUDA device (there is no CPU fallback)");
  cudaDeviceProp prop;
  VD_CUDA(cudaGetDeviceProperties(&prop, device));
  VD_CHECK(prop.major == 10, "vitsdec needs an sm_100 (B200) device: kernels are tcgen05/TMA only");
  FlowDeviceGuard guard(device);
  VD_CHECK(guard.ok, "cudaSetDevice failed");
  std::unique_ptr<vitsdec_flow> f(new vitsdec_flow());
  f->hp = *hp;
  f->device = device;
  f->num_sms = prop.multiProcessorCount;
  const int C2 = hp->channels / 2, H = hp->hidden_channels, nl = hp->n_layers;
  f->skipsum = nl <= kMaxSeg;
  f->cpl.resize(hp->n_flows);
  for (int i = 0; i < hp->n_flows; ++i) {
    FlowCoupling& c = f->cpl[i];
    if (f->skipsum) {   // nl segments of one tap each, all at row offset 0
      if (flow_conv_alloc(c.skip, H, H, nl, 1)) return 1;
      ConvGeom& sg = c.skip.geom;
      for (int l = 0; l < nl; ++l) { sg.tap_off[l] = 0; sg.seg_tap_end[l] = l + 1; }
      sg.nseg = nl;
      VD_CUDA(cudaMalloc(&c.skip_b, (size_t)nl * H * sizeof(float)));
      VD_CUDA(cudaMemset(c.skip_b, 0, (size_t)nl * H * sizeof(float)));
    }
    const std::string p = "flows." + std::to_string(2 * i) + ".";   // odd entries are Flip modules (models.py:199-201)
    if (flow_conv_alloc(c.pre, C2, H, 1, 1)) return 1;
    f->names.push_back(p + "pre");
    c.layers.resize(nl);
    int dil = 1;
    for (int l = 0; l < nl; ++l) {
      FlowLayer& ly = c.layers[l];
      VD_CHECK((hp->kernel_size - 1) / 2 * dil < 4096, "flow: dilation too large");
      if (flow_conv_alloc(ly.in, H, 2 * H, hp->kernel_size, dil)) return 1;
      ly.has_res = l < nl - 1;
      if (flow_conv_alloc(ly.rs, H, (ly.has_res && !f->skipsum) ? 2 * H : H, 1, 1)) return 1;
      dil *= hp->dilation_rate;
    }
    for (int l = 0; l < nl; ++l) f->names.push_back(p + "enc.in_layers." + std::to_string(l));
    for (int l = 0; l < nl; ++l) f->names.push_back(p + "enc.res_skip_layers." + std::to_string(l));
    if (hp->gin_channels) {
      VD_CUDA(cudaMalloc(&c.cond_w, (size_t)nl * 2 * H * hp->gin_channels * sizeof(float)));
      VD_CUDA(cudaMalloc(&c.cond_b, (size_t)nl * 2 * H * sizeof(float)));
      f->names.push_back(p + "enc.cond_layer");
    }
    if (flow_conv_alloc(c.post, H, C2, 1, 1)) return 1;
    f->names.push_back(p + "post");
  }
  VD_CUDA(cudaMalloc(&f->scale_scratch, 8192 * sizeof(float)));
  *out = f.release();
  return 0;
}

void vitsdec_flow_destroy(vitsdec_flow* f) {
  if (!f) return;
  FlowDeviceGuard guard(f->device);
  cudaDeviceSynchronize();
  auto drop = [](FlowConv& c) { cudaFree(c.w); cudaFree(c.bias); };
  for (FlowCoupling& c : f->cpl) {
    drop(c.pre); drop(c.post); drop(c.skip); cudaFree(c.skip_b);
    for (FlowLayer& l : c.layers) { drop(l.in); drop(l.rs); }
    cudaFree(c.cond_w); cudaFree(c.cond_b);
  }
  cudaFree(f->scale_scratch);
  f->plans.clear();
  if (f->cstream) cudaStreamDestroy(f->cstream);
  delete f;
}

int vitsdec_flow_num_layers(const vitsdec_flow* f) { return f ? (int)f->names.size() : 0; }
const char* vitsdec_flow_layer_name(const vitsdec_flow* f, int i) {
  if (!f || i < 0 || i >= (int)f->names.size()) return nullptr;
  return f->names[i].c_str();
}

int vitsdec_flow_load_layer(vitsdec_flow* f, const char* name, const float* w, const float* wg, const float* bias,
                            void* stream) {
  VD_CHECK(f && name && w && bias, "vitsdec_flow_load_layer: null argument");
  FlowDeviceGuard guard(f->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::lock_guard<std::mutex> lock(f->mu);
  const std::string n(name);
  int fi = -1, pos = 0;
  VD_CHECK(sscanf(name, "flows.%d.%n", &fi, &pos) == 1 && fi % 2 == 0 && fi / 2 < (int)f->cpl.size(),
           std::string("vitsdec_flow_load_layer: unknown layer ") + name);
  FlowCoupling& c = f->cpl[fi / 2];
  const std::string rest = n.substr(pos);
  const int H = f->hp.hidden_channels, nl = f->hp.n_layers, gin = f->hp.gin_channels;
  auto load_conv = [&](FlowConv& cv, const float* wsrc, const float* scale, const float* b) -> int {
    return launch_pack_conv(wsrc, scale, cv.w, cv.c_out, cv.c_in, cv.k, st, 0, f->fp16) ||
           launch_replicate_bias(b, cv.bias, cv.c_out, 1, st);
  };
  int li = -1;
  if (rest == "pre" || rest == "post") {
    VD_CHECK(wg == nullptr, "flow: pre / post are not weight-normed in the reference (modules.py:318-320)");
    FlowConv& cv = rest == "pre" ? c.pre : c.post;
    VD_CHECK(cv.c_out <= 8192, "flow: too many channels");
    if (launch_wn_scale(w, nullptr, f->scale_scratch, cv.c_out, cv.c_in * cv.k, st)) return 1;
    if (load_conv(cv, w, f->scale_scratch, bias)) return 1;
  } else if (sscanf(rest.c_str(), "enc.in_layers.%d", &li) == 1 && li >= 0 && li < nl) {
    FlowConv& cv = c.layers[li].in;
    VD_CHECK(cv.c_out <= 8192, "flow: too many channels");
    if (launch_wn_scale(w, wg, f->scale_scratch, cv.c_out, cv.c_in * cv.k, st)) return 1;
    // gate pairs side by side: packed row 2j = tanh-half row j, 2j+1 = sigmoid-half row H + j (ConvEpilogue::gate)
    if (launch_pack_conv(w, f->scale_scratch, cv.w, cv.c_out, cv.c_in, cv.k, st, /*interleave=*/1, f->fp16))
      return 1;
    if (launch_interleave_bias(bias, cv.bias, cv.c_out, st)) return 1;
  } else if (sscanf(rest.c_str(), "enc.res_skip_layers.%d", &li) == 1 && li >= 0 && li < nl) {
    FlowLayer& ly = c.layers[li];
    const int rows = ly.has_res ? 2 * H : H;
    VD_CHECK(rows <= 8192, "flow: too many channels");
    if (launch_wn_scale(w, wg, f->scale_scratch, rows, H, st)) return 1;
    // rows [0, H) update the residual stream, rows [H, 2H) feed the skip sum (modules.py:170-173)
    if (f->skipsum) {
      if (ly.has_res && load_conv(ly.rs, w, f->scale_scratch, bias)) return 1;   // ly.rs has H output rows here
      const int srow = ly.has_res ? H : 0;   // the last layer's H rows are all skip
      if (launch_pack_conv(w + (size_t)srow * H, f->scale_scratch + srow, c.skip.w + (size_t)li * H * H, H, H, 1, st, 0,
                           f->fp16))
        return 1;
      if (launch_replicate_bias(bias + srow, c.skip_b + (size_t)li * H, H, 1, st)) return 1;
      const float* sb[kMaxSeg] = {nullptr, nullptr, nullptr, nullptr};
      for (int l = 0; l < nl; ++l) sb[l] = c.skip_b + (size_t)l * H;   // unloaded layers still hold zeros
      if (launch_sum_bias(sb[0], sb[1], sb[2], sb[3], c.skip.bias, H, st)) return 1;
    } else if (load_conv(ly.rs, w, f->scale_scratch, bias)) {   // one conv, the epilogue splits at column H
      return 1;
    }
  } else if (rest == "enc.cond_layer" && gin > 0) {
    const int rows = nl * 2 * H;
    VD_CHECK(rows <= 8192, "flow: too many conditioning channels");
    if (launch_wn_scale(w, wg, f->scale_scratch, rows, gin, st)) return 1;
    flow_fold_f32_kernel<<<grid1d((long)rows * gin), 256, 0, st>>>(w, f->scale_scratch, c.cond_w, rows, gin);
    VD_CUDA(cudaGetLastError());
    VD_CUDA(cudaMemcpyAsync(c.cond_b, bias, (size_t)rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else {
    set_error(std::string("vitsdec_flow_load_layer: unknown layer ") + name);
    return 1;
  }
  c.loaded[rest] = true;
  return 0;
}

int vitsdec_flow_set_option(vitsdec_flow* f, const char* key, int value) {
  VD_CHECK(f && key, "vitsdec_flow_set_option: null argument");
  std::lock_guard<std::mutex> lock(f->mu);
  if (!strcmp(key, "fp16")) {
    const int v = value ? 1 : 0;
    if (v != f->fp16) {   // packed weights of the other 16-bit format are useless: every layer must be loaded again
      f->fp16 = v;
      for (FlowCoupling& c : f->cpl) c.loaded.clear();
      f->plans.clear();
    }
    return 0;
  }
  if (!strcmp(key, "pdl")) {
    if ((value != 0) != (f->pdl != 0)) f->plans.clear();
    f->pdl = value ? 1 : 0;
    return 0;
  }
  set_error(std::string("unknown option ") + key);
  return 1;
}

size_t vitsdec_flow_workspace_bytes(const vitsdec_flow* f, int batch, int frames) {
  if (!f || batch <= 0 || frames <= 0) return 0;
  return flow_ws(f, batch, frames).total;
}

int vitsdec_flow_apply(vitsdec_flow* f, const float* x, int64_t xsb, int64_t xsc, const float* x_mask, const float* g,
                       float* out, int B, int T, int reverse, void* ws, size_t ws_bytes, void* stream) {
  VD_CHECK(f && x && out && ws, "vitsdec_flow_apply: null argument");
  VD_CHECK(B > 0 && T > 0 && B <= 65535, "vitsdec_flow_apply: bad batch / frames");
  VD_CHECK((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "workspace must be 256-byte aligned");
  VD_CHECK(g == nullptr || f->hp.gin_channels > 0, "g given but the flow was built with gin_channels=0");
  for (size_t i = 0; i < f->names.size(); ++i) {
    int fi = 0, pos = 0;
    sscanf(f->names[i].c_str(), "flows.%d.%n", &fi, &pos);
    // without g the conditioning layer is never evaluated (modules.py:153), so its weights are not needed
    if (!g && f->names[i].substr(pos) == "enc.cond_layer") continue;
    VD_CHECK(f->cpl[fi / 2].loaded.count(f->names[i].substr(pos)),
             "vitsdec_flow_apply: layer " + f->names[i] + " has no weights loaded");
  }
  FlowDeviceGuard guard(f->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const FlowWs w = flow_ws(f, B, T);
  VD_CHECK(ws_bytes >= w.total, "vitsdec_flow_apply: workspace too small");
  uint8_t* base = static_cast<uint8_t*>(ws);
  std::shared_ptr<FlowPlan> plan;
  {
    std::lock_guard<std::mutex> lock(f->mu);
    const auto key = std::make_tuple(B, T, (const void*)ws);
    for (auto it = f->plans.begin(); it != f->plans.end(); ++it)
      if (it->first == key) { plan = it->second; f->plans.splice(f->plans.begin(), f->plans, it); break; }
    if (!plan) {
      plan = std::make_shared<FlowPlan>();
      if (flow_build_plan(f, *plan, B, T, base)) return 1;
      f->plans.emplace_front(key, plan);
      if (f->plans.size() > 32) f->plans.pop_back();
    }
  }
  const int C = f->hp.channels, H = f->hp.hidden_channels, nl = f->hp.n_layers, nf = f->hp.n_flows;
  const long rows = (long)B * T;
  float* X = reinterpret_cast<float*>(base + w.x);
  bf16* X0 = reinterpret_cast<bf16*>(base + w.x0);
  float* M = reinterpret_cast<float*>(base + w.m);
  float* CB = reinterpret_cast<float*>(base + w.cb);
  float* mask = reinterpret_cast<float*>(base + w.mask);
  float* gws = reinterpret_cast<float*>(base + w.g);
  // per-call inputs are staged into the workspace so that everything in between is pointer-stable (graph replay)
  if (x_mask) VD_CUDA(cudaMemcpyAsync(mask, x_mask, (size_t)rows * 4, cudaMemcpyDeviceToDevice, st));
  else flow_fill_mask_kernel<<<grid1d(rows), 256, 0, st>>>(mask, rows);
  if (g) VD_CUDA(cudaMemcpyAsync(gws, g, (size_t)B * f->hp.gin_channels * 4, cudaMemcpyDeviceToDevice, st));
  {
    // (+ the Flip in front of the first coupling of a reverse pass, and that coupling's 16-bit operand)
    dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
    flow_in_kernel<<<grid, block, 0, st>>>(x, xsb, xsc, X, X0, C, T, reverse ? 1 : 0, f->fp16);
  }
  VD_CUDA(cudaGetLastError());
  // reverse: Flip, coupling n-1, Flip, coupling n-2, ...   forward: coupling 0, Flip, coupling 1, Flip, ...
  auto enqueue = [&](cudaStream_t qs) -> int {
    if (g) {   // the conditioning of every coupling: one launch, off the per-coupling chain
      FlowCondPtrs ptrs{};
      for (int i = 0; i < nf; ++i) { ptrs.w[i] = f->cpl[i].cond_w; ptrs.b[i] = f->cpl[i].cond_b; }
      dim3 grid((nl * 2 * H * 32 + 255) / 256, B, nf);
      flow_cond_kernel<<<grid, 256, 0, qs>>>(ptrs, gws, CB, B, 2 * H, nl, f->hp.gin_channels);
      VD_CUDA(cudaGetLastError());
    }
    for (int step = 0; step < nf; ++step) {
      const int ci = reverse ? nf - 1 - step : step;
      std::vector<FlowStep>& steps = plan->steps[ci];
      size_t si = 0;
      auto run = [&](bool with_cond) -> int {
        FlowStep s = steps[si++];   // copy: per-call fields, re-entrant across threads
        if (!with_cond) s.ep.bias_b = nullptr;
        return launch_conv_tc(s.tc, s.ep, qs);
      };
      if (run(false)) return 1;                               // pre
      for (int l = 0; l < nl; ++l) {
        if (run(g != nullptr)) return 1;                      // in_layer (+ cond) with the gate in its epilogue
        if (!f->skipsum || f->cpl[ci].layers[l].has_res)
          if (run(false)) return 1;                           // res_skip: residual rows (skipsum) or split epilogue
      }
      if (f->skipsum && run(false)) return 1;                 // the WN's skip sum: one multi-segment launch
      if (run(false)) return 1;                               // post -> M
      if (step + 1 < nf) {   // couple, the Flip in front of the next coupling and its operand X0 in one pass
        flow_couple_flip_split_kernel<<<grid1d(rows * (C / 2)), 256, 0, qs>>>(X, M, mask, X0, rows, C, reverse ? 1 : 0,
                                                                             f->fp16);
        VD_CUDA(cudaGetLastError());
      }                      // (the last coupling's couple happens inside the output transpose)
    }
    return 0;
  };
  bool launched = false;
  if (!plan->graph_failed && ++plan->uses >= 3) {
    std::lock_guard<std::mutex> lock(f->mu);
    cudaGraphExec_t& exec = plan->graph_exec[(reverse ? 2 : 0) + (g ? 1 : 0)];
    if (!exec) {
      if (!f->cstream) VD_CUDA(cudaStreamCreateWithFlags(&f->cstream, cudaStreamNonBlocking));
      cudaGraph_t graph = nullptr;
      bool ok = cudaStreamBeginCapture(f->cstream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) {
        const int rc = enqueue(f->cstream);
        ok = cudaStreamEndCapture(f->cstream, &graph) == cudaSuccess && rc == 0 && graph != nullptr;
      }
      if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
      if (graph) cudaGraphDestroy(graph);
      if (!ok) {
        cudaGetLastError();
        exec = nullptr;
        plan->graph_failed = true;  // plain launches for this plan from now on
      }
    }
    if (exec) {
      VD_CUDA(cudaGraphLaunch(exec, st));
      launched = true;
    }
  }
  if (!launched && enqueue(st)) return 1;
  {
    dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
    flow_out_kernel<<<grid, block, 0, st>>>(X, M, mask, out, C, T, reverse ? 0 : 1, reverse ? 1 : 0);   // forward ends with a Flip
  }
  VD_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
