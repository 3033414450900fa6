/* vitsdec -- C ABI of the B200-native VITS waveform decoder (HiFi-GAN Generator).
 *
 * Drop-in boundary for ONE path of MedivhJin01/Personalized_Text-to-Speech:
 *     models.Generator.forward(x, g)          /root/reference/models.py:270-289
 * as called by SynthesizerTrn.infer            /root/reference/models.py:522
 *          and SynthesizerTrn.voice_conversion /root/reference/models.py:532
 *
 * The reference is Python; its "FFI" for this path is the nn.Module protocol.  The host-side mirror
 * (personalized_text-to-speech_b200/generator.py) binds these entry points with ctypes and passes
 * raw device pointers (tensor.data_ptr()) and the current CUDA stream.  No torch types cross this
 * boundary.  INTEGRATION.md shows the reference-side stub.
 *
 * Conventions: every function returns 0 on success, non-zero on error; vitsdec_last_error() returns a
 * thread-local message.  All pointers named *_dev are device pointers on the decoder's device; the
 * library never synchronises the stream except in the *_host convenience entry.  There is no CPU
 * fallback: on a machine without an sm_100 GPU vitsdec_create fails.
 */
#ifndef VITSDEC_H_
#define VITSDEC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VITSDEC_API __attribute__((visibility("default")))
#else
#define VITSDEC_API
#endif

#define VITSDEC_ABI_VERSION 1
#define VITSDEC_MAX_UPSAMPLES 8
#define VITSDEC_MAX_KERNELS 8
#define VITSDEC_MAX_DILATIONS 8

/* Generator.__init__ arguments (models.py:245; values from configs/<name>.json "model" block) */
typedef struct vitsdec_hparams {
  int32_t initial_channel;                                   /* inter_channels (192) */
  int32_t resblock;                                          /* 1 -> ResBlock1 ('1'), 2 -> ResBlock2 */
  int32_t num_kernels;                                       /* len(resblock_kernel_sizes) */
  int32_t resblock_kernel_sizes[VITSDEC_MAX_KERNELS];        /* [3,7,11] */
  int32_t num_dilations[VITSDEC_MAX_KERNELS];                /* len(resblock_dilation_sizes[j]) */
  int32_t resblock_dilation_sizes[VITSDEC_MAX_KERNELS][VITSDEC_MAX_DILATIONS]; /* [[1,3,5]]*3 */
  int32_t num_upsamples;                                     /* len(upsample_rates) */
  int32_t upsample_rates[VITSDEC_MAX_UPSAMPLES];             /* [8,8,2,2] */
  int32_t upsample_initial_channel;                          /* 512 */
  int32_t upsample_kernel_sizes[VITSDEC_MAX_UPSAMPLES];      /* [16,16,4,4] */
  int32_t gin_channels;                                      /* 256, 0 = no speaker conditioning */
} vitsdec_hparams;

typedef struct vitsdec_decoder vitsdec_decoder;

VITSDEC_API int vitsdec_abi_version(void);
VITSDEC_API const char* vitsdec_last_error(void);

/* replaces Generator.__init__ (models.py:245-268).  `device` is a CUDA ordinal. */
VITSDEC_API int vitsdec_create(const vitsdec_hparams* hp, int device, vitsdec_decoder** out);
VITSDEC_API void vitsdec_destroy(vitsdec_decoder* dec);

/* Number of layers that take weights, and the state_dict prefix of layer i ("conv_pre", "ups.0",
 * "resblocks.0.convs1.0", ..., "conv_post", "cond").  Mirrors the parameter tree of models.py:249-268. */
VITSDEC_API int vitsdec_num_layers(const vitsdec_decoder* dec);
VITSDEC_API const char* vitsdec_layer_name(const vitsdec_decoder* dec, int layer);

/* replaces load_state_dict + the weight_norm pre-forward hook (torch.nn.utils.weight_norm, dim=0;
 * models.py:254, modules.py:191-206).  fp32 device pointers in the reference's native layouts:
 *   Conv1d           weight [C_out, C_in, k]      ConvTranspose1d  weight [C_in, C_out, k]
 * weight_g_dev == NULL  -> `weight_dev` is the plain (already folded) weight;
 * weight_g_dev != NULL  -> `weight_dev` is weight_v and the fold w = v * g / ||v|| (norm over all dims
 *                          but 0) happens on the GPU, once, here.
 * bias_dev may be NULL (conv_post has no bias).  Enqueued on `stream`. */
VITSDEC_API int vitsdec_load_layer(vitsdec_decoder* dec, const char* name, const float* weight_dev, const float* weight_g_dev,
                       const float* bias_dev, void* stream);

/* Scratch the caller must provide to vitsdec_decode for a [batch, C, frames] latent. */
VITSDEC_API size_t vitsdec_workspace_bytes(const vitsdec_decoder* dec, int batch, int frames);

/* replaces Generator.forward(x, g) (models.py:270-289).
 *   z_dev   fp32 [batch, initial_channel, frames]; element strides z_stride_b / z_stride_c, time stride 1
 *           (infer passes the slice (z*y_mask)[:,:,:max_len], models.py:522)
 *   g_dev   fp32 [batch, gin_channels] (the [B, gin, 1] tensor of models.py:502) or NULL
 *   out_dev fp32 [batch, 1, frames * prod(upsample_rates)] contiguous
 * Asynchronous on `stream`; `workspace_dev` must stay untouched until the stream reaches this point. */
VITSDEC_API int vitsdec_decode(vitsdec_decoder* dec, const float* z_dev, int64_t z_stride_b, int64_t z_stride_c,
                   const float* g_dev, float* out_dev, int batch, int frames, void* workspace_dev,
                   size_t workspace_bytes, void* stream);

/* End-to-end convenience for hosts without a device allocator: pageable or pinned HOST buffers in and
 * out (contiguous), H2D + decode + D2H on an internal stream, synchronous. */
VITSDEC_API int vitsdec_decode_host(vitsdec_decoder* dec, const float* z_host, const float* g_host, float* out_host, int batch,
                        int frames);

/* Output side of cmd_inference.py:114-117 / VC_inference.py:49-51 (`.cpu().float().numpy()` -> wavfile.write): convert a
 * decoded waveform (fp32, [-1, 1], `samples` contiguous values) to 16-bit PCM on the device --
 * round-to-nearest(clip(x, -1, 1) * 32767) -- so that the D2H copy moves half the bytes and can land directly behind the
 * WAV header of a pinned host buffer (personalized_text-to-speech_b200/wavout.py).  Asynchronous on `stream`. */
VITSDEC_API int vitsdec_wav_pcm16(int device, const float* wav_dev, int16_t* pcm_dev, int64_t samples, void* stream);

/* Options: "impl" = 0 tcgen05 tensor-core kernels (the only backend of libvitsdec.so); 1 = CUDA-core cross-check kernels,
 *          accepted by the test build libvitsdec_test.so only (build.py: -DVITSDEC_TESTING);
 *          "desc_mode" = debug / experiment knobs (0 is the product setting): bit 0 UMMA descriptor base
 *          offset, bit 2 (4) TMA L2 prefetch of residual tiles (on in round 1, measured 1 % slower since), bit 11 (2048) LSU
 *          instead of TMA output stores in the channels-as-M epilogues, bit 12 (4096) no paired
 *          (tcgen05.mma.cta_group::2) tiles for 256-channel layers, bit 13 (8192) relay variant of their operand barriers;
 *          "debug_keep" = 1 keep named intermediates for vitsdec_debug_read; "profile" = 1 see below;
 *          "fuse_pairs" = 0 run every ResBlock conv as its own launch (default 1: fused pairs where they fit);
 *          "graph" = 0 plain kernel launches (default 1: the conv steps of a plan replay as one CUDA graph from the
 *          plan's third use on; 2: capture at the first use);
 *          "fold" = 0 keeps narrow layers on plain tiles (default 1: time-folded);
 *          "pairf" = 1 (default) runs every non-final ResBlock pair of a 128-channel stage and the dilation-1, k >= 7
 *          pairs of a 64-channel stage (2-sample folded view) through conv_pairf.cu (both convs as 128-virtual-channel
 *          channels-as-M tiles, h kept in shared memory, two tiles in flight); 2: every pair that kernel supports, also
 *          the C = 32 form (tests / experiments); 3: only the k <= 5 pairs of a 128-channel stage (A/B); 0: never;
 *          "mrfp": bit 0 (C = 32 stage): every ResBlock pair on the 2-sample folded view, and the last pair of every MRF
 *          branch + the branch sum + the average as ONE launch (conv_mrfp.cu) instead of three c1 launches and a fused-MRF
 *          launch; bit 1: the C = 64 pairs whose weights fit (k <= 7) through the same kernel on plain rows instead of
 *          conv_pair.cu (measured neutral: 9.05 vs 9.07 ms per step, so off by default); bit 2: the same last-pairs + MRF
 *          fusion for a 128-channel stage (conv_mrf128.cu, streamed weights); bit 3: also the k < 9 pairs of the C = 32
 *          stage through conv_mrfp.cu (by default they run on conv_pair.cu, which is faster for short kernels).  Default 5;
 *          0 = the round-1 schedule;
 *          "par" = 0 serial MRF branches (default 1: the branches of a stage run concurrently under the graph);
 *          "pdl" = 0 no programmatic dependent launch (default 1: launches whose grid leaves SMs idle let the next launch
 *          of the stream start its prologue early; 2: every launch);
 *          "fp16" = 1 packs weights and stores activations as IEEE fp16 instead of bf16 (default 0).  Same tensor-core
 *          rate (tcgen05 kind::f16), fp32 accumulation, 10 instead of 7 stored mantissa bits; stored activations saturate
 *          at +-65504.  Changing it invalidates the loaded weights: every vitsdec_load_layer must be repeated before the
 *          next decode (vitsdec_decode fails with "no weights loaded" otherwise).
 * Options are configuration, not per-call arguments: set them while no decode of this decoder is in flight (a schedule
 * option drops the cached plans; "fp16" rewrites the packed weights a running decode would still be reading). */
VITSDEC_API int vitsdec_set_option(vitsdec_decoder* dec, const char* key, int value);
/* get_option also reads: "hop", "num_sms", "testing_build", and "graph_failed" = number of launch plans whose CUDA-graph
 * capture or instantiation failed (those plans run as plain launches; results are identical, batch-1 latency is not). */
VITSDEC_API int vitsdec_get_option(const vitsdec_decoder* dec, const char* key, int* value);

/* Option "profile"=1 brackets the convolution launches (the tcgen05 kernel) of every decode with CUDA
 * events on the launching stream; this returns the accumulated device time and launch count since the
 * option was set (synchronises on the last event only).  Used by bench.py's roofline figure. */
VITSDEC_API int vitsdec_profile_read(vitsdec_decoder* dec, double* conv_ms, int64_t* conv_launches);

/* Kernel launches issued by the last vitsdec_decode on this decoder (for bench.py's gpu_launches). */
VITSDEC_API int vitsdec_last_launch_count(const vitsdec_decoder* dec);

/* Test hook: copy a named intermediate of the LAST decode out of the workspace as fp32 NCL
 * [batch, C, L].  Names: "conv_pre", "ups.<i>", "mrf.<i>".  Values are the stored a-form
 * (post-leaky-relu) mapped back to the residual-stream value.  out_elems = capacity of out_dev. */
VITSDEC_API int vitsdec_debug_read(vitsdec_decoder* dec, const char* name, float* out_dev, size_t out_elems, int* channels,
                       int* length, void* stream);

/* Profiling hook for vitsdec_op_*: device buffer of 256 x 12 uint64 that CTA 0 of the tcgen05 kernel fills with
 * clock64() stamps per tile (producer / MMA issuer / epilogue hand-offs); NULL disables.  tools/trace_probe.py. */
VITSDEC_API int vitsdec_debug_set_trace(void* trace_dev);

/* Single fused convolution on channels-last bf16 activations (per-kernel parity tests):
 *   y[b,t,co] = lrelu( bias[co] + sum_j sum_ci w[co,ci,j] * x[b, t + (j-(k-1)/2)*dilation, ci] (+ res), out_slope )
 *   x_dev / res_dev / y_dev: bf16 [batch, length, channels]; w_dev fp32 [c_out, c_in, k]; bias fp32 [c_out].
 *   res (optional) is stored post-leaky-relu with slope 1/res_gain.  impl: 0 tcgen05, 1 CUDA cores.
 *   desc_mode bit 1024: x / res / y are fp16 instead of bf16 (the decoder's "fp16" option); other bits: test knobs. */
VITSDEC_API int vitsdec_op_conv1d(int device, const void* x_dev, const float* w_dev, const float* bias_dev, const void* res_dev,
                      float res_gain, float out_slope, void* y_dev, int batch, int length, int c_in, int c_out,
                      int k, int dilation, int impl, int desc_mode, void* stream);

/* One fused ResBlock1 iteration (modules.py:211-221) on channels-last bf16 a-form activations:
 *   y = lrelu( conv(k,1)( lrelu( conv(k,dilation)(x) + b1 ) ) + b2 + unlrelu(x), slope ),  x stored as lrelu(., slope).
 *   w1/w2 fp32 [C, C, k].  Only shapes for which the fused kernel exists (C in {32, 64}). */
VITSDEC_API int vitsdec_op_resblock_pair(int device, const void* x_dev, const float* w1_dev, const float* b1_dev,
                                         const float* w2_dev, const float* b2_dev, void* y_dev, int batch, int length,
                                         int channels, int k, int dilation, float slope, void* stream);

/* The same ResBlock1 iteration through the time-folded fused kernel (conv_pairf.cu: both convs as 128-virtual-channel
 * channels-as-M tiles, weights streamed).  length must be a multiple of 128 / channels. */
VITSDEC_API int vitsdec_op_resblock_pair_folded(int device, const void* x_dev, const float* w1_dev, const float* b1_dev,
                                                const float* w2_dev, const float* b2_dev, void* y_dev, int batch,
                                                int length, int channels, int k, int dilation, float slope,
                                                void* stream);

/* The fused narrow-stage kernel (conv_mrfp.cu) on its own: nbr ResBlock1 iterations whose results are summed and averaged,
 *   y = lrelu( ( sum_j  c2_j( lrelu( c1_j(x_j) + b1_j ) ) + b2_j + unlrelu(x_j) ) / nbr , out_slope )
 * nbr = 1 is one ResBlock1 loop iteration (modules.py:211-221); nbr = 3 is the last iteration of the three MRF branches plus
 * the branch average (models.py:278-284).  xs[j]: bf16 [batch, length, C] a-form (slope `slope`), C = 32 (length even) or
 * C = 64 (as long as both weight sets fit in shared memory: k <= 7 for one branch); w1[j], w2[j]: fp32 [C, C, k[j]]; c1_j
 * has dilation[j], c2_j dilation 1.  Arrays of nbr host pointers to device data. */
VITSDEC_API int vitsdec_op_mrf_pairs(int device, int nbr, const void* const* xs, const float* const* w1, const float* const* b1,
                                     const float* const* w2, const float* const* b2, void* y_dev, int batch, int length,
                                     int channels, const int* k, const int* dilation, float slope, float out_slope,
                                     void* stream);

/* Same for ConvTranspose1d(c_in, c_out, k, stride, padding=(k-stride)/2) in polyphase form:
 *   w_dev fp32 [c_in, c_out, k]; y_dev bf16 [batch, length*stride, c_out]. */
VITSDEC_API int vitsdec_op_conv_transpose1d(int device, const void* x_dev, const float* w_dev, const float* bias_dev,
                                float out_slope, void* y_dev, int batch, int length, int c_in, int c_out, int k,
                                int stride, int impl, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * The normalising flow that produces the decoder's latent: ResidualCouplingBlock.forward(x, x_mask, g, reverse)
 * (/root/reference/models.py:179-209; SynthesizerTrn.infer calls it in reverse at models.py:521, voice_conversion in
 * both directions at models.py:530-531).  Constructor arguments as models.py:449 passes them. */
typedef struct vitsdec_flow vitsdec_flow;
typedef struct vitsdec_flow_hparams {
  int32_t channels;          /* inter_channels (192) */
  int32_t hidden_channels;   /* 192 */
  int32_t kernel_size;       /* 5 */
  int32_t dilation_rate;     /* 1 */
  int32_t n_layers;          /* 4 */
  int32_t n_flows;           /* 4 */
  int32_t gin_channels;      /* 0 = no speaker conditioning */
} vitsdec_flow_hparams;

VITSDEC_API int vitsdec_flow_create(const vitsdec_flow_hparams* hp, int device, vitsdec_flow** out);
VITSDEC_API void vitsdec_flow_destroy(vitsdec_flow* flow);
/* Parametrised sub-modules in state_dict order: "flows.<2i>.pre", "flows.<2i>.enc.in_layers.<l>",
 * "flows.<2i>.enc.res_skip_layers.<l>", "flows.<2i>.enc.cond_layer" (gin_channels > 0), "flows.<2i>.post". */
VITSDEC_API int vitsdec_flow_num_layers(const vitsdec_flow* flow);
VITSDEC_API const char* vitsdec_flow_layer_name(const vitsdec_flow* flow, int index);
/* Device fp32 tensors in the reference's shapes: weight (or weight_v), weight_g (NULL for pre / post, which carry no
 * weight norm, modules.py:318-320) and bias.  Weight norm (dim 0) is folded here, once. */
VITSDEC_API int vitsdec_flow_load_layer(vitsdec_flow* flow, const char* name, const float* w_dev, const float* wg_dev,
                                        const float* bias_dev, void* stream);
/* Options: "fp16" = 1 conv operands and stored activations are fp16 instead of bf16 (as the decoder's option; the latent
 * itself stays fp32 either way).  Changing it invalidates the loaded weights: load every layer again.
 * "pdl" = 0 switches programmatic dependent launch off (default 1, as for the decoder). */
VITSDEC_API int vitsdec_flow_set_option(vitsdec_flow* flow, const char* key, int value);
VITSDEC_API size_t vitsdec_flow_workspace_bytes(const vitsdec_flow* flow, int batch, int frames);
/* x_dev: fp32 [batch, channels, frames] with element strides (x_stride_b, x_stride_c, 1); x_mask_dev: fp32
 * [batch, frames] (the reference's x_mask [B,1,T], binary: commons.sequence_mask) or NULL = all ones; g_dev: fp32
 * [batch, gin_channels] or NULL; out_dev: contiguous fp32 [batch, channels, frames].  reverse != 0 runs the inverse
 * pass (models.py:207-209).  Asynchronous on `stream`. */
VITSDEC_API int vitsdec_flow_apply(vitsdec_flow* flow, const float* x_dev, int64_t x_stride_b, int64_t x_stride_c,
                                   const float* x_mask_dev, const float* g_dev, float* out_dev, int batch, int frames,
                                   int reverse, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITSDEC_H_ */
