/* audiopure.h -- C ABI of libaudiopure_b200.so: the B200-native (sm_100a) AudioPure
 * purification-and-classify hot path.
 *
 * The reference (cychomatica/Diffusion-Model-for-Audio-Defense) has no FFI layer: its "operator API" is a set of
 * duck-typed torch.nn.Modules.  Each entry point below is what a binding for ONE of those reference functions
 * would call; the reference interface it replaces is cited as file:line (paths relative to the reference root).
 * The Python host mirror (package audiopure_b200) binds these with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every `const float*` / `float*` below marked "device" is caller-owned GPU memory
 *    (fp32, contiguous); "host" pointers are CPU memory read during the call only.
 *  - waveforms are (B, 1, L) fp32 contiguous == (B, L); spectrograms (B, 1, n_mels, frames); logits (B, K).
 *  - all device work is enqueued on `stream` (a cudaStream_t passed as void*); no host synchronisation inside
 *    except in *_create / *_reserve.  Handles are not thread-safe: one per (device, stream).
 *  - return value: 0 = ok, negative = error (AP_ERR_*); ap_last_error() gives the message (thread-local).
 *  - there is NO CPU fallback: on a machine without an sm_100 GPU every compute entry point returns AP_ERR_CUDA.
 */
#ifndef AUDIOPURE_H_
#define AUDIOPURE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AP_OK 0
#define AP_ERR_INVALID (-1) /* bad argument / unsupported configuration */
#define AP_ERR_CUDA (-2)    /* CUDA runtime / driver error, or no usable GPU */
#define AP_ERR_STATE (-3)   /* handle not in a state that allows the call */

/* arithmetic mode of the DiffWave network */
#define AP_MODE_BF16 0 /* tcgen05 tensor cores: bf16 operands, fp32 accumulate (TMEM), fp32 x/update arithmetic */
#define AP_MODE_FP32 1 /* fp32 FFMA path (parity mode, <=1e-5 rel-L2 vs the reference) */
#define AP_MODE_FP16 3 /* DiffWave only: the same tcgen05 kernels with fp16 operands (11-bit mantissa: ~8x smaller eps error than
                          bf16 at the same speed; conversions saturate at +-65504) */
#define AP_MODE_BF16X3 4 /* DiffWave only: fp32-class arithmetic on the bf16 tensor cores.  Activations and weights are kept as
                            bf16 hi/lo plane pairs (16 significand bits) and every product is accumulated as
                            hi*hi + lo*hi + hi*lo in fp32: three MMAs per K step, ~1e-5 eps error, <=1e-5 on the purified
                            waveform (the fp32-mode bar) at several times the speed of the FFMA path */
#define AP_MODE_TF32 2 /* classifiers only: tcgen05 kind::tf32 convolutions (the precision of the reference's cuDNN path) */

typedef struct ap_diffwave_s* ap_diffwave_t;
typedef struct ap_mel_s* ap_mel_t;
typedef struct ap_classifier_s* ap_classifier_t;

const char* ap_last_error(void);
int ap_version(void);
/* number of kernels this library has launched since load (process-wide, all handles); bench.py's gpu_launches */
unsigned long long ap_launch_count(void);
/* changes whenever the library allocates or frees device memory (workspaces grow with the batch, the sequence length and the
 * arithmetic mode): a CUDA graph captured around these entry points holds raw workspace pointers and must be re-captured when
 * the value differs from the one read right after the capture */
unsigned long long ap_alloc_generation(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Host helpers (no GPU needed)
 * ------------------------------------------------------------------------------------------------------------- */
/* w[o,:] = g[o] * v[o,:] / ||v[o,:]||_2 -- torch.nn.utils.weight_norm fold of WaveNet.py:27-28,66-73 (host, fp32 in/out,
 * fp64 accumulate). */
int ap_fold_weight_norm(const float* g, const float* v, float* w, int cout, int fan_in);

/* ---------------------------------------------------------------------------------------------------------------
 * DiffWave network (replaces WaveNet_Speech_Commands.forward, DiffWave_Unconditional/WaveNet.py:138-172)
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  int in_channels;   /* 1 */
  int res_channels;  /* 256 */
  int skip_channels; /* 256 (must equal res_channels) */
  int out_channels;  /* 1 */
  int num_res_layers;
  int dilation_cycle;
  int embed_dim_in, embed_dim_mid, embed_dim_out; /* 128, 512, 512 */
} ap_wavenet_cfg; /* == configs/config.json "wavenet_config" */

/* weights: host fp32 pointers, weight-norm already folded, in this order (n = layer index):
 *   [0] init_w (C)            [1] init_b (C)
 *   [2] fc_t1_w (mid x in)    [3] fc_t1_b (mid)     [4] fc_t2_w (out x mid)   [5] fc_t2_b (out)
 *   [6+8n+0] fc_t_w (C x out) [6+8n+1] fc_t_b (C)
 *   [6+8n+2] dil_w (2C x C x 3, torch Conv1d layout)  [6+8n+3] dil_b (2C)
 *   [6+8n+4] res_w (C x C)    [6+8n+5] res_b (C)    [6+8n+6] skip_w (S x C)   [6+8n+7] skip_b (S)
 *   [6+8N+0] final1_w (S x S) [6+8N+1] final1_b (S) [6+8N+2] final2_w (S)     [6+8N+3] final2_b (1)
 * n_weights must be 6 + 8*num_res_layers + 4.   Replaces create_diffwave_model (diffwave_ddpm.py:395-411). */
int ap_diffwave_create(ap_diffwave_t* out, const ap_wavenet_cfg* cfg, const float* const* weights, int n_weights,
                       int device);
void ap_diffwave_destroy(ap_diffwave_t h);
/* AP_MODE_BF16 (default when the configuration supports the tensor-core kernels: C == S == 256), AP_MODE_FP16,
 * AP_MODE_BF16X3 or AP_MODE_FP32 */
int ap_diffwave_set_mode(ap_diffwave_t h, int mode);
int ap_diffwave_get_mode(ap_diffwave_t h);
/* Pre-allocate the activation workspace for up to `chunk` waveforms of length L processed at once (larger batches are
 * processed in chunks).  Called implicitly (with a default chunk) by the first compute call if omitted. */
int ap_diffwave_reserve(ap_diffwave_t h, int chunk, int L);

/* eps = eps_theta(x, t): WaveNet forward with diffusion_steps == t for every row
 * (DiffWave.compute_eps_t, diffwave_ddpm.py:166-172).  x, eps: device (B, L). */
int ap_diffwave_eps(ap_diffwave_t h, const float* x, float t, float* eps, int B, int L, void* stream);

/* Vector-Jacobian product of the network wrt its input: g_x = (d eps_theta(x, t) / d x)^T g_eps -- what autograd computes
 * through WaveNet_Speech_Commands.forward (WaveNet.py:164-172) when a white-box attack back-propagates through the purifier
 * (DiffWave.forward is differentiable, diffwave_ddpm.py:36-47; robustness_eval/white_box_attack.py:438).  x, g_eps, g_x
 * (and eps_out, optional: the network output at x) are device fp32 (B, L).  AP_MODE_BF16 or AP_MODE_BF16X3 (the forward
 * in the handle's mode, the backward GEMMs with bf16 operands); the forward is recomputed with the gate's two local
 * derivatives of every layer kept (~0.6 GB per 1 s waveform), in sub-batches bounded to ~24 GB.  The head's ReLU makes the
 * gradient discontinuous in the forward values: with a bf16 forward ~1 % of its masks differ from an fp32 forward's. */
int ap_diffwave_eps_vjp(ap_diffwave_t h, const float* x, float t, const float* g_eps, float* g_x, float* eps_out, int B,
                        int L, void* stream);
/* The same in two halves, for callers that know at forward time that a backward will follow (autograd): a forward that keeps
 * the backward's inputs and returns a token, and the backward from that state.  Any later saving forward on the handle
 * (including the recomputation inside ap_diffwave_eps_vjp) overwrites the state: a stale token gives AP_ERR_STATE and the
 * caller falls back to ap_diffwave_eps_vjp.  ap_diffwave_eps_save returns AP_ERR_STATE when B exceeds the ~24 GB sub-batch. */
int ap_diffwave_eps_save(ap_diffwave_t h, const float* x, float t, float* eps, int B, int L, void* stream,
                         unsigned long long* token);
int ap_diffwave_eps_vjp_saved(ap_diffwave_t h, unsigned long long token, const float* x, const float* g_eps, float* g_x, int B,
                              int L, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Per-step updates (HBM-bound elementwise kernels; in-kernel Philox4x32-10 + Box-Muller when z == NULL)
 * Noise element i of a call uses Philox counter (offset + i/4), key = seed; callers advance `offset` by
 * ceil(B*L/4) per call (ap_noise_offset_stride) so streams never overlap.
 * ------------------------------------------------------------------------------------------------------------- */
uint64_t ap_noise_offset_stride(int B, int L);
/* x_t = sqrt_ab * x0 + sqrt_1mab * z                               DiffWave._diffusion, diffwave_ddpm.py:49-73 */
int ap_diffuse(const float* x0, float sqrt_ab, float sqrt_1mab, const float* z_or_null, uint64_t seed, uint64_t offset,
               float* xt, int B, int L, void* stream);
/* x = (x - c_eps * eps) / sqrt_alpha + sigma * z   (sigma == 0: no noise read/drawn)
 *                                                                   compute_coefficients/_reverse, :95-103,:159-160 */
int ap_ddpm_step(float* x, const float* eps, float c_eps, float sqrt_alpha, float sigma, const float* z_or_null,
                 uint64_t seed, uint64_t offset, int B, int L, void* stream);
/* one Euler-Maruyama step of the reverse VP-SDE                      RevVPSDE.f/.g, diffwave_sde.py:73-133 (+ torchsde euler)
 *   drift = (-0.5*beta)*x ; score = -eps/sqrt_1mab ; f = -(drift - diff2*score) ; x = (x + f*dt) + g*(sqrt_dt*z)
 * The host builds one row per Euler step with the reference's own float32 arithmetic (diff2 = sqrt(beta)^2 in fp32). */
typedef struct {
  float beta, diff2, sqrt_1mab, dt, g, sqrt_dt;
} ap_sde_coef;
int ap_sde_step(float* x, const float* eps, const ap_sde_coef* c, const float* z_or_null, uint64_t seed,
                uint64_t offset, int B, int L, void* stream);
/* x0 = sqrt_recip_ab * xt - sqrt_recipm1_ab * eps                    _predict_x0_from_eps, diffwave_ddpm.py:195-205 */
int ap_predict_x0(const float* xt, const float* eps, float sqrt_recip_ab, float sqrt_recipm1_ab, float* x0, int B,
                  int L, void* stream);
/* out[b,:] = scale * (x[0,:] + sigma * z[b,:]) : the noisy, rescaled copies of ONE input that
 * RobustCertificate.smooth_predict builds (certified_robust.py:44-54).  x: device (L), out: device (B, L). */
int ap_smooth_inputs(const float* x, float sigma, float scale, const float* z_or_null, uint64_t seed, uint64_t offset,
                     float* out, int B, int L, void* stream);
/* The certification front end in ONE call (RobustCertificate.smooth_predict's input construction, certified_robust.py:44-54,
 * followed by DiffWave.one_shot_denoise, diffwave_ddpm.py:174-205): for b < B
 *     x_in[b,:] = scale * (x[:] + sigma * z[b,:]) ;  x0[b,:] = sqrt_recip_ab * x_in[b,:] - sqrt_recipm1_ab * eps_theta(x_in[b,:], t)
 * x: device (L), x0: device (B, L).  In the tensor-core modes with L % 4 == 0 the noisy copies are built inside the network's
 * first kernel and x0 is formed in the epilogue of its last one; otherwise the unfused kernels run (same values either way:
 * ap_smooth_inputs -> ap_diffwave_eps -> ap_predict_x0).  Noise element e = b * L + l uses Philox block offset + e / 4.
 * offset_dev: optional device word ADDED to `offset`, so that a captured CUDA graph of a micro-batch can be replayed with a
 * moving noise offset (advance it on the stream with ap_u64_add); NULL = `offset` alone. */
int ap_diffwave_smooth_denoise(ap_diffwave_t h, const float* x, float sigma, float scale, const float* z_or_null,
                               uint64_t seed, uint64_t offset, const uint64_t* offset_dev, float t, float sqrt_recip_ab,
                               float sqrt_recipm1_ab, float* x0, int B, int L, void* stream);
/* *dev += inc, enqueued on `stream` (one thread) */
int ap_u64_add(uint64_t* dev, uint64_t inc, void* stream);
/* standard-normal fill (diagnostics / statistical tests of the in-kernel RNG) */
int ap_randn(float* out, uint64_t n, uint64_t seed, uint64_t offset, void* stream);

/* Whole DDPM purifier in one call: diffuse to t* then t* ancestral steps (DiffWave.forward, diffwave_ddpm.py:36-47).
 * coef: host array of n_steps+1 rows {a, b, c, 0}: row 0 = {sqrt_ab, sqrt_1mab} for the diffusion; row 1+i (i-th reverse step,
 * t = t*-1-i) = {c_eps, sqrt_alpha, sigma, t}.  z_or_null: device (n_steps+1... see below) host-generated noise laid out
 * as (n_noise, B, L) with n_noise = t* (z_diffuse, z_{t*-1}, ..., z_1), or NULL for in-kernel Philox. */
int ap_diffwave_purify_ddpm(ap_diffwave_t h, const float* x0, float* out, int t_star, const float* coef4,
                            const float* z_or_null, uint64_t seed, uint64_t offset, int B, int L, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Log-mel front end (replaces torchaudio MelSpectrogram + AmplitudeToDB built at certified_robustness_eval.py:85-87,
 * kws_adaptive_attack_eval.py:74-76): DFT-as-GEMM -> power -> mel filterbank -> 10*log10(clamp(., 1e-10)).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  int sample_rate, n_fft, hop_length, n_mels;
  int slaney_norm; /* 1: norm='slaney', 0: None */
  int slaney_scale; /* 1: mel_scale='slaney', 0: 'htk' */
  int reflect_pad;  /* 1: pad_mode='reflect', 0: 'constant' (zeros) */
} ap_mel_cfg;
int ap_mel_create(ap_mel_t* out, const ap_mel_cfg* cfg, int device);
void ap_mel_destroy(ap_mel_t h);
int ap_mel_frames(ap_mel_t h, int L); /* 1 + L / hop */
/* wav: device (B, L); spec: device (B, n_mels, frames) */
int ap_mel_db(ap_mel_t h, const float* wav, float* spec, int B, int L, void* stream);
/* Vector-Jacobian product wrt the waveform: g_wav = (d spec / d wav)^T g_spec, the backward of torchaudio's
 * MelSpectrogram + AmplitudeToDB (power 2, 10 log10 with the 1e-10 clamp) as autograd computes it.  wav, g_wav: device
 * (B, L); g_spec: device (B, n_mels, frames). */
int ap_mel_vjp(ap_mel_t h, const float* wav, const float* g_spec, float* g_wav, int B, int L, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Classifiers (replace Classifier(spec) at acoustic_system.py:49 / certified_robust.py:30)
 * ------------------------------------------------------------------------------------------------------------- */
#define AP_CLS_RESNEXT 0 /* CifarResNeXt, models/resnext.py:67-142 : (B,1,32,32) -> (B,nlabels) logits   */
#define AP_CLS_M5 1      /* M5, audio_models/M5/M5Net.py:4-38       : (B,1,L)     -> (B,n) log-probs      */
#define AP_CLS_KWS 2     /* KWSModel, audio_models/RCNN_KWS/model.py:66-113 : (B,1,32,W) -> (B,4) log-probs */
#define AP_CLS_RESNET 3  /* ResNet-18/34/50/101/152, models/resnet.py:103-220 : (B,1,32,32) -> (B,num_classes) logits; `depth` selects */
#define AP_CLS_VGG 4     /* VGG-11/13/16/19 with batch norm, models/vgg.py:32-95 : (B,1,32,32) -> (B,num_classes) logits; `depth` selects */
#define AP_CLS_WRN 5     /* WideResNet-depth-widen_factor, models/wideresnet.py:15-92 : (B,1,32,32) -> (B,num_classes) logits */
#define AP_CLS_DENSENET 6 /* DenseNet-BC-depth-growth, models/densenet.py:15-147 : (B,1,32,32) -> (B,num_classes) logits;
                           * base_width = growthRate, widen_factor = compressionRate */
typedef struct {
  int kind;
  int num_classes;
  /* ResNeXt */
  int cardinality, depth, base_width, widen_factor, in_channels;
  /* M5 */
  int m5_first_kernel, m5_stride, m5_channels;
  /* KWS */
  int kws_in_size, kws_hidden;
} ap_classifier_cfg;
/* weights: host fp32 pointers in the reference module's state_dict order (num_batches_tracked entries skipped);
 * BatchNorm (eval) is folded into the preceding convolution at create. */
int ap_classifier_create(ap_classifier_t* out, const ap_classifier_cfg* cfg, const float* const* weights,
                         int n_weights, int device);
void ap_classifier_destroy(ap_classifier_t h);
/* input: device (B, 1, 32, 32) spectrogram (ResNeXt, ResNet; in_len = 32), (B, L) waveform (M5; in_len = L) or (B, 32, W) (KWS; in_len = W) */
int ap_classifier_forward(ap_classifier_t h, const float* input, float* logits, int B, int in_len, void* stream);
/* Vector-Jacobian product wrt the input: g_input = (d logits / d input)^T g_logits, what autograd computes through
 * CifarResNeXt.forward (models/resnext.py:134-142) or M5.forward (audio_models/M5/M5Net.py:21-38), BatchNorm in eval mode,
 * or KWSModel.forward (audio_models/RCNN_KWS/model.py:90-113), when an attack back-propagates the loss
 * (robustness_eval/white_box_attack.py:438).  Every classifier kind; the forward is recomputed with what the backward needs
 * kept (ReLU outputs / GRU gate values).  input, g_input: device (B, 1, 32, 32), (B, L) or (B, 32, W);
 * g_logits: device (B, num_classes). */
int ap_classifier_vjp(ap_classifier_t h, const float* input, const float* g_logits, float* g_input, int B, int in_len,
                      void* stream);
/* AP_MODE_TF32 (default for ResNeXt, ResNet, VGG and WideResNet: tensor-core convolutions where a tile shape exists, in the forward pass,
 * the recomputed forward of ap_classifier_vjp (AP_CLS_VJP_FWD_FP32=1 keeps that one on the fp32 path) and the data gradients) or AP_MODE_FP32 (every convolution on the FFMA path) */
int ap_classifier_set_mode(ap_classifier_t h, int mode);
int ap_classifier_get_mode(ap_classifier_t h);

/* ---------------------------------------------------------------------------------------------------------------
 * Spectrogram-domain purifier ("Diffusion-Spec"): the UNet eps-network of RevImprovedDiffusion
 * (diffusion_models/improved_diffusion_sde.py:140-226 builds it with create_model_and_diffusion; UNetModel.forward,
 * Improved_Diffusion_Unconditional/improved_diffusion/unet.py:462-497).  The reverse-SDE update itself is ap_sde_step.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct ap_unet_s* ap_unet_t;
typedef struct {
  int image_size;     /* 32 */
  int in_channels;    /* 1 */
  int model_channels; /* 128 */
  int out_channels;   /* 1 */
  int num_res_blocks; /* 3 */
  int num_heads;      /* 4 (64-channel heads) */
  int use_scale_shift_norm; /* 1 */
} ap_unet_cfg; /* == script_util.py model_and_diffusion_defaults / create_model */
enum { AP_UNET_CONV_IN = 0, AP_UNET_RES = 1, AP_UNET_ATTN = 2, AP_UNET_PUSH = 3, AP_UNET_POP = 4, AP_UNET_DOWN = 5, AP_UNET_UP = 6,
       AP_UNET_OUT = 7 };
/* ops: the module walk of UNetModel.__init__ (unet.py:341-443) as n_ops rows {kind, cin, cout} (PUSH / POP = the skip-connection
 * stack of forward(), :483-494; POP's cout = channels after th.cat).  weights: host fp32 pointers in the reference's state-dict
 * order (time_embed.0.weight, .bias, time_embed.2.*, then per op: conv weight / bias; ResBlock: in_layers.0.{weight,bias},
 * in_layers.2.{weight,bias}, emb_layers.1.{weight,bias}, out_layers.0.*, out_layers.3.*, [skip_connection.*]; AttentionBlock:
 * norm.*, qkv.*, proj_out.*; out.0.*, out.2.*). */
int ap_unet_create(ap_unet_t* out, const ap_unet_cfg* cfg, const int* ops, int n_ops, const float* const* weights,
                   int n_weights, int device);
void ap_unet_destroy(ap_unet_t h);
/* AP_MODE_TF32 (default: every convolution with a tensor-core tile shape on tcgen05 kind::tf32 -- the precision of the reference's
 * cuDNN path) or AP_MODE_FP32 (FFMA everywhere; the parity mode) */
int ap_unet_set_mode(ap_unet_t h, int mode);
/* eps = model(x, timesteps = t for every row).  x, eps: device fp32 (B, 1, image_size, image_size). */
int ap_unet_eps(ap_unet_t h, const float* x, float t, float* eps, int B, void* stream);
/* Vector-Jacobian product wrt the input: g_x = (d eps / d x)^T g_eps -- what autograd computes through UNetModel.forward when a
 * white-box attack back-propagates through RevImprovedDiffusion (the reference calls the model without no_grad,
 * improved_diffusion_sde.py:104-105).  The forward is recomputed with its operations recorded and walked in reverse (GroupNorm /
 * SiLU / attention / up-sampling backward kernels, data-gradient twins of every convolution).  eps_out: optional, the network
 * output at x.  All device fp32 (B, 1, image_size, image_size). */
int ap_unet_eps_vjp(ap_unet_t h, const float* x, float t, const float* g_eps, float* g_x, float* eps_out, int B, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Votes (replaces the argmax + per-class .sum().item() loop of smooth_predict, certified_robust.py:59-67)
 * counts: device int64[counts_len], ACCUMULATED (caller zeroes it); K = logits.shape[-1] must not exceed counts_len (the
 * reference sizes counts from output.shape[-1], certified_robust.py:60-63); ties resolve to the lowest class index like torch.max.
 * ------------------------------------------------------------------------------------------------------------- */
int ap_vote_counts(const float* logits, int B, int K, long long* counts, int counts_len, void* stream);
int ap_argmax(const float* logits, int B, int K, int* pred, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Black-box query serving: the NES gradient estimator FAKEBOB runs around the defended system
 * (robustness_eval/_NES.py:14-56, called from black_box_attack.py:186-190) and the per-query loss / decision of the
 * EOT wrapper (_EOT.py:39-42, _utils.py:113-125).  S = samples_per_draw_batch_size (even), H = S / 2, R = S + first.
 *
 * ap_nes_perturb:  out (A, R, L) <- the query batch of one NES draw (_NES.py:18-25):
 *     out[a, 0]             = x[a]                         (only when first = 1: the un-noised query of draw 0)
 *     out[a, first + j]     = z[a, j] * sigma + x[a]       j < H
 *     out[a, first + H + j] = (-z[a, j]) * sigma + x[a]
 *   z_or_null: device (A, H, L) noise, or NULL for Philox normals (element e = lane e % 4 of block offset + e / 4;
 *   a draw consumes ap_nes_noise_blocks(A, S, L) blocks).
 * ap_nes_gradient: grad (A, L) (+)= scale * sum_j (loss[a, first + j] - loss[a, first + H + j]) * z[a, j]
 *   = scale * S * torch.mean(loss * noise, 1) of _NES.py:47,51 with the noise REGENERATED from the same (seed, offset)
 *   instead of read back from HBM.  loss: device (A, R).  accumulate = 0 overwrites grad, 1 adds to it.
 * ap_query_loss:   loss[b] = CrossEntropyLoss(reduction='none') (AP_LOSS_ENTROPY; resolve_loss task 'SCR',
 *   _utils.py:116-117) or score_real + confidence - max_other (AP_LOSS_MARGIN; SEC4SR_MarginLoss CSI branch,
 *   _utils.py:73-84; other - real when targeted; max(., 0) when clip); pred[b] = argmax (first index on ties).
 *   labels: device int64 (B); loss / pred may be NULL (not both).  A label outside [0, K) yields NaN.
 * ------------------------------------------------------------------------------------------------------------- */
enum { AP_LOSS_ENTROPY = 0, AP_LOSS_MARGIN = 1 };
uint64_t ap_nes_noise_blocks(int A, int S, int L);
int ap_nes_perturb(const float* x, float sigma, const float* z_or_null, uint64_t seed, uint64_t offset, int first,
                   float* out, int A, int S, int L, void* stream);
int ap_nes_gradient(const float* loss, const float* z_or_null, uint64_t seed, uint64_t offset, int first, float scale,
                    int accumulate, float* grad, int A, int S, int L, void* stream);
int ap_query_loss(const float* scores, const long long* labels, int B, int K, int kind, int targeted, float confidence,
                  int clip, float* loss, int* pred, void* stream);
/* g_scores (B, K) = g_loss[b] * d loss[b] / d scores[b, :]  (EOT with use_grad=True, _EOT.py:43-44) */
int ap_query_loss_vjp(const float* scores, const long long* labels, const float* g_loss, int B, int K, int kind,
                      int targeted, float confidence, int clip, float* g_scores, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Self tests of the tcgen05/TMA building blocks (used by tests/ and smoke): run a single-tile UMMA GEMM
 * D[128x256] = A[128xK] * B[256xK]^T (bf16 in, fp32 out) through TMA + TMEM and write D to `d_out` (device, 128*256).
 * a_bf16 / b_bf16: device, K-major, raw bf16 bits.  K must be a multiple of 64.
 * ------------------------------------------------------------------------------------------------------------- */
int ap_selftest_umma(const uint16_t* a_bf16, const uint16_t* b_bf16, float* d_out, int K, void* stream);
/* Test hook: run the network up to and including residual layer `layer` (current mode) and return, as fp32 (B, L, C)
 * channels-last device arrays, the next layer's input u = h + fc_t(emb) (Residual_block input after WaveNet.py:84) and
 * the gate output tanh*sigmoid (WaveNet.py:90).  Either output may be NULL.  B must fit one workspace chunk. */
int ap_diffwave_debug_layer(ap_diffwave_t h, const float* x, float t, int layer, float* u_next, float* gate, int B,
                            int L, void* stream);

/* Measurement hook (bench.py roofline): when enabled, every launch of the two tensor-core kernels is bracketed by CUDA
 * events recorded on the launching stream.  ap_diffwave_profile_read synchronises on them and returns the summed
 * device time in ms and the launch count of k1_layer (index 0) and k2_head (index 1) since the last enable. */
int ap_diffwave_profile(ap_diffwave_t h, int enable);
int ap_diffwave_profile_read(ap_diffwave_t h, double* ms2, int* count2);
/* Development aid (env AP_TC_DEBUG=1 at create): per-CTA wait-cycle counters of the last k1_layer launch,
 * host array of 256 x 16 int64 (see csrc/ap_wavenet_tc.cu for the slot meaning). */
int ap_diffwave_debug_counters(ap_diffwave_t h, long long* host16x256);

#ifdef __cplusplus
}
#endif
#endif /* AUDIOPURE_H_ */
