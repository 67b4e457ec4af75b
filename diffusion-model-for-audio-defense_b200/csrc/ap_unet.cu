// Spectrogram-domain diffusion purifier ("Diffusion-Spec"): the UNet eps-network of
// diffusion_models/Improved_Diffusion_Unconditional/improved_diffusion/unet.py:278-497 (ResBlock :107-197 with scale-shift
// GroupNorm, AttentionBlock / QKVAttention :200-253, Downsample / Upsample :50-104, timestep embedding nn.py:103-121) on NHWC
// fp32 activations.
//   * every convolution (3x3, stride-2 3x3, 1x1 qkv / proj / skip) is the implicit GEMM of ap_conv_layer.cuh with bias and the
//     block's residual add fused into its epilogue;
//   * GroupNorm(32) + [scale-shift] + SiLU is one kernel per use (one CTA per (sample, group): statistics and normalisation in
//     a single pass over the group, fp32);
//   * the timestep path is batch-constant in the purifier (every row is at the same discrete step), so the embedding MLP and
//     all 30 per-block projections are one packed GEMV per network evaluation;
//   * attention over T = H W <= 256 positions with 64-channel heads: one CTA per (sample, head), K and V in shared memory,
//     one thread per query with an online softmax.
// The module walk (which block follows which, channel counts, skip-connection stack) is built by the host from the same
// flat op list the Python side derives from UNetModel.__init__ (synthetic.unet_structure), so names, order and shapes agree
// with the reference's state dict by construction.
#include <cmath>
#include <cstdlib>
#include <memory>
#include <vector>

#include "ap_common.cuh"
#include "ap_conv_layer.cuh"
#include "ap_internal.h"

namespace ap {

__device__ __forceinline__ float silu_f(float v) { return v / (1.f + __expf(-v)); }

// out = [silu]( ((x - mean) * rstd * gamma + beta) [* (1 + scale) + shift] ), x NHWC [B][HW][C], one CTA per (b, group).
// scale_shift: [2 C] for this block (scale = first C entries, shift = last C: th.chunk(emb_out, 2, dim=1), unet.py:190), or null.
__global__ void __launch_bounds__(256) gn_kernel(const float* __restrict__ x, float* __restrict__ out, const float* __restrict__ gamma,
                                                 const float* __restrict__ beta, const float* __restrict__ scale_shift, int HW, int C,
                                                 int cpg, int act, int round_tf32) {
  const int b = blockIdx.x / 32, g = blockIdx.x % 32;
  const float* xb = x + static_cast<size_t>(b) * HW * C + g * cpg;
  float* ob = out + static_cast<size_t>(b) * HW * C + g * cpg;
  const int n = HW * cpg;
  // pass 1: mean; pass 2: variance about the mean (two-pass: matches torch's float32 group_norm to ~1e-7)
  __shared__ float red[8];
  __shared__ float s_mean, s_rstd;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += xb[static_cast<size_t>(i / cpg) * C + (i % cpg)];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    s_mean = t / n;
  }
  __syncthreads();
  const float mean = s_mean;
  float v = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = xb[static_cast<size_t>(i / cpg) * C + (i % cpg)] - mean;
    v = fmaf(d, d, v);
  }
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    s_rstd = rsqrtf(t / n + 1e-5f);
  }
  __syncthreads();
  const float rstd = s_rstd;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i % cpg, ch = g * cpg + c;
    const size_t idx = static_cast<size_t>(i / cpg) * C + c;
    float y = (xb[idx] - mean) * rstd * gamma[ch] + beta[ch];
    if (scale_shift) y = y * (1.f + scale_shift[ch]) + scale_shift[C + ch];
    y = act ? silu_f(y) : y;
    if (round_tf32) {   // the tensor core truncates fp32 operands to tf32: round to nearest here instead
      uint32_t r;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(y));
      y = __uint_as_float(r);
    }
    ob[idx] = y;
  }
}

// emb = time_embed(timestep_embedding(t, mc)) (unet.py:335-339,476; nn.py:103-121), then ss[row] = W[row] . silu(emb) + b[row] for
// the packed rows of every ResBlock's emb_layers (unet.py:140-146).  Two launches: <<<1, 512>>> then one warp per packed row.
__global__ void __launch_bounds__(512) unet_time_embed_kernel(float t, int mc, const float* __restrict__ w0, const float* __restrict__ b0,
                                                              const float* __restrict__ w2, const float* __restrict__ b2,
                                                              float* __restrict__ emb_silu) {
  extern __shared__ float sm[];          // [mc] sinusoid | [4 mc] hidden
  const int ted = 4 * mc, half = mc / 2;
  float* e0 = sm;
  float* h = sm + mc;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float f = expf(-logf(10000.f) * static_cast<float>(i) / static_cast<float>(half));
    e0[i] = cosf(t * f), e0[half + i] = sinf(t * f);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < ted; o += blockDim.x) {
    float acc = b0[o];
    for (int k = 0; k < mc; ++k) acc = fmaf(w0[static_cast<size_t>(o) * mc + k], e0[k], acc);
    h[o] = silu_f(acc);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < ted; o += blockDim.x) {
    float acc = b2[o];
    for (int k = 0; k < ted; ++k) acc = fmaf(w2[static_cast<size_t>(o) * ted + k], h[k], acc);
    emb_silu[o] = silu_f(acc);          // every consumer applies SiLU first (emb_layers[0])
  }
}
__global__ void __launch_bounds__(256) unet_emb_proj_kernel(const float* __restrict__ emb_silu, int ted, const float* __restrict__ w,
                                                            const float* __restrict__ b, int rows, float* __restrict__ ss) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  float acc = 0.f;
  for (int k = lane; k < ted; k += 32) acc = fmaf(w[static_cast<size_t>(row) * ted + k], emb_silu[k], acc);
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) ss[row] = acc + b[row];
}

// QKVAttention (unet.py:238-253) on the NHWC qkv tensor [B][T][3 C]: head h owns channels [h 3 ch, (h + 1) 3 ch) = q | k | v
// (the reference reshapes (B, 3C, T) to (B heads, 3 ch, T), :226).  out [B][T][C], channel h ch + c.  One CTA per (b, head),
// one thread per query position; weight = softmax_s(q_t . k_s / sqrt(ch)).
template <int CH> __global__ void __launch_bounds__(256) unet_attn_kernel(const float* __restrict__ qkv, float* __restrict__ out, int T,
                                                                         int C, int heads) {
  extern __shared__ float kv[];          // K [T][CH] | V [T][CH]
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const float* base = qkv + static_cast<size_t>(b) * T * 3 * C + h * 3 * CH;
  float* ks = kv;
  float* vs = kv + static_cast<size_t>(T) * CH;
  for (int i = threadIdx.x; i < T * CH; i += blockDim.x) {
    const int s = i / CH, c = i - s * CH;
    ks[i] = base[static_cast<size_t>(s) * 3 * C + CH + c];
    vs[i] = base[static_cast<size_t>(s) * 3 * C + 2 * CH + c];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float q[CH], acc[CH];
    const float sc = rsqrtf(sqrtf(static_cast<float>(CH)));        // 1 / sqrt(sqrt(ch)), applied to q and to k (:247-250)
#pragma unroll
    for (int c = 0; c < CH; ++c) q[c] = base[static_cast<size_t>(t) * 3 * C + c] * sc, acc[c] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int s = 0; s < T; ++s) {
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) d = fmaf(q[c], ks[s * CH + c] * sc, d);
      const float mn = fmaxf(m, d), corr = __expf(m - mn), pw = __expf(d - mn);
      l = l * corr + pw;
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = fmaf(acc[c], corr, pw * vs[s * CH + c]);
      m = mn;
    }
    const float inv = 1.f / l;
    float* o = out + (static_cast<size_t>(b) * T + t) * C + h * CH;
#pragma unroll
    for (int c = 0; c < CH; ++c) o[c] = acc[c] * inv;
  }
}

// out[b][2i + di][2j + dj][c] = in[b][i][j][c]   (F.interpolate(scale_factor=2, mode='nearest'), unet.py:76)
__global__ void __launch_bounds__(256) nearest_up2_kernel(const float4* __restrict__ in, float4* __restrict__ out, int B, int H, int W,
                                                          int C4) {
  const long long total = static_cast<long long>(B) * (2 * H) * (2 * W) * C4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C4);
    long long t = i / C4;
    const int ow = static_cast<int>(t % (2 * W));
    t /= 2 * W;
    const int oh = static_cast<int>(t % (2 * H));
    const long long b = t / (2 * H);
    out[i] = in[((b * H + (oh >> 1)) * W + (ow >> 1)) * C4 + c];
  }
}
// out[b][p][0:C1] = a[b][p][:], out[b][p][C1:C1+C2] = s[b][p][:]     (th.cat([h, hs.pop()], dim=1), unet.py:493)
__global__ void __launch_bounds__(256) concat_kernel(const float4* __restrict__ a, const float4* __restrict__ s, float4* __restrict__ out,
                                                     long long pixels, int C1_4, int C2_4) {
  const int Ct = C1_4 + C2_4;
  const long long total = pixels * Ct;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Ct);
    const long long px = i / Ct;
    out[i] = c < C1_4 ? a[px * C1_4 + c] : s[px * C2_4 + (c - C1_4)];
  }
}
// NCHW (C == 1: same memory as NHWC) in / out is handled by the caller; nothing to do here.

static int grid_for_n(long long n, int threads) {
  long long b = ceil_div_ll(n, threads);
  const long long cap = static_cast<long long>(num_sms()) * 8;
  return static_cast<int>(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace ap

using namespace ap;

enum { OP_CONV_IN = 0, OP_RES = 1, OP_ATTN = 2, OP_PUSH = 3, OP_POP = 4, OP_DOWN = 5, OP_UP = 6, OP_OUT = 7 };

namespace {
struct UOp {
  int kind = 0, cin = 0, cout = 0;
  ConvLayer c1, c2, skip;      // res: in conv, out conv, skip 1x1 | attn: qkv (c1), proj (c2) | conv_in / down / up / out: c1
  bool has_skip = false;
  DevBuf g1, b1, g2, b2;       // GroupNorm affine parameters (res: in_layers.0 / out_layers.0; attn: norm; out: out.0)
  int ss_off = 0;              // res: first packed row of this block's (scale | shift)
};
}  // namespace

struct ap_unet_s {
  ap_unet_cfg cfg{};
  int device = 0;
  int mode = AP_MODE_TF32;      // AP_MODE_TF32: tcgen05 kind::tf32 convolutions where a tile shape exists; AP_MODE_FP32: FFMA everywhere
  std::vector<std::unique_ptr<UOp>> ops;
  DevBuf te_w0, te_b0, te_w2, te_b2, emb_w, emb_b, emb_silu, ss;
  int ss_rows = 0;
  // activation arena (bump allocator, sized for `cap_B` samples) -- every buffer of one evaluation lives here
  DevBuf arena;
  size_t arena_floats = 0;
  int cap_B = 0;
  size_t need_per_sample = 0;   // floats
  bool attn_attr = false;
};

static int upload_vec(DevBuf& d, const float* p, size_t n) {
  AP_CUDA(d.upload(p, n * sizeof(float)));
  return AP_OK;
}

// weights: host fp32 pointers in state-dict order (the reference's 446 tensors for the default configuration); ops: the flat
// op list [n_ops][3] = (kind, cin, cout) from synthetic.unet_structure / improved_diffusion.UNet.
extern "C" int ap_unet_create(ap_unet_t* out, const ap_unet_cfg* cfg, const int* op_list, int n_ops, const float* const* w,
                              int n_weights, int device) {
  AP_REQUIRE(out && cfg && op_list && w, "ap_unet_create: null argument");
  *out = nullptr;
  AP_REQUIRE(cfg->model_channels % 32 == 0 && cfg->model_channels >= 32 && cfg->num_heads > 0 && cfg->in_channels == 1 &&
             cfg->out_channels == 1 && cfg->use_scale_shift_norm == 1,
             "ap_unet_create: unsupported configuration (1 input / output channel, scale-shift norm, model_channels %% 32 == 0)");
  int rc = select_device(device);
  if (rc != AP_OK) return rc;
  auto h = std::unique_ptr<ap_unet_s>(new ap_unet_s());
  h->cfg = *cfg, h->device = device;
  const int mc = cfg->model_channels, ted = 4 * mc;
  int wi = 0;
  auto next = [&]() -> const float* { return wi < n_weights ? w[wi++] : nullptr; };
#define TRY(x) do { int rc__ = (x); if (rc__ != AP_OK) return rc__; } while (0)
  {
    const float *w0 = next(), *b0 = next(), *w2 = next(), *b2 = next();
    AP_REQUIRE(b2, "ap_unet_create: too few weights");
    TRY(upload_vec(h->te_w0, w0, static_cast<size_t>(ted) * mc));
    TRY(upload_vec(h->te_b0, b0, ted));
    TRY(upload_vec(h->te_w2, w2, static_cast<size_t>(ted) * ted));
    TRY(upload_vec(h->te_b2, b2, ted));
  }
  std::vector<float> emb_w, emb_b;
  for (int i = 0; i < n_ops; ++i) {
    auto op = std::unique_ptr<UOp>(new UOp());
    op->kind = op_list[3 * i], op->cin = op_list[3 * i + 1], op->cout = op_list[3 * i + 2];
    const int ci = op->cin, co = op->cout;
    if (op->kind == OP_CONV_IN || op->kind == OP_DOWN || op->kind == OP_UP) {
      const float *cw = next(), *cb = next();
      AP_REQUIRE(cb, "ap_unet_create: too few weights");
      TRY(op->c1.init(ci, co, 3, 3, op->kind == OP_DOWN ? 2 : 1, 1, 1, cw, cb, nullptr, nullptr, nullptr, nullptr, true));
    } else if (op->kind == OP_RES) {
      const float *g1 = next(), *b1 = next(), *w1 = next(), *bb1 = next(), *ew = next(), *eb = next(), *g2 = next(), *b2 = next(),
                  *w2 = next(), *bb2 = next();
      AP_REQUIRE(bb2, "ap_unet_create: too few weights");
      AP_REQUIRE(ci % 32 == 0 && co % 32 == 0, "ap_unet_create: channel counts must be multiples of 32 (GroupNorm32)");
      TRY(upload_vec(op->g1, g1, ci));
      TRY(upload_vec(op->b1, b1, ci));
      TRY(op->c1.init(ci, co, 3, 3, 1, 1, 1, w1, bb1, nullptr, nullptr, nullptr, nullptr, true));
      op->ss_off = static_cast<int>(emb_b.size());
      emb_w.insert(emb_w.end(), ew, ew + static_cast<size_t>(2 * co) * ted);
      emb_b.insert(emb_b.end(), eb, eb + 2 * co);
      TRY(upload_vec(op->g2, g2, co));
      TRY(upload_vec(op->b2, b2, co));
      TRY(op->c2.init(co, co, 3, 3, 1, 1, 1, w2, bb2, nullptr, nullptr, nullptr, nullptr, true));
      if (ci != co) {
        const float *sw = next(), *sb = next();
        AP_REQUIRE(sb, "ap_unet_create: too few weights");
        TRY(op->skip.init(ci, co, 1, 1, 1, 0, 1, sw, sb, nullptr, nullptr, nullptr, nullptr, true));
        op->has_skip = true;
      }
    } else if (op->kind == OP_ATTN) {
      const float *g = next(), *b = next(), *qw = next(), *qb = next(), *pw = next(), *pb = next();
      AP_REQUIRE(pb, "ap_unet_create: too few weights");
      AP_REQUIRE(ci % cfg->num_heads == 0 && ci / cfg->num_heads == 64, "ap_unet_create: attention heads must be 64 channels wide");
      TRY(upload_vec(op->g1, g, ci));
      TRY(upload_vec(op->b1, b, ci));
      TRY(op->c1.init(ci, 3 * ci, 1, 1, 1, 0, 1, qw, qb, nullptr, nullptr, nullptr, nullptr, true));
      TRY(op->c2.init(ci, ci, 1, 1, 1, 0, 1, pw, pb, nullptr, nullptr, nullptr, nullptr, true));
    } else if (op->kind == OP_OUT) {
      const float *g = next(), *b = next(), *cw = next(), *cb = next();
      AP_REQUIRE(cb, "ap_unet_create: too few weights");
      TRY(upload_vec(op->g1, g, ci));
      TRY(upload_vec(op->b1, b, ci));
      TRY(op->c1.init(ci, co, 3, 3, 1, 1, 1, cw, cb, nullptr, nullptr, nullptr, nullptr, true));
    } else {
      AP_REQUIRE(op->kind == OP_PUSH || op->kind == OP_POP, "ap_unet_create: unknown op kind %d", op->kind);
    }
    h->ops.push_back(std::move(op));
  }
  AP_REQUIRE(wi == n_weights, "ap_unet_create: %d weights given, the op list consumes %d", n_weights, wi);
  h->ss_rows = static_cast<int>(emb_b.size());
  TRY(upload_vec(h->emb_w, emb_w.data(), emb_w.size()));
  TRY(upload_vec(h->emb_b, emb_b.data(), emb_b.size()));
  AP_CUDA(h->emb_silu.alloc(sizeof(float) * ted));
  AP_CUDA(h->ss.alloc(sizeof(float) * h->ss_rows));
#undef TRY
  // arena size per sample: walk the ops once with a symbolic allocator (same order as the forward pass)
  {
    size_t need = 0;
    int H = cfg->image_size;
    std::vector<int> stackH;
    for (auto& op : h->ops) {
      const size_t px = static_cast<size_t>(H) * H;
      switch (op->kind) {
        case OP_CONV_IN: need += px * op->cout; break;
        case OP_RES: need += px * (op->cin + 3 * static_cast<size_t>(op->cout)) + (op->has_skip ? px * op->cout : 0); break;
        case OP_ATTN: need += px * (op->cin + 3 * static_cast<size_t>(op->cin) + 2 * static_cast<size_t>(op->cin)); break;
        case OP_POP: need += px * op->cout; break;
        case OP_DOWN: H /= 2; need += static_cast<size_t>(H) * H * op->cout; break;
        case OP_UP: need += 4 * px * op->cin + 4 * px * op->cout; H *= 2; break;
        case OP_OUT: need += px * (op->cin + op->cout); break;
        default: break;
      }
    }
    h->need_per_sample = need + 64;
  }
  *out = h.release();
  return AP_OK;
}

extern "C" void ap_unet_destroy(ap_unet_t h) { delete h; }
extern "C" int ap_unet_set_mode(ap_unet_t h, int mode) {
  AP_REQUIRE(h && (mode == AP_MODE_TF32 || mode == AP_MODE_FP32), "ap_unet_set_mode: AP_MODE_TF32 or AP_MODE_FP32");
  h->mode = mode;
  return AP_OK;
}

// eps[b] = UNet(x[b], t) with the same discrete step t for every sample (RevVPSDE.rvpsde_fn passes one step per Euler step,
// improved_diffusion_sde.py:104-105).  x, eps: device fp32 (B, 1, S, S).
extern "C" int ap_unet_eps(ap_unet_t h, const float* x, float t, float* eps, int B, void* stream) {
  AP_REQUIRE(h && x && eps && B > 0, "ap_unet_eps: bad arguments");
  AP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int mc = h->cfg.model_channels, ted = 4 * mc, S = h->cfg.image_size, heads = h->cfg.num_heads;
  // activations: sub-batches that keep the arena below ~8 GB
  const size_t per = h->need_per_sample;
  int chunk = static_cast<int>(std::max<size_t>(1, (static_cast<size_t>(2) << 30) / per));   // floats: 2 Gi floats = 8 GB
  if (chunk > B) chunk = B;
  if (h->cap_B < chunk) {
    h->cap_B = 0;
    AP_CUDA(h->arena.alloc(per * chunk * sizeof(float)));
    h->cap_B = chunk, h->arena_floats = per * chunk;
  }
  unet_time_embed_kernel<<<1, 512, sizeof(float) * (mc + ted), st>>>(t, mc, h->te_w0.as<float>(), h->te_b0.as<float>(),
                                                                     h->te_w2.as<float>(), h->te_b2.as<float>(), h->emb_silu.as<float>());
  AP_LAUNCH_CHECK();
  unet_emb_proj_kernel<<<ceil_div(h->ss_rows * 32, 256), 256, 0, st>>>(h->emb_silu.as<float>(), ted, h->emb_w.as<float>(),
                                                                       h->emb_b.as<float>(), h->ss_rows, h->ss.as<float>());
  AP_LAUNCH_CHECK();
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bn = std::min(chunk, B - b0);
    float* arena = h->arena.as<float>();
    size_t top = 0;
    auto alloc = [&](size_t n) {
      float* p = arena + top;
      top += (n + 3) & ~static_cast<size_t>(3);
      return p;
    };
    struct Act { float* p; int H, C; };
    std::vector<Act> stack;
    Act cur{const_cast<float*>(x) + static_cast<size_t>(b0) * S * S, S, 1};
    auto gn = [&](const Act& a, float* out, const DevBuf& g, const DevBuf& b, const float* ss, int act) -> int {
      gn_kernel<<<bn * 32, 256, 0, st>>>(a.p, out, g.as<float>(), b.as<float>(), ss, a.H * a.H, a.C, a.C / 32, act,
                                         h->mode == AP_MODE_TF32);
      AP_LAUNCH_CHECK();
      return AP_OK;
    };
    auto conv = [&](const ConvLayer& L, const float* in, int H, float* out, const float* res) -> int {
      if (h->mode == AP_MODE_TF32 && L.has_tc && conv_tc_supported(L.Cin, L.Cout, L.groups, H, H, L.kh, L.kw, L.stride, L.pad)) {
        ConvTcBinding bnd;
        int rcb = L.tc.bind(&bnd, in, bn, H, H, out, res, 0, 0);
        return rcb != AP_OK ? rcb : L.tc.run(bnd, st);
      }
      return L.run(in, bn, H, H, out, res, 0, st);
    };
    int rc = AP_OK;
    for (auto& opp : h->ops) {
      UOp& op = *opp;
      const size_t px = static_cast<size_t>(bn) * cur.H * cur.H;
      if (op.kind == OP_CONV_IN) {
        float* o = alloc(px * op.cout);
        rc = conv(op.c1, cur.p, cur.H, o, nullptr);
        cur = {o, cur.H, op.cout};
        stack.push_back(cur);                              // hs.append(h) of input_blocks[0] (unet.py:483-485)
      } else if (op.kind == OP_RES) {
        float* n1 = alloc(px * op.cin);
        float* y1 = alloc(px * op.cout);
        float* n2 = alloc(px * op.cout);
        float* o = alloc(px * op.cout);
        rc = gn(cur, n1, op.g1, op.b1, nullptr, 1);                                             // in_layers: GN, SiLU
        if (rc == AP_OK) rc = conv(op.c1, n1, cur.H, y1, nullptr);              //            conv
        Act a1{y1, cur.H, op.cout};
        if (rc == AP_OK) rc = gn(a1, n2, op.g2, op.b2, h->ss.as<float>() + op.ss_off, 1);       // out_layers[0] * (1 + scale) + shift, SiLU
        const float* res = cur.p;
        if (rc == AP_OK && op.has_skip) {
          float* sk = alloc(px * op.cout);
          rc = conv(op.skip, cur.p, cur.H, sk, nullptr);
          res = sk;
        }
        if (rc == AP_OK) rc = conv(op.c2, n2, cur.H, o, res);                   // conv + skip_connection(x)
        cur = {o, cur.H, op.cout};
      } else if (op.kind == OP_ATTN) {
        const int T = cur.H * cur.H, Cc = cur.C;
        float* n1 = alloc(px * Cc);
        float* qkv = alloc(px * 3 * Cc);
        float* av = alloc(px * Cc);
        float* o = alloc(px * Cc);
        rc = gn(cur, n1, op.g1, op.b1, nullptr, 0);
        if (rc == AP_OK) rc = conv(op.c1, n1, cur.H, qkv, nullptr);
        if (rc == AP_OK) {
          const size_t smem = static_cast<size_t>(2) * T * 64 * sizeof(float);
          if (!h->attn_attr) {     // per handle, i.e. per device: function attributes belong to the device's context
            AP_CUDA(cudaFuncSetAttribute(unet_attn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * 64 * 4));
            h->attn_attr = true;
          }
          AP_REQUIRE(T <= 256, "ap_unet_eps: attention over more than 256 positions is not supported");
          unet_attn_kernel<64><<<bn * heads, T < 256 ? ((T + 31) / 32) * 32 : 256, smem, st>>>(qkv, av, T, Cc, heads);
          AP_LAUNCH_CHECK();
          rc = conv(op.c2, av, cur.H, o, cur.p);                                // proj_out + x
        }
        cur = {o, cur.H, Cc};
      } else if (op.kind == OP_PUSH) {
        stack.push_back(cur);
      } else if (op.kind == OP_POP) {
        AP_REQUIRE(!stack.empty(), "ap_unet_eps: skip stack underflow");
        const Act s = stack.back();
        stack.pop_back();
        AP_REQUIRE(s.H == cur.H && s.C + cur.C == op.cout, "ap_unet_eps: skip connection shape mismatch");
        float* o = alloc(px * op.cout);
        concat_kernel<<<grid_for_n(static_cast<long long>(px) * op.cout / 4, 256), 256, 0, st>>>(
            reinterpret_cast<const float4*>(cur.p), reinterpret_cast<const float4*>(s.p), reinterpret_cast<float4*>(o),
            static_cast<long long>(px), cur.C / 4, s.C / 4);
        AP_LAUNCH_CHECK();
        cur = {o, cur.H, op.cout};
      } else if (op.kind == OP_DOWN) {
        const int Ho = cur.H / 2;
        float* o = alloc(static_cast<size_t>(bn) * Ho * Ho * op.cout);
        rc = conv(op.c1, cur.p, cur.H, o, nullptr);
        cur = {o, Ho, op.cout};
      } else if (op.kind == OP_UP) {
        float* up = alloc(4 * px * op.cin);
        float* o = alloc(4 * px * op.cout);
        nearest_up2_kernel<<<grid_for_n(static_cast<long long>(px) * op.cin, 256), 256, 0, st>>>(
            reinterpret_cast<const float4*>(cur.p), reinterpret_cast<float4*>(up), bn, cur.H, cur.H, op.cin / 4);
        AP_LAUNCH_CHECK();
        rc = conv(op.c1, up, 2 * cur.H, o, nullptr);
        cur = {o, 2 * cur.H, op.cout};
      } else if (op.kind == OP_OUT) {
        float* n1 = alloc(px * op.cin);
        rc = gn(cur, n1, op.g1, op.b1, nullptr, 1);
        if (rc == AP_OK) rc = conv(op.c1, n1, cur.H, eps + static_cast<size_t>(b0) * S * S, nullptr);
      }
      if (rc != AP_OK) return rc;
      if (top > h->arena_floats) return fail(AP_ERR_STATE, "ap_unet_eps: activation arena overflow (%zu > %zu floats)", top, h->arena_floats);
    }
  }
  return AP_OK;
}
