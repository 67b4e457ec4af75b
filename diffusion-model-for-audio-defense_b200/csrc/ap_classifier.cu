// K5 / K6: classifier forward passes (replace Classifier(spec) at acoustic_system.py:49 / certified_robust.py:30).
//   AP_CLS_RESNEXT  CifarResNeXt (models/resnext.py:23-142): every convolution is an implicit GEMM over NHWC activations
//                   (rows = output pixels, K = taps x input channels of the group) with eval-mode BatchNorm folded into
//                   the weights/bias and the residual add + ReLU fused into the epilogue.
//   AP_CLS_M5       M5 (audio_models/M5/M5Net.py:4-38)            conv1d/BN/ReLU/maxpool x4, avgpool, fc, log_softmax
//   AP_CLS_KWS      KWSModel (audio_models/RCNN_KWS/model.py:5-113)  sepconv, 2-layer bi-GRU, attention, fc, log_softmax
#include <algorithm>
#include <cmath>
#include <memory>

#include <cstdlib>
#include <map>

#include "ap_common.cuh"
#include "ap_conv_layer.cuh"
#include "ap_conv_tc.h"
#include "ap_internal.h"
#include "ap_sgemm.cuh"

namespace ap {

// g[i] = act[i] > 0 ? g[i] : 0      (backward of ReLU, from the saved activation)
__global__ void __launch_bounds__(256) relu_mask_kernel(float4* __restrict__ g, const float4* __restrict__ act, long long n4) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 v = g[i];
    const float4 a = act[i];
    v.x = a.x > 0.f ? v.x : 0.f, v.y = a.y > 0.f ? v.y : 0.f, v.z = a.z > 0.f ? v.z : 0.f, v.w = a.w > 0.f ? v.w : 0.f;
    g[i] = v;
  }
}
// up[b][2i][2j][c] = g[b][i][j][c], zero elsewhere   (NHWC; gradient of a stride-2 convolution before its dgrad)
__global__ void __launch_bounds__(256) upsample2_kernel(const float4* __restrict__ g, float4* __restrict__ up, int B, int Ho, int Wo,
                                                        int C4) {
  const long long total = static_cast<long long>(B) * (2 * Ho) * (2 * Wo) * C4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C4);
    long long t = i / C4;
    const int w = static_cast<int>(t % (2 * Wo));
    t /= 2 * Wo;
    const int hh = static_cast<int>(t % (2 * Ho));
    const long long b = t / (2 * Ho);
    up[i] = ((hh | w) & 1) ? make_float4(0.f, 0.f, 0.f, 0.f) : g[((b * Ho + (hh >> 1)) * Wo + (w >> 1)) * C4 + c];
  }
}
// backward of avg_pool(P positions) + linear: g_y[b][p][c] = (sum_k fc_w[k][c] g_logits[b][k]) / P
__global__ void __launch_bounds__(256) pool_fc_bwd_kernel(const float* __restrict__ g_logits, const float* __restrict__ fw, int Kc,
                                                          int Cc, int P, float* __restrict__ g_y) {
  const int b = blockIdx.x;
  const float inv = 1.f / static_cast<float>(P);
  for (int c = threadIdx.x; c < Cc; c += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < Kc; ++k) acc = fmaf(fw[k * Cc + c], g_logits[b * Kc + k], acc);
    acc *= inv;
    for (int pp = 0; pp < P; ++pp) g_y[(static_cast<long long>(b) * P + pp) * Cc + c] = acc;
  }
}

// 1-D zero-upsampling: up[b][i*s][c] = g[b][i][c] (i < lo), zero elsewhere; lu >= (lo-1)*s + 1   (stride-s conv1d dgrad input)
__global__ void __launch_bounds__(256) upsample1d_kernel(const float* __restrict__ g, float* __restrict__ up, int B, int lo, int lu,
                                                         int s, int Cc) {
  const long long total = static_cast<long long>(B) * lu * Cc;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cc);
    const long long t = i / Cc;
    const int l = static_cast<int>(t % lu);
    const long long b = t / lu;
    const int q = l / s;
    up[i] = (l - q * s == 0 && q < lo) ? g[(b * lo + q) * Cc + c] : 0.f;
  }
}
// backward of max_pool1d(4, 4) after ReLU: the gradient goes to the first maximum of each window if it is positive
__global__ void __launch_bounds__(256) maxpool4_relu_bwd_kernel(const float* __restrict__ g_p, const float* __restrict__ a,
                                                                float* __restrict__ g_a, int B, int lo, int Cc) {
  const int lp = lo / 4;
  const long long total = static_cast<long long>(B) * lo * Cc;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cc);
    const long long t = i / Cc;
    const int l = static_cast<int>(t % lo);
    const long long b = t / lo;
    const int w = l / 4, j = l - 4 * w;
    float out = 0.f;
    if (w < lp) {
      const float* p = a + ((b * lo + 4 * w) * Cc) + c;
      int arg = 0;
      float m = p[0];
      for (int q = 1; q < 4; ++q)
        if (p[q * Cc] > m) m = p[q * Cc], arg = q;
      if (arg == j && m > 0.f) out = g_p[(b * lp + w) * Cc + c];
    }
    g_a[i] = out;
  }
}
// backward of avg-pool(P) + linear + log_softmax: g_z = g_out - softmax * sum(g_out); g_x[b][p][c] = (fc_w^T g_z)[c] / P
__global__ void __launch_bounds__(256) logsoftmax_pool_fc_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ logp,
                                                                     const float* __restrict__ fw, int Kc, int Cc, int P,
                                                                     float* __restrict__ g_x) {
  __shared__ float gz[64];
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    float sum = 0.f;
    for (int k = 0; k < Kc; ++k) sum += g_out[b * Kc + k];
    for (int k = 0; k < Kc; ++k) gz[k] = g_out[b * Kc + k] - expf(logp[b * Kc + k]) * sum;
  }
  __syncthreads();
  const float inv = 1.f / static_cast<float>(P);
  for (int c = threadIdx.x; c < Cc; c += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < Kc; ++k) acc = fmaf(fw[k * Cc + c], gz[k], acc);
    acc *= inv;
    for (int pp = 0; pp < P; ++pp) g_x[(static_cast<long long>(b) * P + pp) * Cc + c] = acc;
  }
}

// global average pool over P positions + linear: logits[b][k] = fc_b[k] + sum_c fc_w[k][c] * mean_p x[b][p][c]
__global__ void __launch_bounds__(256) pool_fc_kernel(const float* __restrict__ x, int P, int Cc, const float* __restrict__ fw,
                                                      const float* __restrict__ fb, int Kc, float* __restrict__ logits,
                                                      int log_softmax) {
  extern __shared__ float sm[];   // Cc pooled + Kc logits
  float* pooled = sm;
  float* lg = sm + Cc;
  const int b = blockIdx.x;
  const float inv = 1.f / static_cast<float>(P);
  for (int c = threadIdx.x; c < Cc; c += blockDim.x) {
    float acc = 0.f;
    for (int pp = 0; pp < P; ++pp) acc += x[(static_cast<long long>(b) * P + pp) * Cc + c];
    pooled[c] = acc * inv;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < Kc; k += blockDim.x >> 5) {
    float acc = 0.f;
    for (int c = lane; c < Cc; c += 32) acc = fmaf(fw[k * Cc + c], pooled[c], acc);
    for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) lg[k] = acc + (fb ? fb[k] : 0.f);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float off = 0.f;
    if (log_softmax) {
      float mx = lg[0];
      for (int k = 1; k < Kc; ++k) mx = fmaxf(mx, lg[k]);
      float s = 0.f;
      for (int k = 0; k < Kc; ++k) s += expf(lg[k] - mx);
      off = mx + logf(s);
    }
    for (int k = 0; k < Kc; ++k) logits[b * Kc + k] = lg[k] - off;
  }
}

// max_pool1d(kernel 4, stride 4) over channels-last [B][Lin][C] -> [B][Lin/4][C]
__global__ void __launch_bounds__(256) maxpool4_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int Lin,
                                                       int Cc) {
  const int Lout = Lin / 4;
  const long long total = static_cast<long long>(B) * Lout * Cc;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cc);
    const long long t = i / Cc;
    const int l = static_cast<int>(t % Lout);
    const long long b = t / Lout;
    const float* p = in + ((b * Lin + 4 * l) * Cc) + c;
    out[i] = fmaxf(fmaxf(p[0], p[Cc]), fmaxf(p[2 * Cc], p[3 * Cc]));
  }
}

// MaxPool2d(kernel 3, stride 2, padding 1) over NHWC [B][H][W][C] -> [B][Ho][Wo][C]   (torchvision ResNet stem, resnet.py:112)
// y = relu(x * scale[c] + shift[c]) on NHWC: a pre-activation BatchNorm (eval) + ReLU that cannot fold into a convolution
// because it acts on a residual sum (models/wideresnet.py:31-35,87)
__device__ __forceinline__ float round_tf32_rne(float x) {   // what the tensor-core path's consumers expect (ap_conv_tc.cu)
  uint32_t u = __float_as_uint(x);
  u = (u + 0x00000fffu + ((u >> 13) & 1u)) & 0xffffe000u;
  return __uint_as_float(u);
}
__global__ void __launch_bounds__(256) bn_relu_kernel(const float4* __restrict__ x, const float4* __restrict__ scale,
                                                      const float4* __restrict__ shift, float4* __restrict__ y, long long n4,
                                                      int C4, int round_out) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C4);
    const float4 v = x[i], a = scale[c], b = shift[c];
    float4 o = make_float4(fmaxf(fmaf(v.x, a.x, b.x), 0.f), fmaxf(fmaf(v.y, a.y, b.y), 0.f), fmaxf(fmaf(v.z, a.z, b.z), 0.f),
                           fmaxf(fmaf(v.w, a.w, b.w), 0.f));
    if (round_out) o = make_float4(round_tf32_rne(o.x), round_tf32_rne(o.y), round_tf32_rne(o.z), round_tf32_rne(o.w));
    y[i] = o;
  }
}
// its backward: g_x = (act > 0 ? g_act * scale[c] : 0) + skip   (skip: the identity-shortcut gradient, or null)
__global__ void __launch_bounds__(256) bn_relu_bwd_kernel(const float4* g_act, const float4* __restrict__ act,
                                                          const float4* __restrict__ scale, const float4* skip, float4* g_x,
                                                          long long n4, int C4) {   // g_x may alias g_act (same index only)
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C4);
    const float4 g = g_act[i], a = act[i], sc = scale[c];
    float4 o = skip ? skip[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    o.x += a.x > 0.f ? g.x * sc.x : 0.f;
    o.y += a.y > 0.f ? g.y * sc.y : 0.f;
    o.z += a.z > 0.f ? g.z * sc.z : 0.f;
    o.w += a.w > 0.f ? g.w * sc.w : 0.f;
    g_x[i] = o;
  }
}
// DenseNet: pre-activation BatchNorm + ReLU of a channel PREFIX of the concatenated tensor x (pixel stride xs floats) into
// a compact tensor y (pixel stride C4 * 4); scale / shift are zero on layout-padding channels  (models/densenet.py:27-28)
__global__ void __launch_bounds__(256) bn_relu_prefix_kernel(const float* __restrict__ x, int xs, const float4* __restrict__ scale,
                                                             const float4* __restrict__ shift, float4* __restrict__ y,
                                                             long long npix, int C4) {
  const long long total = npix * C4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C4);
    const long long pix = i / C4;
    const float4 v = *reinterpret_cast<const float4*>(x + pix * xs + 4 * c), a = scale[c], b = shift[c];
    y[i] = make_float4(fmaxf(fmaf(v.x, a.x, b.x), 0.f), fmaxf(fmaf(v.y, a.y, b.y), 0.f), fmaxf(fmaf(v.z, a.z, b.z), 0.f),
                       fmaxf(fmaf(v.w, a.w, b.w), 0.f));
  }
}
// its backward, with the ReLU mask recomputed from x: gx[prefix] (+)= (x * scale + shift > 0) ? g * scale : 0
__global__ void __launch_bounds__(256) bn_relu_prefix_bwd_kernel(const float4* __restrict__ g, const float* __restrict__ x, int xs,
                                                                 const float4* __restrict__ scale, const float4* __restrict__ shift,
                                                                 float* __restrict__ gx, long long npix, int C4, int accumulate) {
  const long long total = npix * C4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C4);
    const long long pix = i / C4;
    const float4 v = *reinterpret_cast<const float4*>(x + pix * xs + 4 * c), a = scale[c], b = shift[c], gg = g[i];
    float4* dst = reinterpret_cast<float4*>(gx + pix * xs + 4 * c);
    float4 o = accumulate ? *dst : make_float4(0.f, 0.f, 0.f, 0.f);
    o.x += fmaf(v.x, a.x, b.x) > 0.f ? gg.x * a.x : 0.f;
    o.y += fmaf(v.y, a.y, b.y) > 0.f ? gg.y * a.y : 0.f;
    o.z += fmaf(v.z, a.z, b.z) > 0.f ? gg.z * a.z : 0.f;
    o.w += fmaf(v.w, a.w, b.w) > 0.f ? gg.w * a.w : 0.f;
    *dst = o;
  }
}
// avg_pool2d(2) of a compact (B, H, W, cs) tensor into the first C channels of a (B, H/2, W/2, os) tensor (densenet.py:70),
// and its backward (every input pixel receives a quarter of its window's gradient)
__global__ void __launch_bounds__(256) avgpool2_kernel(const float* __restrict__ in, int cs, float* __restrict__ out, int os, int B,
                                                       int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long t = i / C;
    const int ow = static_cast<int>(t % Wo);
    t /= Wo;
    const int oh = static_cast<int>(t % Ho);
    const long long b = t / Ho;
    const float* p = in + ((b * H + 2 * oh) * W + 2 * ow) * cs + c;
    out[((b * Ho + oh) * Wo + ow) * os + c] = (p[0] + p[cs] + p[static_cast<long long>(W) * cs] + p[static_cast<long long>(W + 1) * cs]) * 0.25f;
  }
}
__global__ void __launch_bounds__(256) avgpool2_bwd_kernel(const float* __restrict__ g_out, int os, float* __restrict__ g_in, int cs,
                                                           int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * H * W * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long t = i / C;
    const int w = static_cast<int>(t % W);
    t /= W;
    const int hh = static_cast<int>(t % H);
    const long long b = t / H;
    g_in[((b * H + hh) * W + w) * cs + c] = 0.25f * g_out[((b * Ho + (hh >> 1)) * Wo + (w >> 1)) * os + c];
  }
}
// Data gradient of a strided single-input-channel conv1d (M5's first layer: k = 160, stride 16, 1 -> 32 channels):
//   g_x[b][n] = sum_{i : 0 <= n - i s < k} sum_c g_y[b][i][c] * w[c][n - i s]
// The generic dgrad twin would run an implicit GEMM with ONE useful output column of its 128-wide tile over a 16x
// zero-upsampled gradient (263 ms for 512 clips); here each thread owns one sample n and walks its <= k / s windows.
__global__ void __launch_bounds__(256) conv1d_cin1_dgrad_kernel(const float* __restrict__ g_y, const float* __restrict__ w,
                                                                float* __restrict__ g_x, int L, int lo, int k, int s, int Cc) {
  extern __shared__ float ws[];   // [c][j]
  for (int i = threadIdx.x; i < Cc * k; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const float* gy = g_y + static_cast<long long>(b) * lo * Cc;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < L; n += gridDim.x * blockDim.x) {
    int i_hi = n / s;
    if (i_hi > lo - 1) i_hi = lo - 1;
    int i_lo = n - k + 1 <= 0 ? 0 : (n - k + s) / s;      // ceil((n - k + 1) / s)
    float acc = 0.f;
    for (int i = i_lo; i <= i_hi; ++i) {
      const int j = n - i * s;
      const float* row = gy + static_cast<long long>(i) * Cc;
      for (int c = 0; c < Cc; ++c) acc = fmaf(row[c], ws[c * k + j], acc);
    }
    g_x[static_cast<long long>(b) * L + n] = acc;
  }
}
// MaxPool2d(kernel 2, stride 2) on NHWC (models/vgg.py:73), H and W even
__global__ void __launch_bounds__(256) maxpool2x2_kernel(const float4* __restrict__ in, float4* __restrict__ out, int B, int H, int W,
                                                         int C4) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo * C4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C4);
    long long t = i / C4;
    const int ow = static_cast<int>(t % Wo);
    t /= Wo;
    const int oh = static_cast<int>(t % Ho);
    const long long b = t / Ho;
    const float4* p = in + ((b * H + 2 * oh) * W + 2 * ow) * C4 + c;
    const float4 v0 = p[0], v1 = p[C4], v2 = p[static_cast<long long>(W) * C4], v3 = p[static_cast<long long>(W + 1) * C4];
    out[i] = make_float4(fmaxf(fmaxf(v0.x, v1.x), fmaxf(v2.x, v3.x)), fmaxf(fmaxf(v0.y, v1.y), fmaxf(v2.y, v3.y)),
                         fmaxf(fmaxf(v0.z, v1.z), fmaxf(v2.z, v3.z)), fmaxf(fmaxf(v0.w, v1.w), fmaxf(v2.w, v3.w)));
  }
}
// its backward: the window's gradient goes to the first maximal element in (row, column) scan order, like torch
__global__ void __launch_bounds__(256) maxpool2x2_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ in,
                                                             float* __restrict__ g_in, int B, int H, int W, int Cc) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo * Cc;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cc);
    long long t = i / Cc;
    const int ow = static_cast<int>(t % Wo);
    t /= Wo;
    const int oh = static_cast<int>(t % Ho);
    const long long b = t / Ho;
    const long long base = ((b * H + 2 * oh) * W + 2 * ow) * Cc + c;
    const long long off[4] = {0, Cc, static_cast<long long>(W) * Cc, static_cast<long long>(W + 1) * Cc};
    float m = in[base];
    int arg = 0;
#pragma unroll
    for (int q = 1; q < 4; ++q) {
      const float v = in[base + off[q]];
      if (v > m) m = v, arg = q;
    }
    const float g = g_out[i];
#pragma unroll
    for (int q = 0; q < 4; ++q) g_in[base + off[q]] = q == arg ? g : 0.f;
  }
}
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int H,
                                                           int W, int Cc) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = static_cast<long long>(B) * Ho * Wo * Cc;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cc);
    long long t = i / Cc;
    const int ow = static_cast<int>(t % Wo);
    t /= Wo;
    const int oh = static_cast<int>(t % Ho);
    const long long b = t / Ho;
    float m = -INFINITY;
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s) {
        const int ih = 2 * oh + r - 1, iw = 2 * ow + s - 1;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) m = fmaxf(m, in[((b * H + ih) * W + iw) * Cc + c]);
      }
    out[i] = m;
  }
}

// backward of MaxPool2d(3, 2, 1): each input position gathers the gradient of the (at most four) windows whose FIRST maximum
// (row-major scan, like torch) it is
__global__ void __launch_bounds__(256) maxpool3x3s2_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ in,
                                                               float* __restrict__ g_in, int B, int H, int W, int Cc) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = static_cast<long long>(B) * H * W * Cc;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cc);
    long long t = i / Cc;
    const int iw = static_cast<int>(t % W);
    t /= W;
    const int ih = static_cast<int>(t % H);
    const long long b = t / H;
    float acc = 0.f;
    for (int oh = (ih + 1) / 2 - 1; oh <= (ih + 1) / 2; ++oh) {       // windows with 2 oh - 1 <= ih <= 2 oh + 1
      if (oh < 0 || oh >= Ho || ih < 2 * oh - 1 || ih > 2 * oh + 1) continue;
      for (int ow = (iw + 1) / 2 - 1; ow <= (iw + 1) / 2; ++ow) {
        if (ow < 0 || ow >= Wo || iw < 2 * ow - 1 || iw > 2 * ow + 1) continue;
        float m = -INFINITY;
        int ar = -1, as = -1;
        for (int r = 0; r < 3; ++r)
          for (int q = 0; q < 3; ++q) {
            const int jh = 2 * oh + r - 1, jw = 2 * ow + q - 1;
            if (jh < 0 || jh >= H || jw < 0 || jw >= W) continue;
            const float v = in[((b * H + jh) * W + jw) * Cc + c];
            if (v > m) m = v, ar = jh, as = jw;
          }
        if (ar == ih && as == iw) acc += g_out[((b * Ho + oh) * Wo + ow) * Cc + c];
      }
    }
    g_in[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------- KWS (RCNN + attention)
// One CTA per sample; everything lives in shared memory (W <= 512 spectrogram frames).
struct KwsWeights {
  const float *dw_w, *dw_b, *pw_w, *pw_b;                 // sepconv.0 (32,1,5) / sepconv.1 (64,32,1)
  const float* gru[2][2][4];                              // [layer][dir]{w_ih, w_hh, b_ih, b_hh}
  const float *wx_w, *wx_b, *vt_w, *u_w;                  // attention
};
__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256) kws_kernel(const float* __restrict__ spec, int Wf, int in_size, int H,
                                                  int num_classes, KwsWeights w, float* __restrict__ out) {
  extern __shared__ float sm[];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int W1 = (Wf - 5) / 2 + 1;          // depthwise conv k=5 stride 2          (model.py:7-9)
  const int T = (W1 - 1) / 8 + 1;           // pointwise conv k=1 stride 8          (model.py:10-11)
  float* dw = sm;                           // [in_size][W1] (only columns 8*t are needed, kept simple)
  float* x0 = dw + in_size * W1;            // [T][H]
  float* x1 = x0 + T * H;                   // [T][2H]
  float* x2 = x1 + T * 2 * H;               // [T][2H]
  float* hbuf = x2 + T * 2 * H;             // [2 dirs][H] hidden + [2][3H] gates scratch
  float* gi = hbuf + 2 * H;                 // [2][3H]
  float* gh = gi + 2 * 3 * H;               // [2][3H]
  float* att = gh + 2 * 3 * H;              // [T]
  const float* sp = spec + static_cast<long long>(b) * in_size * Wf;
  for (int i = tid; i < in_size * W1; i += nt) {
    const int c = i / W1, t = i - c * W1;
    float acc = w.dw_b[c];
    for (int k = 0; k < 5; ++k) acc = fmaf(w.dw_w[c * 5 + k], sp[c * Wf + 2 * t + k], acc);
    dw[i] = acc;
  }
  __syncthreads();
  for (int i = tid; i < T * H; i += nt) {
    const int t = i / H, o = i - t * H;
    float acc = w.pw_b[o];
    for (int c = 0; c < in_size; ++c) acc = fmaf(w.pw_w[o * in_size + c], dw[c * W1 + 8 * t], acc);
    x0[i] = acc;
  }
  __syncthreads();
  const float* xin = x0;
  int in_l = H;
  float* xout = x1;
  for (int layer = 0; layer < 2; ++layer) {
    for (int i = tid; i < 2 * H; i += nt) hbuf[i] = 0.f;
    __syncthreads();
    for (int step = 0; step < T; ++step) {
      // both directions in parallel: dir 0 reads time `step`, dir 1 reads time T-1-step
      for (int i = tid; i < 2 * 3 * H; i += nt) {
        const int dir = i / (3 * H), g = i - dir * 3 * H;
        const int t = dir ? T - 1 - step : step;
        const float* wih = w.gru[layer][dir][0] + static_cast<long long>(g) * in_l;
        const float* whh = w.gru[layer][dir][1] + static_cast<long long>(g) * H;
        float a = w.gru[layer][dir][2][g], c = w.gru[layer][dir][3][g];
        for (int k = 0; k < in_l; ++k) a = fmaf(wih[k], xin[t * in_l + k], a);
        for (int k = 0; k < H; ++k) c = fmaf(whh[k], hbuf[dir * H + k], c);
        gi[i] = a, gh[i] = c;
      }
      __syncthreads();
      for (int i = tid; i < 2 * H; i += nt) {
        const int dir = i / H, j = i - dir * H;
        const int t = dir ? T - 1 - step : step;
        const float* a = gi + dir * 3 * H;
        const float* c = gh + dir * 3 * H;
        const float r = sigm(a[j] + c[j]), z = sigm(a[H + j] + c[H + j]);
        const float n = tanhf(a[2 * H + j] + r * c[2 * H + j]);
        const float hn = (1.f - z) * n + z * hbuf[i];
        hbuf[i] = hn;
        xout[t * 2 * H + dir * H + j] = hn;
      }
      __syncthreads();
    }
    xin = xout, in_l = 2 * H, xout = x2;
  }
  // attention (model.py:38-62,105-108): e_t = Vt . tanh(Wx x_t + b); a = softmax_t(e); c = sum_t a_t x_t; out = log_softmax(U c)
  const float* xf = xin;
  const int warp = tid >> 5, lane = tid & 31;
  for (int t = warp; t < T; t += nt >> 5) {
    float acc = 0.f;
    for (int o = lane; o < 2 * H; o += 32) {
      float d = w.wx_b[o];
      for (int k = 0; k < 2 * H; ++k) d = fmaf(w.wx_w[o * 2 * H + k], xf[t * 2 * H + k], d);
      acc = fmaf(w.vt_w[o], tanhf(d), acc);
    }
    for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) att[t] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    float mx = att[0];
    for (int t = 1; t < T; ++t) mx = fmaxf(mx, att[t]);
    float s = 0.f;
    for (int t = 0; t < T; ++t) att[t] = expf(att[t] - mx), s += att[t];
    for (int t = 0; t < T; ++t) att[t] /= s;
  }
  __syncthreads();
  float* ctx = gi;  // reuse: [2H]
  for (int j = tid; j < 2 * H; j += nt) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc = fmaf(att[t], xf[t * 2 * H + j], acc);
    ctx[j] = acc;
  }
  __syncthreads();
  float* lg = gh;
  for (int k = warp; k < num_classes; k += nt >> 5) {
    float acc = 0.f;
    for (int j = lane; j < 2 * H; j += 32) acc = fmaf(w.u_w[k * 2 * H + j], ctx[j], acc);
    for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) lg[k] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    float mx = lg[0];
    for (int k = 1; k < num_classes; ++k) mx = fmaxf(mx, lg[k]);
    float s = 0.f;
    for (int k = 0; k < num_classes; ++k) s += expf(lg[k] - mx);
    const float off = mx + logf(s);
    for (int k = 0; k < num_classes; ++k) out[b * num_classes + k] = lg[k] - off;
  }
}

// ---- backward of the KWS forward (autograd over audio_models/RCNN_KWS/model.py:5-113): one CTA per sample recomputes the
//      forward with the gate values of every GRU step kept in a global scratch area, then runs attention / BPTT / separable
//      convolution backward.  g_spec = (d logp / d spec)^T g_out.
struct KwsScratch {   // offsets (floats) into the per-sample scratch block
  int dw, x0, x1, x2, sv, th, att, ctx, lg, g_x2, g_x1, g_x0, g_dw, total;
};
__host__ __device__ inline KwsScratch kws_scratch(int I, int W1, int T, int H, int K) {
  KwsScratch s;
  int o = 0;
  s.dw = o, o += I * W1;
  s.x0 = o, o += T * H;
  s.x1 = o, o += T * 2 * H;
  s.x2 = o, o += T * 2 * H;
  s.sv = o, o += 2 * 2 * T * 5 * H;     // [layer][dir][step]{r, z, n, h_prev, gh_n}[H]
  s.th = o, o += T * 2 * H;             // tanh(Wx x_t + b)
  s.att = o, o += T;
  s.ctx = o, o += 2 * H;
  s.lg = o, o += K;
  s.g_x2 = o, o += T * 2 * H;
  s.g_x1 = o, o += T * 2 * H;
  s.g_x0 = o, o += T * H;
  s.g_dw = o, o += I * W1;
  s.total = (o + 3) & ~3;
  return s;
}

__global__ void __launch_bounds__(256) kws_vjp_kernel(const float* __restrict__ spec, int Wf, int in_size, int H, int num_classes,
                                                      KwsWeights w, const float* __restrict__ g_out, float* __restrict__ scratch,
                                                      float* __restrict__ g_spec) {
  extern __shared__ float sm[];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int W1 = (Wf - 5) / 2 + 1, T = (W1 - 1) / 8 + 1;
  const KwsScratch so = kws_scratch(in_size, W1, T, H, num_classes);
  float* S = scratch + static_cast<long long>(b) * so.total;
  float *dw = S + so.dw, *x0 = S + so.x0, *x1 = S + so.x1, *x2 = S + so.x2, *sv = S + so.sv, *th = S + so.th, *att = S + so.att,
        *ctx = S + so.ctx, *lg = S + so.lg, *g_x2 = S + so.g_x2, *g_x1 = S + so.g_x1, *g_x0 = S + so.g_x0, *g_dw = S + so.g_dw;
  float* hbuf = sm;                 // [2][H]   hidden state (forward) / carried gradient (backward)
  float* gi = hbuf + 2 * H;         // [2][3H]
  float* gh = gi + 2 * 3 * H;       // [2][3H]
  float* red = gh + 2 * 3 * H;      // [64] small reductions
  const float* sp = spec + static_cast<long long>(b) * in_size * Wf;
  const int warp = tid >> 5, lane = tid & 31;

  // ------------------------------------------------------------------ forward (as kws_kernel), keeping the gate values
  for (int i = tid; i < in_size * W1; i += nt) {
    const int c = i / W1, t = i - c * W1;
    float acc = w.dw_b[c];
    for (int k = 0; k < 5; ++k) acc = fmaf(w.dw_w[c * 5 + k], sp[c * Wf + 2 * t + k], acc);
    dw[i] = acc;
  }
  __syncthreads();
  for (int i = tid; i < T * H; i += nt) {
    const int t = i / H, o = i - t * H;
    float acc = w.pw_b[o];
    for (int c = 0; c < in_size; ++c) acc = fmaf(w.pw_w[o * in_size + c], dw[c * W1 + 8 * t], acc);
    x0[i] = acc;
  }
  __syncthreads();
  {
    const float* xin = x0;
    int in_l = H;
    float* xout = x1;
    for (int layer = 0; layer < 2; ++layer) {
      for (int i = tid; i < 2 * H; i += nt) hbuf[i] = 0.f;
      __syncthreads();
      for (int step = 0; step < T; ++step) {
        for (int i = tid; i < 2 * 3 * H; i += nt) {
          const int dir = i / (3 * H), g = i - dir * 3 * H;
          const int t = dir ? T - 1 - step : step;
          const float* wih = w.gru[layer][dir][0] + static_cast<long long>(g) * in_l;
          const float* whh = w.gru[layer][dir][1] + static_cast<long long>(g) * H;
          float a = w.gru[layer][dir][2][g], c = w.gru[layer][dir][3][g];
          for (int k = 0; k < in_l; ++k) a = fmaf(wih[k], xin[t * in_l + k], a);
          for (int k = 0; k < H; ++k) c = fmaf(whh[k], hbuf[dir * H + k], c);
          gi[i] = a, gh[i] = c;
        }
        __syncthreads();
        for (int i = tid; i < 2 * H; i += nt) {
          const int dir = i / H, j = i - dir * H;
          const int t = dir ? T - 1 - step : step;
          const float* a = gi + dir * 3 * H;
          const float* c = gh + dir * 3 * H;
          const float r = sigm(a[j] + c[j]), z = sigm(a[H + j] + c[H + j]);
          const float n = tanhf(a[2 * H + j] + r * c[2 * H + j]);
          const float hp = hbuf[i];
          const float hn = (1.f - z) * n + z * hp;
          float* q = sv + (((layer * 2 + dir) * T + step) * 5) * H + j;
          q[0] = r, q[H] = z, q[2 * H] = n, q[3 * H] = hp, q[4 * H] = c[2 * H + j];
          hbuf[i] = hn;
          xout[t * 2 * H + dir * H + j] = hn;
        }
        __syncthreads();
      }
      xin = xout, in_l = 2 * H, xout = x2;
    }
  }
  const float* xf = x2;
  for (int t = warp; t < T; t += nt >> 5) {
    float acc = 0.f;
    for (int o = lane; o < 2 * H; o += 32) {
      float d = w.wx_b[o];
      for (int k = 0; k < 2 * H; ++k) d = fmaf(w.wx_w[o * 2 * H + k], xf[t * 2 * H + k], d);
      const float tv = tanhf(d);
      th[t * 2 * H + o] = tv;
      acc = fmaf(w.vt_w[o], tv, acc);
    }
    for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) att[t] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    float mx = att[0];
    for (int t = 1; t < T; ++t) mx = fmaxf(mx, att[t]);
    float s = 0.f;
    for (int t = 0; t < T; ++t) att[t] = expf(att[t] - mx), s += att[t];
    for (int t = 0; t < T; ++t) att[t] /= s;
  }
  __syncthreads();
  for (int j = tid; j < 2 * H; j += nt) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc = fmaf(att[t], xf[t * 2 * H + j], acc);
    ctx[j] = acc;
  }
  __syncthreads();
  for (int k = warp; k < num_classes; k += nt >> 5) {
    float acc = 0.f;
    for (int j = lane; j < 2 * H; j += 32) acc = fmaf(w.u_w[k * 2 * H + j], ctx[j], acc);
    for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) lg[k] = acc;
  }
  __syncthreads();

  // ------------------------------------------------------------------ backward
  // log_softmax: g_lg = g_out - softmax * sum(g_out)      (red[0..K) = g_lg)
  if (tid == 0) {
    float mx = lg[0];
    for (int k = 1; k < num_classes; ++k) mx = fmaxf(mx, lg[k]);
    float s = 0.f, gs = 0.f;
    for (int k = 0; k < num_classes; ++k) s += expf(lg[k] - mx), gs += g_out[b * num_classes + k];
    for (int k = 0; k < num_classes; ++k) red[k] = g_out[b * num_classes + k] - expf(lg[k] - mx) / s * gs;
  }
  __syncthreads();
  // g_ctx = U^T g_lg  (kept in gi[0..2H))
  float* g_ctx = gi;
  for (int j = tid; j < 2 * H; j += nt) {
    float acc = 0.f;
    for (int k = 0; k < num_classes; ++k) acc = fmaf(w.u_w[k * 2 * H + j], red[k], acc);
    g_ctx[j] = acc;
  }
  __syncthreads();
  // g_att[t] = g_ctx . x_t   (gh[0..T)); g_x2[t] = att[t] * g_ctx
  float* g_att = gh;
  for (int t = warp; t < T; t += nt >> 5) {
    float acc = 0.f;
    for (int j = lane; j < 2 * H; j += 32) {
      acc = fmaf(g_ctx[j], xf[t * 2 * H + j], acc);
      g_x2[t * 2 * H + j] = att[t] * g_ctx[j];
    }
    for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) g_att[t] = acc;
  }
  __syncthreads();
  if (tid == 0) {   // softmax over time: g_e = att * (g_att - sum att g_att)   (in place in g_att)
    float dot = 0.f;
    for (int t = 0; t < T; ++t) dot = fmaf(att[t], g_att[t], dot);
    for (int t = 0; t < T; ++t) g_att[t] = att[t] * (g_att[t] - dot);
  }
  __syncthreads();
  // e_t = Vt . tanh(d_t), d_t = Wx x_t + b:  g_d[t][o] = g_e[t] Vt[o] (1 - th^2)  (overwrites th);  g_x2[t] += Wx^T g_d[t]
  for (int i = tid; i < T * 2 * H; i += nt) {
    const int t = i / (2 * H), o = i - t * 2 * H;
    const float tv = th[i];
    th[i] = g_att[t] * w.vt_w[o] * (1.f - tv * tv);
  }
  __syncthreads();
  for (int i = tid; i < T * 2 * H; i += nt) {
    const int t = i / (2 * H), k = i - t * 2 * H;
    float acc = 0.f;
    for (int o = 0; o < 2 * H; ++o) acc = fmaf(w.wx_w[o * 2 * H + k], th[t * 2 * H + o], acc);
    g_x2[i] += acc;
  }
  for (int i = tid; i < T * 2 * H; i += nt) g_x1[i] = 0.f;
  for (int i = tid; i < T * H; i += nt) g_x0[i] = 0.f;
  __syncthreads();
  // GRU layers, back-propagation through time (both directions in parallel, steps in reverse order)
  for (int layer = 1; layer >= 0; --layer) {
    const int in_l = layer == 0 ? H : 2 * H;
    const float* g_y = layer == 1 ? g_x2 : g_x1;       // gradient wrt this layer's outputs
    float* g_in = layer == 1 ? g_x1 : g_x0;            // gradient wrt this layer's inputs (accumulated)
    for (int i = tid; i < 2 * H; i += nt) hbuf[i] = 0.f;   // carried d loss / d h
    __syncthreads();
    for (int step = T - 1; step >= 0; --step) {
      for (int i = tid; i < 2 * H; i += nt) {
        const int dir = i / H, j = i - dir * H;
        const int t = dir ? T - 1 - step : step;
        const float* q = sv + (((layer * 2 + dir) * T + step) * 5) * H + j;
        const float r = q[0], z = q[H], n = q[2 * H], hp = q[3 * H], ghn = q[4 * H];
        const float gh_tot = hbuf[i] + g_y[t * 2 * H + dir * H + j];
        const float g_an = gh_tot * (1.f - z) * (1.f - n * n);
        const float g_ar = g_an * ghn * r * (1.f - r);
        const float g_az = gh_tot * (hp - n) * z * (1.f - z);
        float* a = gi + dir * 3 * H;
        float* c = gh + dir * 3 * H;
        a[j] = g_ar, a[H + j] = g_az, a[2 * H + j] = g_an;
        c[j] = g_ar, c[H + j] = g_az, c[2 * H + j] = g_an * r;
        hbuf[i] = gh_tot * z;                               // direct path to h_prev; the W_hh path is added below
      }
      __syncthreads();
      for (int i = tid; i < 2 * H; i += nt) {               // d h_prev += W_hh^T g_gh
        const int dir = i / H, j = i - dir * H;
        const float* whh = w.gru[layer][dir][1];
        const float* c = gh + dir * 3 * H;
        float acc = 0.f;
        for (int g = 0; g < 3 * H; ++g) acc = fmaf(whh[static_cast<long long>(g) * H + j], c[g], acc);
        hbuf[i] += acc;
      }
      for (int i = tid; i < 2 * in_l; i += nt) {            // d x_t += W_ih^T g_gi (the two directions may hit the same t)
        const int dir = i / in_l, k = i - dir * in_l;
        const int t = dir ? T - 1 - step : step;
        const float* wih = w.gru[layer][dir][0];
        const float* a = gi + dir * 3 * H;
        float acc = 0.f;
        for (int g = 0; g < 3 * H; ++g) acc = fmaf(wih[static_cast<long long>(g) * in_l + k], a[g], acc);
        atomicAdd(g_in + t * in_l + k, acc);
      }
      __syncthreads();
    }
  }
  // separable convolution: x0[t][o] = pw_b[o] + sum_c pw_w[o][c] dw[c][8 t];  dw[c][t'] = dw_b[c] + sum_k dw_w[c][k] spec[c][2 t' + k]
  for (int i = tid; i < in_size * T; i += nt) {
    const int c = i / T, t = i - c * T;
    float acc = 0.f;
    for (int o = 0; o < H; ++o) acc = fmaf(w.pw_w[o * in_size + c], g_x0[t * H + o], acc);
    g_dw[c * W1 + 8 * t] = acc;
  }
  float* gs = g_spec + static_cast<long long>(b) * in_size * Wf;
  for (int i = tid; i < in_size * Wf; i += nt) gs[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < in_size * T * 5; i += nt) {       // windows of different t do not overlap (16 columns apart, 5 wide)
    const int c = i / (T * 5), rem = i - c * T * 5, t = rem / 5, k = rem - t * 5;
    gs[c * Wf + 16 * t + k] = w.dw_w[c * 5 + k] * g_dw[c * W1 + 8 * t];
  }
}

// NCHW with C == 1 is already NHWC; this transposes (B, C, W) -> (B, W, C) for the M5 input when C > 1 (unused: C == 1)

}  // namespace ap

using namespace ap;

struct Bottleneck {
  ConvLayer reduce, conv, expand, shortcut;
  ConvLayer t_reduce, t_conv, t_expand, t_shortcut;   // data-gradient twins (built by the first ap_classifier_vjp call)
  bool has_shortcut = false;
  int stride = 1, D = 0, cout = 0, cin = 0;
};

struct ConvStep {   // one convolution of the ResNeXt forward bound to workspace buffers
  const ConvLayer* L;
  int in, out, res;     // workspace buffer indices (res < 0: none)
  int H, W, relu;
  bool tc;
  ConvTcBinding bnd;
};

struct ap_classifier_s {
  ap_classifier_cfg cfg{};
  int device = 0;
  int mode = AP_MODE_FP32;                        // AP_MODE_TF32: tensor-core convolutions where the shape allows
  unsigned tc_mask = 7;                           // bit 0: 1x1, bit 1: 3x3 stride 1, bit 2: strided (AP_CLS_TC_MASK, bisecting aid)
  std::map<int, std::vector<ConvStep>> plans;     // keyed by the number of images in the chunk
  struct TcSlot {   // one convolution of a VGG / WideResNet pass bound to the buffers it was last run on
    const float* in = nullptr;
    float* out = nullptr;
    const float* res = nullptr;
    int relu = -1, bn = 0;
    bool tc = false;
    ConvTcBinding bnd;
  };
  std::vector<TcSlot> tc_slots;
  // ResNeXt
  ConvLayer stem;
  std::vector<std::unique_ptr<Bottleneck>> blocks;
  DevBuf fc_w, fc_b;
  int feat = 0;
  // ResNet family
  struct ResBlock {
    ConvLayer c1, c2, c3, down;
    ConvLayer t_c1, t_c2, t_c3, t_down;   // data-gradient twins (backward pass)
    bool bottleneck = false, has_down = false;
    int stride = 1, cout = 0, cin = 0, planes = 0;
  };
  std::vector<std::unique_ptr<ResBlock>> resblocks;
  // M5
  ConvLayer m5conv[4], m5conv_t[4];   // forward layers and their data-gradient twins
  DevBuf m5_w0;                       // first layer's folded weights [c][j] for conv1d_cin1_dgrad_kernel
  // WideResNet: pre-activation blocks; bn1 of every block (and the final BatchNorm) as device scale / shift vectors
  struct WrnBlock {
    ConvLayer c1, c2, sc, t_c1, t_c2, t_sc;
    DevBuf scale, shift;
    bool equal = true;
    int stride = 1, cin = 0, cout = 0;
  };
  std::vector<std::unique_ptr<WrnBlock>> wrn;
  DevBuf wrn_scale, wrn_shift;   // final BatchNorm
  // DenseNet-BC: three dense blocks; each keeps its concatenated activations in ONE (B, H, W, stride) tensor whose first
  // C0p = round4(C0) channels are the block input (padding channels stay zero) followed by `growth` channels per layer
  struct DnLayer {
    DevBuf scale, shift;                 // bn1 over the layer's input prefix, in the physical channel layout
    ConvLayer c1, c2, t_c1, t_c2;        // 1x1 (bn2 folded, ReLU) and 3x3; data-gradient twins
    int cp = 0, off = 0;                 // physical input channels; where this layer's `growth` outputs go
  };
  struct DnBlock {
    int C0 = 0, C0p = 0, stride = 0, H = 0;
    std::vector<std::unique_ptr<DnLayer>> layers;
    DevBuf t_scale, t_shift;             // transition after the block (blocks 0, 1) or the final BatchNorm (block 2)
    ConvLayer t_conv, t_conv_t;
    int t_cout = 0;
  };
  std::vector<std::unique_ptr<DnBlock>> dn;
  int dn_growth = 0, dn_mid = 0;
  DevBuf dn_x[3], dn_gx[3], dn_a, dn_h, dn_t;
  size_t dn_elems = 0;
  // VGG (batch-norm variants): `vgg_plan` lists output channels per convolution, -1 for a max-pool; three fully connected layers
  std::vector<int> vgg_plan;
  std::vector<std::unique_ptr<ConvLayer>> vgg_conv, vgg_conv_t;
  ConvLayer vgg_fc[3], vgg_fc_t[3];
  // KWS
  std::vector<std::unique_ptr<DevBuf>> kws_bufs;
  KwsWeights kws{};
  // workspace
  DevBuf buf[5];
  size_t buf_elems = 0;
  int chunk = 0;
  // backward pass (ResNeXt): transposed stem, the tape of saved activations and five gradient buffers for `bwd_bn` images
  ConvLayer t_stem;
  bool bwd_ready = false;
  int bwd_bn = 0;
  DevBuf tape_x0, gbuf[5];
  std::vector<std::unique_ptr<DevBuf>> tape;   // per block: r, c, y
};

static int ensure_ws(ap_classifier_t h, size_t elems) {
  if (elems <= h->buf_elems) return AP_OK;
  h->buf_elems = 0;   // alloc() releases the old buffers first: a failure must not leave the old size behind
  for (auto& b : h->buf) AP_CUDA(b.alloc(elems * sizeof(float)));
  h->buf_elems = elems;
  return AP_OK;
}

static int create_resnext(ap_classifier_t h, const float* const* w, int n_weights) {
  const ap_classifier_cfg& c = h->cfg;
  AP_REQUIRE(c.cardinality > 0 && c.base_width > 0 && c.widen_factor > 0 && c.depth >= 11 && c.in_channels == 1,
             "ap_classifier_create: bad ResNeXt configuration (in_channels must be 1: NCHW == NHWC)");
  const int block_depth = (c.depth - 2) / 9;
  const int stages[4] = {64, 64 * c.widen_factor, 128 * c.widen_factor, 256 * c.widen_factor};
  int expected = 5 + 2;
  for (int s = 0; s < 3; ++s)
    for (int b = 0; b < block_depth; ++b) expected += 15 + ((b == 0 && stages[s] != stages[s + 1]) ? 5 : 0);
  AP_REQUIRE(n_weights == expected, "ap_classifier_create: ResNeXt expects %d weight tensors, got %d", expected, n_weights);
  int i = 0;
  h->stem.keep_host = true;
  int rc = h->stem.init(c.in_channels, 64, 3, 3, 1, 1, 1, w[0], nullptr, w[1], w[2], w[3], w[4]);
  if (rc != AP_OK) return rc;
  i = 5;
  for (int s = 0; s < 3; ++s)
    for (int b = 0; b < block_depth; ++b) {
      auto blk = std::make_unique<Bottleneck>();
      const int cin = b == 0 ? stages[s] : stages[s + 1], cout = stages[s + 1];
      const int stride = (b == 0 && s > 0) ? 2 : 1;
      const double width_ratio = cout / (c.widen_factor * 64.0);
      const int D = c.cardinality * static_cast<int>(c.base_width * width_ratio);
      blk->stride = stride, blk->D = D, blk->cout = cout, blk->cin = cin;
      blk->reduce.keep_host = blk->conv.keep_host = blk->expand.keep_host = blk->shortcut.keep_host = true;
      rc = blk->reduce.init(cin, D, 1, 1, 1, 0, 1, w[i], nullptr, w[i + 1], w[i + 2], w[i + 3], w[i + 4], true);
      if (rc == AP_OK)
        rc = blk->conv.init(D, D, 3, 3, stride, 1, c.cardinality, w[i + 5], nullptr, w[i + 6], w[i + 7], w[i + 8], w[i + 9], true);
      if (rc == AP_OK)
        rc = blk->expand.init(D, cout, 1, 1, 1, 0, 1, w[i + 10], nullptr, w[i + 11], w[i + 12], w[i + 13], w[i + 14], true);
      i += 15;
      if (rc == AP_OK && cin != cout) {
        blk->has_shortcut = true;
        rc = blk->shortcut.init(cin, cout, 1, 1, stride, 0, 1, w[i], nullptr, w[i + 1], w[i + 2], w[i + 3], w[i + 4], true);
        i += 5;
      }
      if (rc != AP_OK) return rc;
      h->blocks.push_back(std::move(blk));
    }
  h->feat = stages[3];
  AP_CUDA(h->fc_w.upload(w[i], sizeof(float) * c.num_classes * h->feat));
  AP_CUDA(h->fc_b.upload(w[i + 1], sizeof(float) * c.num_classes));
  return AP_OK;
}

// Build (once per chunk size) the list of convolutions bound to the five workspace buffers:
//   0/1 = block input / output (ping-pong), 2 = reduce output, 3 = grouped-conv output, 4 = shortcut
static int build_plan(ap_classifier_t h, int bn, std::vector<ConvStep>& plan) {
  plan.clear();
  int H = 32, W = 32, x = 0, xo = 1;
  auto add = [&](const ConvLayer* L, int in, int out, int res, int Hin, int Win, int relu) -> int {
    ConvStep st{};
    st.L = L, st.in = in, st.out = out, st.res = res, st.H = Hin, st.W = Win, st.relu = relu;
    const unsigned kind = (L->stride != 1) ? 4u : (L->kh == 1 ? 1u : 2u);
    st.tc = h->mode == AP_MODE_TF32 && L->has_tc && (h->tc_mask & kind) &&
            conv_tc_supported(L->Cin, L->Cout, L->groups, Hin, Win, L->kh, L->kw, L->stride, L->pad);
    if (st.tc) {
      int rc = L->tc.bind(&st.bnd, h->buf[in].as<float>(), bn, Hin, Win, h->buf[out].as<float>(),
                          res >= 0 ? h->buf[res].as<float>() : nullptr, relu, 1);
      if (rc != AP_OK) return rc;
    }
    plan.push_back(st);
    return AP_OK;
  };
  for (auto& blk : h->blocks) {
    const int Ho = (H - 1) / blk->stride + 1, Wo = (W - 1) / blk->stride + 1;
    int rc = add(&blk->reduce, x, 2, -1, H, W, 1);
    if (rc == AP_OK) rc = add(&blk->conv, 2, 3, -1, H, W, 1);
    int res = x;
    if (rc == AP_OK && blk->has_shortcut) {
      rc = add(&blk->shortcut, x, 4, -1, H, W, 0);
      res = 4;
    }
    if (rc == AP_OK) rc = add(&blk->expand, 3, xo, res, Ho, Wo, 1);
    if (rc != AP_OK) return rc;
    std::swap(x, xo);
    H = Ho, W = Wo;
  }
  AP_REQUIRE(H == 8 && W == 8, "ResNeXt: unexpected final spatial size %dx%d", H, W);
  return AP_OK;
}

static int forward_resnext(ap_classifier_t h, const float* spec, float* logits, int B, cudaStream_t st) {
  // spatial size fixed by avg_pool2d(x, 8, 1) + view(-1, C): 32x32 input -> 8x8 after two stride-2 stages
  const int H0 = 32, W0 = 32;
  int maxc = 64;
  for (auto& b : h->blocks) maxc = std::max(maxc, std::max(b->D, b->cout));
  static const int chunk = [] {   // images per pass through the five activation buffers (AP_CLS_CHUNK overrides)
    const char* e = std::getenv("AP_CLS_CHUNK");
    const int v = e ? std::atoi(e) : 0;
    return v > 0 ? v : 256;   // 64 -> 256 images: 15.8 -> 13.6 ms per 512 images (fewer ragged waves in the 8x8 stage)
  }();
  const size_t need = static_cast<size_t>(std::min(B, chunk)) * H0 * W0 * maxc;
  if (need > h->buf_elems) h->plans.clear();     // buffers move: the tensor maps must be re-encoded
  int rc = ensure_ws(h, need);
  if (rc != AP_OK) return rc;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bn = std::min(chunk, B - b0);
    auto it = h->plans.find(bn);
    if (it == h->plans.end()) {
      std::vector<ConvStep> plan;
      rc = build_plan(h, bn, plan);
      if (rc != AP_OK) return rc;
      it = h->plans.emplace(bn, std::move(plan)).first;
    }
    rc = h->stem.run(spec + static_cast<size_t>(b0) * h->cfg.in_channels * H0 * W0, bn, H0, W0, h->buf[0].as<float>(), nullptr, 1, st);
    if (rc != AP_OK) return rc;
    int last_out = 0;
    for (const ConvStep& cs : it->second) {
      if (cs.tc) rc = cs.L->tc.run(cs.bnd, st);
      else
        rc = cs.L->run(h->buf[cs.in].as<float>(), bn, cs.H, cs.W, h->buf[cs.out].as<float>(),
                       cs.res >= 0 ? h->buf[cs.res].as<float>() : nullptr, cs.relu, st);
      if (rc != AP_OK) return rc;
      last_out = cs.out;
    }
    const size_t smem = sizeof(float) * (h->feat + h->cfg.num_classes);
    pool_fc_kernel<<<bn, 256, smem, st>>>(h->buf[last_out].as<float>(), 64, h->feat, h->fc_w.as<float>(), h->fc_b.as<float>(),
                                          h->cfg.num_classes, logits + static_cast<size_t>(b0) * h->cfg.num_classes, 0);
    AP_LAUNCH_CHECK();
  }
  return AP_OK;
}

// ---- backward of the ResNeXt forward: g_spec = (d logits / d spec)^T g_logits (autograd over resnext.py:56-64,134-142 with
//      BatchNorm in eval mode).  The forward is recomputed on the fp32 FFMA path with every ReLU output kept (the tape);
//      each convolution's data gradient is the forward convolution of the (zero-upsampled, for stride 2) output gradient with
//      the transposed, 180-degree-rotated folded weights.
static int vjp_resnext(ap_classifier_t h, const float* spec, const float* g_logits, float* g_spec, int B, cudaStream_t st) {
  const int H0 = 32, W0 = 32, chunk = 32;
  if (!h->bwd_ready) {
    int rc = init_dgrad(h->t_stem, h->stem);
    for (auto& b : h->blocks) {
      if (rc == AP_OK) rc = init_dgrad(b->t_reduce, b->reduce, true);
      if (rc == AP_OK) rc = init_dgrad(b->t_conv, b->conv, true);
      if (rc == AP_OK) rc = init_dgrad(b->t_expand, b->expand, true);
      if (rc == AP_OK && b->has_shortcut) rc = init_dgrad(b->t_shortcut, b->shortcut, true);
    }
    if (rc != AP_OK) return rc;
    h->bwd_ready = true;
  }
  const int bn_max = std::min(B, chunk);
  if (bn_max > h->bwd_bn) {   // (re)allocate the tape and the gradient buffers
    size_t max_e = static_cast<size_t>(H0) * W0 * 64;
    int H = H0, W = W0;
    h->tape.clear();
    AP_CUDA(h->tape_x0.alloc(static_cast<size_t>(bn_max) * H0 * W0 * 64 * sizeof(float)));
    for (auto& b : h->blocks) {
      const int Ho = (H - 1) / b->stride + 1, Wo = (W - 1) / b->stride + 1;
      const size_t er = static_cast<size_t>(H) * W * b->D, ec = static_cast<size_t>(Ho) * Wo * b->D,
                   ey = static_cast<size_t>(Ho) * Wo * b->cout, ex = static_cast<size_t>(H) * W * std::max(b->cin, b->cout);
      max_e = std::max(std::max(max_e, er), std::max(std::max(ec, ey), ex));
      for (size_t e : {er, ec, ey}) {
        auto d = std::make_unique<DevBuf>();
        AP_CUDA(d->alloc(e * bn_max * sizeof(float)));
        h->tape.push_back(std::move(d));
      }
      H = Ho, W = Wo;
    }
    for (auto& g : h->gbuf) AP_CUDA(g.alloc(max_e * bn_max * sizeof(float)));
    h->bwd_bn = bn_max;
  }
  auto grid_for = [](long long work) {
    long long b = ceil_div_ll(work, 256);
    const long long cap = static_cast<long long>(num_sms()) * 16;
    return static_cast<unsigned>(b < cap ? (b > 0 ? b : 1) : cap);
  };
  auto mask = [&](float* g, const float* act, size_t elems) -> int {
    relu_mask_kernel<<<grid_for(static_cast<long long>(elems / 4)), 256, 0, st>>>(reinterpret_cast<float4*>(g),
                                                                                  reinterpret_cast<const float4*>(act),
                                                                                  static_cast<long long>(elems / 4));
    AP_LAUNCH_CHECK();
    return AP_OK;
  };
  auto upsample = [&](const float* g, float* up, int bn, int Ho, int Wo, int Cc) -> int {
    upsample2_kernel<<<grid_for(static_cast<long long>(bn) * 4 * Ho * Wo * (Cc / 4)), 256, 0, st>>>(
        reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(up), bn, Ho, Wo, Cc / 4);
    AP_LAUNCH_CHECK();
    return AP_OK;
  };
  // AP_MODE_TF32: the recomputed forward and the data-gradient convolutions run on the tcgen05 kind::tf32 kernel wherever the
  // shape allows (outputs rounded to tf32 for the next layer, like the inference path); AP_MODE_FP32: everything on FFMA
  // (a ReLU network's gradient is discontinuous in the forward values: the tf32 forward flips ~0.1 % of the masks per layer
  // relative to an fp32 forward, which moves the gradient by ~6 %; AP_CLS_VJP_FWD_FP32=1 keeps the recomputed forward on
  // FFMA so that tests can check the tensor-core data-gradient kernels alone against fp32 autograd)
  const bool tc_bwd = h->mode == AP_MODE_TF32;
  const char* env_fwd = std::getenv("AP_CLS_VJP_FWD_FP32");
  const bool tc_fwd = tc_bwd && !(env_fwd && env_fwd[0] == '1');
  bool use_tc = tc_fwd;
  int bn = 0;
  auto conv = [&](const ConvLayer& Lr, const float* in, int Hh, int Ww, float* out, const float* res, int relu) -> int {
    if (use_tc && Lr.has_tc && conv_tc_supported(Lr.Cin, Lr.Cout, Lr.groups, Hh, Ww, Lr.kh, Lr.kw, Lr.stride, Lr.pad)) {
      ConvTcBinding bnd;
      int rc = Lr.tc.bind(&bnd, in, bn, Hh, Ww, out, res, relu, 1);
      return rc != AP_OK ? rc : Lr.tc.run(bnd, st);
    }
    return Lr.run(in, bn, Hh, Ww, out, res, relu, st);
  };
  for (int b0 = 0; b0 < B; b0 += chunk) {
    bn = std::min(chunk, B - b0);
    // ---- forward with the tape
    use_tc = tc_fwd;
    float* x0 = h->tape_x0.as<float>();
    int rc = h->stem.run(spec + static_cast<size_t>(b0) * H0 * W0, bn, H0, W0, x0, nullptr, 1, st);
    if (rc != AP_OK) return rc;
    const float* x = x0;
    int H = H0, W = W0;
    for (size_t i = 0; i < h->blocks.size(); ++i) {
      Bottleneck& b = *h->blocks[i];
      float *r = h->tape[3 * i]->as<float>(), *c = h->tape[3 * i + 1]->as<float>(), *y = h->tape[3 * i + 2]->as<float>();
      const int Ho = (H - 1) / b.stride + 1, Wo = (W - 1) / b.stride + 1;
      rc = conv(b.reduce, x, H, W, r, nullptr, 1);
      if (rc == AP_OK) rc = conv(b.conv, r, H, W, c, nullptr, 1);
      const float* res = x;
      if (rc == AP_OK && b.has_shortcut) {
        rc = conv(b.shortcut, x, H, W, h->gbuf[4].as<float>(), nullptr, 0);
        res = h->gbuf[4].as<float>();
      }
      if (rc == AP_OK) rc = conv(b.expand, c, Ho, Wo, y, res, 1);
      if (rc != AP_OK) return rc;
      x = y, H = Ho, W = Wo;
    }
    // ---- backward
    use_tc = tc_bwd;
    float *GA = h->gbuf[0].as<float>(), *GB = h->gbuf[1].as<float>(), *GC = h->gbuf[2].as<float>(), *GD = h->gbuf[3].as<float>(),
          *GE = h->gbuf[4].as<float>();
    pool_fc_bwd_kernel<<<bn, 256, 0, st>>>(g_logits + static_cast<size_t>(b0) * h->cfg.num_classes, h->fc_w.as<float>(),
                                           h->cfg.num_classes, h->feat, H * W, GA);
    AP_LAUNCH_CHECK();
    for (int i = static_cast<int>(h->blocks.size()) - 1; i >= 0; --i) {
      Bottleneck& b = *h->blocks[i];
      const float *r = h->tape[3 * i]->as<float>(), *c = h->tape[3 * i + 1]->as<float>(), *y = h->tape[3 * i + 2]->as<float>();
      const int Ho = H, Wo = W, Hi = H * b.stride, Wi = W * b.stride;
      const size_t npix_o = static_cast<size_t>(bn) * Ho * Wo, npix_i = static_cast<size_t>(bn) * Hi * Wi;
      rc = mask(GA, y, npix_o * b.cout);                                          // through the block's final ReLU
      if (rc == AP_OK) rc = conv(b.t_expand, GA, Ho, Wo, GB, nullptr, 0);   // -> d c
      if (rc == AP_OK) rc = mask(GB, c, npix_o * b.D);
      const float* src = GB;
      if (rc == AP_OK && b.stride == 2) {
        rc = upsample(GB, GC, bn, Ho, Wo, b.D);
        src = GC;
      }
      if (rc == AP_OK) rc = conv(b.t_conv, src, Hi, Wi, GD, nullptr, 0);    // -> d r
      if (rc == AP_OK) rc = mask(GD, r, npix_i * b.D);
      const float* res = GA;                                                      // identity shortcut
      if (rc == AP_OK && b.has_shortcut) {
        const float* ssrc = GA;
        if (b.stride == 2) {
          rc = upsample(GA, GC, bn, Ho, Wo, b.cout);
          ssrc = GC;
        }
        if (rc == AP_OK) rc = conv(b.t_shortcut, ssrc, Hi, Wi, GE, nullptr, 0);
        res = GE;
      }
      if (rc == AP_OK) rc = conv(b.t_reduce, GD, Hi, Wi, GB, res, 0);       // d x = reduce^T(d r) + shortcut path
      if (rc != AP_OK) return rc;
      std::swap(GA, GB);
      H = Hi, W = Wi;
    }
    rc = mask(GA, x0, static_cast<size_t>(bn) * H0 * W0 * 64);
    if (rc == AP_OK) rc = h->t_stem.run(GA, bn, H0, W0, g_spec + static_cast<size_t>(b0) * H0 * W0, nullptr, 0, st);
    if (rc != AP_OK) return rc;
  }
  return AP_OK;
}

// The recomputed forward of a VGG / WideResNet backward pass runs on the tensor cores in AP_MODE_TF32 (the gradient is then that of
// the tf32 network: its ReLU masks differ from an fp32 forward's in a few units, which moves the input gradient by a few per cent,
// see DESIGN.md "Backward pass") unless AP_CLS_VJP_FWD_FP32=1 keeps it on the fp32 path, as for ResNeXt.
static bool tape_on_tensor_cores(const ap_classifier_s* h) {
  const char* e = std::getenv("AP_CLS_VJP_FWD_FP32");   // read per call, like vjp_resnext
  return h->mode == AP_MODE_TF32 && !(e && std::atoi(e) != 0);
}

// A convolution of a VGG / WideResNet pass: tf32 tensor cores when the classifier is in AP_MODE_TF32 and the layer
// has a tensor-core twin for this geometry, else the fp32 FFMA implicit GEMM.  Slot `si` caches the tensor maps of the buffers the
// convolution ran on last time (they are re-encoded when a buffer moved or the chunk size changed).
static int conv_auto(ap_classifier_t h, size_t si, const ConvLayer& L, const float* in, int bn, int H, int W, float* out,
                     const float* res, int relu, int round_out, cudaStream_t st, bool allow_tc = true) {
  if (!allow_tc || h->mode != AP_MODE_TF32 || !L.has_tc || !conv_tc_supported(L.Cin, L.Cout, L.groups, H, W, L.kh, L.kw, L.stride, L.pad))
    return L.run(in, bn, H, W, out, res, relu, st);
  if (h->tc_slots.size() <= si) h->tc_slots.resize(si + 1);
  auto& s = h->tc_slots[si];
  if (!s.tc || s.in != in || s.out != out || s.res != res || s.relu != relu || s.bn != bn) {
    int rc = L.tc.bind(&s.bnd, in, bn, H, W, out, res, relu, round_out);
    if (rc != AP_OK) return rc;
    s.in = in, s.out = out, s.res = res, s.relu = relu, s.bn = bn, s.tc = true;
  }
  return L.tc.run(s.bnd, st);
}

// ---- ResNet family (models/resnet.py:103-220): state_dict order conv1.weight, bn1.{w,b,mean,var}, then per block
//      conv1, bn1, conv2, bn2, [conv3, bn3,] [downsample.0.weight, downsample.1.{...}], finally fc.weight, fc.bias
static int create_resnet(ap_classifier_t h, const float* const* w, int n_weights) {
  const ap_classifier_cfg& c = h->cfg;
  int counts[4];
  bool bott = false;
  switch (c.depth) {
    case 18: counts[0] = 2, counts[1] = 2, counts[2] = 2, counts[3] = 2; break;
    case 34: counts[0] = 3, counts[1] = 4, counts[2] = 6, counts[3] = 3; break;
    case 50: counts[0] = 3, counts[1] = 4, counts[2] = 6, counts[3] = 3, bott = true; break;
    case 101: counts[0] = 3, counts[1] = 4, counts[2] = 23, counts[3] = 3, bott = true; break;
    case 152: counts[0] = 3, counts[1] = 8, counts[2] = 36, counts[3] = 3, bott = true; break;
    default: return fail(AP_ERR_INVALID, "ap_classifier_create: ResNet depth must be 18/34/50/101/152 (got %d)", c.depth);
  }
  AP_REQUIRE(c.in_channels == 1, "ap_classifier_create: ResNet in_channels must be 1 (NCHW == NHWC)");
  const int exp = bott ? 4 : 1;
  int expected = 5 + 2, inpl = 64;
  for (int l = 0; l < 4; ++l)
    for (int b = 0; b < counts[l]; ++b) {
      const int planes = 64 << l, stride = (b == 0 && l > 0) ? 2 : 1;
      expected += (bott ? 15 : 10) + ((b == 0 && (stride != 1 || inpl != planes * exp)) ? 5 : 0);
      inpl = planes * exp;
    }
  AP_REQUIRE(n_weights == expected, "ap_classifier_create: ResNet-%d expects %d weight tensors, got %d", c.depth, expected, n_weights);
  h->stem.keep_host = true;
  int rc = h->stem.init(1, 64, 7, 7, 2, 3, 1, w[0], nullptr, w[1], w[2], w[3], w[4]);
  if (rc != AP_OK) return rc;
  int i = 5;
  inpl = 64;
  for (int l = 0; l < 4; ++l)
    for (int b = 0; b < counts[l]; ++b) {
      auto blk = std::make_unique<ap_classifier_s::ResBlock>();
      const int planes = 64 << l, stride = (b == 0 && l > 0) ? 2 : 1;
      blk->bottleneck = bott, blk->stride = stride, blk->cout = planes * exp, blk->cin = inpl, blk->planes = planes;
      blk->c1.keep_host = blk->c2.keep_host = blk->c3.keep_host = blk->down.keep_host = true;
      if (bott) {
        rc = blk->c1.init(inpl, planes, 1, 1, 1, 0, 1, w[i], nullptr, w[i + 1], w[i + 2], w[i + 3], w[i + 4], true);
        if (rc == AP_OK) rc = blk->c2.init(planes, planes, 3, 3, stride, 1, 1, w[i + 5], nullptr, w[i + 6], w[i + 7], w[i + 8], w[i + 9], true);
        if (rc == AP_OK) rc = blk->c3.init(planes, planes * 4, 1, 1, 1, 0, 1, w[i + 10], nullptr, w[i + 11], w[i + 12], w[i + 13], w[i + 14], true);
        i += 15;
      } else {
        rc = blk->c1.init(inpl, planes, 3, 3, stride, 1, 1, w[i], nullptr, w[i + 1], w[i + 2], w[i + 3], w[i + 4], true);
        if (rc == AP_OK) rc = blk->c2.init(planes, planes, 3, 3, 1, 1, 1, w[i + 5], nullptr, w[i + 6], w[i + 7], w[i + 8], w[i + 9], true);
        i += 10;
      }
      if (rc == AP_OK && b == 0 && (stride != 1 || inpl != planes * exp)) {
        blk->has_down = true;
        rc = blk->down.init(inpl, planes * exp, 1, 1, stride, 0, 1, w[i], nullptr, w[i + 1], w[i + 2], w[i + 3], w[i + 4], true);
        i += 5;
      }
      if (rc != AP_OK) return rc;
      inpl = planes * exp;
      h->resblocks.push_back(std::move(blk));
    }
  h->feat = 512 * exp;
  AP_CUDA(h->fc_w.upload(w[i], sizeof(float) * c.num_classes * h->feat));
  AP_CUDA(h->fc_b.upload(w[i + 1], sizeof(float) * c.num_classes));
  return AP_OK;
}

static int forward_resnet(ap_classifier_t h, const float* spec, float* logits, int B, int H0, int W0, cudaStream_t st) {
  // conv1 7x7 s2 -> maxpool 3x3 s2 -> 4 stages; the reference's x.view(B, -1) needs a 1x1 final map (32x32 input)
  auto half = [](int v) { return (v - 1) / 2 + 1; };
  const int H1 = half(H0), W1 = half(W0), H2 = half(H1), W2 = half(W1);
  AP_REQUIRE(half(half(half(H2))) == 1 && half(half(half(W2))) == 1, "ResNet: input %dx%d does not reduce to 1x1", H0, W0);
  const int chunk = 256;
  const size_t per = std::max<size_t>(static_cast<size_t>(H1) * W1 * 64, static_cast<size_t>(H2) * W2 * 64 * (h->feat / 512) * 4);
  int rc = ensure_ws(h, static_cast<size_t>(std::min(B, chunk)) * per);
  if (rc != AP_OK) return rc;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bn = std::min(chunk, B - b0);
    float* x = h->buf[0].as<float>();
    float* xo = h->buf[1].as<float>();
    float* y1 = h->buf[2].as<float>();
    float* y2 = h->buf[3].as<float>();
    float* sc = h->buf[4].as<float>();
    rc = h->stem.run(spec + static_cast<size_t>(b0) * H0 * W0, bn, H0, W0, y1, nullptr, 1, st);      // resnet.py:146-148
    if (rc != AP_OK) return rc;
    {
      const long long total = static_cast<long long>(bn) * H2 * W2 * 64;
      long long blocks = ceil_div_ll(total, 256);
      if (blocks > num_sms() * 8) blocks = num_sms() * 8;
      maxpool3x3s2_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(y1, x, bn, H1, W1, 64);          // :149
      AP_LAUNCH_CHECK();
    }
    int H = H2, W = W2;
    size_t si = 0;                                  // tensor-map slots: four per block
    for (auto& blk : h->resblocks) {
      const int Ho = (H - 1) / blk->stride + 1, Wo = (W - 1) / blk->stride + 1;
      const float* res = x;
      if (blk->has_down) {
        rc = conv_auto(h, si + 3, blk->down, x, bn, H, W, sc, nullptr, 0, 1, st);
        if (rc != AP_OK) return rc;
        res = sc;
      }
      if (blk->bottleneck) {
        rc = conv_auto(h, si, blk->c1, x, bn, H, W, y1, nullptr, 1, 1, st);
        if (rc == AP_OK) rc = conv_auto(h, si + 1, blk->c2, y1, bn, H, W, y2, nullptr, 1, 1, st);
        if (rc == AP_OK) rc = conv_auto(h, si + 2, blk->c3, y2, bn, Ho, Wo, xo, res, 1, 1, st);
      } else {
        rc = conv_auto(h, si, blk->c1, x, bn, H, W, y1, nullptr, 1, 1, st);
        if (rc == AP_OK) rc = conv_auto(h, si + 1, blk->c2, y1, bn, Ho, Wo, xo, res, 1, 1, st);
      }
      si += 4;
      if (rc != AP_OK) return rc;
      std::swap(x, xo);
      H = Ho, W = Wo;
    }
    const size_t smem = sizeof(float) * (h->feat + h->cfg.num_classes);
    pool_fc_kernel<<<bn, 256, smem, st>>>(x, 1, h->feat, h->fc_w.as<float>(), h->fc_b.as<float>(), h->cfg.num_classes,
                                          logits + static_cast<size_t>(b0) * h->cfg.num_classes, 0);   // AvgPool2d(1) + fc, :156-158
    AP_LAUNCH_CHECK();
  }
  return AP_OK;
}

// ---- backward of the ResNet forward (autograd over models/resnet.py:140-160, BatchNorm in eval mode), fp32 on the FFMA path
static int vjp_resnet(ap_classifier_t h, const float* spec, const float* g_logits, float* g_spec, int B, int H0, int W0,
                      cudaStream_t st) {
  auto half = [](int v) { return (v - 1) / 2 + 1; };
  const int H1 = half(H0), W1 = half(W0), H2 = half(H1), W2 = half(W1);
  AP_REQUIRE(H0 == 2 * H1 && W0 == 2 * W1 && half(half(half(H2))) == 1 && half(half(half(W2))) == 1,
             "ResNet backward: input %dx%d does not reduce to 1x1 by exact halvings", H0, W0);
  const int chunk = 64;
  const size_t NS = 4 * h->resblocks.size();   // tensor-map slots: [0, NS) inference, [NS, 2 NS) recomputed forward, [2 NS, 3 NS) backward
  const bool tape_tc = tape_on_tensor_cores(h);
  if (!h->bwd_ready) {
    int rc = init_dgrad(h->t_stem, h->stem);
    for (auto& b : h->resblocks) {
      if (rc == AP_OK) rc = init_dgrad(b->t_c1, b->c1, true);
      if (rc == AP_OK) rc = init_dgrad(b->t_c2, b->c2, true);
      if (rc == AP_OK && b->bottleneck) rc = init_dgrad(b->t_c3, b->c3, true);
      if (rc == AP_OK && b->has_down) rc = init_dgrad(b->t_down, b->down, true);
    }
    if (rc != AP_OK) return rc;
    h->bwd_ready = true;
  }
  const int bn_max = std::min(B, chunk);
  if (bn_max > h->bwd_bn) {
    h->tape.clear();
    size_t max_e = static_cast<size_t>(H0) * W0 * 64;    // the zero-upsampled stem gradient
    AP_CUDA(h->tape_x0.alloc(static_cast<size_t>(bn_max) * H1 * W1 * 64 * sizeof(float)));     // stem conv + ReLU output
    auto add = [&](size_t e) -> int {
      auto d = std::make_unique<DevBuf>();
      AP_CUDA(d->alloc(e * bn_max * sizeof(float)));
      h->tape.push_back(std::move(d));
      max_e = std::max(max_e, e);
      return AP_OK;
    };
    int rc = add(static_cast<size_t>(H2) * W2 * 64);       // tape[0]: pooled stem output = input of the first block
    int H = H2, W = W2;
    for (auto& b : h->resblocks) {
      const int Ho = (H - 1) / b->stride + 1, Wo = (W - 1) / b->stride + 1;
      if (rc == AP_OK) rc = add(static_cast<size_t>(b->bottleneck ? H * W : Ho * Wo) * b->planes);     // c1 output
      if (rc == AP_OK) rc = add(static_cast<size_t>(Ho) * Wo * b->planes);                              // c2 output (bottleneck)
      if (rc == AP_OK) rc = add(static_cast<size_t>(Ho) * Wo * b->cout);                                // block output
      max_e = std::max(max_e, static_cast<size_t>(H) * W * std::max(b->cin, std::max(b->cout, b->planes)));
      if (rc != AP_OK) return rc;
      H = Ho, W = Wo;
    }
    for (auto& g : h->gbuf) AP_CUDA(g.alloc(max_e * bn_max * sizeof(float)));
    h->bwd_bn = bn_max;
  }
  auto grid_for = [](long long work) {
    long long b = ceil_div_ll(work, 256);
    const long long cap = static_cast<long long>(num_sms()) * 16;
    return static_cast<unsigned>(b < cap ? (b > 0 ? b : 1) : cap);
  };
  int bn = 0;
  auto mask = [&](float* g, const float* act, size_t elems) -> int {
    relu_mask_kernel<<<grid_for(static_cast<long long>(elems / 4)), 256, 0, st>>>(reinterpret_cast<float4*>(g),
                                                                                  reinterpret_cast<const float4*>(act),
                                                                                  static_cast<long long>(elems / 4));
    AP_LAUNCH_CHECK();
    return AP_OK;
  };
  auto upsample = [&](const float* g, float* up, int Ho, int Wo, int Cc) -> int {
    upsample2_kernel<<<grid_for(static_cast<long long>(bn) * 4 * Ho * Wo * (Cc / 4)), 256, 0, st>>>(
        reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(up), bn, Ho, Wo, Cc / 4);
    AP_LAUNCH_CHECK();
    return AP_OK;
  };
  for (int b0 = 0; b0 < B; b0 += chunk) {
    bn = std::min(chunk, B - b0);
    // ---- forward with the tape
    float* y1s = h->tape_x0.as<float>();
    int rc = h->stem.run(spec + static_cast<size_t>(b0) * H0 * W0, bn, H0, W0, y1s, nullptr, 1, st);
    if (rc != AP_OK) return rc;
    float* x = h->tape[0]->as<float>();
    maxpool3x3s2_kernel<<<grid_for(static_cast<long long>(bn) * H2 * W2 * 64), 256, 0, st>>>(y1s, x, bn, H1, W1, 64);
    AP_LAUNCH_CHECK();
    int H = H2, W = W2;
    for (size_t i = 0; i < h->resblocks.size(); ++i) {
      auto& b = *h->resblocks[i];
      float *a1 = h->tape[1 + 3 * i]->as<float>(), *a2 = h->tape[2 + 3 * i]->as<float>(), *y = h->tape[3 + 3 * i]->as<float>();
      const int Ho = (H - 1) / b.stride + 1, Wo = (W - 1) / b.stride + 1;
      const float* res = x;
      const size_t sf = NS + 4 * i;                 // tensor-map slots of the recomputed forward
      if (b.has_down) {
        rc = conv_auto(h, sf + 3, b.down, x, bn, H, W, h->gbuf[4].as<float>(), nullptr, 0, 1, st, tape_tc);
        if (rc != AP_OK) return rc;
        res = h->gbuf[4].as<float>();
      }
      if (b.bottleneck) {
        rc = conv_auto(h, sf, b.c1, x, bn, H, W, a1, nullptr, 1, 1, st, tape_tc);
        if (rc == AP_OK) rc = conv_auto(h, sf + 1, b.c2, a1, bn, H, W, a2, nullptr, 1, 1, st, tape_tc);
        if (rc == AP_OK) rc = conv_auto(h, sf + 2, b.c3, a2, bn, Ho, Wo, y, res, 1, 1, st, tape_tc);
      } else {
        rc = conv_auto(h, sf, b.c1, x, bn, H, W, a1, nullptr, 1, 1, st, tape_tc);
        if (rc == AP_OK) rc = conv_auto(h, sf + 1, b.c2, a1, bn, Ho, Wo, y, res, 1, 1, st, tape_tc);
      }
      if (rc != AP_OK) return rc;
      x = y, H = Ho, W = Wo;
    }
    // ---- backward
    float *GA = h->gbuf[0].as<float>(), *GB = h->gbuf[1].as<float>(), *GC = h->gbuf[2].as<float>(), *GD = h->gbuf[3].as<float>(),
          *GE = h->gbuf[4].as<float>();
    pool_fc_bwd_kernel<<<bn, 256, 0, st>>>(g_logits + static_cast<size_t>(b0) * h->cfg.num_classes, h->fc_w.as<float>(),
                                           h->cfg.num_classes, h->feat, H * W, GA);
    AP_LAUNCH_CHECK();
    for (int i = static_cast<int>(h->resblocks.size()) - 1; i >= 0; --i) {
      auto& b = *h->resblocks[i];
      const float *a1 = h->tape[1 + 3 * i]->as<float>(), *a2 = h->tape[2 + 3 * i]->as<float>(), *y = h->tape[3 + 3 * i]->as<float>();
      const int Ho = H, Wo = W, Hi = H * b.stride, Wi = W * b.stride;
      const size_t npo = static_cast<size_t>(bn) * Ho * Wo, npi = static_cast<size_t>(bn) * Hi * Wi;
      const size_t sb = 2 * NS + 4 * static_cast<size_t>(i);   // tensor-map slots of this block's data-gradient twins
      rc = mask(GA, y, npo * b.cout);
      const float* res = GA;                                    // identity shortcut
      if (rc == AP_OK && b.has_down) {                          // shortcut path first: it borrows GC for the upsampled gradient
        const float* ssrc = GA;
        if (b.stride == 2) {
          rc = upsample(GA, GC, Ho, Wo, b.cout);
          ssrc = GC;
        }
        if (rc == AP_OK) rc = conv_auto(h, sb + 3, b.t_down, ssrc, bn, Hi, Wi, GE, nullptr, 0, 0, st);
        res = GE;
      }
      if (b.bottleneck) {
        if (rc == AP_OK) rc = conv_auto(h, sb + 2, b.t_c3, GA, bn, Ho, Wo, GB, nullptr, 0, 0, st);
        if (rc == AP_OK) rc = mask(GB, a2, npo * b.planes);
        const float* src = GB;
        if (rc == AP_OK && b.stride == 2) {
          rc = upsample(GB, GC, Ho, Wo, b.planes);
          src = GC;
        }
        if (rc == AP_OK) rc = conv_auto(h, sb + 1, b.t_c2, src, bn, Hi, Wi, GD, nullptr, 0, 0, st);
        if (rc == AP_OK) rc = mask(GD, a1, npi * b.planes);
        if (rc == AP_OK) rc = conv_auto(h, sb, b.t_c1, GD, bn, Hi, Wi, GB, res, 0, 0, st);
      } else {
        if (rc == AP_OK) rc = conv_auto(h, sb + 1, b.t_c2, GA, bn, Ho, Wo, GB, nullptr, 0, 0, st);
        if (rc == AP_OK) rc = mask(GB, a1, npo * b.planes);
        const float* src = GB;
        if (rc == AP_OK && b.stride == 2) {
          rc = upsample(GB, GC, Ho, Wo, b.planes);
          src = GC;
        }
        if (rc == AP_OK) rc = conv_auto(h, sb, b.t_c1, src, bn, Hi, Wi, GD, res, 0, 0, st);
        if (rc == AP_OK) std::swap(GB, GD);                     // the block's input gradient is in GB either way
      }
      if (rc != AP_OK) return rc;
      std::swap(GA, GB);
      H = Hi, W = Wi;
    }
    // stem: max-pool, ReLU, 7x7 stride-2 convolution
    maxpool3x3s2_bwd_kernel<<<grid_for(static_cast<long long>(bn) * H1 * W1 * 64), 256, 0, st>>>(GA, y1s, GB, bn, H1, W1, 64);
    AP_LAUNCH_CHECK();
    rc = mask(GB, y1s, static_cast<size_t>(bn) * H1 * W1 * 64);
    if (rc == AP_OK) rc = upsample(GB, GC, H1, W1, 64);
    if (rc == AP_OK) rc = h->t_stem.run(GC, bn, H0, W0, g_spec + static_cast<size_t>(b0) * H0 * W0, nullptr, 0, st);
    if (rc != AP_OK) return rc;
  }
  return AP_OK;
}

// ---- VGG-11/13/16/19 with batch norm (models/vgg.py:32-95; the SC09 factory builds vgg19_bn, models/__init__.py:44-45).
// state_dict order: features.{i}.{weight,bias} + BatchNorm {weight,bias,running_mean,running_var} per convolution, then
// classifier.{0,3,6}.{weight,bias}.  Every 3x3 convolution folds its bias and BatchNorm; Dropout is the identity in eval mode;
// on a 32x32 input the five pools leave a 1x1 map, so the three Linear layers are 1x1 convolutions on it.
static int create_vgg(ap_classifier_t h, const float* const* w, int n_weights) {
  const ap_classifier_cfg& c = h->cfg;
  static const int A[] = {64, -1, 128, -1, 256, 256, -1, 512, 512, -1, 512, 512, -1, 0};
  static const int Bc[] = {64, 64, -1, 128, 128, -1, 256, 256, -1, 512, 512, -1, 512, 512, -1, 0};
  static const int D[] = {64, 64, -1, 128, 128, -1, 256, 256, 256, -1, 512, 512, 512, -1, 512, 512, 512, -1, 0};
  static const int E[] = {64, 64, -1, 128, 128, -1, 256, 256, 256, 256, -1, 512, 512, 512, 512, -1, 512, 512, 512, 512, -1, 0};
  const int* plan = c.depth == 11 ? A : c.depth == 13 ? Bc : c.depth == 16 ? D : c.depth == 19 ? E : nullptr;
  AP_REQUIRE(plan, "ap_classifier_create: VGG depth must be 11/13/16/19 (got %d)", c.depth);
  AP_REQUIRE(c.in_channels == 1, "ap_classifier_create: VGG in_channels must be 1 (NCHW == NHWC)");
  int n_conv = 0;
  for (const int* q = plan; *q; ++q) {
    h->vgg_plan.push_back(*q);
    n_conv += *q > 0;
  }
  AP_REQUIRE(n_weights == 6 * n_conv + 6, "ap_classifier_create: VGG-%d (batch norm) expects %d weight tensors, got %d", c.depth,
             6 * n_conv + 6, n_weights);
  int i = 0, cin = 1;
  for (int v : h->vgg_plan) {
    if (v < 0) continue;
    auto L = std::make_unique<ConvLayer>();
    L->keep_host = true;
    int rc = L->init(cin, v, 3, 3, 1, 1, 1, w[i], w[i + 1], w[i + 2], w[i + 3], w[i + 4], w[i + 5], true);
    if (rc != AP_OK) return rc;
    h->vgg_conv.push_back(std::move(L));
    i += 6, cin = v;
  }
  const int dims[4] = {512, 4096, 4096, c.num_classes};
  for (int j = 0; j < 3; ++j) {
    h->vgg_fc[j].keep_host = true;
    int rc = h->vgg_fc[j].init(dims[j], dims[j + 1], 1, 1, 1, 0, 1, w[i], w[i + 1], nullptr, nullptr, nullptr, nullptr, true);
    if (rc != AP_OK) return rc;
    i += 2;
  }
  h->feat = 512;
  return AP_OK;
}

static unsigned vgg_grid(long long work) {
  long long b = ceil_div_ll(work, 256);
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<unsigned>(b < cap ? (b > 0 ? b : 1) : cap);
}

static int forward_vgg(ap_classifier_t h, const float* spec, float* logits, int B, int H0, int W0, cudaStream_t st) {
  AP_REQUIRE(H0 == 32 && W0 == 32, "VGG: the classifier head needs a 1x1 final map, i.e. a 32x32 input (got %dx%d)", H0, W0);
  const int chunk = 256;
  int rc = ensure_ws(h, static_cast<size_t>(std::min(B, chunk)) * H0 * W0 * 64);
  if (rc != AP_OK) return rc;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bn = std::min(chunk, B - b0);
    const float* x = spec + static_cast<size_t>(b0) * H0 * W0;
    float* nxt[2] = {h->buf[0].as<float>(), h->buf[1].as<float>()};
    int H = H0, W = W0, C = 1, which = 0;
    size_t ci = 0;
    for (int v : h->vgg_plan) {
      float* y = nxt[which];
      if (v > 0) {
        rc = conv_auto(h, ci, *h->vgg_conv[ci], x, bn, H, W, y, nullptr, 1, 1, st);                      // vgg.py:75-78
        if (rc != AP_OK) return rc;
        ++ci, C = v;
      } else {
        maxpool2x2_kernel<<<vgg_grid(static_cast<long long>(bn) * (H / 2) * (W / 2) * (C / 4)), 256, 0, st>>>(
            reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), bn, H, W, C / 4);
        AP_LAUNCH_CHECK();
        H /= 2, W /= 2;
      }
      x = y, which ^= 1;
    }
    for (int j = 0; j < 3; ++j) {                                                                        // vgg.py:36-44
      float* y = j == 2 ? logits + static_cast<size_t>(b0) * h->cfg.num_classes : nxt[which];
      rc = conv_auto(h, ci + j, h->vgg_fc[j], x, bn, 1, 1, y, nullptr, j < 2, 1, st);
      if (rc != AP_OK) return rc;
      x = y, which ^= 1;
    }
  }
  return AP_OK;
}

// backward of forward_vgg (autograd over models/vgg.py:49-53, BatchNorm / Dropout in eval mode), fp32 on the FFMA path:
// the forward is recomputed with every layer output kept, then walked in reverse
static int vjp_vgg(ap_classifier_t h, const float* spec, const float* g_logits, float* g_spec, int B, int H0, int W0,
                   cudaStream_t st) {
  AP_REQUIRE(H0 == 32 && W0 == 32, "VGG backward: input must be 32x32 (got %dx%d)", H0, W0);
  const int chunk = 64;
  const size_t n_layers = h->vgg_plan.size();
  const size_t NS = h->vgg_conv.size() + 3;     // tensor-map slots: [0, NS) inference, [NS, 2 NS) recomputed forward, [2 NS, 3 NS) backward
  const bool tape_tc = tape_on_tensor_cores(h);
  if (!h->bwd_ready) {
    for (auto& L : h->vgg_conv) {
      auto T = std::make_unique<ConvLayer>();
      int rc = init_dgrad(*T, *L, true);
      if (rc != AP_OK) return rc;
      h->vgg_conv_t.push_back(std::move(T));
    }
    for (int j = 0; j < 3; ++j) {
      int rc = init_dgrad(h->vgg_fc_t[j], h->vgg_fc[j], true);
      if (rc != AP_OK) return rc;
    }
    h->bwd_ready = true;
  }
  const int bn_max = std::min(B, chunk);
  if (bn_max > h->bwd_bn) {
    h->tape.clear();
    int H = H0, W = W0, C = 1;
    size_t max_e = 4096;
    for (int v : h->vgg_plan) {
      if (v > 0) C = v;
      else H /= 2, W /= 2;
      const size_t e = static_cast<size_t>(H) * W * C;
      auto d = std::make_unique<DevBuf>();
      AP_CUDA(d->alloc(e * bn_max * sizeof(float)));
      h->tape.push_back(std::move(d));
      max_e = std::max(max_e, e);
    }
    for (int j = 0; j < 2; ++j) {
      auto d = std::make_unique<DevBuf>();
      AP_CUDA(d->alloc(static_cast<size_t>(4096) * bn_max * sizeof(float)));
      h->tape.push_back(std::move(d));
    }
    for (int j = 0; j < 2; ++j) AP_CUDA(h->gbuf[j].alloc(max_e * bn_max * sizeof(float)));
    h->bwd_bn = bn_max;
  }
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bn = std::min(chunk, B - b0);
    // ---- forward with the tape
    const float* x = spec + static_cast<size_t>(b0) * H0 * W0;
    int H = H0, W = W0, C = 1, rc = AP_OK;
    size_t ci = 0;
    for (size_t l = 0; l < n_layers; ++l) {
      float* y = h->tape[l]->as<float>();
      const int v = h->vgg_plan[l];
      if (v > 0) {
        rc = conv_auto(h, NS + ci, *h->vgg_conv[ci], x, bn, H, W, y, nullptr, 1, 1, st, tape_tc);
        if (rc != AP_OK) return rc;
        ++ci, C = v;
      } else {
        maxpool2x2_kernel<<<vgg_grid(static_cast<long long>(bn) * (H / 2) * (W / 2) * (C / 4)), 256, 0, st>>>(
            reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), bn, H, W, C / 4);
        AP_LAUNCH_CHECK();
        H /= 2, W /= 2;
      }
      x = y;
    }
    float *f1 = h->tape[n_layers]->as<float>(), *f2 = h->tape[n_layers + 1]->as<float>();
    rc = conv_auto(h, NS + ci, h->vgg_fc[0], x, bn, 1, 1, f1, nullptr, 1, 1, st, tape_tc);
    if (rc == AP_OK) rc = conv_auto(h, NS + ci + 1, h->vgg_fc[1], f1, bn, 1, 1, f2, nullptr, 1, 1, st, tape_tc);
    if (rc != AP_OK) return rc;
    // ---- backward
    float *GA = h->gbuf[0].as<float>(), *GB = h->gbuf[1].as<float>();
    auto mask = [&](float* g, const float* act, size_t elems) -> int {
      relu_mask_kernel<<<vgg_grid(static_cast<long long>(elems / 4)), 256, 0, st>>>(
          reinterpret_cast<float4*>(g), reinterpret_cast<const float4*>(act), static_cast<long long>(elems / 4));
      AP_LAUNCH_CHECK();
      return AP_OK;
    };
    rc = h->vgg_fc_t[2].run(g_logits + static_cast<size_t>(b0) * h->cfg.num_classes, bn, 1, 1, GA, nullptr, 0, st);
    if (rc == AP_OK) rc = mask(GA, f2, static_cast<size_t>(bn) * 4096);
    if (rc == AP_OK) rc = conv_auto(h, 2 * NS + ci + 1, h->vgg_fc_t[1], GA, bn, 1, 1, GB, nullptr, 0, 0, st);
    if (rc == AP_OK) rc = mask(GB, f1, static_cast<size_t>(bn) * 4096);
    if (rc == AP_OK) rc = conv_auto(h, 2 * NS + ci, h->vgg_fc_t[0], GB, bn, 1, 1, GA, nullptr, 0, 0, st);   // gradient of the pooled 1x1x512 features
    if (rc != AP_OK) return rc;
    for (int l = static_cast<int>(n_layers) - 1; l >= 0; --l) {
      const int v = h->vgg_plan[l];
      const float* in_act = l > 0 ? h->tape[l - 1]->as<float>() : nullptr;          // input of layer l
      if (v > 0) {
        --ci;
        rc = mask(GA, h->tape[l]->as<float>(), static_cast<size_t>(bn) * H * W * C);
        float* dst = l == 0 ? g_spec + static_cast<size_t>(b0) * H0 * W0 : GB;
        if (rc == AP_OK) rc = conv_auto(h, 2 * NS + ci, *h->vgg_conv_t[ci], GA, bn, H, W, dst, nullptr, 0, 0, st);
        if (rc != AP_OK) return rc;
        C = h->vgg_conv[ci]->Cin;
      } else {
        maxpool2x2_bwd_kernel<<<vgg_grid(static_cast<long long>(bn) * H * W * C), 256, 0, st>>>(GA, in_act, GB, bn, 2 * H, 2 * W, C);
        AP_LAUNCH_CHECK();
        H *= 2, W *= 2;
      }
      std::swap(GA, GB);
    }
  }
  return AP_OK;
}

// ---- WideResNet-d-k (models/wideresnet.py:15-92; `wideresnet28_10` is a --classifier_model choice, adaptive_attack_eval.py:21).
// Weights: conv1.weight; per block bn1.{weight,bias,running_mean,running_var}, conv1.weight, bn2.{...}, conv2.weight,
// [convShortcut.weight]; bn1.{...}; fc.weight, fc.bias.  bn2 folds into conv1 (ReLU in its epilogue); bn1 acts on the residual sum
// and runs as bn_relu_kernel.  A block whose width changes feeds relu(bn1(x)) to BOTH the convolutions and the shortcut (:31-32,39).
static int bn_vectors(DevBuf& scale, DevBuf& shift, const float* const* w, int C) {
  std::vector<float> a(C), b(C);
  for (int c = 0; c < C; ++c) {
    a[c] = w[0][c] / std::sqrt(w[3][c] + 1e-5f);
    b[c] = w[1][c] - w[2][c] * a[c];
  }
  AP_CUDA(scale.upload(a.data(), sizeof(float) * C));
  AP_CUDA(shift.upload(b.data(), sizeof(float) * C));
  return AP_OK;
}

static int create_wrn(ap_classifier_t h, const float* const* w, int n_weights) {
  const ap_classifier_cfg& c = h->cfg;
  AP_REQUIRE(c.depth >= 10 && (c.depth - 4) % 6 == 0, "ap_classifier_create: WideResNet depth must be 6n + 4 (got %d)", c.depth);
  AP_REQUIRE(c.widen_factor >= 1 && c.widen_factor <= 16, "ap_classifier_create: WideResNet widen_factor %d out of range", c.widen_factor);
  AP_REQUIRE(c.in_channels == 1, "ap_classifier_create: WideResNet in_channels must be 1 (NCHW == NHWC)");
  const int n = (c.depth - 4) / 6;
  const int ch[4] = {16, 16 * c.widen_factor, 32 * c.widen_factor, 64 * c.widen_factor};
  int expected = 1 + 4 + 2;
  for (int s = 0; s < 3; ++s)
    for (int b = 0; b < n; ++b) expected += 10 + ((b == 0 && ch[s] != ch[s + 1]) ? 1 : 0);
  AP_REQUIRE(n_weights == expected, "ap_classifier_create: WideResNet-%d-%d expects %d weight tensors, got %d", c.depth,
             c.widen_factor, expected, n_weights);
  h->stem.keep_host = true;
  int rc = h->stem.init(1, 16, 3, 3, 1, 1, 1, w[0], nullptr, nullptr, nullptr, nullptr, nullptr);
  if (rc != AP_OK) return rc;
  int i = 1;
  for (int s = 0; s < 3; ++s)
    for (int b = 0; b < n; ++b) {
      auto blk = std::make_unique<ap_classifier_s::WrnBlock>();
      blk->cin = b == 0 ? ch[s] : ch[s + 1], blk->cout = ch[s + 1], blk->stride = (b == 0 && s > 0) ? 2 : 1;
      blk->equal = blk->cin == blk->cout;
      blk->c1.keep_host = blk->c2.keep_host = blk->sc.keep_host = true;
      rc = bn_vectors(blk->scale, blk->shift, w + i, blk->cin);
      if (rc == AP_OK)
        rc = blk->c1.init(blk->cin, blk->cout, 3, 3, blk->stride, 1, 1, w[i + 4], nullptr, w[i + 5], w[i + 6], w[i + 7], w[i + 8], true);
      if (rc == AP_OK) rc = blk->c2.init(blk->cout, blk->cout, 3, 3, 1, 1, 1, w[i + 9], nullptr, nullptr, nullptr, nullptr, nullptr, true);
      i += 10;
      if (rc == AP_OK && !blk->equal) {
        rc = blk->sc.init(blk->cin, blk->cout, 1, 1, blk->stride, 0, 1, w[i], nullptr, nullptr, nullptr, nullptr, nullptr, true);
        i += 1;
      }
      if (rc != AP_OK) return rc;
      h->wrn.push_back(std::move(blk));
    }
  rc = bn_vectors(h->wrn_scale, h->wrn_shift, w + i, ch[3]);
  if (rc != AP_OK) return rc;
  h->feat = ch[3];
  AP_CUDA(h->fc_w.upload(w[i + 4], sizeof(float) * c.num_classes * h->feat));
  AP_CUDA(h->fc_b.upload(w[i + 5], sizeof(float) * c.num_classes));
  return AP_OK;
}

static int wrn_bn_relu(const float* x, const DevBuf& scale, const DevBuf& shift, float* y, long long elems, int C, int round_out,
                       cudaStream_t st) {
  bn_relu_kernel<<<vgg_grid(elems / 4), 256, 0, st>>>(reinterpret_cast<const float4*>(x), scale.as<float4>(), shift.as<float4>(),
                                                      reinterpret_cast<float4*>(y), elems / 4, C / 4, round_out);
  AP_LAUNCH_CHECK();
  return AP_OK;
}

// One pass over `bn` images.  tape == nullptr: inference (activations ping-pong through buf[0..4]);
// otherwise a_l = relu(bn1(x_l)) and h_l = relu(bn2(conv1(a_l))) of every block and the final relu(bn(x)) are kept in tape[2l], [2l+1], [2n].
static int wrn_pass(ap_classifier_t h, const float* spec, float* logits, int bn, std::vector<std::unique_ptr<DevBuf>>* tape,
                    cudaStream_t st) {
  float* x = h->buf[0].as<float>();
  float* xo = h->buf[1].as<float>();
  int rc = h->stem.run(spec, bn, 32, 32, x, nullptr, 0, st);                                               // wideresnet.py:83
  if (rc != AP_OK) return rc;
  int H = 32, W = 32;
  for (size_t l = 0; l < h->wrn.size(); ++l) {
    auto& b = *h->wrn[l];
    float* a = tape ? (*tape)[2 * l]->as<float>() : h->buf[2].as<float>();
    float* hh = tape ? (*tape)[2 * l + 1]->as<float>() : h->buf[3].as<float>();
    const int Ho = H / b.stride, Wo = W / b.stride;
    const bool tc = h->mode == AP_MODE_TF32 && (!tape || tape_on_tensor_cores(h));
    const size_t sb = tape ? 3 * h->wrn.size() : 0;       // tensor-map slots of the recomputed forward follow the inference ones
    rc = wrn_bn_relu(x, b.scale, b.shift, a, static_cast<long long>(bn) * H * W * b.cin, b.cin, tc, st);   // :31-34
    const float* res = x;
    if (rc == AP_OK && !b.equal) {
      rc = tc ? conv_auto(h, sb + 3 * l, b.sc, a, bn, H, W, h->buf[4].as<float>(), nullptr, 0, 0, st)
              : b.sc.run(a, bn, H, W, h->buf[4].as<float>(), nullptr, 0, st);                              // :39
      res = h->buf[4].as<float>();
    }
    if (rc == AP_OK) rc = tc ? conv_auto(h, sb + 3 * l + 1, b.c1, a, bn, H, W, hh, nullptr, 1, 1, st) : b.c1.run(a, bn, H, W, hh, nullptr, 1, st);   // :35
    if (rc == AP_OK) rc = tc ? conv_auto(h, sb + 3 * l + 2, b.c2, hh, bn, Ho, Wo, xo, res, 0, 0, st) : b.c2.run(hh, bn, Ho, Wo, xo, res, 0, st);     // :38-39
    if (rc != AP_OK) return rc;
    std::swap(x, xo);
    H = Ho, W = Wo;
  }
  float* f = tape ? (*tape)[2 * h->wrn.size()]->as<float>() : xo;
  rc = wrn_bn_relu(x, h->wrn_scale, h->wrn_shift, f, static_cast<long long>(bn) * H * W * h->feat, h->feat, 0, st);   // :87
  if (rc != AP_OK) return rc;
  const size_t smem = sizeof(float) * (h->feat + h->cfg.num_classes);
  pool_fc_kernel<<<bn, 256, smem, st>>>(f, H * W, h->feat, h->fc_w.as<float>(), h->fc_b.as<float>(), h->cfg.num_classes, logits, 0);
  AP_LAUNCH_CHECK();                                                                                       // :88-90
  return AP_OK;
}

static int forward_wrn(ap_classifier_t h, const float* spec, float* logits, int B, int H0, int W0, cudaStream_t st) {
  AP_REQUIRE(H0 == 32 && W0 == 32, "WideResNet: avg_pool2d(8) + view needs a 32x32 input (got %dx%d)", H0, W0);
  const int chunk = 128;
  int rc = ensure_ws(h, static_cast<size_t>(std::min(B, chunk)) * 32 * 32 * std::max(16, h->wrn[0]->cout));
  if (rc != AP_OK) return rc;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bn = std::min(chunk, B - b0);
    rc = wrn_pass(h, spec + static_cast<size_t>(b0) * 1024, logits + static_cast<size_t>(b0) * h->cfg.num_classes, bn, nullptr, st);
    if (rc != AP_OK) return rc;
  }
  return AP_OK;
}

// backward of the WideResNet forward (autograd over models/wideresnet.py:30-39,82-90, BatchNorm in eval mode), fp32 FFMA path
static int vjp_wrn(ap_classifier_t h, const float* spec, const float* g_logits, float* g_spec, int B, int H0, int W0,
                   cudaStream_t st) {
  AP_REQUIRE(H0 == 32 && W0 == 32, "WideResNet backward: input must be 32x32 (got %dx%d)", H0, W0);
  const int chunk = 32;
  const size_t nb = h->wrn.size();
  if (!h->bwd_ready) {
    int rc = init_dgrad(h->t_stem, h->stem);
    for (auto& b : h->wrn) {
      if (rc == AP_OK) rc = init_dgrad(b->t_c1, b->c1, true);
      if (rc == AP_OK) rc = init_dgrad(b->t_c2, b->c2, true);
      if (rc == AP_OK && !b->equal) rc = init_dgrad(b->t_sc, b->sc, true);
    }
    if (rc != AP_OK) return rc;
    h->bwd_ready = true;
  }
  const int bn_max = std::min(B, chunk);
  int rc = ensure_ws(h, static_cast<size_t>(bn_max) * 32 * 32 * std::max(16, h->wrn[0]->cout));
  if (rc != AP_OK) return rc;
  size_t widest = 0;    // largest gradient tensor per image: a stride-2 block's zero-upsampled gradient is (2H, 2W, cout)
  for (int H = 32; auto& b : h->wrn) {
    widest = std::max(widest, static_cast<size_t>(H) * H * std::max(b->cin, b->cout));
    H /= b->stride;
  }
  if (bn_max > h->bwd_bn) {
    h->tape.clear();
    int H = 32;
    auto add = [&](size_t e) -> int {
      auto d = std::make_unique<DevBuf>();
      AP_CUDA(d->alloc(e * bn_max * sizeof(float)));
      h->tape.push_back(std::move(d));
      return AP_OK;
    };
    for (auto& b : h->wrn) {
      rc = add(static_cast<size_t>(H) * H * b->cin);
      H /= b->stride;
      if (rc == AP_OK) rc = add(static_cast<size_t>(H) * H * b->cout);
      if (rc != AP_OK) return rc;
    }
    rc = add(static_cast<size_t>(H) * H * h->feat);
    if (rc != AP_OK) return rc;
    for (auto& g : h->gbuf) AP_CUDA(g.alloc(widest * bn_max * sizeof(float)));
    AP_CUDA(h->tape_x0.alloc(sizeof(float) * bn_max * h->cfg.num_classes));     // logits of the recomputed forward (unused)
    h->bwd_bn = bn_max;
  }
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bn = std::min(chunk, B - b0);
    rc = wrn_pass(h, spec + static_cast<size_t>(b0) * 1024, h->tape_x0.as<float>(), bn, &h->tape, st);
    if (rc != AP_OK) return rc;
    // G: gradient of the current block's output; T1..T4: scratch with fixed roles (g_h, upsampled g_h, upsampled G, shortcut gradient)
    float *G = h->gbuf[0].as<float>(), *T1 = h->gbuf[1].as<float>(), *T2 = h->gbuf[2].as<float>(), *T3 = h->gbuf[3].as<float>(),
          *T4 = h->gbuf[4].as<float>();
    int H = 8, W = 8;
    auto bwd_act = [&](float* g_act, const float* act, const DevBuf& scale, const float* skip, long long elems, int C) -> int {
      bn_relu_bwd_kernel<<<vgg_grid(elems / 4), 256, 0, st>>>(reinterpret_cast<const float4*>(g_act),
                                                              reinterpret_cast<const float4*>(act), scale.as<float4>(),
                                                              reinterpret_cast<const float4*>(skip), reinterpret_cast<float4*>(g_act),
                                                              elems / 4, C / 4);
      AP_LAUNCH_CHECK();
      return AP_OK;
    };
    auto upsample = [&](const float* g, float* up, int Ho, int Wo, int Cc) -> int {
      upsample2_kernel<<<vgg_grid(static_cast<long long>(bn) * 4 * Ho * Wo * (Cc / 4)), 256, 0, st>>>(
          reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(up), bn, Ho, Wo, Cc / 4);
      AP_LAUNCH_CHECK();
      return AP_OK;
    };
    // head: avg_pool2d(8) + fc, then the final BatchNorm + ReLU (in place)
    pool_fc_bwd_kernel<<<bn, 256, 0, st>>>(g_logits + static_cast<size_t>(b0) * h->cfg.num_classes, h->fc_w.as<float>(),
                                           h->cfg.num_classes, h->feat, H * W, G);
    AP_LAUNCH_CHECK();
    rc = bwd_act(G, h->tape[2 * nb]->as<float>(), h->wrn_scale, nullptr, static_cast<long long>(bn) * H * W * h->feat, h->feat);
    if (rc != AP_OK) return rc;
    for (int l = static_cast<int>(nb) - 1; l >= 0; --l) {
      auto& b = *h->wrn[l];
      const float *a = h->tape[2 * l]->as<float>(), *hh = h->tape[2 * l + 1]->as<float>();
      const int Hi = H * b.stride, Wi = W * b.stride;
      const long long npo = static_cast<long long>(bn) * H * W;
      const size_t sb = 6 * nb + 3 * static_cast<size_t>(l);                    // tensor-map slots of this block's three twins
      rc = conv_auto(h, sb, b.t_c2, G, bn, H, W, T1, nullptr, 0, 0, st);         // g_h = conv2^T G, then the ReLU of bn2
      if (rc != AP_OK) return rc;
      relu_mask_kernel<<<vgg_grid(npo * b.cout / 4), 256, 0, st>>>(reinterpret_cast<float4*>(T1), reinterpret_cast<const float4*>(hh),
                                                                   npo * b.cout / 4);
      AP_LAUNCH_CHECK();
      const float* res = nullptr;                                               // what reaches `a` through the shortcut convolution
      if (!b.equal) {
        const float* src = G;
        if (b.stride == 2) {
          rc = upsample(G, T3, H, W, b.cout);
          src = T3;
        }
        if (rc == AP_OK) rc = conv_auto(h, sb + 1, b.t_sc, src, bn, Hi, Wi, T4, nullptr, 0, 0, st);
        res = T4;
      }
      const float* src = T1;
      float* g_a = T2;
      if (rc == AP_OK && b.stride == 2) {
        rc = upsample(T1, T2, H, W, b.cout);
        src = T2, g_a = T1;
      }
      if (rc == AP_OK) rc = conv_auto(h, sb + 2, b.t_c1, src, bn, Hi, Wi, g_a, res, 0, 0, st);   // g_a = conv1^T g_h (+ shortcut part)
      // x feeds bn1 + ReLU -> a, and for an equal-width block also the identity shortcut (+ G); in place in g_a
      if (rc == AP_OK) rc = bwd_act(g_a, a, b.scale, b.equal ? G : nullptr, static_cast<long long>(bn) * Hi * Wi * b.cin, b.cin);
      if (rc != AP_OK) return rc;
      if (g_a == T1) std::swap(G, T1);
      else std::swap(G, T2);
      H = Hi, W = Wi;
    }
    rc = h->t_stem.run(G, bn, 32, 32, g_spec + static_cast<size_t>(b0) * 1024, nullptr, 0, st);
    if (rc != AP_OK) return rc;
  }
  return AP_OK;
}

// ---- DenseNet-BC-depth-growth (models/densenet.py:15-147; `densenet_bc_100_12` is a --classifier_model choice,
// adaptive_attack_eval.py:21).  cfg.base_width = growthRate, cfg.widen_factor = compressionRate.
// Weights: conv1.weight; per dense layer bn1.{weight,bias,running_mean,running_var}, conv1.weight, bn2.{...}, conv2.weight;
// after dense1 / dense2 the transition's bn1.{...}, conv1.weight; bn.{...}; fc.weight, fc.bias.
// torch.cat((x, out), 1) (:36) never copies here: every dense block owns one NHWC tensor of its final width, each layer's 3x3
// convolution writes its `growth` channels straight into its slice, and each layer's BatchNorm reads a channel prefix.
static std::vector<float> dn_phys(const float* src, int rows, int Cl, int C0, int C0p) {
  const int cp = Cl + (C0p - C0);
  std::vector<float> out(static_cast<size_t>(rows) * cp, 0.f);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < Cl; ++c) out[static_cast<size_t>(r) * cp + (c < C0 ? c : c + (C0p - C0))] = src[static_cast<size_t>(r) * Cl + c];
  return out;
}
static int dn_bn(DevBuf& scale, DevBuf& shift, const float* const* w, int Cl, int C0, int C0p) {
  std::vector<float> a(Cl), b(Cl);
  for (int c = 0; c < Cl; ++c) {
    a[c] = w[0][c] / std::sqrt(w[3][c] + 1e-5f);
    b[c] = w[1][c] - w[2][c] * a[c];
  }
  const std::vector<float> ap = dn_phys(a.data(), 1, Cl, C0, C0p), bp = dn_phys(b.data(), 1, Cl, C0, C0p);
  AP_CUDA(scale.upload(ap.data(), sizeof(float) * ap.size()));
  AP_CUDA(shift.upload(bp.data(), sizeof(float) * bp.size()));
  return AP_OK;
}

static int create_dn(ap_classifier_t h, const float* const* w, int n_weights) {
  const ap_classifier_cfg& c = h->cfg;
  const int g = c.base_width, comp = c.widen_factor;
  AP_REQUIRE(c.depth >= 10 && (c.depth - 4) % 6 == 0, "ap_classifier_create: DenseNet-BC depth must be 6n + 4 (got %d)", c.depth);
  AP_REQUIRE(g >= 4 && g % 4 == 0 && g <= 64, "ap_classifier_create: DenseNet growthRate must be a multiple of 4 in [4, 64] (got %d)", g);
  AP_REQUIRE(comp >= 1 && comp <= 4, "ap_classifier_create: DenseNet compressionRate %d out of range", comp);
  AP_REQUIRE(c.in_channels == 1, "ap_classifier_create: DenseNet in_channels must be 1 (NCHW == NHWC)");
  const int n = (c.depth - 4) / 6;
  AP_REQUIRE(n_weights == 1 + 30 * n + 10 + 6, "ap_classifier_create: DenseNet-BC-%d-%d expects %d weight tensors, got %d", c.depth, g,
             1 + 30 * n + 16, n_weights);
  h->dn_growth = g, h->dn_mid = 4 * g;
  int C = 2 * g, i = 1, rc;
  h->stem.keep_host = true;
  rc = h->stem.init(1, C, 3, 3, 1, 1, 1, w[0], nullptr, nullptr, nullptr, nullptr, nullptr);
  if (rc != AP_OK) return rc;
  for (int blk = 0; blk < 3; ++blk) {
    auto B = std::make_unique<ap_classifier_s::DnBlock>();
    B->C0 = C, B->C0p = (C + 3) / 4 * 4, B->stride = B->C0p + g * n, B->H = 32 >> blk;
    for (int l = 0; l < n; ++l) {
      auto L = std::make_unique<ap_classifier_s::DnLayer>();
      const int Cl = C + g * l;
      L->cp = B->C0p + g * l, L->off = L->cp;
      rc = dn_bn(L->scale, L->shift, w + i, Cl, B->C0, B->C0p);
      if (rc != AP_OK) return rc;
      const std::vector<float> w1 = dn_phys(w[i + 4], 4 * g, Cl, B->C0, B->C0p);
      L->c1.keep_host = L->c2.keep_host = true;
      rc = L->c1.init(L->cp, 4 * g, 1, 1, 1, 0, 1, w1.data(), nullptr, w[i + 5], w[i + 6], w[i + 7], w[i + 8]);
      if (rc == AP_OK) rc = L->c2.init(4 * g, g, 3, 3, 1, 1, 1, w[i + 9], nullptr, nullptr, nullptr, nullptr, nullptr);
      if (rc != AP_OK) return rc;
      B->layers.push_back(std::move(L));
      i += 10;
    }
    C += g * n;
    rc = dn_bn(B->t_scale, B->t_shift, w + i, C, B->C0, B->C0p);
    if (rc != AP_OK) return rc;
    if (blk < 2) {                                                                  // Transition, densenet.py:56-71
      B->t_cout = C / comp;
      const std::vector<float> wt = dn_phys(w[i + 4], B->t_cout, C, B->C0, B->C0p);
      B->t_conv.keep_host = true;
      rc = B->t_conv.init(B->stride, B->t_cout, 1, 1, 1, 0, 1, wt.data(), nullptr, nullptr, nullptr, nullptr, nullptr);
      if (rc != AP_OK) return rc;
      i += 5;
      C = B->t_cout;
    } else {                                                                        // bn, relu, avgpool(8), fc: :139-144
      const std::vector<float> fw = dn_phys(w[i + 4], c.num_classes, C, B->C0, B->C0p);
      h->feat = B->stride;
      AP_CUDA(h->fc_w.upload(fw.data(), sizeof(float) * fw.size()));
      AP_CUDA(h->fc_b.upload(w[i + 5], sizeof(float) * c.num_classes));
    }
    h->dn.push_back(std::move(B));
  }
  return AP_OK;
}

static int dn_ensure(ap_classifier_t h, int bn, bool bwd) {
  const size_t need = static_cast<size_t>(bn);
  if (need > h->dn_elems) {
    size_t a_max = 0;
    for (int blk = 0; blk < 3; ++blk) {
      auto& B = *h->dn[blk];
      AP_CUDA(h->dn_x[blk].alloc(need * B.H * B.H * B.stride * sizeof(float)));
      h->dn_gx[blk].release();
      a_max = std::max(a_max, static_cast<size_t>(B.H) * B.H * B.stride);
    }
    AP_CUDA(h->dn_a.alloc(need * a_max * sizeof(float)));
    AP_CUDA(h->dn_h.alloc(need * 1024 * h->dn_mid * sizeof(float)));
    AP_CUDA(h->dn_t.alloc(need * 1024 * ((h->dn[0]->t_cout + 3) / 4 * 4) * sizeof(float)));
    h->dn_elems = need;
    h->bwd_bn = 0;
  }
  if (bwd && bn > h->bwd_bn) {
    h->tape.clear();
    for (int blk = 0; blk < 3; ++blk) {
      auto& B = *h->dn[blk];
      AP_CUDA(h->dn_gx[blk].alloc(h->dn_elems * B.H * B.H * B.stride * sizeof(float)));
      for (size_t l = 0; l < B.layers.size(); ++l) {
        auto d = std::make_unique<DevBuf>();
        AP_CUDA(d->alloc(h->dn_elems * B.H * B.H * h->dn_mid * sizeof(float)));
        h->tape.push_back(std::move(d));
      }
    }
    AP_CUDA(h->tape_x0.alloc(sizeof(float) * h->dn_elems * h->cfg.num_classes));
    h->bwd_bn = static_cast<int>(h->dn_elems);
  }
  return AP_OK;
}

static int dn_prefix(const float* x, int xs, const DevBuf& scale, const DevBuf& shift, float* y, long long npix, int cp, cudaStream_t st) {
  bn_relu_prefix_kernel<<<vgg_grid(npix * (cp / 4)), 256, 0, st>>>(x, xs, scale.as<float4>(), shift.as<float4>(),
                                                                   reinterpret_cast<float4*>(y), npix, cp / 4);
  AP_LAUNCH_CHECK();
  return AP_OK;
}

// One pass over `bn` images; with `tape`, the bottleneck activation relu(bn2(conv1(.))) of every dense layer is kept.
static int dn_pass(ap_classifier_t h, const float* spec, float* logits, int bn, bool tape, cudaStream_t st) {
  const int g = h->dn_growth, mid = h->dn_mid;
  for (int blk = 0; blk < 3; ++blk)
    AP_CUDA(cudaMemsetAsync(h->dn_x[blk].p, 0, static_cast<size_t>(bn) * h->dn[blk]->H * h->dn[blk]->H * h->dn[blk]->stride * sizeof(float), st));
  int rc = h->stem.run_strided(spec, 1, bn, 32, 32, h->dn_x[0].as<float>(), h->dn[0]->stride, nullptr, 0, st);   // densenet.py:134
  if (rc != AP_OK) return rc;
  float* A = h->dn_a.as<float>();
  size_t ti = 0;
  for (int blk = 0; blk < 3; ++blk) {
    auto& B = *h->dn[blk];
    float* X = h->dn_x[blk].as<float>();
    const int H = B.H;
    const long long npix = static_cast<long long>(bn) * H * H;
    for (auto& Lp : B.layers) {                                                      // Bottleneck.forward, :26-36
      auto& L = *Lp;
      float* h1 = tape ? h->tape[ti++]->as<float>() : h->dn_h.as<float>();
      rc = dn_prefix(X, B.stride, L.scale, L.shift, A, npix, L.cp, st);
      if (rc == AP_OK) rc = L.c1.run_strided(A, L.cp, bn, H, H, h1, mid, nullptr, 1, st);
      if (rc == AP_OK) rc = L.c2.run_strided(h1, mid, bn, H, H, X + L.off, B.stride, nullptr, 0, st);
      if (rc != AP_OK) return rc;
    }
    rc = dn_prefix(X, B.stride, B.t_scale, B.t_shift, A, npix, B.stride, st);
    if (rc != AP_OK) return rc;
    if (blk < 2) {                                                                   // Transition.forward, :65-71
      const int cs = (B.t_cout + 3) / 4 * 4;
      rc = B.t_conv.run_strided(A, B.stride, bn, H, H, h->dn_t.as<float>(), cs, nullptr, 0, st);
      if (rc != AP_OK) return rc;
      avgpool2_kernel<<<vgg_grid(npix / 4 * B.t_cout), 256, 0, st>>>(h->dn_t.as<float>(), cs, h->dn_x[blk + 1].as<float>(),
                                                                     h->dn[blk + 1]->stride, bn, H, H, B.t_cout);
      AP_LAUNCH_CHECK();
    } else {
      const size_t smem = sizeof(float) * (h->feat + h->cfg.num_classes);
      pool_fc_kernel<<<bn, 256, smem, st>>>(A, H * H, h->feat, h->fc_w.as<float>(), h->fc_b.as<float>(), h->cfg.num_classes, logits, 0);
      AP_LAUNCH_CHECK();
    }
  }
  (void)g;
  return AP_OK;
}

static int forward_dn(ap_classifier_t h, const float* spec, float* logits, int B, int H0, int W0, cudaStream_t st) {
  AP_REQUIRE(H0 == 32 && W0 == 32, "DenseNet: AvgPool2d(8) + view needs a 32x32 input (got %dx%d)", H0, W0);
  const int chunk = 128;
  int rc = dn_ensure(h, std::min(B, chunk), false);
  if (rc != AP_OK) return rc;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bn = std::min(chunk, B - b0);
    rc = dn_pass(h, spec + static_cast<size_t>(b0) * 1024, logits + static_cast<size_t>(b0) * h->cfg.num_classes, bn, false, st);
    if (rc != AP_OK) return rc;
  }
  return AP_OK;
}

// backward of the DenseNet forward (autograd over models/densenet.py:26-36,65-71,133-146, BatchNorm in eval mode).  The
// gradient of a block's concatenated tensor lives in one tensor of the same layout: walking the layers in reverse, the slice a
// layer wrote has by then received the contributions of every later reader.
static int vjp_dn(ap_classifier_t h, const float* spec, const float* g_logits, float* g_spec, int B, int H0, int W0, cudaStream_t st) {
  AP_REQUIRE(H0 == 32 && W0 == 32, "DenseNet backward: input must be 32x32 (got %dx%d)", H0, W0);
  const int chunk = 32, mid = h->dn_mid;
  if (!h->bwd_ready) {
    int rc = init_dgrad(h->t_stem, h->stem);
    for (auto& Bk : h->dn) {
      for (auto& L : Bk->layers) {
        if (rc == AP_OK) rc = init_dgrad(L->t_c1, L->c1);
        if (rc == AP_OK) rc = init_dgrad(L->t_c2, L->c2);
      }
      if (rc == AP_OK && Bk->t_cout) rc = init_dgrad(Bk->t_conv_t, Bk->t_conv);
    }
    if (rc != AP_OK) return rc;
    h->bwd_ready = true;
  }
  int rc = dn_ensure(h, std::min(B, chunk), true);
  if (rc != AP_OK) return rc;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bn = std::min(chunk, B - b0);
    rc = dn_pass(h, spec + static_cast<size_t>(b0) * 1024, h->tape_x0.as<float>(), bn, true, st);
    if (rc != AP_OK) return rc;
    float *A = h->dn_a.as<float>(), *Hb = h->dn_h.as<float>(), *T = h->dn_t.as<float>();
    auto prefix_bwd = [&](const float* g, const float* x, int xs, const DevBuf& scale, const DevBuf& shift, float* gx, long long npix,
                          int cp, int accumulate) -> int {
      bn_relu_prefix_bwd_kernel<<<vgg_grid(npix * (cp / 4)), 256, 0, st>>>(reinterpret_cast<const float4*>(g), x, xs, scale.as<float4>(),
                                                                          shift.as<float4>(), gx, npix, cp / 4, accumulate);
      AP_LAUNCH_CHECK();
      return AP_OK;
    };
    {   // head: AvgPool2d(8) + fc, then the final BatchNorm + ReLU
      auto& Bk = *h->dn[2];
      pool_fc_bwd_kernel<<<bn, 256, 0, st>>>(g_logits + static_cast<size_t>(b0) * h->cfg.num_classes, h->fc_w.as<float>(),
                                             h->cfg.num_classes, h->feat, Bk.H * Bk.H, A);
      AP_LAUNCH_CHECK();
      rc = prefix_bwd(A, h->dn_x[2].as<float>(), Bk.stride, Bk.t_scale, Bk.t_shift, h->dn_gx[2].as<float>(),
                      static_cast<long long>(bn) * Bk.H * Bk.H, Bk.stride, 0);
      if (rc != AP_OK) return rc;
    }
    size_t ti = h->tape.size();
    for (int blk = 2; blk >= 0; --blk) {
      auto& Bk = *h->dn[blk];
      const int H = Bk.H;
      const long long npix = static_cast<long long>(bn) * H * H;
      float *X = h->dn_x[blk].as<float>(), *GX = h->dn_gx[blk].as<float>();
      for (int l = static_cast<int>(Bk.layers.size()) - 1; l >= 0; --l) {
        auto& L = *Bk.layers[l];
        const float* h1 = h->tape[--ti]->as<float>();
        rc = L.t_c2.run_strided(GX + L.off, Bk.stride, bn, H, H, Hb, mid, nullptr, 0, st);
        if (rc != AP_OK) return rc;
        relu_mask_kernel<<<vgg_grid(npix * mid / 4), 256, 0, st>>>(reinterpret_cast<float4*>(Hb), reinterpret_cast<const float4*>(h1),
                                                                   npix * mid / 4);
        AP_LAUNCH_CHECK();
        rc = L.t_c1.run_strided(Hb, mid, bn, H, H, A, L.cp, nullptr, 0, st);
        if (rc == AP_OK) rc = prefix_bwd(A, X, Bk.stride, L.scale, L.shift, GX, npix, L.cp, 1);
        if (rc != AP_OK) return rc;
      }
      if (blk > 0) {   // through the transition that produced this block's input
        auto& P = *h->dn[blk - 1];
        const int Hp = P.H, cs = (P.t_cout + 3) / 4 * 4;
        const long long npp = static_cast<long long>(bn) * Hp * Hp;
        avgpool2_bwd_kernel<<<vgg_grid(npp * P.t_cout), 256, 0, st>>>(GX, Bk.stride, T, cs, bn, Hp, Hp, P.t_cout);
        AP_LAUNCH_CHECK();
        rc = P.t_conv_t.run_strided(T, cs, bn, Hp, Hp, A, P.stride, nullptr, 0, st);
        if (rc == AP_OK)
          rc = prefix_bwd(A, h->dn_x[blk - 1].as<float>(), P.stride, P.t_scale, P.t_shift, h->dn_gx[blk - 1].as<float>(), npp, P.stride, 0);
        if (rc != AP_OK) return rc;
      }
    }
    rc = h->t_stem.run_strided(h->dn_gx[0].as<float>(), h->dn[0]->stride, bn, 32, 32, g_spec + static_cast<size_t>(b0) * 1024, 1,
                               nullptr, 0, st);
    if (rc != AP_OK) return rc;
  }
  return AP_OK;
}

// ---- M5: state_dict order conv{i}.weight, conv{i}.bias, bn{i}.{weight,bias,running_mean,running_var} (i=1..4), fc1.weight, fc1.bias
static int create_m5(ap_classifier_t h, const float* const* w, int n_weights) {
  const ap_classifier_cfg& c = h->cfg;
  AP_REQUIRE(n_weights == 26, "ap_classifier_create: M5 expects 26 weight tensors, got %d", n_weights);
  AP_REQUIRE(c.m5_first_kernel > 0 && c.m5_stride > 0 && c.m5_channels > 0 && c.m5_channels % 4 == 0,
             "ap_classifier_create: bad M5 configuration");
  const int n = c.m5_channels;
  const int cin[4] = {1, n, n, 2 * n}, cout[4] = {n, n, 2 * n, 2 * n}, ks[4] = {c.m5_first_kernel, 3, 3, 3};
  for (int i = 0; i < 4; ++i) {
    const float* const* p = w + 6 * i;
    h->m5conv[i].keep_host = true;
    int rc = h->m5conv[i].init(cin[i], cout[i], 1, ks[i], i == 0 ? c.m5_stride : 1, 0, 1, p[0], p[1], p[2], p[3], p[4], p[5]);
    if (rc != AP_OK) return rc;
  }
  h->feat = 2 * n;
  AP_CUDA(h->fc_w.upload(w[24], sizeof(float) * c.num_classes * h->feat));
  AP_CUDA(h->fc_b.upload(w[25], sizeof(float) * c.num_classes));
  return AP_OK;
}

static int forward_m5(ap_classifier_t h, const float* wav, float* out, int B, int L, cudaStream_t st) {
  // treat the waveform as an image of height 1: NHWC [B][1][L][1]
  const int n = h->cfg.m5_channels;
  const int L1 = (L - h->cfg.m5_first_kernel) / h->cfg.m5_stride + 1;
  AP_REQUIRE(L1 >= 4, "M5: input too short (L=%d)", L);
  const int chunk = 256;
  int rc = ensure_ws(h, static_cast<size_t>(std::min(B, chunk)) * L1 * 2 * n);
  if (rc != AP_OK) return rc;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bn = std::min(chunk, B - b0);
    float* a = h->buf[0].as<float>();   // conv + BN + ReLU output
    float* p = h->buf[1].as<float>();   // pooled output == next conv input
    const float* in = wav + static_cast<size_t>(b0) * L;
    int len = L;
    for (int i = 0; i < 4; ++i) {
      const ConvLayer& cv = h->m5conv[i];
      const int lo = (len - cv.kw) / cv.stride + 1;
      AP_REQUIRE(lo >= 4, "M5: sequence too short at conv%d", i + 1);
      rc = cv.run(in, bn, 1, len, a, nullptr, 1, st);      // conv + BN + ReLU (M5Net.py:23-25,...)
      if (rc != AP_OK) return rc;
      const long long total = static_cast<long long>(bn) * (lo / 4) * cv.Cout;
      long long blocks = ceil_div_ll(total, 256);
      if (blocks > num_sms() * 8) blocks = num_sms() * 8;
      maxpool4_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(a, p, bn, lo, cv.Cout);   // M5Net.py:26,...
      AP_LAUNCH_CHECK();
      in = p;
      len = lo / 4;
    }
    const size_t smem = sizeof(float) * (h->feat + h->cfg.num_classes);
    pool_fc_kernel<<<bn, 256, smem, st>>>(in, len, h->feat, h->fc_w.as<float>(), h->fc_b.as<float>(), h->cfg.num_classes,
                                          out + static_cast<size_t>(b0) * h->cfg.num_classes, 1);   // M5Net.py:35-38
    AP_LAUNCH_CHECK();
  }
  return AP_OK;
}

// ---- backward of the M5 forward (autograd over M5Net.py:21-38, BatchNorm in eval mode): g_wav = (d logp / d wav)^T g_logp
static int vjp_m5(ap_classifier_t h, const float* wav, const float* g_out, float* g_wav, int B, int L, cudaStream_t st) {
  const int chunk = 64, n = h->cfg.m5_channels, K = h->cfg.num_classes;
  AP_REQUIRE(K <= 64, "M5 backward: at most 64 classes");
  if (!h->bwd_ready) {
    for (int i = 0; i < 4; ++i) {
      // dedicated kernel (see conv1d_cin1_dgrad_kernel) when its weight tile fits the default 48 KB of dynamic shared memory
      if (i == 0 && h->m5conv[0].Cin == 1 && h->m5conv[0].wf_host.size() * sizeof(float) <= 48 * 1024) {
        AP_CUDA(h->m5_w0.upload(h->m5conv[0].wf_host.data(), h->m5conv[0].wf_host.size() * sizeof(float)));
        continue;
      }
      int rc = init_dgrad(h->m5conv_t[i], h->m5conv[i]);
      if (rc != AP_OK) return rc;
    }
    h->bwd_ready = true;
  }
  int len[5], lo[4];
  len[0] = L;
  for (int i = 0; i < 4; ++i) {
    lo[i] = (len[i] - h->m5conv[i].kw) / h->m5conv[i].stride + 1;
    AP_REQUIRE(lo[i] >= 4, "M5: sequence too short at conv%d", i + 1);
    len[i + 1] = lo[i] / 4;
  }
  const int bn_max = std::min(B, chunk);
  // elements per image of the largest buffer: the zero-upsampled gradient of the strided first convolution, (L - k + 1) x n
  const size_t big = static_cast<size_t>(L) * std::max(n, 2);
  if (bn_max > h->bwd_bn) {
    h->tape.clear();
    for (int i = 0; i < 4; ++i) {   // conv + BN + ReLU outputs
      auto d = std::make_unique<DevBuf>();
      AP_CUDA(d->alloc(static_cast<size_t>(bn_max) * lo[i] * h->m5conv[i].Cout * sizeof(float)));
      h->tape.push_back(std::move(d));
    }
    for (auto& g : h->gbuf) AP_CUDA(g.alloc(big * bn_max * sizeof(float)));
    h->bwd_bn = bn_max;
  }
  auto grid_for = [](long long work) {
    long long b = ceil_div_ll(work, 256);
    const long long cap = static_cast<long long>(num_sms()) * 16;
    return static_cast<unsigned>(b < cap ? (b > 0 ? b : 1) : cap);
  };
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bn = std::min(chunk, B - b0);
    float *P0 = h->gbuf[0].as<float>(), *GA = h->gbuf[1].as<float>(), *GB = h->gbuf[2].as<float>(), *LP = h->gbuf[3].as<float>();
    // forward with the tape
    const float* in = wav + static_cast<size_t>(b0) * L;
    for (int i = 0; i < 4; ++i) {
      const ConvLayer& cv = h->m5conv[i];
      float* a = h->tape[i]->as<float>();
      int rc = cv.run(in, bn, 1, len[i], a, nullptr, 1, st);
      if (rc != AP_OK) return rc;
      maxpool4_kernel<<<grid_for(static_cast<long long>(bn) * len[i + 1] * cv.Cout), 256, 0, st>>>(a, P0, bn, lo[i], cv.Cout);
      AP_LAUNCH_CHECK();
      in = P0;   // the pooled tensor is consumed by the next convolution before it is overwritten
      if (i < 3) {   // ping-pong the pooled buffers
        std::swap(P0, GA);
      }
    }
    const size_t smem = sizeof(float) * (h->feat + K);
    pool_fc_kernel<<<bn, 256, smem, st>>>(in, len[4], h->feat, h->fc_w.as<float>(), h->fc_b.as<float>(), K, LP, 1);
    AP_LAUNCH_CHECK();
    // backward
    float* g = h->gbuf[0].as<float>();
    float* t1 = h->gbuf[1].as<float>();
    logsoftmax_pool_fc_bwd_kernel<<<bn, 256, 0, st>>>(g_out + static_cast<size_t>(b0) * K, LP, h->fc_w.as<float>(), K, h->feat, len[4], g);
    AP_LAUNCH_CHECK();
    for (int i = 3; i >= 0; --i) {
      const ConvLayer& cv = h->m5conv[i];
      const float* a = h->tape[i]->as<float>();
      // g: (bn, len[i+1], Cout) -> t1: (bn, lo[i], Cout) through max-pool + ReLU
      maxpool4_relu_bwd_kernel<<<grid_for(static_cast<long long>(bn) * lo[i] * cv.Cout), 256, 0, st>>>(g, a, t1, bn, lo[i], cv.Cout);
      AP_LAUNCH_CHECK();
      if (i == 0 && h->m5_w0.p) {
        const size_t smem = static_cast<size_t>(cv.Cout) * cv.kw * sizeof(float);
        dim3 grid(static_cast<unsigned>(std::min(ceil_div(L, 256), 64)), static_cast<unsigned>(bn));
        conv1d_cin1_dgrad_kernel<<<grid, 256, smem, st>>>(t1, h->m5_w0.as<float>(), g_wav + static_cast<size_t>(b0) * L, L, lo[0],
                                                         cv.kw, cv.stride, cv.Cout);
        AP_LAUNCH_CHECK();
        continue;
      }
      const float* src = t1;
      int lsrc = lo[i];
      if (cv.stride > 1) {   // zero-upsample to length len - k + 1 (stride-1 geometry of the transposed convolution)
        lsrc = len[i] - cv.kw + 1;
        upsample1d_kernel<<<grid_for(static_cast<long long>(bn) * lsrc * cv.Cout), 256, 0, st>>>(t1, GB, bn, lo[i], lsrc, cv.stride, cv.Cout);
        AP_LAUNCH_CHECK();
        src = GB;
      } else if (lo[i] + cv.kw - 1 != len[i]) {
        return fail(AP_ERR_STATE, "M5 backward: unexpected geometry");
      }
      float* dst = i == 0 ? g_wav + static_cast<size_t>(b0) * L : g;
      int rc = h->m5conv_t[i].run(src, bn, 1, lsrc, dst, nullptr, 0, st);   // -> (bn, len[i], Cin)
      if (rc != AP_OK) return rc;
    }
  }
  return AP_OK;
}

// ---- KWS: state_dict order (24 tensors): sepconv.0.{weight,bias}, sepconv.1.{weight,bias},
//      gru.{weight_ih,weight_hh,bias_ih,bias_hh}_l0, ..._l0_reverse, ..._l1, ..._l1_reverse,
//      attn_layer.Wx_b.{weight,bias}, attn_layer.Vt.weight, apply_attn.U.weight
static int create_kws(ap_classifier_t h, const float* const* w, int n_weights) {
  const ap_classifier_cfg& c = h->cfg;
  AP_REQUIRE(n_weights == 24, "ap_classifier_create: KWS expects 24 weight tensors, got %d", n_weights);
  AP_REQUIRE(c.kws_in_size == 32 && c.kws_hidden > 0 && c.kws_hidden <= 128, "ap_classifier_create: bad KWS configuration");
  const int I = c.kws_in_size, H = c.kws_hidden;
  size_t sizes[24];
  sizes[0] = I * 5, sizes[1] = I, sizes[2] = static_cast<size_t>(H) * I, sizes[3] = H;
  int k = 4;
  for (int layer = 0; layer < 2; ++layer)
    for (int dir = 0; dir < 2; ++dir) {
      const int in_l = layer == 0 ? H : 2 * H;
      sizes[k++] = static_cast<size_t>(3) * H * in_l, sizes[k++] = static_cast<size_t>(3) * H * H, sizes[k++] = 3 * H, sizes[k++] = 3 * H;
    }
  sizes[20] = static_cast<size_t>(4) * H * H, sizes[21] = 2 * H, sizes[22] = 2 * H, sizes[23] = static_cast<size_t>(c.num_classes) * 2 * H;
  const float* d[24];
  for (int i = 0; i < 24; ++i) {
    auto buf = std::make_unique<DevBuf>();
    AP_CUDA(buf->upload(w[i], sizes[i] * sizeof(float)));
    d[i] = buf->as<float>();
    h->kws_bufs.push_back(std::move(buf));
  }
  h->kws.dw_w = d[0], h->kws.dw_b = d[1], h->kws.pw_w = d[2], h->kws.pw_b = d[3];
  k = 4;
  for (int layer = 0; layer < 2; ++layer)
    for (int dir = 0; dir < 2; ++dir)
      for (int j = 0; j < 4; ++j) h->kws.gru[layer][dir][j] = d[k++];
  h->kws.wx_w = d[20], h->kws.wx_b = d[21], h->kws.vt_w = d[22], h->kws.u_w = d[23];
  return AP_OK;
}

static int forward_kws(ap_classifier_t h, const float* spec, float* out, int B, int Wf, cudaStream_t st) {
  const int I = h->cfg.kws_in_size, H = h->cfg.kws_hidden;
  AP_REQUIRE(Wf >= 5, "KWS: need at least 5 spectrogram frames (got %d)", Wf);
  const int W1 = (Wf - 5) / 2 + 1, T = (W1 - 1) / 8 + 1;
  const size_t smem = sizeof(float) * (static_cast<size_t>(I) * W1 + T * H + 2 * T * 2 * H + 2 * H + 2 * 2 * 3 * H + T);
  AP_REQUIRE(smem <= 200 * 1024, "KWS: spectrogram too long (%d frames)", Wf);
  static bool attr = false;
  if (!attr) {
    AP_CUDA(cudaFuncSetAttribute(kws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  kws_kernel<<<B, 256, smem, st>>>(spec, Wf, I, H, h->cfg.num_classes, h->kws, out);
  AP_LAUNCH_CHECK();
  return AP_OK;
}

static int vjp_kws(ap_classifier_t h, const float* spec, const float* g_out, float* g_spec, int B, int Wf, cudaStream_t st) {
  const int I = h->cfg.kws_in_size, H = h->cfg.kws_hidden, K = h->cfg.num_classes;
  AP_REQUIRE(Wf >= 5, "KWS: need at least 5 spectrogram frames (got %d)", Wf);
  AP_REQUIRE(K <= 64, "KWS backward: at most 64 classes");
  const int W1 = (Wf - 5) / 2 + 1, T = (W1 - 1) / 8 + 1;
  AP_REQUIRE(T <= 6 * H, "KWS backward: spectrogram too long (%d frames)", Wf);   // g_att lives in a [2][3H] scratch row
  const KwsScratch so = kws_scratch(I, W1, T, H, K);
  const size_t need = static_cast<size_t>(B) * so.total * sizeof(float);
  if (h->gbuf[0].bytes < need) AP_CUDA(h->gbuf[0].alloc(need));
  const size_t smem = sizeof(float) * (2 * H + 2 * 2 * 3 * H + 64);
  kws_vjp_kernel<<<B, 256, smem, st>>>(spec, Wf, I, H, K, h->kws, g_out, h->gbuf[0].as<float>(), g_spec);
  AP_LAUNCH_CHECK();
  return AP_OK;
}

extern "C" int ap_classifier_create(ap_classifier_t* out, const ap_classifier_cfg* cfg, const float* const* weights,
                                    int n_weights, int device) {
  AP_REQUIRE(out && cfg && weights, "ap_classifier_create: null argument");
  *out = nullptr;
  AP_REQUIRE(cfg->num_classes > 0 && cfg->num_classes <= 1024, "ap_classifier_create: bad num_classes %d", cfg->num_classes);
  for (int i = 0; i < n_weights; ++i) AP_REQUIRE(weights[i], "ap_classifier_create: weight pointer %d is null", i);
  int rc = select_device(device);
  if (rc != AP_OK) return rc;
  auto* h = new ap_classifier_s();
  h->cfg = *cfg;
  h->device = device;
  switch (cfg->kind) {
    case AP_CLS_RESNEXT: {
      rc = create_resnext(h, weights, n_weights);
      h->mode = AP_MODE_TF32;
      if (const char* env = std::getenv("AP_CLS_TC_MASK")) h->tc_mask = static_cast<unsigned>(std::atoi(env));
      break;
    }
    case AP_CLS_M5: rc = create_m5(h, weights, n_weights); break;
    case AP_CLS_RESNET:
      rc = create_resnet(h, weights, n_weights);
      h->mode = AP_MODE_TF32;
      break;
    case AP_CLS_KWS: rc = create_kws(h, weights, n_weights); break;
    case AP_CLS_VGG:   // tf32 by default, as for ResNeXt: what cuDNN runs for the reference's convolutions on this GPU
      rc = create_vgg(h, weights, n_weights);
      h->mode = AP_MODE_TF32;
      break;
    case AP_CLS_WRN:
      rc = create_wrn(h, weights, n_weights);
      h->mode = AP_MODE_TF32;
      break;
    case AP_CLS_DENSENET: rc = create_dn(h, weights, n_weights); break;
    default: rc = fail(AP_ERR_INVALID, "ap_classifier_create: unknown classifier kind %d", cfg->kind);
  }
  if (rc != AP_OK) {
    delete h;
    return rc;
  }
  *out = h;
  return AP_OK;
}

extern "C" void ap_classifier_destroy(ap_classifier_t h) { delete h; }

extern "C" int ap_classifier_forward(ap_classifier_t h, const float* input, float* logits, int B, int in_len, void* stream) {
  AP_REQUIRE(h && input && logits, "ap_classifier_forward: null argument");
  AP_REQUIRE(B > 0, "ap_classifier_forward: B must be positive (got %d)", B);
  AP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (h->cfg.kind) {
    case AP_CLS_RESNEXT: return forward_resnext(h, input, logits, B, st);
    case AP_CLS_M5: return forward_m5(h, input, logits, B, in_len, st);
    case AP_CLS_RESNET: return forward_resnet(h, input, logits, B, 32, in_len, st);
    case AP_CLS_VGG: return forward_vgg(h, input, logits, B, 32, in_len, st);
    case AP_CLS_WRN: return forward_wrn(h, input, logits, B, 32, in_len, st);
    case AP_CLS_DENSENET: return forward_dn(h, input, logits, B, 32, in_len, st);
    default: return forward_kws(h, input, logits, B, in_len, st);
  }
}

// g_input = (d logits / d input)^T g_logits.  ResNeXt (the SC09 default victim, adaptive_attack_eval.py:21), M5 and RCNN_KWS.
extern "C" int ap_classifier_vjp(ap_classifier_t h, const float* input, const float* g_logits, float* g_input, int B, int in_len,
                                 void* stream) {
  AP_REQUIRE(h && input && g_logits && g_input, "ap_classifier_vjp: null argument");
  AP_REQUIRE(B > 0, "ap_classifier_vjp: B must be positive");
  AP_CUDA(cudaSetDevice(h->device));
  if (h->cfg.kind == AP_CLS_RESNET) return vjp_resnet(h, input, g_logits, g_input, B, in_len, in_len, static_cast<cudaStream_t>(stream));
  if (h->cfg.kind == AP_CLS_DENSENET) return vjp_dn(h, input, g_logits, g_input, B, in_len, in_len, static_cast<cudaStream_t>(stream));
  if (h->cfg.kind == AP_CLS_WRN) return vjp_wrn(h, input, g_logits, g_input, B, in_len, in_len, static_cast<cudaStream_t>(stream));
  if (h->cfg.kind == AP_CLS_VGG) return vjp_vgg(h, input, g_logits, g_input, B, in_len, in_len, static_cast<cudaStream_t>(stream));
  if (h->cfg.kind == AP_CLS_KWS) return vjp_kws(h, input, g_logits, g_input, B, in_len, static_cast<cudaStream_t>(stream));
  if (h->cfg.kind == AP_CLS_M5) return vjp_m5(h, input, g_logits, g_input, B, in_len, static_cast<cudaStream_t>(stream));
  AP_REQUIRE(in_len == 32, "ap_classifier_vjp: ResNeXt input is (B, 1, 32, 32)");
  return vjp_resnext(h, input, g_logits, g_input, B, static_cast<cudaStream_t>(stream));
}

extern "C" int ap_classifier_set_mode(ap_classifier_t h, int mode) {
  AP_REQUIRE(h, "ap_classifier_set_mode: null handle");
  AP_REQUIRE(mode == AP_MODE_FP32 || mode == AP_MODE_TF32, "ap_classifier_set_mode: mode must be AP_MODE_FP32 or AP_MODE_TF32");
  AP_REQUIRE(mode == AP_MODE_FP32 || h->cfg.kind == AP_CLS_RESNEXT || h->cfg.kind == AP_CLS_RESNET || h->cfg.kind == AP_CLS_VGG ||
                 h->cfg.kind == AP_CLS_WRN,
             "ap_classifier_set_mode: only ResNeXt, ResNet, VGG and WideResNet have tensor-core convolutions");
  if (mode != h->mode) h->plans.clear(), h->tc_slots.clear();
  h->mode = mode;
  return AP_OK;
}
extern "C" int ap_classifier_get_mode(ap_classifier_t h) { return h ? h->mode : AP_ERR_INVALID; }
