// Convolution as an implicit GEMM over NHWC activations (FFMA path) with its optional TF32 tensor-core twin: shared by the
// classifiers (ap_classifier.cu) and the spectrogram UNet (ap_unet.cu).
#pragma once
#include <cmath>
#include <vector>

#include "ap_common.cuh"
#include "ap_conv_tc.h"
#include "ap_internal.h"
#include "ap_sgemm.cuh"

namespace ap {

// ---------------------------------------------------------------------------------------------- conv as implicit GEMM
// in: NHWC [B][H][W][Ctot]; group z uses channels [z*Cg, (z+1)*Cg); k = (r*kw + s)*Cg + c
template <bool VEC> struct Conv2dLoader {
  const float* in;
  int H, W, Ctot, Cg, kh, kw, stride, pad, pad_h, Ho, Wo;   // pad: width padding; pad_h: height padding (0 for 1 x k kernels)
  __device__ __forceinline__ float one(int z, int b, int oh, int ow, int k) const {
    const int rs = k / Cg, c = k - rs * Cg, r = rs / kw, s = rs - r * kw;
    const int ih = oh * stride + r - pad_h, iw = ow * stride + s - pad;
    if (ih < 0 || ih >= H || iw < 0 || iw >= W) return 0.f;
    return in[((static_cast<long long>(b) * H + ih) * W + iw) * Ctot + z * Cg + c];
  }
  __device__ __forceinline__ float4 load4(int z, int m, int k, int M, int K) const {
    if (m >= M || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
    const int b = m / (Ho * Wo), rem = m - b * (Ho * Wo), oh = rem / Wo, ow = rem - oh * Wo;
    if (VEC) {
      const int rs = k / Cg, c = k - rs * Cg, r = rs / kw, s = rs - r * kw;
      const int ih = oh * stride + r - pad_h, iw = ow * stride + s - pad;
      if (ih < 0 || ih >= H || iw < 0 || iw >= W) return make_float4(0.f, 0.f, 0.f, 0.f);
      return *reinterpret_cast<const float4*>(in + ((static_cast<long long>(b) * H + ih) * W + iw) * Ctot + z * Cg + c);
    }
    float4 v;
    v.x = one(z, b, oh, ow, k);
    v.y = k + 1 < K ? one(z, b, oh, ow, k + 1) : 0.f;
    v.z = k + 2 < K ? one(z, b, oh, ow, k + 2) : 0.f;
    v.w = k + 3 < K ? one(z, b, oh, ow, k + 3) : 0.f;
    return v;
  }
};
// out[m][z*Ng + n] = act(acc + bias [+ residual])
struct ConvEpi {
  float* out;
  const float* bias;
  const float* residual;
  int Ctot, Ng, relu;
  __device__ __forceinline__ void put(int z, int m, int n, float acc) const {
    if (n >= Ng) return;
    const int ch = z * Ng + n;
    const long long idx = static_cast<long long>(m) * Ctot + ch;
    float v = acc + bias[ch];
    if (residual) v += residual[idx];
    out[idx] = relu ? fmaxf(v, 0.f) : v;
  }
  __device__ __forceinline__ void store(int z, int m, int n0, int tx, const float (&lo)[4], const float (&hi)[4], int) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      put(z, m, n0 + tx * 4 + j, lo[j]);
      put(z, m, n0 + 64 + tx * 4 + j, hi[j]);
    }
  }
};

// 3 x 3, stride 1, pad 1 convolution to ONE output channel (the UNet's output layer; the data gradient of a first layer that reads a
// one-channel image): the implicit GEMM would compute a 16-column tile for one useful column.  CTA = (image, strip of R output rows)
// with the R + 2 input rows in shared memory; a warp owns one pixel at a time: 9 taps x (one float4 per lane), then a warp reduction.
static __global__ void __launch_bounds__(256) conv3x3_c1_kernel(const float* __restrict__ in, const float* __restrict__ w9 /*[9][C]*/, float bias,
                                                         const float* __restrict__ residual, float* __restrict__ out, int H, int W, int C,
                                                         int R, int relu) {
  extern __shared__ float4 c1_rows[];                // [(R + 2)][W][C / 4]
  const int strips = (H + R - 1) / R, b = blockIdx.x / strips, h0 = (blockIdx.x % strips) * R;
  const int C4 = C >> 2, rowq = W * C4;
  const float4* in4 = reinterpret_cast<const float4*>(in) + static_cast<size_t>(b) * H * rowq;
#pragma unroll 4
  for (int i = threadIdx.x; i < (R + 2) * rowq; i += blockDim.x) {
    const int r = i / rowq, ih = h0 - 1 + r;
    c1_rows[i] = (ih >= 0 && ih < H) ? in4[static_cast<size_t>(ih) * rowq + (i - r * rowq)] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float4* w4 = reinterpret_cast<const float4*>(w9);
  const int npx = (H - h0 < R ? H - h0 : R) * W;
  const bool reg_w = C4 <= 32;                       // up to 128 channels: the lane's nine weight vectors stay in registers
  float4 wr[9];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) wr[tap] = (reg_w && lane < C4) ? __ldg(w4 + tap * C4 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int px = warp; px < npx; px += nw) {
    const int r = px / W, ow = px - r * W;
    float acc = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int dr = tap / 3, iw = ow + (tap - dr * 3) - 1;
      if (iw < 0 || iw >= W) continue;               // warp-uniform
      const float4* src = c1_rows + ((r + dr) * W + iw) * C4;
      if (reg_w) {
        if (lane < C4) {
          const float4 v = src[lane], k = wr[tap];
          acc = fmaf(v.x, k.x, acc), acc = fmaf(v.y, k.y, acc), acc = fmaf(v.z, k.z, acc), acc = fmaf(v.w, k.w, acc);
        }
      } else {
        for (int c4 = lane; c4 < C4; c4 += 32) {
          const float4 v = src[c4], k = __ldg(w4 + tap * C4 + c4);
          acc = fmaf(v.x, k.x, acc), acc = fmaf(v.y, k.y, acc), acc = fmaf(v.z, k.z, acc), acc = fmaf(v.w, k.w, acc);
        }
      }
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const size_t idx = (static_cast<size_t>(b) * H + h0 + r) * W + ow;
      float v = acc + bias;
      if (residual) v += residual[idx];
      out[idx] = relu ? fmaxf(v, 0.f) : v;
    }
  }
}

struct ConvLayer {
  int Cin = 0, Cout = 0, kh = 1, kw = 1, stride = 1, pad = 0, groups = 1;
  int Cg = 0, Ng = 0, Ngp = 0, K = 0;
  DevBuf w, bias;
  DevBuf w9;            // [9][Cin] tap-major weights of a 3 x 3 / stride 1 / pad 1 layer with one output channel (conv3x3_c1_kernel)
  float bias0 = 0.f;
  std::vector<float> wf_host;   // folded torch-layout weights [Cout][Cin/groups][kh][kw], kept when `keep_host` (backward pass)
  bool keep_host = false;
  ConvTc tc;            // TF32 tensor-core twin (weights [Cout][K] K-major), built when the shape allows
  bool has_tc = false;
  // w_t: torch [Cout][Cin/groups][kh][kw]; optional BatchNorm(eval) folded: scale = gamma/sqrt(var+eps), shift = beta - mean*scale
  int init(int cin, int cout, int kh_, int kw_, int stride_, int pad_, int groups_, const float* w_t, const float* conv_bias,
           const float* bn_w, const float* bn_b, const float* bn_m, const float* bn_v, bool want_tc = false) {
    Cin = cin, Cout = cout, kh = kh_, kw = kw_, stride = stride_, pad = pad_, groups = groups_;
    Cg = cin / groups, Ng = cout / groups, Ngp = ((Ng + 127) / 128) * 128, K = kh * kw * Cg;
    std::vector<float> wp(static_cast<size_t>(groups) * K * Ngp, 0.f), bp(cout), wf(static_cast<size_t>(cout) * K);
    for (int o = 0; o < cout; ++o) {
      float scale = 1.f, shift = conv_bias ? conv_bias[o] : 0.f;
      if (bn_w) {
        scale = bn_w[o] / std::sqrt(bn_v[o] + 1e-5f);
        shift = bn_b[o] + (shift - bn_m[o]) * scale;
      }
      bp[o] = shift;
      const int g = o / Ng, n = o - g * Ng;
      for (int c = 0; c < Cg; ++c)
        for (int r = 0; r < kh; ++r)
          for (int s = 0; s < kw; ++s)
          {
            const size_t ti = ((static_cast<size_t>(o) * Cg + c) * kh + r) * kw + s;
            wf[ti] = w_t[ti] * scale;
            wp[(static_cast<size_t>(g) * K + (r * kw + s) * Cg + c) * Ngp + n] = wf[ti];
          }
    }
    AP_CUDA(w.upload(wp.data(), wp.size() * sizeof(float)));
    AP_CUDA(bias.upload(bp.data(), bp.size() * sizeof(float)));
    if (keep_host) wf_host = wf;
    if (cout == 1 && kh == 3 && kw == 3 && stride == 1 && pad == 1 && groups == 1 && cin % 4 == 0) {
      std::vector<float> t9(static_cast<size_t>(9) * cin);
      for (int c = 0; c < cin; ++c)
        for (int tap = 0; tap < 9; ++tap) t9[static_cast<size_t>(tap) * cin + c] = wf[static_cast<size_t>(c) * 9 + tap];
      AP_CUDA(w9.upload(t9.data(), t9.size() * sizeof(float)));
      bias0 = bp[0];
    }
    // structural part of conv_tc_supported (the spatial part is checked when the layer is bound to buffers)
    if (want_tc && Cg % 32 == 0 && conv_tc_n_tile(Ng) != 0) {
      int rc = tc.init(cin, cout, kh, kw, stride, pad, groups, wf.data(), bp.data());
      if (rc != AP_OK) return rc;
      has_tc = true;
    }
    return AP_OK;
  }
  int run(const float* in, int B, int H, int W, float* out, const float* residual, int relu, cudaStream_t st) const {
    return run_strided(in, Cin, B, H, W, out, Cout, residual, relu, st);
  }
  // in / out are channel slices of wider NHWC tensors: pixel strides in_ctot / out_ctot floats (DenseNet's concatenation)
  int run_strided(const float* in, int in_ctot, int B, int H, int W, float* out, int out_ctot, const float* residual, int relu,
                  cudaStream_t st) const {
    const int pad_h = kh == 1 ? 0 : pad;   // 1 x k kernels (conv1d as a height-1 image) pad the width only
    const int Ho = (H + 2 * pad_h - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
    const long long M = static_cast<long long>(B) * Ho * Wo;
    if (M >= (1ll << 31)) return fail(AP_ERR_INVALID, "conv: too many output pixels");
    if (w9.p && in_ctot == Cin && out_ctot == 1 && (reinterpret_cast<uintptr_t>(in) & 15u) == 0) {
      const size_t row_bytes = static_cast<size_t>(W) * Cin * sizeof(float);
      int R = static_cast<int>((96 * 1024) / row_bytes) - 2;
      if (R > H) R = H;
      if (R >= 1) {
        const size_t smem = (R + 2) * row_bytes;
        if (smem > 48 * 1024) {
          cudaError_t ea = cudaFuncSetAttribute(conv3x3_c1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
          if (ea != cudaSuccess) return fail(AP_ERR_CUDA, "conv3x3_c1_kernel attribute: %s", cudaGetErrorString(ea));
        }
        conv3x3_c1_kernel<<<B * ((H + R - 1) / R), 256, smem, st>>>(in, w9.as<float>(), bias0, residual, out, H, W, Cin, R, relu);
        cudaError_t el = cudaGetLastError();
        if (el != cudaSuccess) return fail(AP_ERR_CUDA, "conv3x3_c1_kernel launch: %s", cudaGetErrorString(el));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return AP_OK;
      }
    }
    ConvEpi ep{out, bias.as<float>(), residual, out_ctot, Ng, relu};
    cudaError_t e;
    const bool narrow = Ng <= 16;   // 128 x 16 tiles instead of 128 x 128: DenseNet's growth convolutions, 1-channel data gradients
    if (Cg % 4 == 0 && in_ctot % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15u) == 0) {
      Conv2dLoader<true> al{in, H, W, in_ctot, Cg, kh, kw, stride, pad, pad_h, Ho, Wo};
      e = narrow ? sgemm::launch_n16(al, w.as<float>(), Ngp, static_cast<long long>(K) * Ngp, groups, static_cast<int>(M), Ng, K, ep, st)
                 : sgemm::launch(al, w.as<float>(), Ngp, static_cast<long long>(K) * Ngp, groups, static_cast<int>(M), Ng, K, ep, st);
    } else {
      Conv2dLoader<false> al{in, H, W, in_ctot, Cg, kh, kw, stride, pad, pad_h, Ho, Wo};
      e = narrow ? sgemm::launch_n16(al, w.as<float>(), Ngp, static_cast<long long>(K) * Ngp, groups, static_cast<int>(M), Ng, K, ep, st)
                 : sgemm::launch(al, w.as<float>(), Ngp, static_cast<long long>(K) * Ngp, groups, static_cast<int>(M), Ng, K, ep, st);
    }
    if (e != cudaSuccess) return fail(AP_ERR_CUDA, "conv launch: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return AP_OK;
  }
};

// Data-gradient twin of a convolution: g_x = conv(g_y [zero-upsampled by the stride], W^T rotated 180 degrees), stride 1,
// padding k - 1 - pad, same groups.  `L` must have kept its folded weights (keep_host).
inline int init_dgrad(ConvLayer& T, const ConvLayer& L, bool want_tc = false) {
  if (L.wf_host.empty()) return fail(AP_ERR_STATE, "init_dgrad: the forward layer did not keep its weights");
  const int Cg = L.Cg, Ng = L.Ng, kh = L.kh, kw = L.kw;
  std::vector<float> wt(static_cast<size_t>(L.Cin) * Ng * kh * kw);
  for (int g = 0; g < L.groups; ++g)
    for (int c = 0; c < Cg; ++c)
      for (int n = 0; n < Ng; ++n)
        for (int r = 0; r < kh; ++r)
          for (int q = 0; q < kw; ++q)
            wt[((static_cast<size_t>(g * Cg + c) * Ng + n) * kh + r) * kw + q] =
                L.wf_host[((static_cast<size_t>(g * Ng + n) * Cg + c) * kh + (kh - 1 - r)) * kw + (kw - 1 - q)];
  return T.init(L.Cout, L.Cin, kh, kw, 1, kw - 1 - L.pad, L.groups, wt.data(), nullptr, nullptr, nullptr, nullptr, nullptr, want_tc);
}

}  // namespace ap
