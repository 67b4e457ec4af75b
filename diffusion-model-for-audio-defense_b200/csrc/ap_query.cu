// Black-box query serving: the NES gradient estimator of FAKEBOB
// (robustness_eval/_NES.py:14-56) around the defended forward pass, and the per-query loss the attack drivers use
// (robustness_eval/_utils.py:113-125, _EOT.py:40-42).
//
//   nes_perturb_kernel : out[a, first + j] = z*sigma + x[a],  out[a, first + H + j] = (-z)*sigma + x[a]   (j < H = S/2)
//   nes_grad_kernel    : grad[a, n] (+)= scale * sum_j (loss[a, first + j] - loss[a, first + H + j]) * z[a, j, n]
//   query_loss_kernel  : per-row cross entropy / margin loss + argmax decision
//
// The antithetic noise is never stored: the gradient kernel REGENERATES z[a, j, :] from the same Philox counters the
// perturb kernel used (element e of the (A, H, L) noise tensor = lane e % 4 of block offset + e / 4), so a NES draw
// costs one write of the query batch and one write of the gradient in HBM instead of the reference's noise tensor
// (write) + cat copy + eval_input (write) + loss*noise product (read + write) + mean (read).  HBM-bound byte work.
#include "ap_common.cuh"
#include "ap_internal.h"
#include "ap_philox.cuh"

namespace ap {

namespace {

struct NoiseSrc {
  const float* z;  // host-supplied noise (A, H, L) or nullptr -> Philox
  uint64_t seed, offset;
  __device__ __forceinline__ void quad(long long e, float (&v)[4]) const {  // e % 4 == 0, whole quad inside the tensor
    if (z) {
      const float4 t = *reinterpret_cast<const float4*>(z + e);
      v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    } else {
      normal4(offset + static_cast<uint64_t>(e >> 2), seed, v);
    }
  }
  __device__ __forceinline__ float at(long long e) const {  // ragged path: one element
    if (z) return z[e];
    float v[4];
    normal4(offset + static_cast<uint64_t>(e >> 2), seed, v);
    return v[e & 3];
  }
};

// grid.x strides over the quads of one (a, j) noise row, grid.y = a * H + j.  VEC: L % 4 == 0 and 16 B aligned pointers.
template <bool VEC> __global__ void __launch_bounds__(256) nes_perturb_kernel(const float* __restrict__ x, float sigma,
                                                                               NoiseSrc ns, int first, int H, int L,
                                                                               float* __restrict__ out) {
  const int row = blockIdx.y;  // a * H + j
  const int a = row / H, j = row - a * H;
  const int R = 2 * H + first;
  const float* xa = x + static_cast<long long>(a) * L;
  float* plus = out + (static_cast<long long>(a) * R + first + j) * L;
  float* minus = plus + static_cast<long long>(H) * L;
  float* clean = (first && j == 0) ? out + static_cast<long long>(a) * R * L : nullptr;
  const long long e0 = static_cast<long long>(row) * L;
  const int nq = (L + 3) >> 2;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
    const int n = q << 2;
    if (VEC) {
      float zz[4];
      ns.quad(e0 + n, zz);
      const float4 xv = *reinterpret_cast<const float4*>(xa + n);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
      float p[4], m[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // _NES.py:24: noise * sigma + x, with noise = cat(z, -z)
        const float zs = __fmul_rn(zz[k], sigma);
        p[k] = __fadd_rn(zs, xs[k]);
        m[k] = __fadd_rn(-zs, xs[k]);
      }
      *reinterpret_cast<float4*>(plus + n) = make_float4(p[0], p[1], p[2], p[3]);
      *reinterpret_cast<float4*>(minus + n) = make_float4(m[0], m[1], m[2], m[3]);
      if (clean) *reinterpret_cast<float4*>(clean + n) = xv;  // 0 * sigma + x
    } else {
      for (int k = 0; k < 4 && n + k < L; ++k) {
        const float zs = __fmul_rn(ns.at(e0 + n + k), sigma);
        const float xs = xa[n + k];
        plus[n + k] = __fadd_rn(zs, xs);
        minus[n + k] = __fadd_rn(-zs, xs);
        if (clean) clean[n + k] = xs;
      }
    }
  }
}

// grid.y = a; each thread owns one quad of positions and walks the H antithetic pairs.  coef[j] = loss_+ - loss_- in smem.
template <bool VEC> __global__ void __launch_bounds__(128) nes_grad_kernel(const float* __restrict__ loss, NoiseSrc ns,
                                                                            int first, int H, int L, float scale,
                                                                            int accumulate, float* __restrict__ grad) {
  extern __shared__ float coef[];
  const int a = blockIdx.y;
  const int R = 2 * H + first;
  const float* la = loss + static_cast<long long>(a) * R + first;
  for (int j = threadIdx.x; j < H; j += blockDim.x) coef[j] = __fsub_rn(la[j], la[H + j]);
  __syncthreads();
  float* ga = grad + static_cast<long long>(a) * L;
  const int nq = (L + 3) >> 2;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
    const int n = q << 2;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    long long e = static_cast<long long>(a) * H * L + n;
    for (int j = 0; j < H; ++j, e += L) {
      const float c = coef[j];
      if (VEC) {
        float zz[4];
        ns.quad(e, zz);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = fmaf(c, zz[k], acc[k]);
      } else {
        for (int k = 0; k < 4 && n + k < L; ++k) acc[k] = fmaf(c, ns.at(e + k), acc[k]);
      }
    }
    if (VEC) {
      float4 g = make_float4(acc[0] * scale, acc[1] * scale, acc[2] * scale, acc[3] * scale);
      if (accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(ga + n);
        g.x += o.x, g.y += o.y, g.z += o.z, g.w += o.w;
      }
      *reinterpret_cast<float4*>(ga + n) = g;
    } else {
      for (int k = 0; k < 4 && n + k < L; ++k) ga[n + k] = acc[k] * scale + (accumulate ? ga[n + k] : 0.f);
    }
  }
}

// One thread per query row.  kind 0: nn.CrossEntropyLoss(reduction='none') (_utils.py:117); kind 1: the CSI margin
// loss of SEC4SR_MarginLoss (_utils.py:73-84: score_real + confidence - score_other, sign flipped when targeted,
// optionally clipped at 0, :98-99).  pred: argmax with torch.max's first-index tie-break (_EOT.py:40).
__global__ void __launch_bounds__(128) query_loss_kernel(const float* __restrict__ scores, const long long* __restrict__ y,
                                                         int B, int K, int kind, int targeted, float confidence,
                                                         int clip, float* __restrict__ loss, int* __restrict__ pred) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= B) return;
  const float* s = scores + static_cast<long long>(row) * K;
  const int label = static_cast<int>(y[row]);
  float best = s[0];
  int arg = 0;
  for (int k = 1; k < K; ++k) {
    const float v = s[k];
    if (v > best || (v != v && best == best)) best = v, arg = k;
  }
  if (pred) pred[row] = arg;
  if (!loss) return;
  float out;
  if (label < 0 || label >= K) {  // the reference raises on such a target; no host sync here, so the row reads NaN
    out = __int_as_float(0x7fc00000);
  } else if (kind == 0) {
    float sum = 0.f;
    for (int k = 0; k < K; ++k) sum += expf(s[k] - best);
    out = -((s[label] - best) - logf(sum));  // -log_softmax(scores)[label]
  } else {
    const float real = s[label];
    float other = -10000.f;  // _utils.py:76: max((1 - onehot) * scores - onehot * 10000)
    for (int k = 0; k < K; ++k)
      if (k != label) other = fmaxf(other, s[k]);
    out = targeted ? __fsub_rn(__fadd_rn(other, confidence), real) : __fsub_rn(__fadd_rn(real, confidence), other);
    if (clip) out = fmaxf(out, 0.f);
  }
  loss[row] = out;
}

// Backward of query_loss_kernel: g_scores[b, :] = g_loss[b] * d loss[b] / d scores[b, :] (the EOT wrapper with
// use_grad=True backpropagates ones through the loss, _EOT.py:43-44).  Cross entropy: softmax - onehot.  Margin:
// +-(onehot(label) - onehot(argmax other)), zero where the clip is active or no other class exceeds -10000.
__global__ void __launch_bounds__(128) query_loss_vjp_kernel(const float* __restrict__ scores,
                                                             const long long* __restrict__ y,
                                                             const float* __restrict__ g_loss, int B, int K, int kind,
                                                             int targeted, float confidence, int clip,
                                                             float* __restrict__ g_scores) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= B) return;
  const float* s = scores + static_cast<long long>(row) * K;
  float* gs = g_scores + static_cast<long long>(row) * K;
  const int label = static_cast<int>(y[row]);
  const float g = g_loss[row];
  if (label < 0 || label >= K) {
    for (int k = 0; k < K; ++k) gs[k] = __int_as_float(0x7fc00000);
    return;
  }
  if (kind == 0) {
    float best = s[0];
    for (int k = 1; k < K; ++k) best = fmaxf(best, s[k]);
    float sum = 0.f;
    for (int k = 0; k < K; ++k) sum += expf(s[k] - best);
    const float inv = 1.f / sum;
    for (int k = 0; k < K; ++k) gs[k] = g * (expf(s[k] - best) * inv - (k == label ? 1.f : 0.f));
  } else {
    float other = -10000.f;
    int arg = -1;
    for (int k = 0; k < K; ++k)
      if (k != label && s[k] > other) other = s[k], arg = k;  // first maximal index, like torch.max's backward
    const float val = targeted ? (other + confidence) - s[label] : (s[label] + confidence) - other;
    const float sign = (clip && !(val > 0.f)) ? 0.f : (targeted ? -g : g);
    for (int k = 0; k < K; ++k) gs[k] = k == label ? sign : (k == arg ? -sign : 0.f);
  }
}

}  // namespace

}  // namespace ap

// ================================================================================================ C ABI
using namespace ap;

extern "C" {

uint64_t ap_nes_noise_blocks(int A, int S, int L) {
  return (static_cast<uint64_t>(A) * static_cast<uint64_t>(S / 2) * static_cast<uint64_t>(L) + 3) / 4;
}

int ap_nes_perturb(const float* x, float sigma, const float* z, uint64_t seed, uint64_t offset, int first, float* out,
                   int A, int S, int L, void* stream) {
  AP_REQUIRE(x && out, "ap_nes_perturb: null pointer");
  AP_REQUIRE(A > 0 && L > 0 && S >= 2 && S % 2 == 0, "ap_nes_perturb: need A, L > 0 and an even S >= 2 (got %d, %d, %d)", A, L, S);
  AP_REQUIRE(first == 0 || first == 1, "ap_nes_perturb: first must be 0 or 1");
  const int H = S / 2;
  AP_REQUIRE(static_cast<long long>(A) * H <= 65535, "ap_nes_perturb: A * S / 2 = %lld exceeds 65535 rows per call",
             static_cast<long long>(A) * H);
  const bool vec = L % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                                   reinterpret_cast<uintptr_t>(z)) & 15u) == 0;
  const int nq = (L + 3) / 4;
  dim3 grid(static_cast<unsigned>(std::min(ceil_div(nq, 256), 64)), static_cast<unsigned>(A * H));
  NoiseSrc ns{z, seed, offset};
  auto st = static_cast<cudaStream_t>(stream);
  if (vec) nes_perturb_kernel<true><<<grid, 256, 0, st>>>(x, sigma, ns, first, H, L, out);
  else nes_perturb_kernel<false><<<grid, 256, 0, st>>>(x, sigma, ns, first, H, L, out);
  AP_LAUNCH_CHECK();
  return AP_OK;
}

int ap_nes_gradient(const float* loss, const float* z, uint64_t seed, uint64_t offset, int first, float scale,
                    int accumulate, float* grad, int A, int S, int L, void* stream) {
  AP_REQUIRE(loss && grad, "ap_nes_gradient: null pointer");
  AP_REQUIRE(A > 0 && L > 0 && S >= 2 && S % 2 == 0, "ap_nes_gradient: need A, L > 0 and an even S >= 2 (got %d, %d, %d)", A, L, S);
  AP_REQUIRE(first == 0 || first == 1, "ap_nes_gradient: first must be 0 or 1");
  AP_REQUIRE(A <= 65535 && S / 2 <= 8192, "ap_nes_gradient: A <= 65535 and S <= 16384 per call");
  const int H = S / 2;
  const bool vec = L % 4 == 0 && ((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(z)) & 15u) == 0;
  const int nq = (L + 3) / 4;
  dim3 grid(static_cast<unsigned>(ceil_div(nq, 128)), static_cast<unsigned>(A));
  NoiseSrc ns{z, seed, offset};
  auto st = static_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(H) * sizeof(float);
  if (vec) nes_grad_kernel<true><<<grid, 128, smem, st>>>(loss, ns, first, H, L, scale, accumulate, grad);
  else nes_grad_kernel<false><<<grid, 128, smem, st>>>(loss, ns, first, H, L, scale, accumulate, grad);
  AP_LAUNCH_CHECK();
  return AP_OK;
}

int ap_query_loss(const float* scores, const long long* labels, int B, int K, int kind, int targeted, float confidence,
                  int clip, float* loss, int* pred, void* stream) {
  AP_REQUIRE(scores && labels && (loss || pred), "ap_query_loss: null pointer");
  AP_REQUIRE(B >= 0 && K > 0 && K <= 4096, "ap_query_loss: bad shape (B %d, K %d)", B, K);
  AP_REQUIRE(kind == AP_LOSS_ENTROPY || kind == AP_LOSS_MARGIN, "ap_query_loss: unknown loss kind %d", kind);
  if (B == 0) return AP_OK;
  query_loss_kernel<<<ceil_div(B, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(scores, labels, B, K, kind, targeted,
                                                                                    confidence, clip, loss, pred);
  AP_LAUNCH_CHECK();
  return AP_OK;
}

int ap_query_loss_vjp(const float* scores, const long long* labels, const float* g_loss, int B, int K, int kind,
                      int targeted, float confidence, int clip, float* g_scores, void* stream) {
  AP_REQUIRE(scores && labels && g_loss && g_scores, "ap_query_loss_vjp: null pointer");
  AP_REQUIRE(B >= 0 && K > 0 && K <= 4096, "ap_query_loss_vjp: bad shape (B %d, K %d)", B, K);
  AP_REQUIRE(kind == AP_LOSS_ENTROPY || kind == AP_LOSS_MARGIN, "ap_query_loss_vjp: unknown loss kind %d", kind);
  if (B == 0) return AP_OK;
  query_loss_vjp_kernel<<<ceil_div(B, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      scores, labels, g_loss, B, K, kind, targeted, confidence, clip, g_scores);
  AP_LAUNCH_CHECK();
  return AP_OK;
}

}  // extern "C"
