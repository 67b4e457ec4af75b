// Spectrogram-domain diffusion purifier ("Diffusion-Spec"): the UNet eps-network of
// diffusion_models/Improved_Diffusion_Unconditional/improved_diffusion/unet.py:278-497 (ResBlock :107-197 with scale-shift
// GroupNorm, AttentionBlock / QKVAttention :200-253, Downsample / Upsample :50-104, timestep embedding nn.py:103-121) on NHWC
// fp32 activations.
//   * every convolution (3x3, stride-2 3x3, 1x1 qkv / proj / skip) is the implicit GEMM of ap_conv_layer.cuh with bias and the
//     block's residual add fused into its epilogue;
//   * GroupNorm(32) + [scale-shift] + SiLU is one kernel per use: gn_tile_kernel keeps a (sample, 8-32 channel) tile in shared
//     memory (one HBM read, one write, two-pass statistics in fp32); gn_kernel is the strided fallback for odd channel counts;
//   * the timestep path is batch-constant in the purifier (every row is at the same discrete step), so the embedding MLP and
//     all 30 per-block projections are one packed GEMV per network evaluation;
//   * attention over T = H W <= 256 positions with 64-channel heads: unet_attn_mma_kernel (tf32 mode: mma.sync tensor cores, a warp
//     per 16 queries, online softmax) or unet_attn_kernel (fp32 parity mode: one thread per query);
//   * ap_unet_eps_vjp: the input gradient -- the forward recorded on a tape, then walked in reverse (gn_bwd_tile_kernel,
//     unet_attn_mma_bwd_kernel, data-gradient twins of every convolution, up-sampling / concatenation / fan-out kernels).
// The module walk (which block follows which, channel counts, skip-connection stack) is built by the host from the same
// flat op list the Python side derives from UNetModel.__init__ (synthetic.unet_structure), so names, order and shapes agree
// with the reference's state dict by construction.
#include <cmath>
#include <cstdlib>
#include <memory>
#include <unordered_map>
#include <vector>

#include "ap_common.cuh"
#include "ap_conv_layer.cuh"
#include "ap_internal.h"

namespace ap {

__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.f + __expf(-v)); }   // ex2 + rcp on the MUFU (2 ulp)

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

// GroupNorm with the (sample, group-chunk) tile resident in shared memory: one HBM read, one write.  A CTA owns `gpc` consecutive
// groups = Cc = gpc * cpg contiguous channels (>= 32 B per pixel, whole sectors); thread = (pixel row p0, float4 column j) with
// blockDim.x a multiple of q = Cc / 4, so a thread's group is fixed.  Sums are reduced in a fixed order (no atomics): results are
// reproducible run to run.  Same two-pass statistics and epilogue as gn_kernel.
struct GnTile {
  int q, j, p0, pstep, gl;
  __device__ GnTile(int Cc, int cpg) {
    q = Cc >> 2, j = threadIdx.x % q, p0 = threadIdx.x / q, pstep = blockDim.x / q, gl = (4 * j) / cpg;
  }
  // per-group sums of v over the CTA: part [blockDim.x] scratch, res [8]
  __device__ void reduce(float v, float* part, float* res, int cpg, int gpc) const {
    if ((q & (q - 1)) == 0 && q <= 32) {
      // q a power of two: lanes with the same column sit q apart -- butterfly inside the warp, then one pass over the warps
      for (int o = q; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
      if (lane < q) part[(threadIdx.x >> 5) * q + lane] = v;
      __syncthreads();
      if (threadIdx.x < gpc) {
        float t = 0.f;
        for (int w = 0; w < nw; ++w)
          for (int jj = threadIdx.x * (cpg >> 2); jj < (threadIdx.x + 1) * (cpg >> 2); ++jj) t += part[w * q + jj];
        res[threadIdx.x] = t;
      }
      __syncthreads();
      return;
    }
    part[threadIdx.x] = v;
    __syncthreads();
    for (int rows = pstep; rows > 1;) {
      const int half = (rows + 1) >> 1;
      if (p0 + half < rows) part[threadIdx.x] += part[threadIdx.x + half * q];
      rows = half;
      __syncthreads();
    }
    if (threadIdx.x < gpc) {
      float t = 0.f;
      for (int jj = threadIdx.x * (cpg >> 2); jj < (threadIdx.x + 1) * (cpg >> 2); ++jj) t += part[jj];
      res[threadIdx.x] = t;
    }
    __syncthreads();
  }
};
__device__ __forceinline__ float4 gn_affine(float4 xh, int ch, const float* __restrict__ gamma, const float* __restrict__ beta,
                                            const float* __restrict__ ss, int C, float4* mul_out) {
  const float4 ga = *reinterpret_cast<const float4*>(gamma + ch), be = *reinterpret_cast<const float4*>(beta + ch);
  float4 z = make_float4(xh.x * ga.x + be.x, xh.y * ga.y + be.y, xh.z * ga.z + be.z, xh.w * ga.w + be.w);
  float4 mul = ga;
  if (ss) {
    const float4 sc = *reinterpret_cast<const float4*>(ss + ch), sh = *reinterpret_cast<const float4*>(ss + C + ch);
    z = make_float4(z.x * (1.f + sc.x) + sh.x, z.y * (1.f + sc.y) + sh.y, z.z * (1.f + sc.z) + sh.z, z.w * (1.f + sc.w) + sh.w);
    mul = make_float4(mul.x * (1.f + sc.x), mul.y * (1.f + sc.y), mul.z * (1.f + sc.z), mul.w * (1.f + sc.w));
  }
  if (mul_out) *mul_out = mul;
  return z;
}
__device__ __forceinline__ float round_tf32_f(float y) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(y));
  return __uint_as_float(r);
}
__global__ void __launch_bounds__(256) gn_tile_kernel(const float* __restrict__ x, float* __restrict__ out, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, const float* __restrict__ scale_shift, int HW, int C,
                                                      int cpg, int gpc, int act, int round_tf32) {
  extern __shared__ float4 gn_sm[];                 // tile [HW][q] | part [blockDim.x] floats
  __shared__ float res[8];
  const int chunks = 32 / gpc, b = blockIdx.x / chunks, Cc = gpc * cpg, c0 = (blockIdx.x % chunks) * Cc;
  const GnTile tl(Cc, cpg);
  float* part = reinterpret_cast<float*>(gn_sm + static_cast<size_t>(HW) * tl.q);
  const float4* xb = reinterpret_cast<const float4*>(x + static_cast<size_t>(b) * HW * C + c0);
  float4* ob = reinterpret_cast<float4*>(out + static_cast<size_t>(b) * HW * C + c0);
  const int C4 = C >> 2, n = HW * cpg;
  float s = 0.f;
#pragma unroll 8
  for (int p = tl.p0; p < HW; p += tl.pstep) {
    const float4 v = xb[static_cast<size_t>(p) * C4 + tl.j];
    gn_sm[p * tl.q + tl.j] = v;
    s += (v.x + v.y) + (v.z + v.w);
  }
  tl.reduce(s, part, res, cpg, gpc);
  const float mean = res[tl.gl] / n;
  float vv = 0.f;
  for (int p = tl.p0; p < HW; p += tl.pstep) {
    const float4 v = gn_sm[p * tl.q + tl.j];
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    vv += fmaf(d0, d0, d1 * d1) + fmaf(d2, d2, d3 * d3);
  }
  __syncthreads();                                  // every thread has read res (mean) before it is overwritten
  tl.reduce(vv, part, res, cpg, gpc);
  const float rstd = rsqrtf(res[tl.gl] / n + 1e-5f);
  const int ch = c0 + 4 * tl.j;
  // the thread's four channels are fixed: y = (x - mean) A + Bc with A = rstd gamma [(1 + scale)], Bc = beta [(1 + scale) + shift]
  float4 A, Bc = gn_affine(make_float4(0.f, 0.f, 0.f, 0.f), ch, gamma, beta, scale_shift, C, &A);
  A = make_float4(A.x * rstd, A.y * rstd, A.z * rstd, A.w * rstd);
#pragma unroll 4
  for (int p = tl.p0; p < HW; p += tl.pstep) {
    const float4 v = gn_sm[p * tl.q + tl.j];
    float4 y = make_float4(fmaf(v.x - mean, A.x, Bc.x), fmaf(v.y - mean, A.y, Bc.y), fmaf(v.z - mean, A.z, Bc.z), fmaf(v.w - mean, A.w, Bc.w));
    if (act) y = make_float4(silu_f(y.x), silu_f(y.y), silu_f(y.z), silu_f(y.w));
    if (round_tf32) y = make_float4(round_tf32_f(y.x), round_tf32_f(y.y), round_tf32_f(y.z), round_tf32_f(y.w));
    ob[static_cast<size_t>(p) * C4 + tl.j] = y;
  }
}
// The same tile held in REGISTERS: NITER float4 per thread (256 threads: tiles of 4-32 KB), so the two statistics passes and the
// normalisation cost no shared-memory traffic at all and the NITER loads are in flight together.  Same thread <-> element mapping and
// the same reduction order as gn_tile_kernel: bitwise the same result.
template <int NITER>
__global__ void __launch_bounds__(256) gn_reg_kernel(const float* __restrict__ x, float* __restrict__ out, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, const float* __restrict__ scale_shift, int HW, int C,
                                                     int cpg, int gpc, int act, int round_tf32) {
  __shared__ float part[256];
  __shared__ float res[8];
  const int chunks = 32 / gpc, b = blockIdx.x / chunks, Cc = gpc * cpg, c0 = (blockIdx.x % chunks) * Cc;
  const GnTile tl(Cc, cpg);
  const float4* xb = reinterpret_cast<const float4*>(x + static_cast<size_t>(b) * HW * C + c0);
  float4* ob = reinterpret_cast<float4*>(out + static_cast<size_t>(b) * HW * C + c0);
  const int C4 = C >> 2, n = HW * cpg;
  float4 v[NITER];
#pragma unroll
  for (int k = 0; k < NITER; ++k) v[k] = xb[static_cast<size_t>(tl.p0 + k * tl.pstep) * C4 + tl.j];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NITER; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  tl.reduce(s, part, res, cpg, gpc);
  const float mean = res[tl.gl] / n;
  float vv = 0.f;
#pragma unroll
  for (int k = 0; k < NITER; ++k) {
    const float d0 = v[k].x - mean, d1 = v[k].y - mean, d2 = v[k].z - mean, d3 = v[k].w - mean;
    vv += fmaf(d0, d0, d1 * d1) + fmaf(d2, d2, d3 * d3);
  }
  __syncthreads();
  tl.reduce(vv, part, res, cpg, gpc);
  const float rstd = rsqrtf(res[tl.gl] / n + 1e-5f);
  const int ch = c0 + 4 * tl.j;
  float4 A, Bc = gn_affine(make_float4(0.f, 0.f, 0.f, 0.f), ch, gamma, beta, scale_shift, C, &A);
  A = make_float4(A.x * rstd, A.y * rstd, A.z * rstd, A.w * rstd);
#pragma unroll
  for (int k = 0; k < NITER; ++k) {
    float4 y = make_float4(fmaf(v[k].x - mean, A.x, Bc.x), fmaf(v[k].y - mean, A.y, Bc.y), fmaf(v[k].z - mean, A.z, Bc.z),
                           fmaf(v[k].w - mean, A.w, Bc.w));
    if (act) y = make_float4(silu_f(y.x), silu_f(y.y), silu_f(y.z), silu_f(y.w));
    if (round_tf32) y = make_float4(round_tf32_f(y.x), round_tf32_f(y.y), round_tf32_f(y.z), round_tf32_f(y.w));
    ob[static_cast<size_t>(tl.p0 + k * tl.pstep) * C4 + tl.j] = y;
  }
}
// backward of the same: x-hat and d x-hat tiles in shared memory, g_y read once.
__global__ void __launch_bounds__(256) gn_bwd_tile_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ scale_shift, int HW, int C, int cpg, int gpc, int act,
                                                          int accumulate) {
  extern __shared__ float4 gn_sm[];                 // x-hat [HW][q] | d x-hat [HW][q] | part [blockDim.x]
  __shared__ float res[8], res2[8];
  const int chunks = 32 / gpc, b = blockIdx.x / chunks, Cc = gpc * cpg, c0 = (blockIdx.x % chunks) * Cc;
  const GnTile tl(Cc, cpg);
  float4* xs = gn_sm;
  float4* ds = gn_sm + static_cast<size_t>(HW) * tl.q;
  float* part = reinterpret_cast<float*>(ds + static_cast<size_t>(HW) * tl.q);
  const size_t off = static_cast<size_t>(b) * HW * C + c0;
  const float4* xb = reinterpret_cast<const float4*>(x + off);
  const float4* gb = reinterpret_cast<const float4*>(gy + off);
  float4* ob = reinterpret_cast<float4*>(gx + off);
  const int C4 = C >> 2, n = HW * cpg;
  float s = 0.f;
#pragma unroll 8
  for (int p = tl.p0; p < HW; p += tl.pstep) {
    const float4 v = xb[static_cast<size_t>(p) * C4 + tl.j];
    xs[p * tl.q + tl.j] = v;
    s += (v.x + v.y) + (v.z + v.w);
  }
  tl.reduce(s, part, res, cpg, gpc);
  const float mean = res[tl.gl] / n;
  float vv = 0.f;
  for (int p = tl.p0; p < HW; p += tl.pstep) {
    const float4 v = xs[p * tl.q + tl.j];
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    vv += fmaf(d0, d0, d1 * d1) + fmaf(d2, d2, d3 * d3);
  }
  __syncthreads();
  tl.reduce(vv, part, res, cpg, gpc);
  const float rstd = rsqrtf(res[tl.gl] / n + 1e-5f);
  const int ch = c0 + 4 * tl.j;
  float s1 = 0.f, s2 = 0.f;
  float4 mul;                                        // the thread's four channels are fixed: z = x-hat mul + Bc
  const float4 Bc = gn_affine(make_float4(0.f, 0.f, 0.f, 0.f), ch, gamma, beta, scale_shift, C, &mul);
#pragma unroll 4
  for (int p = tl.p0; p < HW; p += tl.pstep) {
    const float4 v = xs[p * tl.q + tl.j];
    const float4 xh = make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
    const float4 z = make_float4(fmaf(xh.x, mul.x, Bc.x), fmaf(xh.y, mul.y, Bc.y), fmaf(xh.z, mul.z, Bc.z), fmaf(xh.w, mul.w, Bc.w));
    float4 d = gb[static_cast<size_t>(p) * C4 + tl.j];
    if (act) {
      auto dsilu = [](float zz) {
        const float sg = __fdividef(1.f, 1.f + __expf(-zz));
        return sg * (1.f + zz * (1.f - sg));
      };
      d = make_float4(d.x * dsilu(z.x), d.y * dsilu(z.y), d.z * dsilu(z.z), d.w * dsilu(z.w));
    }
    d = make_float4(d.x * mul.x, d.y * mul.y, d.z * mul.z, d.w * mul.w);
    xs[p * tl.q + tl.j] = xh;
    ds[p * tl.q + tl.j] = d;
    s1 += (d.x + d.y) + (d.z + d.w);
    s2 += fmaf(d.x, xh.x, d.y * xh.y) + fmaf(d.z, xh.z, d.w * xh.w);
  }
  __syncthreads();
  tl.reduce(s1, part, res, cpg, gpc);
  tl.reduce(s2, part, res2, cpg, gpc);
  const float m1 = res[tl.gl] / n, m2 = res2[tl.gl] / n;
  for (int p = tl.p0; p < HW; p += tl.pstep) {
    const float4 xh = xs[p * tl.q + tl.j], d = ds[p * tl.q + tl.j];
    float4 r = make_float4(rstd * (d.x - m1 - xh.x * m2), rstd * (d.y - m1 - xh.y * m2), rstd * (d.z - m1 - xh.z * m2),
                           rstd * (d.w - m1 - xh.w * m2));
    if (accumulate) {
      const float4 o = ob[static_cast<size_t>(p) * C4 + tl.j];
      r = make_float4(r.x + o.x, r.y + o.y, r.z + o.z, r.w + o.w);
    }
    ob[static_cast<size_t>(p) * C4 + tl.j] = r;
  }
}

// groups per CTA for a tile budget: the largest power of two (<= 8) whose `tiles` tiles fit `budget` bytes; 0 = not even one group
static int gn_groups_per_cta(int HW, int cpg, int tiles, size_t budget) {
  if (cpg % 4 != 0) return 0;
  for (int gpc = 8; gpc >= 1; gpc >>= 1)
    if (static_cast<size_t>(HW) * gpc * cpg * sizeof(float) * tiles <= budget) return gpc;
  return 0;
}

// ------------------------------------------------------------------------------------------------ attention on the tensor cores
// (tf32 mode.)  mma.sync m16n8k8 tf32, fp32 accumulation; a warp owns 16 rows (queries, or keys in the key-major backward phase) and
// walks the other dimension in chunks of NB * 8 columns: C = A . M^T ("nt": scores) and C += P . M ("nn": P in the accumulator
// layout is fed back as the A operand through a fixed permutation of the k slots -- slot t <-> column 2t, slot t + 4 <-> column
// 2t + 1 -- applied to the rows of M instead of shuffling registers).  K / V (or Q / g_o) of the head sit in shared memory as tf32
// with a row stride of 68 words: both access patterns are bank-conflict free.
constexpr int ATT_LD = 68;
constexpr int ATT_KT = 128;      // rows of the two shared-memory operand tiles resident at a time (70 KB: three CTAs per SM)
__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float ex2_approx(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// A fragments of rows r0 .. r0 + 15, 64 channels, of a row-major fp32 matrix (row stride `stride` floats), times `scale`
__device__ __forceinline__ void att_load_a(uint32_t (&a)[8][4], const float* __restrict__ base, size_t stride, int r0, int lane, float scale) {
  const int g = lane >> 2, t = lane & 3;
  const float* lo = base + static_cast<size_t>(r0 + g) * stride;
  const float* hi = lo + 8 * stride;
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    a[ks][0] = to_tf32(lo[8 * ks + t] * scale), a[ks][1] = to_tf32(hi[8 * ks + t] * scale);
    a[ks][2] = to_tf32(lo[8 * ks + t + 4] * scale), a[ks][3] = to_tf32(hi[8 * ks + t + 4] * scale);
  }
}
template <int NB> __device__ __forceinline__ void att_mma_nt(float (&c)[NB][4], const uint32_t (&a)[8][4], const uint32_t* __restrict__ M,
                                                             int col0, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const uint32_t* row = M + (col0 + g) * ATT_LD + t;
  // k step outermost: consecutive MMAs go to NB different accumulators (a chain on one accumulator would serialise on the MMA latency)
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) mma_tf32(c[nb], a[ks], row[nb * 8 * ATT_LD + 8 * ks], row[nb * 8 * ATT_LD + 8 * ks + 4]);
}
template <int NB> __device__ __forceinline__ void att_mma_nn(float (&c)[8][4], const float (&p)[NB][4], const uint32_t* __restrict__ M,
                                                             int k0, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kb = 0; kb < NB; ++kb) {
    const uint32_t a[4] = {to_tf32(p[kb][0]), to_tf32(p[kb][2]), to_tf32(p[kb][1]), to_tf32(p[kb][3])};
    const uint32_t* r0 = M + (k0 + kb * 8 + 2 * t) * ATT_LD + g;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) mma_tf32(c[nb], a, r0[nb * 8], r0[ATT_LD + nb * 8]);
  }
}
// rows [0, T) x 64 channels of a row-major fp32 matrix -> shared memory as tf32 (stride ATT_LD)
__device__ __forceinline__ void att_stage(uint32_t* __restrict__ dst, const float* __restrict__ src, size_t stride, int T) {
  for (int i = threadIdx.x; i < T * 16; i += blockDim.x) {
    const int s = i >> 4, c4 = (i & 15) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + static_cast<size_t>(s) * stride + c4);
    *reinterpret_cast<uint4*>(dst + s * ATT_LD + c4) = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
  }
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
constexpr float kLog2e = 1.4426950408889634f;

// forward: grid (B heads, ceil(T / 128)), one warp per 16 queries.  lse (optional, [B heads][T]): log2-domain log-sum-exp of the
// scaled scores, kept for the backward pass.
__global__ void __launch_bounds__(256, 2) unet_attn_mma_kernel(const float* __restrict__ qkv, float* __restrict__ out, float* __restrict__ lse,
                                                            int T, int C, int heads) {
  extern __shared__ uint32_t att_sm[];              // K [KT][68] | V [KT][68], KT = min(T, ATT_KT) keys at a time
  constexpr int NB = 4;                             // 32 keys per softmax step: 128 registers, two CTAs per SM
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads, KT = T < ATT_KT ? T : ATT_KT;
  const size_t qs = 3 * static_cast<size_t>(C);
  const float* base = qkv + static_cast<size_t>(b) * T * qs + h * 192;
  uint32_t* Ks = att_sm;
  uint32_t* Vs = att_sm + KT * ATT_LD;
  const int lane = threadIdx.x & 31, r0 = blockIdx.y * 128 + (threadIdx.x >> 5) * 16;
  const bool active = r0 < T;                       // (warp-uniform; inactive warps only help staging)
  const int g = lane >> 2, t = lane & 3;
  uint32_t qa[8][4];
  att_load_a(qa, base, qs, active ? r0 : 0, lane, 0.125f * kLog2e);                 // 1 / sqrt(64): q and k each carry 64^-1/4
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f, o[8][4];
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f;
  for (int kt = 0; kt < T; kt += KT) {
   if (kt) __syncthreads();                         // every warp is done with the previous tile
   att_stage(Ks, base + 64 + kt * qs, qs, KT);
   att_stage(Vs, base + 128 + kt * qs, qs, KT);
   __syncthreads();
   if (active)
   for (int k0 = 0; k0 < KT; k0 += NB * 8) {
    float s[NB][4];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
    att_mma_nt<NB>(s, qa, Ks, k0, lane);
    float x0 = -INFINITY, x1 = -INFINITY;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) x0 = fmaxf(x0, fmaxf(s[nb][0], s[nb][1])), x1 = fmaxf(x1, fmaxf(s[nb][2], s[nb][3]));
    const float n0 = fmaxf(m0, quad_max(x0)), n1 = fmaxf(m1, quad_max(x1));
    const float c0 = ex2_approx(m0 - n0), c1 = ex2_approx(m1 - n1);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      s[nb][0] = ex2_approx(s[nb][0] - n0), s[nb][1] = ex2_approx(s[nb][1] - n0);
      s[nb][2] = ex2_approx(s[nb][2] - n1), s[nb][3] = ex2_approx(s[nb][3] - n1);
      a0 += s[nb][0] + s[nb][1], a1 += s[nb][2] + s[nb][3];
    }
    l0 = l0 * c0 + a0, l1 = l1 * c1 + a1, m0 = n0, m1 = n1;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) o[nb][0] *= c0, o[nb][1] *= c0, o[nb][2] *= c1, o[nb][3] *= c1;
    att_mma_nn<NB>(o, s, Vs, k0, lane);
   }
  }
  if (!active) return;
  l0 = quad_sum(l0), l1 = quad_sum(l1);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  float* olo = out + (static_cast<size_t>(b) * T + r0 + g) * C + h * 64 + 2 * t;
  float* ohi = olo + 8 * static_cast<size_t>(C);
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    *reinterpret_cast<float2*>(olo + nb * 8) = make_float2(o[nb][0] * i0, o[nb][1] * i0);
    *reinterpret_cast<float2*>(ohi + nb * 8) = make_float2(o[nb][2] * i1, o[nb][3] * i1);
  }
  if (lse && t == 0) {
    lse[static_cast<size_t>(bh) * T + r0 + g] = m0 + log2f(l0);
    lse[static_cast<size_t>(bh) * T + r0 + g + 8] = m1 + log2f(l1);
  }
}

// backward: grid (B heads, 2 ceil(T / 128)).  blockIdx.y < nblk: query-major phase -- P = 2^(S - lse), g_P = g_o V^T,
// g_S = P (g_P - D), D = rowsum(g_o o); g_q = g_S K / 8.  Otherwise key-major phase on the transposed problem: g_v = P^T g_o,
// g_k = g_S^T Q / 8.  No atomics, no T x T scratch: each phase recomputes the scores it needs from the stored log-sum-exp.
__global__ void __launch_bounds__(256) unet_attn_mma_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ o_fwd,
                                                                const float* __restrict__ lse, const float* __restrict__ g_out,
                                                                float* __restrict__ g_qkv, int T, int C, int heads) {
  extern __shared__ uint32_t att_sm[];              // two [KT][68] operand tiles | lse [T] | D [T]
  constexpr int NB = 4;
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads, nblk = gridDim.y >> 1, KT = T < ATT_KT ? T : ATT_KT;
  const bool key_major = static_cast<int>(blockIdx.y) >= nblk;
  const size_t qs = 3 * static_cast<size_t>(C);
  const float* base = qkv + static_cast<size_t>(b) * T * qs + h * 192;
  const float* gob = g_out + static_cast<size_t>(b) * T * C + h * 64;
  const float* ofb = o_fwd + static_cast<size_t>(b) * T * C + h * 64;
  float* gq = g_qkv + static_cast<size_t>(b) * T * qs + h * 192;
  uint32_t* M0 = att_sm;
  uint32_t* M1 = att_sm + KT * ATT_LD;
  float* Ls = reinterpret_cast<float*>(att_sm + 2 * KT * ATT_LD);
  float* Ds = Ls + T;
  auto stage_tile = [&](int r) {                    // rows r .. r + KT of (K, V) or of (Q, g_o)
    if (!key_major) {
      att_stage(M0, base + 64 + r * qs, qs, KT);
      att_stage(M1, base + 128 + r * qs, qs, KT);
    } else {
      att_stage(M0, base + r * qs, qs, KT);
      att_stage(M1, gob + static_cast<size_t>(r) * C, C, KT);
    }
  };
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    const float4* a = reinterpret_cast<const float4*>(gob + static_cast<size_t>(i) * C);
    const float4* c = reinterpret_cast<const float4*>(ofb + static_cast<size_t>(i) * C);
    float d = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 u = a[k], v = c[k];
      d += fmaf(u.x, v.x, u.y * v.y) + fmaf(u.z, v.z, u.w * v.w);
    }
    Ds[i] = d;
    Ls[i] = lse[static_cast<size_t>(bh) * T + i];
  }
  const int lane = threadIdx.x & 31, r0a = (blockIdx.y % nblk) * 128 + (threadIdx.x >> 5) * 16;
  const bool active = r0a < T;
  const int r0 = active ? r0a : 0, g = lane >> 2, t = lane & 3;
  if (!key_major) {
    uint32_t qa[8][4], da[8][4];
    att_load_a(qa, base, qs, r0, lane, 0.125f * kLog2e);
    att_load_a(da, gob, C, r0, lane, 1.f);
    float L0 = 0.f, L1 = 0.f, D0 = 0.f, D1 = 0.f;
    float dq[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) dq[nb][0] = dq[nb][1] = dq[nb][2] = dq[nb][3] = 0.f;
    for (int kt = 0; kt < T; kt += KT) {
     if (kt) __syncthreads();
     stage_tile(kt);
     __syncthreads();                               // also orders the Ls / Ds writes above before the reads below
     if (kt == 0) L0 = Ls[r0 + g], L1 = Ls[r0 + g + 8], D0 = Ds[r0 + g], D1 = Ds[r0 + g + 8];
     if (active)
     for (int k0 = 0; k0 < KT; k0 += NB * 8) {
      float s[NB][4], dp[NB][4];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = dp[nb][0] = dp[nb][1] = dp[nb][2] = dp[nb][3] = 0.f;
      att_mma_nt<NB>(s, qa, M0, k0, lane);
      att_mma_nt<NB>(dp, da, M1, k0, lane);
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        s[nb][0] = ex2_approx(s[nb][0] - L0) * (dp[nb][0] - D0), s[nb][1] = ex2_approx(s[nb][1] - L0) * (dp[nb][1] - D0);
        s[nb][2] = ex2_approx(s[nb][2] - L1) * (dp[nb][2] - D1), s[nb][3] = ex2_approx(s[nb][3] - L1) * (dp[nb][3] - D1);
      }
      att_mma_nn<NB>(dq, s, M0, k0, lane);
     }
    }
    if (!active) return;
    float* lo = gq + static_cast<size_t>(r0 + g) * qs + 2 * t;
    float* hi = lo + 8 * qs;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      *reinterpret_cast<float2*>(lo + nb * 8) = make_float2(dq[nb][0] * 0.125f, dq[nb][1] * 0.125f);
      *reinterpret_cast<float2*>(hi + nb * 8) = make_float2(dq[nb][2] * 0.125f, dq[nb][3] * 0.125f);
    }
  } else {
    uint32_t ka[8][4], va[8][4];
    att_load_a(ka, base + 64, qs, r0, lane, 0.125f * kLog2e);
    att_load_a(va, base + 128, qs, r0, lane, 1.f);
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) dk[nb][0] = dk[nb][1] = dk[nb][2] = dk[nb][3] = dv[nb][0] = dv[nb][1] = dv[nb][2] = dv[nb][3] = 0.f;
    for (int qt = 0; qt < T; qt += KT) {
     if (qt) __syncthreads();
     stage_tile(qt);
     __syncthreads();
     if (active)
     for (int q0 = 0; q0 < KT; q0 += NB * 8) {
      float s[NB][4], dp[NB][4];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = dp[nb][0] = dp[nb][1] = dp[nb][2] = dp[nb][3] = 0.f;
      att_mma_nt<NB>(s, ka, M0, q0, lane);          // S^T[key][query]
      att_mma_nt<NB>(dp, va, M1, q0, lane);         // g_P^T[key][query]
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const int qc = qt + q0 + nb * 8 + 2 * t;
        const float La = Ls[qc], Lb = Ls[qc + 1], Da = Ds[qc], Db = Ds[qc + 1];
        s[nb][0] = ex2_approx(s[nb][0] - La), s[nb][1] = ex2_approx(s[nb][1] - Lb);
        s[nb][2] = ex2_approx(s[nb][2] - La), s[nb][3] = ex2_approx(s[nb][3] - Lb);
        dp[nb][0] = s[nb][0] * (dp[nb][0] - Da), dp[nb][1] = s[nb][1] * (dp[nb][1] - Db);
        dp[nb][2] = s[nb][2] * (dp[nb][2] - Da), dp[nb][3] = s[nb][3] * (dp[nb][3] - Db);
      }
      att_mma_nn<NB>(dv, s, M1, q0, lane);
      att_mma_nn<NB>(dk, dp, M0, q0, lane);
     }
    }
    if (!active) return;
    float* lo = gq + static_cast<size_t>(r0 + g) * qs + 64 + 2 * t;
    float* hi = lo + 8 * qs;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      *reinterpret_cast<float2*>(lo + nb * 8) = make_float2(dk[nb][0] * 0.125f, dk[nb][1] * 0.125f);
      *reinterpret_cast<float2*>(hi + nb * 8) = make_float2(dk[nb][2] * 0.125f, dk[nb][3] * 0.125f);
      *reinterpret_cast<float2*>(lo + 64 + nb * 8) = make_float2(dv[nb][0], dv[nb][1]);
      *reinterpret_cast<float2*>(hi + 64 + nb * 8) = make_float2(dv[nb][2], dv[nb][3]);
    }
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

// ------------------------------------------------------------------------------------------------ backward kernels (VJP wrt x)
// The reference back-propagates through this UNet (improved_diffusion_sde.py:104-105 calls the model without no_grad), so a
// white-box attack on Diffusion-Spec needs g_x = (d eps / d x)^T g_eps.  Chain rule over the recorded forward (activations are
// still in the arena): convolution data-gradient twins (transposed, 180-degree-rotated weights), and the kernels below.

// backward of gn_kernel: y = act(z), z = ((x - mean) rstd gamma + beta) [(1 + scale) + shift]; statistics recomputed from x.
//   dxhat = g_y act'(z) [(1 + scale)] gamma ;  g_x = rstd (dxhat - mean(dxhat) - xhat mean(dxhat xhat))   over the (sample, group)
__global__ void __launch_bounds__(256) gn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     const float* __restrict__ scale_shift, int HW, int C, int cpg, int act, int accumulate) {
  const int b = blockIdx.x / 32, g = blockIdx.x % 32;
  const size_t off = static_cast<size_t>(b) * HW * C + g * cpg;
  const float* xb = x + off;
  const float* gb = gy + off;
  float* ob = gx + off;
  const int n = HW * cpg;
  __shared__ float red[2][8];
  __shared__ float s_a, s_b;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  auto block_sum2 = [&](float v0, float v1, float& o0, float& o1) {
    for (int o = 16; o; o >>= 1) v0 += __shfl_xor_sync(0xffffffffu, v0, o), v1 += __shfl_xor_sync(0xffffffffu, v1, o);
    __syncthreads();
    if (lane == 0) red[0][wid] = v0, red[1][wid] = v1;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t0 = 0.f, t1 = 0.f;
      for (int i = 0; i < 8; ++i) t0 += red[0][i], t1 += red[1][i];
      s_a = t0, s_b = t1;
    }
    __syncthreads();
    o0 = s_a, o1 = s_b;
  };
  float s = 0.f, dummy;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += xb[static_cast<size_t>(i / cpg) * C + (i % cpg)];
  float mean;
  block_sum2(s, 0.f, mean, dummy);
  mean /= n;
  float v = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = xb[static_cast<size_t>(i / cpg) * C + (i % cpg)] - mean;
    v = fmaf(d, d, v);
  }
  float var;
  block_sum2(v, 0.f, var, dummy);
  const float rstd = rsqrtf(var / n + 1e-5f);
  auto dxhat_of = [&](int i, float& xhat) -> float {
    const int c = i % cpg, ch = g * cpg + c;
    const size_t idx = static_cast<size_t>(i / cpg) * C + c;
    xhat = (xb[idx] - mean) * rstd;
    float z = xhat * gamma[ch] + beta[ch];
    float dz = gb[idx];
    float mul = gamma[ch];
    if (scale_shift) {
      z = z * (1.f + scale_shift[ch]) + scale_shift[C + ch];
      mul *= 1.f + scale_shift[ch];
    }
    if (act) {
      const float sg = 1.f / (1.f + __expf(-z));
      dz *= sg * (1.f + z * (1.f - sg));                    // d silu(z) / dz
    }
    return dz * mul;
  };
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float xhat;
    const float d = dxhat_of(i, xhat);
    s1 += d, s2 = fmaf(d, xhat, s2);
  }
  float m1, m2;
  block_sum2(s1, s2, m1, m2);
  m1 /= n, m2 /= n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float xhat;
    const float d = dxhat_of(i, xhat);
    const size_t idx = static_cast<size_t>(i / cpg) * C + (i % cpg);
    const float r = rstd * (d - m1 - xhat * m2);
    ob[idx] = accumulate ? ob[idx] + r : r;
  }
}

// backward of unet_attn_kernel.  One CTA per (sample, head), one thread per position.  Phase A (thread = query t): softmax row,
// D_t = sum_s P g_P, g_S = P (g_P - D_t) -> g_q; P and g_S of the head go to global scratch [T][T].  Phase B (thread = key s):
// g_k[s] = sc^2 sum_t g_S[t][s] q_t, g_v[s] = sum_t P[t][s] g_o[t].
template <int CH> __global__ void __launch_bounds__(256) unet_attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ g_out,
                                                                             float* __restrict__ g_qkv, float* __restrict__ Pm,
                                                                             float* __restrict__ Gs, int T, int C, int heads) {
  extern __shared__ float kv[];
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const float* base = qkv + static_cast<size_t>(b) * T * 3 * C + h * 3 * CH;
  const float* gob = g_out + static_cast<size_t>(b) * T * C + h * CH;
  float* gq = g_qkv + static_cast<size_t>(b) * T * 3 * C + h * 3 * CH;
  float* P = Pm + static_cast<size_t>(blockIdx.x) * T * T;
  float* G = Gs + static_cast<size_t>(blockIdx.x) * T * T;
  float* ks = kv;
  float* vs = kv + static_cast<size_t>(T) * CH;
  for (int i = threadIdx.x; i < T * CH; i += blockDim.x) {
    const int s = i / CH, c = i - s * CH;
    ks[i] = base[static_cast<size_t>(s) * 3 * C + CH + c];
    vs[i] = base[static_cast<size_t>(s) * 3 * C + 2 * CH + c];
  }
  __syncthreads();
  const float sc = rsqrtf(sqrtf(static_cast<float>(CH))), sc2 = sc * sc;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float q[CH], go[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) q[c] = base[static_cast<size_t>(t) * 3 * C + c], go[c] = gob[static_cast<size_t>(t) * C + c];
    float m = -INFINITY, l = 0.f;
    for (int s = 0; s < T; ++s) {
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) d = fmaf(q[c], ks[s * CH + c], d);
      d *= sc2;
      const float mn = fmaxf(m, d);
      l = l * __expf(m - mn) + __expf(d - mn);
      m = mn;
    }
    const float inv = 1.f / l;
    float D = 0.f;
    for (int s = 0; s < T; ++s) {
      float d = 0.f, gp = 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) d = fmaf(q[c], ks[s * CH + c], d), gp = fmaf(go[c], vs[s * CH + c], gp);
      const float pw = __expf(d * sc2 - m) * inv;
      P[static_cast<size_t>(t) * T + s] = pw;
      G[static_cast<size_t>(t) * T + s] = gp;
      D = fmaf(pw, gp, D);
    }
    float acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = 0.f;
    for (int s = 0; s < T; ++s) {
      const float gs = P[static_cast<size_t>(t) * T + s] * (G[static_cast<size_t>(t) * T + s] - D);
      G[static_cast<size_t>(t) * T + s] = gs;
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = fmaf(gs, ks[s * CH + c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) gq[static_cast<size_t>(t) * 3 * C + c] = acc[c] * sc2;
  }
  __syncthreads();     // P and G of this head (written by this CTA) are complete
  for (int s = threadIdx.x; s < T; s += blockDim.x) {
    float gk[CH], gv[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) gk[c] = 0.f, gv[c] = 0.f;
    for (int t = 0; t < T; ++t) {
      const float gs = G[static_cast<size_t>(t) * T + s], pw = P[static_cast<size_t>(t) * T + s];
      const float* qt = base + static_cast<size_t>(t) * 3 * C;
      const float* got = gob + static_cast<size_t>(t) * C;
#pragma unroll
      for (int c = 0; c < CH; ++c) gk[c] = fmaf(gs, qt[c], gk[c]), gv[c] = fmaf(pw, got[c], gv[c]);
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      gq[static_cast<size_t>(s) * 3 * C + CH + c] = gk[c] * sc2;
      gq[static_cast<size_t>(s) * 3 * C + 2 * CH + c] = gv[c];
    }
  }
}

// backward of nearest_up2_kernel: g_in[b][i][j][c] (+)= sum of the 2 x 2 block of g_out
__global__ void __launch_bounds__(256) up2_bwd_kernel(const float4* __restrict__ g_out, float4* __restrict__ g_in, int B, int H, int W,
                                                      int C4, int accumulate) {
  const long long total = static_cast<long long>(B) * H * W * C4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C4);
    long long t = i / C4;
    const int w = static_cast<int>(t % W);
    t /= W;
    const int hh = static_cast<int>(t % H);
    const long long b = t / H;
    float4 a = accumulate ? g_in[i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const float4 v = g_out[((b * 2 * H + 2 * hh + (d >> 1)) * 2 * W + 2 * w + (d & 1)) * C4 + c];
      a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
    }
    g_in[i] = a;
  }
}
// backward of concat_kernel: g_a (+)= g_out[..., :C1], g_s (+)= g_out[..., C1:]
__global__ void __launch_bounds__(256) split_bwd_kernel(const float4* __restrict__ g_out, float4* __restrict__ g_a, float4* __restrict__ g_s,
                                                        long long pixels, int C1_4, int C2_4, int acc_a, int acc_s) {
  const int Ct = C1_4 + C2_4;
  const long long total = pixels * Ct;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Ct);
    const long long px = i / Ct;
    const float4 v = g_out[i];
    float4* dst = c < C1_4 ? g_a + px * C1_4 + c : g_s + px * C2_4 + (c - C1_4);
    if (c < C1_4 ? acc_a : acc_s) {
      float4 o = *dst;
      o.x += v.x, o.y += v.y, o.z += v.z, o.w += v.w;
      *dst = o;
    } else {
      *dst = v;
    }
  }
}
__global__ void __launch_bounds__(256) add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] += src[i];
}
// up[b][2i][2j][c] = g[b][i][j][c], zero elsewhere (the gradient of a stride-2 convolution before its stride-1 data-gradient twin)
__global__ void __launch_bounds__(256) zero_up2_kernel(const float4* __restrict__ g, float4* __restrict__ up, int B, int Ho, int Wo, int C4) {
  const long long total = static_cast<long long>(B) * (2 * Ho) * (2 * Wo) * C4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C4);
    long long t = i / C4;
    const int ow = static_cast<int>(t % (2 * Wo));
    t /= 2 * Wo;
    const int oh = static_cast<int>(t % (2 * Ho));
    const long long b = t / (2 * Ho);
    up[i] = ((oh | ow) & 1) ? make_float4(0.f, 0.f, 0.f, 0.f) : g[((b * Ho + (oh >> 1)) * Wo + (ow >> 1)) * C4 + c];
  }
}

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
  UOp() { c1.keep_host = c2.keep_host = skip.keep_host = true; }   // the backward pass builds data-gradient twins from the host copy
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
  // activation arena (bump allocator over `arena_floats` floats) -- every buffer of one sub-batch lives here
  DevBuf arena;
  size_t arena_floats = 0;
  // forward-only evaluations keep the block outputs (skip connections) and recycle one scratch region for a block's temporaries
  size_t need_fwd_persist = 0, need_fwd_tmp = 0, budget_floats = 0;
  size_t need_per_sample = 0;   // floats
  bool attn_attr = false;
  // backward pass: gradient arena, data-gradient twins of every convolution, scratch for the recomputed eps
  DevBuf garena, eps_scratch;
  size_t garena_floats = 0, need_grad_per_sample = 0;

  bool twins_ready = false;
  std::unordered_map<const ConvLayer*, std::unique_ptr<ConvLayer>> twins;
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
        case OP_ATTN:      // normalised input, qkv, attention output, block output, log-sum-exp rows kept for the backward pass
          need += px * (op->cin + 3 * static_cast<size_t>(op->cin) + 2 * static_cast<size_t>(op->cin)) + px * cfg->num_heads;
          break;
        case OP_POP: need += px * op->cout; break;
        case OP_DOWN: H /= 2; need += static_cast<size_t>(H) * H * op->cout; break;
        case OP_UP: need += 4 * px * op->cin + 4 * px * op->cout; H *= 2; break;
        case OP_OUT: need += px * (op->cin + op->cout); break;
        default: break;
      }
    }
    h->need_per_sample = need + 64;
    {
      size_t pers = 0, tmp = 0;
      int Hf = cfg->image_size;
      for (auto& op : h->ops) {
        const size_t px = static_cast<size_t>(Hf) * Hf;
        switch (op->kind) {
          case OP_CONV_IN: pers += px * op->cout; break;
          case OP_RES: pers += px * op->cout, tmp = std::max(tmp, px * (op->cin + 2 * static_cast<size_t>(op->cout) + (op->has_skip ? op->cout : 0))); break;
          case OP_ATTN: pers += px * op->cin, tmp = std::max(tmp, px * 5 * static_cast<size_t>(op->cin)); break;
          case OP_POP: pers += px * op->cout; break;
          case OP_DOWN: Hf /= 2; pers += static_cast<size_t>(Hf) * Hf * op->cout; break;
          case OP_UP: pers += 4 * px * op->cout, tmp = std::max(tmp, 4 * px * op->cin); Hf *= 2; break;
          case OP_OUT: tmp = std::max(tmp, px * op->cin); break;
          default: break;
        }
      }
      h->need_fwd_persist = pers + 32, h->need_fwd_tmp = tmp + 32;
      size_t free_b = 0, total_b = 0;
      AP_CUDA(cudaMemGetInfo(&free_b, &total_b));
      double cap_gb = 20.0;                           // both arenas together; AP_UNET_ARENA_GB overrides (tests use it to force sub-batches)
      if (const char* e = std::getenv("AP_UNET_ARENA_GB")) cap_gb = std::max(0.01, std::atof(e));
      h->budget_floats = std::min<size_t>(static_cast<size_t>(cap_gb * (1u << 30)), free_b / 2) / sizeof(float);
    }
    // gradients: one buffer per activation (<= the forward's arena), zero-upsampled gradients of the three stride-2 convolutions,
    // and the attention scratch (P and g_S: 2 x heads x T x T per attention block)
    size_t gneed = need;
    int Hh = cfg->image_size;
    for (auto& op : h->ops) {
      if (op->kind == OP_DOWN) gneed += static_cast<size_t>(Hh) * Hh * op->cout, Hh /= 2;
      else if (op->kind == OP_UP) Hh *= 2;
      else if (op->kind == OP_ATTN) gneed += static_cast<size_t>(2) * cfg->num_heads * Hh * Hh * Hh * Hh;
    }
    h->need_grad_per_sample = gneed + 64;
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

// ---- one recorded operation of a forward pass (the activations stay in the arena until the chunk is done)
namespace {
enum { T_CONV = 0, T_GN = 1, T_ATTN = 2, T_UP2 = 3, T_CONCAT = 4 };
struct TapeEntry {
  int kind = 0;
  const ConvLayer* L = nullptr;       // T_CONV
  const float* in = nullptr;
  float* out = nullptr;
  const float* res = nullptr;         // T_CONV: tensor added in the epilogue; T_CONCAT: the second input
  int Hin = 0, Cin = 0, Cout = 0;     // spatial size of `in`; channels of in / out (T_CONCAT: C1 = Cin, C2 = Cout - Cin)
  const DevBuf* gamma = nullptr;      // T_GN
  const DevBuf* beta = nullptr;
  const float* ss = nullptr;
  int act = 0;
  const float* lse = nullptr;         // T_ATTN on the tensor-core kernel: log-sum-exp rows [bn heads][T]
};
constexpr size_t kGnSmemMax = 208 * 1024, kAttSmemMax = 2 * ATT_KT * ATT_LD * 4 + 2 * 256 * 4;
// GroupNorm launch: the shared-memory-tile kernel when a tile fits, else the strided one-group-per-CTA kernel
int launch_gn(const float* x, float* out, const float* gamma, const float* beta, const float* ss, int bn, int HW, int C, int act,
              int round_tf32, cudaStream_t st) {
  const int cpg = C / 32;
  int gpc = gn_groups_per_cta(HW, cpg, 1, 32 * 1024);        // small tiles: 6 CTAs per SM keep HBM busy across the kernel's phases
  if (!gpc && gn_groups_per_cta(HW, cpg, 1, kGnSmemMax - 2048)) gpc = 1;   // one group per CTA if that is all that fits
  if (gpc) {
    const int q = gpc * cpg / 4, threads = (256 / q) * q;
    const int items = HW * q;                        // float4 per tile
    const bool regs = threads == 256 && items % 256 == 0 && (q & (q - 1)) == 0;
    const int niter = regs ? items / 256 : 0;
    const int grid = bn * (32 / gpc);
    if (niter == 8) gn_reg_kernel<8><<<grid, 256, 0, st>>>(x, out, gamma, beta, ss, HW, C, cpg, gpc, act, round_tf32);
    else if (niter == 4) gn_reg_kernel<4><<<grid, 256, 0, st>>>(x, out, gamma, beta, ss, HW, C, cpg, gpc, act, round_tf32);
    else if (niter == 2) gn_reg_kernel<2><<<grid, 256, 0, st>>>(x, out, gamma, beta, ss, HW, C, cpg, gpc, act, round_tf32);
    else if (niter == 1) gn_reg_kernel<1><<<grid, 256, 0, st>>>(x, out, gamma, beta, ss, HW, C, cpg, gpc, act, round_tf32);
    else
      gn_tile_kernel<<<grid, threads, static_cast<size_t>(HW) * q * 16 + threads * 4, st>>>(x, out, gamma, beta, ss, HW, C, cpg, gpc, act,
                                                                                           round_tf32);
  } else {
    gn_kernel<<<bn * 32, 256, 0, st>>>(x, out, gamma, beta, ss, HW, C, cpg, act, round_tf32);
  }
  AP_LAUNCH_CHECK();
  return AP_OK;
}
int launch_gn_bwd(const float* x, const float* gy, float* gx, const float* gamma, const float* beta, const float* ss, int bn, int HW, int C,
                  int act, int accumulate, cudaStream_t st) {
  const int cpg = C / 32;
  int gpc = gn_groups_per_cta(HW, cpg, 2, 48 * 1024);
  if (!gpc && gn_groups_per_cta(HW, cpg, 2, kGnSmemMax - 2048)) gpc = 1;
  if (gpc) {
    const int q = gpc * cpg / 4, threads = (256 / q) * q;
    gn_bwd_tile_kernel<<<bn * (32 / gpc), threads, static_cast<size_t>(HW) * q * 32 + threads * 4, st>>>(x, gy, gx, gamma, beta, ss, HW, C, cpg,
                                                                                                        gpc, act, accumulate);
  } else {
    gn_bwd_kernel<<<bn * 32, 256, 0, st>>>(x, gy, gx, gamma, beta, ss, HW, C, cpg, act, accumulate);
  }
  AP_LAUNCH_CHECK();
  return AP_OK;
}
}  // namespace

static int unet_time_path(ap_unet_t h, float t, cudaStream_t st) {
  const int mc = h->cfg.model_channels, ted = 4 * mc;
  unet_time_embed_kernel<<<1, 512, sizeof(float) * (mc + ted), st>>>(t, mc, h->te_w0.as<float>(), h->te_b0.as<float>(),
                                                                     h->te_w2.as<float>(), h->te_b2.as<float>(), h->emb_silu.as<float>());
  AP_LAUNCH_CHECK();
  unet_emb_proj_kernel<<<ceil_div(h->ss_rows * 32, 256), 256, 0, st>>>(h->emb_silu.as<float>(), ted, h->emb_w.as<float>(),
                                                                       h->emb_b.as<float>(), h->ss_rows, h->ss.as<float>());
  AP_LAUNCH_CHECK();
  return AP_OK;
}
static int unet_kernel_attrs(ap_unet_t h) {     // per handle, i.e. per device: function attributes belong to the device's context
  if (h->attn_attr) return AP_OK;
  AP_CUDA(cudaFuncSetAttribute(gn_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGnSmemMax)));
  AP_CUDA(cudaFuncSetAttribute(gn_bwd_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGnSmemMax)));
  AP_CUDA(cudaFuncSetAttribute(unet_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kAttSmemMax)));
  AP_CUDA(cudaFuncSetAttribute(unet_attn_mma_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kAttSmemMax)));
  AP_CUDA(cudaFuncSetAttribute(unet_attn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * 64 * 4));
  AP_CUDA(cudaFuncSetAttribute(unet_attn_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * 64 * 4));
  h->attn_attr = true;
  return AP_OK;
}
static int unet_chunk_size(ap_unet_t h, int B, bool with_backward) {
  const size_t per = with_backward ? h->need_per_sample + h->need_grad_per_sample : h->need_fwd_persist + h->need_fwd_tmp;
  const int max_chunk = static_cast<int>(std::min<size_t>(std::max<size_t>(1, h->budget_floats / per), 1u << 20));
  const int n_chunks = (B + max_chunk - 1) / max_chunk;
  return (B + n_chunks - 1) / n_chunks;                       // balanced sub-batches
}
static int unet_conv(ap_unet_t h, const ConvLayer& L, const float* in, int bn, int H, float* out, const float* res, cudaStream_t st) {
  if (h->mode == AP_MODE_TF32 && L.has_tc && conv_tc_supported(L.Cin, L.Cout, L.groups, H, H, L.kh, L.kw, L.stride, L.pad)) {
    ConvTcBinding bnd;
    int rcb = L.tc.bind(&bnd, in, bn, H, H, out, res, 0, 0);
    return rcb != AP_OK ? rcb : L.tc.run(bnd, st);
  }
  return L.run(in, bn, H, H, out, res, 0, st);
}

// forward of `bn` samples starting at x (NHWC == NCHW for one channel); records the operations when `tape` is given
static int unet_forward_chunk(ap_unet_t h, const float* x, float* eps, int bn, cudaStream_t st, std::vector<TapeEntry>* tape) {
  const int S = h->cfg.image_size, heads = h->cfg.num_heads;
  float* arena = h->arena.as<float>();
  // alloc: block outputs (alive until the chunk is done); talloc: a block's temporaries -- recycled per block in a forward-only
  // evaluation (scratch region at the end of the arena), kept like everything else when the operations are recorded
  const size_t scratch0 = tape ? h->arena_floats : h->arena_floats - h->need_fwd_tmp * bn;
  size_t top = 0, ttop = 0;
  auto alloc = [&](size_t n) {
    float* p = arena + top;
    top += (n + 3) & ~static_cast<size_t>(3);
    return p;
  };
  auto talloc = [&](size_t n) {
    if (tape) return alloc(n);
    float* p = arena + scratch0 + ttop;
    ttop += (n + 3) & ~static_cast<size_t>(3);
    return p;
  };
  struct Act { float* p; int H, C; };
  std::vector<Act> stack;
  Act cur{const_cast<float*>(x), S, 1};
  auto gn = [&](const Act& a, float* out, const DevBuf& g, const DevBuf& b, const float* ss, int act) -> int {
    int rcg = launch_gn(a.p, out, g.as<float>(), b.as<float>(), ss, bn, a.H * a.H, a.C, act, h->mode == AP_MODE_TF32, st);
    if (rcg != AP_OK) return rcg;
    if (tape) {
      TapeEntry e;
      e.kind = T_GN, e.in = a.p, e.out = out, e.Hin = a.H, e.Cin = e.Cout = a.C, e.gamma = &g, e.beta = &b, e.ss = ss, e.act = act;
      tape->push_back(e);
    }
    return AP_OK;
  };
  auto conv = [&](const ConvLayer& L, const float* in, int H, float* out, const float* res) -> int {
    if (tape) {
      TapeEntry e;
      e.kind = T_CONV, e.L = &L, e.in = in, e.out = out, e.res = res, e.Hin = H, e.Cin = L.Cin, e.Cout = L.Cout;
      tape->push_back(e);
    }
    return unet_conv(h, L, in, bn, H, out, res, st);
  };
  int rc = AP_OK;
  for (auto& opp : h->ops) {
    UOp& op = *opp;
    const size_t px = static_cast<size_t>(bn) * cur.H * cur.H;
    ttop = 0;
    if (op.kind == OP_CONV_IN) {
      float* o = alloc(px * op.cout);
      rc = conv(op.c1, cur.p, cur.H, o, nullptr);
      cur = {o, cur.H, op.cout};
      stack.push_back(cur);                              // hs.append(h) of input_blocks[0] (unet.py:483-485)
    } else if (op.kind == OP_RES) {
      float* n1 = talloc(px * op.cin);
      float* y1 = talloc(px * op.cout);
      float* n2 = talloc(px * op.cout);
      float* o = alloc(px * op.cout);
      rc = gn(cur, n1, op.g1, op.b1, nullptr, 1);                                             // in_layers: GN, SiLU
      if (rc == AP_OK) rc = conv(op.c1, n1, cur.H, y1, nullptr);                              //            conv
      Act a1{y1, cur.H, op.cout};
      if (rc == AP_OK) rc = gn(a1, n2, op.g2, op.b2, h->ss.as<float>() + op.ss_off, 1);       // out_layers[0] * (1 + scale) + shift, SiLU
      const float* res = cur.p;
      if (rc == AP_OK && op.has_skip) {
        float* sk = talloc(px * op.cout);
        rc = conv(op.skip, cur.p, cur.H, sk, nullptr);
        res = sk;
      }
      if (rc == AP_OK) rc = conv(op.c2, n2, cur.H, o, res);                                   // conv + skip_connection(x)
      cur = {o, cur.H, op.cout};
    } else if (op.kind == OP_ATTN) {
      const int T = cur.H * cur.H, Cc = cur.C;
      float* n1 = talloc(px * Cc);
      float* qkv = talloc(px * 3 * Cc);
      float* av = talloc(px * Cc);
      float* o = alloc(px * Cc);
      rc = gn(cur, n1, op.g1, op.b1, nullptr, 0);
      if (rc == AP_OK) rc = conv(op.c1, n1, cur.H, qkv, nullptr);
      if (rc == AP_OK) {
        const size_t smem = static_cast<size_t>(2) * T * 64 * sizeof(float);
        AP_REQUIRE(T <= 256, "ap_unet_eps: attention over more than 256 positions is not supported");
        AP_REQUIRE(Cc == 64 * heads, "ap_unet_eps: attention heads must have 64 channels");
        float* lse = nullptr;
        if (h->mode == AP_MODE_TF32 && T % 64 == 0) {      // tensor cores; the log-sum-exp rows are kept when a backward pass follows
          if (tape) lse = alloc(static_cast<size_t>(bn) * heads * T);
          unet_attn_mma_kernel<<<dim3(bn * heads, (T + 127) / 128), T < 128 ? T * 2 : 256, static_cast<size_t>(2) * std::min(T, ATT_KT) * ATT_LD * 4, st>>>(
              qkv, av, lse, T, Cc, heads);
        } else {
          unet_attn_kernel<64><<<bn * heads, T < 256 ? ((T + 31) / 32) * 32 : 256, smem, st>>>(qkv, av, T, Cc, heads);
        }
        AP_LAUNCH_CHECK();
        if (tape) {
          TapeEntry e;
          e.kind = T_ATTN, e.in = qkv, e.out = av, e.Hin = cur.H, e.Cin = 3 * Cc, e.Cout = Cc, e.lse = lse;
          tape->push_back(e);
        }
        rc = conv(op.c2, av, cur.H, o, cur.p);                                                // proj_out + x
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
      if (tape) {
        TapeEntry e;
        e.kind = T_CONCAT, e.in = cur.p, e.res = s.p, e.out = o, e.Hin = cur.H, e.Cin = cur.C, e.Cout = op.cout;
        tape->push_back(e);
      }
      cur = {o, cur.H, op.cout};
    } else if (op.kind == OP_DOWN) {
      const int Ho = cur.H / 2;
      float* o = alloc(static_cast<size_t>(bn) * Ho * Ho * op.cout);
      rc = conv(op.c1, cur.p, cur.H, o, nullptr);
      cur = {o, Ho, op.cout};
    } else if (op.kind == OP_UP) {
      float* up = talloc(4 * px * op.cin);
      float* o = alloc(4 * px * op.cout);
      nearest_up2_kernel<<<grid_for_n(static_cast<long long>(px) * op.cin, 256), 256, 0, st>>>(
          reinterpret_cast<const float4*>(cur.p), reinterpret_cast<float4*>(up), bn, cur.H, cur.H, op.cin / 4);
      AP_LAUNCH_CHECK();
      if (tape) {
        TapeEntry e;
        e.kind = T_UP2, e.in = cur.p, e.out = up, e.Hin = cur.H, e.Cin = e.Cout = op.cin;
        tape->push_back(e);
      }
      rc = conv(op.c1, up, 2 * cur.H, o, nullptr);
      cur = {o, 2 * cur.H, op.cout};
    } else if (op.kind == OP_OUT) {
      float* n1 = talloc(px * op.cin);
      rc = gn(cur, n1, op.g1, op.b1, nullptr, 1);
      if (rc == AP_OK) rc = conv(op.c1, n1, cur.H, eps, nullptr);
    }
    if (rc != AP_OK) return rc;
    if (top > scratch0 || scratch0 + ttop > h->arena_floats)
      return fail(AP_ERR_STATE, "ap_unet_eps: activation arena overflow (%zu + %zu > %zu floats)", top, ttop, h->arena_floats);
  }
  return AP_OK;
}

static int unet_reserve(ap_unet_t h, int chunk, bool with_backward) {
  const size_t want = (with_backward ? h->need_per_sample : h->need_fwd_persist + h->need_fwd_tmp) * chunk;
  if (h->arena_floats < want) {
    h->arena_floats = 0;
    AP_CUDA(h->arena.alloc(want * sizeof(float)));
    h->arena_floats = want;
  }
  const size_t gwant = with_backward ? h->need_grad_per_sample * chunk : 0;
  if (h->garena_floats < gwant) {
    h->garena_floats = 0;
    AP_CUDA(h->garena.alloc(gwant * sizeof(float)));
    h->garena_floats = gwant;
  }
  return AP_OK;
}

// eps[b] = UNet(x[b], t) with the same discrete step t for every sample (RevVPSDE.rvpsde_fn passes one step per Euler step,
// improved_diffusion_sde.py:104-105).  x, eps: device fp32 (B, 1, S, S).
extern "C" int ap_unet_eps(ap_unet_t h, const float* x, float t, float* eps, int B, void* stream) {
  AP_REQUIRE(h && x && eps && B > 0, "ap_unet_eps: bad arguments");
  AP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int S = h->cfg.image_size;
  const int chunk = unet_chunk_size(h, B, false);          // activations: sub-batches that keep the arena below ~8 GB
  int rc = unet_kernel_attrs(h);
  if (rc == AP_OK) rc = unet_reserve(h, chunk, false);
  if (rc == AP_OK) rc = unet_time_path(h, t, st);
  for (int b0 = 0; rc == AP_OK && b0 < B; b0 += chunk)
    rc = unet_forward_chunk(h, x + static_cast<size_t>(b0) * S * S, eps + static_cast<size_t>(b0) * S * S, std::min(chunk, B - b0), st,
                            nullptr);
  return rc;
}

// g_x = (d eps / d x)^T g_eps at (x, t): what autograd computes through UNetModel.forward when a white-box attack back-propagates
// through RevImprovedDiffusion.  The forward is recomputed with its operations recorded, then walked in reverse.  eps_out may be
// null.  x, g_eps, g_x, eps_out: device fp32 (B, 1, S, S).
extern "C" int ap_unet_eps_vjp(ap_unet_t h, const float* x, float t, const float* g_eps, float* g_x, float* eps_out, int B, void* stream) {
  AP_REQUIRE(h && x && g_eps && g_x && B > 0, "ap_unet_eps_vjp: bad arguments");
  AP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int S = h->cfg.image_size, heads = h->cfg.num_heads;
  if (!h->twins_ready) {      // data-gradient twins of every convolution (transposed, rotated weights), built on first use
    for (auto& opp : h->ops)
      for (ConvLayer* L : {&opp->c1, &opp->c2, &opp->skip}) {
        if (L->Cin == 0) continue;
        auto tw = std::unique_ptr<ConvLayer>(new ConvLayer());
        int rc = init_dgrad(*tw, *L, true);
        if (rc != AP_OK) return rc;
        h->twins[L] = std::move(tw);
      }
    h->twins_ready = true;
  }
  const int chunk = unet_chunk_size(h, B, true);
  int rc = unet_kernel_attrs(h);
  if (rc == AP_OK) rc = unet_reserve(h, chunk, true);
  if (rc == AP_OK && h->eps_scratch.bytes < static_cast<size_t>(chunk) * S * S * sizeof(float))
    AP_CUDA(h->eps_scratch.alloc(static_cast<size_t>(chunk) * S * S * sizeof(float)));
  if (rc == AP_OK) rc = unet_time_path(h, t, st);
  for (int b0 = 0; rc == AP_OK && b0 < B; b0 += chunk) {
    const int bn = std::min(chunk, B - b0);
    const size_t o0 = static_cast<size_t>(b0) * S * S;
    float* eps = eps_out ? eps_out + o0 : h->eps_scratch.as<float>();
    std::vector<TapeEntry> tape;
    rc = unet_forward_chunk(h, x + o0, eps, bn, st, &tape);
    if (rc != AP_OK) break;
    // ---- reverse walk
    float* garena = h->garena.as<float>();
    size_t gtop = 0;
    auto galloc = [&](size_t n) {
      float* p = garena + gtop;
      gtop += (n + 3) & ~static_cast<size_t>(3);
      return p;
    };
    std::unordered_map<const float*, float*> grad;     // activation -> its gradient buffer (present = holds a value)
    grad[eps] = const_cast<float*>(g_eps) + o0;
    const float* x_in = x + o0;
    // destination for the gradient of `act`: (buffer, whether it already holds a value to accumulate into)
    auto dest = [&](const float* act, size_t n, bool& had) -> float* {
      auto it = grad.find(act);
      if (it != grad.end()) {
        had = true;
        return it->second;
      }
      had = false;
      float* p = act == x_in ? g_x + o0 : galloc(n);
      grad[act] = p;
      return p;
    };
    for (size_t ei = tape.size(); rc == AP_OK && ei-- > 0;) {
      const TapeEntry& e = tape[ei];
      auto go = grad.find(e.out);
      if (go == grad.end()) continue;                  // no gradient reaches this output
      float* g_out = go->second;
      const size_t px_in = static_cast<size_t>(bn) * e.Hin * e.Hin;
      bool had = false;
      if (e.kind == T_CONV) {
        const ConvLayer& L = *e.L;
        const ConvLayer& T = *h->twins.at(e.L);
        const int Ho = (e.Hin + 2 * L.pad - L.kh) / L.stride + 1;
        const float* src = g_out;
        int Hs = Ho;
        if (L.stride == 2) {
          float* up = galloc(static_cast<size_t>(bn) * (2 * Ho) * (2 * Ho) * L.Cout);
          zero_up2_kernel<<<grid_for_n(static_cast<long long>(bn) * 4 * Ho * Ho * L.Cout / 4, 256), 256, 0, st>>>(
              reinterpret_cast<const float4*>(g_out), reinterpret_cast<float4*>(up), bn, Ho, Ho, L.Cout / 4);
          AP_LAUNCH_CHECK();
          src = up, Hs = 2 * Ho;
        }
        float* dst = dest(e.in, px_in * L.Cin, had);
        rc = unet_conv(h, T, src, bn, Hs, dst, had ? dst : nullptr, st);
        if (rc == AP_OK && e.res) {                    // the tensor added in the epilogue receives g_out unchanged
          auto ir = grad.find(e.res);
          if (ir == grad.end()) {
            if (e.res == x_in) {                       // (cannot happen: x is never a residual)
              AP_CUDA(cudaMemcpyAsync(g_x + o0, g_out, sizeof(float) * bn * Ho * Ho * L.Cout, cudaMemcpyDeviceToDevice, st));
              grad[e.res] = g_x + o0;
            } else {
              grad[e.res] = g_out;                     // alias: g_out is not needed after this entry
            }
          } else {
            add_inplace_kernel<<<grid_for_n(static_cast<long long>(bn) * Ho * Ho * L.Cout, 256), 256, 0, st>>>(
                ir->second, g_out, static_cast<long long>(bn) * Ho * Ho * L.Cout);
            AP_LAUNCH_CHECK();
          }
        }
      } else if (e.kind == T_GN) {
        float* dst = dest(e.in, px_in * e.Cin, had);
        rc = launch_gn_bwd(e.in, g_out, dst, e.gamma->as<float>(), e.beta->as<float>(), e.ss, bn, e.Hin * e.Hin, e.Cin, e.act, had ? 1 : 0,
                           st);
      } else if (e.kind == T_ATTN) {
        const int T = e.Hin * e.Hin, Cc = e.Cout;
        float* dst = dest(e.in, px_in * e.Cin, had);   // qkv has a single consumer: written in full
        if (e.lse) {
          const int nblk = (T + 127) / 128;
          unet_attn_mma_bwd_kernel<<<dim3(bn * heads, 2 * nblk), T < 128 ? T * 2 : 256,
                                     static_cast<size_t>(2) * std::min(T, ATT_KT) * ATT_LD * 4 + 2 * T * 4, st>>>(e.in, e.out, e.lse, g_out, dst, T, Cc, heads);
        } else {
          float* Pm = galloc(static_cast<size_t>(bn) * heads * T * T);
          float* Gs = galloc(static_cast<size_t>(bn) * heads * T * T);
          unet_attn_bwd_kernel<64><<<bn * heads, T < 256 ? ((T + 31) / 32) * 32 : 256, static_cast<size_t>(2) * T * 64 * sizeof(float), st>>>(
              e.in, g_out, dst, Pm, Gs, T, Cc, heads);
        }
        AP_LAUNCH_CHECK();
      } else if (e.kind == T_UP2) {
        float* dst = dest(e.in, px_in * e.Cin, had);
        up2_bwd_kernel<<<grid_for_n(static_cast<long long>(px_in) * e.Cin / 4, 256), 256, 0, st>>>(
            reinterpret_cast<const float4*>(g_out), reinterpret_cast<float4*>(dst), bn, e.Hin, e.Hin, e.Cin / 4, had ? 1 : 0);
        AP_LAUNCH_CHECK();
      } else if (e.kind == T_CONCAT) {
        bool had_s = false;
        float* da = dest(e.in, px_in * e.Cin, had);
        float* ds = dest(e.res, px_in * (e.Cout - e.Cin), had_s);
        split_bwd_kernel<<<grid_for_n(static_cast<long long>(px_in) * e.Cout / 4, 256), 256, 0, st>>>(
            reinterpret_cast<const float4*>(g_out), reinterpret_cast<float4*>(da), reinterpret_cast<float4*>(ds),
            static_cast<long long>(px_in), e.Cin / 4, (e.Cout - e.Cin) / 4, had ? 1 : 0, had_s ? 1 : 0);
        AP_LAUNCH_CHECK();
      }
      if (gtop > h->garena_floats)
        return fail(AP_ERR_STATE, "ap_unet_eps_vjp: gradient arena overflow (%zu > %zu floats)", gtop, h->garena_floats);
    }
    if (rc == AP_OK && grad.find(x_in) == grad.end()) AP_CUDA(cudaMemsetAsync(g_x + o0, 0, sizeof(float) * bn * S * S, st));
  }
  return rc;
}
