// K3 / K7: per-step diffusion updates with in-kernel counter-based Philox4x32-10 + Box-Muller noise,
// smoothing-input construction, and the argmax / vote-count kernel.  All HBM-bound elementwise work:
// float4 accesses, grid sized to a multiple of the SM count, arithmetic in fp32 in the reference's operation order
// (explicit __f*_rn so nvcc cannot contract into FMAs the reference does not perform).
#include "ap_common.cuh"
#include "ap_internal.h"
#include "ap_philox.cuh"

namespace ap {

thread_local std::string g_last_error;
std::atomic<unsigned long long> g_launches{0};
std::atomic<unsigned long long> g_alloc_generation{0};

int select_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(AP_ERR_CUDA, "no CUDA device available (%s); libaudiopure_b200 has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(AP_ERR_INVALID, "device %d out of range [0,%d)", device, n);
  AP_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  AP_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(AP_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                prop.minor);
  return AP_OK;
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// ------------------------------------------------------------------------------------------------ generic driver
// One thread handles 4 consecutive elements [4i, 4i+4).  VEC: every pointer is 16 B aligned and n % 4 == 0.
struct Quad {
  float v[4];
};
template <bool VEC> __device__ __forceinline__ Quad ldq(const float* p, long long i, int cnt) {
  Quad q;
  if (VEC) {
    const float4 t = *reinterpret_cast<const float4*>(p + i);
    q.v[0] = t.x, q.v[1] = t.y, q.v[2] = t.z, q.v[3] = t.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) q.v[j] = j < cnt ? p[i + j] : 0.f;
  }
  return q;
}
template <bool VEC> __device__ __forceinline__ void stq(float* p, long long i, int cnt, const Quad& q) {
  if (VEC) {
    *reinterpret_cast<float4*>(p + i) = make_float4(q.v[0], q.v[1], q.v[2], q.v[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < cnt) p[i + j] = q.v[j];
  }
}

enum NoiseMode { NOISE_NONE = 0, NOISE_HOST = 1, NOISE_PHILOX = 2 };

template <class Op, bool VEC> __global__ void __launch_bounds__(256) ew_kernel(Op op, long long n, int noise_mode,
                                                                               const float* __restrict__ z,
                                                                               uint64_t seed, uint64_t offset) {
  const long long nq = (n + 3) >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < nq; q += stride) {
    const long long i = q << 2;
    const int cnt = (n - i) >= 4 ? 4 : static_cast<int>(n - i);
    Quad zz;
    zz.v[0] = zz.v[1] = zz.v[2] = zz.v[3] = 0.f;
    if (noise_mode == NOISE_HOST) zz = ldq<VEC>(z, i, cnt);
    else if (noise_mode == NOISE_PHILOX) normal4(offset + static_cast<uint64_t>(q), seed, zz.v);
    op.template apply<VEC>(i, cnt, zz);
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <class Op> static int launch_ew(Op op, long long n, bool vec_ok, int noise_mode, const float* z, uint64_t seed,
                                         uint64_t offset, cudaStream_t st) {
  if (n <= 0) return AP_OK;
  const long long nq = (n + 3) >> 2;
  const int threads = 256;
  long long blocks = ceil_div_ll(nq, threads);
  const long long cap = static_cast<long long>(num_sms()) * 8;  // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  const bool vec = vec_ok && (n % 4 == 0) && (noise_mode != NOISE_HOST || aligned16(z));
  if (vec) ew_kernel<Op, true><<<static_cast<unsigned>(blocks), threads, 0, st>>>(op, n, noise_mode, z, seed, offset);
  else ew_kernel<Op, false><<<static_cast<unsigned>(blocks), threads, 0, st>>>(op, n, noise_mode, z, seed, offset);
  AP_LAUNCH_CHECK();
  return AP_OK;
}

// ------------------------------------------------------------------------------------------------ operations
struct DiffuseOp {  // diffwave_ddpm.py:67
  const float* x0;
  float* xt;
  float a, b;
  template <bool VEC> __device__ void apply(long long i, int cnt, const Quad& z) const {
    Quad x = ldq<VEC>(x0, i, cnt), o;
#pragma unroll
    for (int j = 0; j < 4; ++j) o.v[j] = __fadd_rn(__fmul_rn(a, x.v[j]), __fmul_rn(b, z.v[j]));
    stq<VEC>(xt, i, cnt, o);
  }
};
struct DdpmStepOp {  // diffwave_ddpm.py:159 then :100
  float* x;
  const float* eps;
  float c_eps, sqrt_alpha, sigma;
  int add_noise;
  template <bool VEC> __device__ void apply(long long i, int cnt, const Quad& z) const {
    Quad xv = ldq<VEC>(x, i, cnt), e = ldq<VEC>(eps, i, cnt), o;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float mu = __fdiv_rn(__fsub_rn(xv.v[j], __fmul_rn(c_eps, e.v[j])), sqrt_alpha);
      o.v[j] = add_noise ? __fadd_rn(mu, __fmul_rn(sigma, z.v[j])) : mu;
    }
    stq<VEC>(x, i, cnt, o);
  }
};
struct SdeStepOp {  // diffwave_sde.py:80,98,103,124 + Euler-Maruyama
  float* x;
  const float* eps;
  ap_sde_coef c;
  template <bool VEC> __device__ void apply(long long i, int cnt, const Quad& z) const {
    Quad xv = ldq<VEC>(x, i, cnt), e = ldq<VEC>(eps, i, cnt), o;
    const float nhb = __fmul_rn(-0.5f, c.beta);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float drift = __fmul_rn(nhb, xv.v[j]);
      const float score = -__fdiv_rn(e.v[j], c.sqrt_1mab);
      const float f = -__fsub_rn(drift, __fmul_rn(c.diff2, score));
      const float y = __fadd_rn(xv.v[j], __fmul_rn(f, c.dt));
      o.v[j] = __fadd_rn(y, __fmul_rn(c.g, __fmul_rn(c.sqrt_dt, z.v[j])));
    }
    stq<VEC>(x, i, cnt, o);
  }
};
struct PredictX0Op {  // diffwave_ddpm.py:203
  const float* xt;
  const float* eps;
  float* x0;
  float a, b;
  template <bool VEC> __device__ void apply(long long i, int cnt, const Quad&) const {
    Quad xv = ldq<VEC>(xt, i, cnt), e = ldq<VEC>(eps, i, cnt), o;
#pragma unroll
    for (int j = 0; j < 4; ++j) o.v[j] = __fsub_rn(__fmul_rn(a, xv.v[j]), __fmul_rn(b, e.v[j]));
    stq<VEC>(x0, i, cnt, o);
  }
};
struct SmoothOp {  // certified_robust.py:46-48,54 ; x broadcast over the batch (L % 4 == 0 in the VEC instantiation)
  const float* x;
  float* out;
  float sigma, scale;
  long long L;
  template <bool VEC> __device__ void apply(long long i, int cnt, const Quad& z) const {
    Quad o;
    if (VEC) {
      Quad xv = ldq<true>(x, i % L, 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) o.v[j] = __fmul_rn(scale, __fadd_rn(xv.v[j], __fmul_rn(sigma, z.v[j])));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o.v[j] = j < cnt ? __fmul_rn(scale, __fadd_rn(x[(i + j) % L], __fmul_rn(sigma, z.v[j]))) : 0.f;
    }
    stq<VEC>(out, i, cnt, o);
  }
};
struct RandnOp {
  float* out;
  template <bool VEC> __device__ void apply(long long i, int cnt, const Quad& z) const { stq<VEC>(out, i, cnt, z); }
};

static inline int noise_mode_for(const float* z) { return z ? NOISE_HOST : NOISE_PHILOX; }

int diffuse(const float* x0, float a, float b, const float* z, uint64_t seed, uint64_t offset, float* xt, long long n,
            cudaStream_t st) {
  DiffuseOp op{x0, xt, a, b};
  return launch_ew(op, n, aligned16(x0) && aligned16(xt), noise_mode_for(z), z, seed, offset, st);
}
int ddpm_step(float* x, const float* eps, float c_eps, float sqrt_alpha, float sigma, const float* z, uint64_t seed,
              uint64_t offset, long long n, cudaStream_t st) {
  const int add_noise = sigma != 0.f;
  DdpmStepOp op{x, eps, c_eps, sqrt_alpha, sigma, add_noise};
  return launch_ew(op, n, aligned16(x) && aligned16(eps), add_noise ? noise_mode_for(z) : NOISE_NONE, z, seed, offset, st);
}
int sde_step(float* x, const float* eps, const ap_sde_coef& c, const float* z, uint64_t seed, uint64_t offset,
             long long n, cudaStream_t st) {
  SdeStepOp op{x, eps, c};
  const bool noise = c.g != 0.f;
  return launch_ew(op, n, aligned16(x) && aligned16(eps), noise ? noise_mode_for(z) : NOISE_NONE, z, seed, offset, st);
}
int predict_x0(const float* xt, const float* eps, float a, float b, float* x0, long long n, cudaStream_t st) {
  PredictX0Op op{xt, eps, x0, a, b};
  return launch_ew(op, n, aligned16(xt) && aligned16(eps) && aligned16(x0), NOISE_NONE, nullptr, 0, 0, st);
}

// ------------------------------------------------------------------------------------------------ votes
// One warp per 32 rows; argmax with first-index tie-break (torch.max semantics), block histogram in smem,
// one 64-bit atomic per non-empty class per block.
__global__ void __launch_bounds__(256) vote_kernel(const float* __restrict__ logits, int B, int K,
                                                   unsigned long long* __restrict__ counts, int* __restrict__ pred) {
  extern __shared__ int hist[];
  for (int k = threadIdx.x; k < K; k += blockDim.x) hist[k] = 0;
  __syncthreads();
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < B; row += gridDim.x * blockDim.x) {
    const float* p = logits + static_cast<long long>(row) * K;
    float best = p[0];
    int arg = 0;
    for (int k = 1; k < K; ++k) {
      const float v = p[k];
      if (v > best || (v != v && best == best)) best = v, arg = k;  // NaN counts as the maximum, like torch.max
    }
    if (pred) pred[row] = arg;
    if (counts) atomicAdd(&hist[arg], 1);
  }
  __syncthreads();
  if (counts)
    for (int k = threadIdx.x; k < K; k += blockDim.x)
      if (hist[k]) atomicAdd(&counts[k], static_cast<unsigned long long>(hist[k]));
}

int vote(const float* logits, int B, int K, long long* counts, int* pred, cudaStream_t st) {
  if (B <= 0) return AP_OK;
  int blocks = ceil_div(B, 256);
  if (blocks > num_sms() * 4) blocks = num_sms() * 4;
  vote_kernel<<<blocks, 256, K * sizeof(int), st>>>(logits, B, K, reinterpret_cast<unsigned long long*>(counts), pred);
  AP_LAUNCH_CHECK();
  return AP_OK;
}

}  // namespace ap

// ================================================================================================ C ABI
using namespace ap;

extern "C" {

const char* ap_last_error(void) { return g_last_error.c_str(); }
int ap_version(void) { return 100; }
unsigned long long ap_launch_count(void) { return g_launches.load(); }
unsigned long long ap_alloc_generation(void) { return g_alloc_generation.load(); }

int ap_fold_weight_norm(const float* g, const float* v, float* w, int cout, int fan_in) {
  AP_REQUIRE(g && v && w && cout > 0 && fan_in > 0, "ap_fold_weight_norm: bad arguments");
  for (int o = 0; o < cout; ++o) {
    const float* vo = v + static_cast<size_t>(o) * fan_in;
    double ss = 0.0;
    for (int i = 0; i < fan_in; ++i) ss += static_cast<double>(vo[i]) * vo[i];
    // torch: v * (g / norm) with a float32 norm
    const float norm = static_cast<float>(std::sqrt(ss));
    const float s = g[o] / norm;
    for (int i = 0; i < fan_in; ++i) w[static_cast<size_t>(o) * fan_in + i] = vo[i] * s;
  }
  return AP_OK;
}

uint64_t ap_noise_offset_stride(int B, int L) {
  return (static_cast<uint64_t>(B) * static_cast<uint64_t>(L) + 3) / 4;
}

#define AP_CHECK_BL(name) AP_REQUIRE(B > 0 && L > 0, name ": B and L must be positive (got %d, %d)", B, L)

int ap_diffuse(const float* x0, float sqrt_ab, float sqrt_1mab, const float* z, uint64_t seed, uint64_t offset,
               float* xt, int B, int L, void* stream) {
  AP_CHECK_BL("ap_diffuse");
  AP_REQUIRE(x0 && xt, "ap_diffuse: null pointer");
  return diffuse(x0, sqrt_ab, sqrt_1mab, z, seed, offset, xt, static_cast<long long>(B) * L,
                 static_cast<cudaStream_t>(stream));
}
int ap_ddpm_step(float* x, const float* eps, float c_eps, float sqrt_alpha, float sigma, const float* z, uint64_t seed,
                 uint64_t offset, int B, int L, void* stream) {
  AP_CHECK_BL("ap_ddpm_step");
  AP_REQUIRE(x && eps, "ap_ddpm_step: null pointer");
  AP_REQUIRE(sqrt_alpha != 0.f, "ap_ddpm_step: sqrt_alpha == 0");
  return ddpm_step(x, eps, c_eps, sqrt_alpha, sigma, z, seed, offset, static_cast<long long>(B) * L,
                   static_cast<cudaStream_t>(stream));
}
int ap_sde_step(float* x, const float* eps, const ap_sde_coef* c, const float* z, uint64_t seed, uint64_t offset,
                int B, int L, void* stream) {
  AP_CHECK_BL("ap_sde_step");
  AP_REQUIRE(x && eps && c, "ap_sde_step: null pointer");
  return sde_step(x, eps, *c, z, seed, offset, static_cast<long long>(B) * L, static_cast<cudaStream_t>(stream));
}
int ap_predict_x0(const float* xt, const float* eps, float a, float b, float* x0, int B, int L, void* stream) {
  AP_CHECK_BL("ap_predict_x0");
  AP_REQUIRE(xt && eps && x0, "ap_predict_x0: null pointer");
  return predict_x0(xt, eps, a, b, x0, static_cast<long long>(B) * L, static_cast<cudaStream_t>(stream));
}
int ap_smooth_inputs(const float* x, float sigma, float scale, const float* z, uint64_t seed, uint64_t offset,
                     float* out, int B, int L, void* stream) {
  AP_CHECK_BL("ap_smooth_inputs");
  AP_REQUIRE(x && out, "ap_smooth_inputs: null pointer");
  SmoothOp op{x, out, sigma, scale, static_cast<long long>(L)};
  const bool vec = aligned16(x) && aligned16(out) && (L % 4 == 0);
  return launch_ew(op, static_cast<long long>(B) * L, vec, noise_mode_for(z), z, seed, offset,
                   static_cast<cudaStream_t>(stream));
}
int ap_randn(float* out, uint64_t n, uint64_t seed, uint64_t offset, void* stream) {
  AP_REQUIRE(out, "ap_randn: null pointer");
  RandnOp op{out};
  return launch_ew(op, static_cast<long long>(n), aligned16(out), NOISE_PHILOX, nullptr, seed, offset,
                   static_cast<cudaStream_t>(stream));
}
int ap_vote_counts(const float* logits, int B, int K, long long* counts, int counts_len, void* stream) {
  AP_REQUIRE(logits && counts && B >= 0 && K > 0 && K <= 4096, "ap_vote_counts: bad arguments");
  AP_REQUIRE(K <= counts_len, "ap_vote_counts: %d classes do not fit a count vector of %d entries", K, counts_len);
  return vote(logits, B, K, counts, nullptr, static_cast<cudaStream_t>(stream));
}
int ap_argmax(const float* logits, int B, int K, int* pred, void* stream) {
  AP_REQUIRE(logits && pred && B >= 0 && K > 0 && K <= 4096, "ap_argmax: bad arguments");
  return vote(logits, B, K, nullptr, pred, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
