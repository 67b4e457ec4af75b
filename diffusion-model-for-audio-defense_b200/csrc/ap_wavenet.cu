// DiffWave network handle (replaces WaveNet_Speech_Commands / DiffWave, reference WaveNet.py:138-172 and
// diffwave_ddpm.py:36-205): weight packing, step-embedding kernels (K0), the fp32 FFMA parity path, the chunk loop
// and the whole-purifier entry point.  The bf16 tensor-core path lives in ap_wavenet_tc.cu.
//
// Activation layout (both modes): channels-last [waveform][position][channel], so a dilated tap is a contiguous
// channel vector at position l +- d and the implicit GEMM has positions as rows (M) and channels as K.
#include <cmath>
#include <cstdlib>

#include "ap_common.cuh"
#include "ap_internal.h"
#include "ap_sgemm.cuh"

namespace ap {

// ---------------------------------------------------------------------------------------------- K0: step embedding
// util.py:84-93 + WaveNet.py:124-126.  freq[j] = exp(-j ln(1e4)/(half-1)) is computed on the host exactly as torch does
// (float32 exp of a float32 argument) and passed in, so the sin/cos arguments are bit-identical to the reference.
__device__ __forceinline__ float swishf(float x) { return x * (1.f / (1.f + expf(-x))); }

__global__ void __launch_bounds__(512) embed_kernel(float t, const float* __restrict__ freq, int e_in, int e_mid,
                                                    int e_out, const float* __restrict__ w1,
                                                    const float* __restrict__ b1, const float* __restrict__ w2,
                                                    const float* __restrict__ b2, float* __restrict__ emb) {
  extern __shared__ float sm[];
  float* e = sm;           // e_in
  float* h1 = sm + e_in;   // e_mid
  const int half = e_in / 2;
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    const float arg = t * freq[j];
    e[j] = sinf(arg);
    e[half + j] = cosf(arg);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int o = warp; o < e_mid; o += nwarps) {
    float acc = 0.f;
    for (int k = lane; k < e_in; k += 32) acc = fmaf(w1[o * e_in + k], e[k], acc);
    for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) h1[o] = swishf(acc + b1[o]);
  }
  __syncthreads();
  for (int o = warp; o < e_out; o += nwarps) {
    float acc = 0.f;
    for (int k = lane; k < e_mid; k += 32) acc = fmaf(w2[o * e_mid + k], h1[k], acc);
    for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) emb[o] = swishf(acc + b2[o]);
  }
}

// ptab[n][c] = fc_t_n(emb)[c]  (WaveNet.py:82); one warp per output, grid = (C/8, N); row N of ptab stays zero.
__global__ void __launch_bounds__(256) ptab_kernel(const float* __restrict__ emb, int e_out, int C,
                                                   const float* __restrict__ w, const float* __restrict__ b,
                                                   float* __restrict__ ptab) {
  const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + warp;
  if (c >= C) return;
  const float* wr = w + (static_cast<size_t>(n) * C + c) * e_out;
  float acc = 0.f;
  for (int k = lane; k < e_out; k += 32) acc = fmaf(wr[k], emb[k], acc);
  for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) ptab[n * C + c] = acc + b[n * C + c];
}

// ---------------------------------------------------------------------------------------------- fp32 path kernels
// u0[m][c] = max(w[c] * x[m] + b[c], 0) + p0[c]        (init_conv + custom ReLU, WaveNet.py:147,13-19; then :84)
__global__ void __launch_bounds__(256) init_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ b, const float* __restrict__ p0,
                                                       float* __restrict__ u, long long M, int C) {
  const long long total = M * (C / 4);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / (C / 4);
    const int c = static_cast<int>(i - m * (C / 4)) * 4;
    const float xv = x[m];
    float4 o;
    o.x = __fadd_rn(fmaxf(__fadd_rn(__fmul_rn(w[c + 0], xv), b[c + 0]), 0.f), p0[c + 0]);
    o.y = __fadd_rn(fmaxf(__fadd_rn(__fmul_rn(w[c + 1], xv), b[c + 1]), 0.f), p0[c + 1]);
    o.z = __fadd_rn(fmaxf(__fadd_rn(__fmul_rn(w[c + 2], xv), b[c + 2]), 0.f), p0[c + 2]);
    o.w = __fadd_rn(fmaxf(__fadd_rn(__fmul_rn(w[c + 3], xv), b[c + 3]), 0.f), p0[c + 3]);
    *reinterpret_cast<float4*>(u + m * C + c) = o;
  }
}

__device__ __forceinline__ float sigmoidf_precise(float x) { return 1.f / (1.f + expf(-x)); }

// GEMM-1 epilogue: column tile t holds [tanh channels 64t..64t+63 | sigmoid channels 64t..64t+63]   (WaveNet.py:90)
struct GateEpi {
  float* out;        // [M][C]
  const float* bias; // packed column order, 2C
  int C;
  __device__ __forceinline__ void store(int, int m, int n0, int tx, const float (&lo)[4], const float (&hi)[4], int) const {
    const int c = (n0 >> 1) + tx * 4;
    float4 o;
    float* po = reinterpret_cast<float*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = lo[j] + bias[n0 + tx * 4 + j];
      const float g = hi[j] + bias[n0 + 64 + tx * 4 + j];
      po[j] = tanhf(a) * sigmoidf_precise(g);
    }
    *reinterpret_cast<float4*>(out + static_cast<long long>(m) * C + c) = o;
  }
};

// GEMM-2 epilogue: columns [0,C) = res conv, [C,2C) = skip conv      (WaveNet.py:93-97,133)
struct ResSkipEpi {
  const float* u_in;
  float* u_out;
  float* skip;
  const float* bias;    // res bias (C) | skip bias (C)
  const float* p_next;  // step-embedding projection of the next layer (zeros after the last)
  int C;
  float sqrt_half;
  __device__ __forceinline__ void one(int m, int n, float acc) const {
    const long long row = static_cast<long long>(m) * C;
    if (n < C) {
      const float h = __fmul_rn(__fadd_rn(u_in[row + n], __fadd_rn(acc, bias[n])), sqrt_half);
      u_out[row + n] = __fadd_rn(h, p_next[n]);
    } else {
      const int c = n - C;
      skip[row + c] = __fadd_rn(skip[row + c], __fadd_rn(acc, bias[n]));
    }
  }
  __device__ __forceinline__ void store(int, int m, int n0, int tx, const float (&lo)[4], const float (&hi)[4], int) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      one(m, n0 + tx * 4 + j, lo[j]);
      one(m, n0 + 64 + tx * 4 + j, hi[j]);
    }
  }
};

// head GEMM: A = skip * sqrt(1/N) (WaveNet.py:135), epilogue y = relu(acc + b) (WaveNet.py:160-161)
struct RowScaled {
  const float* a;
  int ld;
  float scale;
  __device__ __forceinline__ float4 load4(int, int m, int k, int M, int K) const {
    if (m >= M || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v = *reinterpret_cast<const float4*>(a + static_cast<long long>(m) * ld + k);
    v.x = __fmul_rn(v.x, scale), v.y = __fmul_rn(v.y, scale), v.z = __fmul_rn(v.z, scale), v.w = __fmul_rn(v.w, scale);
    return v;
  }
};
struct ReluEpi {
  float* y;
  const float* bias;
  int ld;
  __device__ __forceinline__ void store(int, int m, int n0, int tx, const float (&lo)[4], const float (&hi)[4], int N) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n1 = n0 + tx * 4 + j, n2 = n0 + 64 + tx * 4 + j;
      if (n1 < N) y[static_cast<long long>(m) * ld + n1] = fmaxf(lo[j] + bias[n1], 0.f);
      if (n2 < N) y[static_cast<long long>(m) * ld + n2] = fmaxf(hi[j] + bias[n2], 0.f);
    }
  }
};
// eps[m] = w2 . y[m] + b2   (zero-conv 256 -> 1, WaveNet.py:162); one warp per row
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ y, const float* __restrict__ w2,
                                                     const float* __restrict__ b2, float* __restrict__ eps, long long M,
                                                     int S) {
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long m = warp; m < M; m += nwarps) {
    float acc = 0.f;
    for (int c = lane; c < S; c += 32) acc = fmaf(y[m * S + c], w2[c], acc);
    for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) eps[m] = acc + b2[0];
  }
}

}  // namespace ap

using namespace ap;

// ================================================================================================ handle
struct ap_diffwave_s {
  ap_wavenet_cfg cfg{};
  int device = 0;
  int mode = AP_MODE_FP32;
  bool tc_capable = false;
  // embedding
  DevBuf freq, fc1_w, fc1_b, fc2_w, fc2_b, fct_w, fct_b, emb, ptab;
  // fp32 network
  DevBuf init_w, init_b, wd32, bd32, wrs32, brs32, f1w32, f1b, f2w, f2b;
  // fp32 workspace
  DevBuf u0, u1, outb, skip;
  int chunk = 0, L = 0;
  // purifier scratch
  DevBuf eps_buf;
  // bf16 tensor-core network
  TcNet* tc = nullptr;
  int tc_chunk = 0, tc_L = 0;
  bool user_reserved = false;   // the workspace size was chosen by ap_diffwave_reserve: implicit calls do not change it
};

static int upload(DevBuf& d, const std::vector<float>& v) {
  AP_CUDA(d.upload(v.data(), v.size() * sizeof(float)));
  return AP_OK;
}
static int upload(DevBuf& d, const float* p, size_t n) {
  AP_CUDA(d.upload(p, n * sizeof(float)));
  return AP_OK;
}

extern "C" int ap_diffwave_create(ap_diffwave_t* out, const ap_wavenet_cfg* cfg, const float* const* weights,
                                  int n_weights, int device) {
  AP_REQUIRE(out && cfg && weights, "ap_diffwave_create: null argument");
  *out = nullptr;
  const int C = cfg->res_channels, S = cfg->skip_channels, N = cfg->num_res_layers;
  AP_REQUIRE(cfg->in_channels == 1 && cfg->out_channels == 1, "ap_diffwave_create: in/out channels must be 1");
  AP_REQUIRE(C == S, "ap_diffwave_create: skip_channels must equal res_channels (got %d, %d)", S, C);
  AP_REQUIRE(C > 0 && C % 64 == 0, "ap_diffwave_create: res_channels must be a positive multiple of 64 (got %d)", C);
  AP_REQUIRE(N > 0 && cfg->dilation_cycle > 0 && cfg->dilation_cycle <= 24, "ap_diffwave_create: bad layer/dilation config");
  AP_REQUIRE(cfg->embed_dim_in > 0 && cfg->embed_dim_in % 2 == 0 && cfg->embed_dim_mid > 0 && cfg->embed_dim_out > 0,
             "ap_diffwave_create: bad embedding dims");
  AP_REQUIRE(n_weights == 6 + 8 * N + 4, "ap_diffwave_create: expected %d weight tensors, got %d", 6 + 8 * N + 4, n_weights);
  for (int i = 0; i < n_weights; ++i) AP_REQUIRE(weights[i], "ap_diffwave_create: weight pointer %d is null", i);
  int rc = select_device(device);
  if (rc != AP_OK) return rc;

  auto* h = new ap_diffwave_s();
  h->cfg = *cfg;
  h->device = device;
  const int Ein = cfg->embed_dim_in, Emid = cfg->embed_dim_mid, Eout = cfg->embed_dim_out;
#define AP_TRY(expr)        \
  do {                      \
    int rc__ = (expr);      \
    if (rc__ != AP_OK) {    \
      ap_diffwave_destroy(h); \
      return rc__;          \
    }                       \
  } while (0)

  {  // util.py:86-87: _embed = exp(arange(half) * -(ln(1e4)/(half-1))) evaluated in float32
    const int half = Ein / 2;
    std::vector<float> f(half);
    const double scale = std::log(10000.0) / (half - 1);
    for (int j = 0; j < half; ++j) f[j] = expf(static_cast<float>(j) * static_cast<float>(-scale));
    AP_TRY(upload(h->freq, f));
  }
  AP_TRY(upload(h->init_w, weights[0], C));
  AP_TRY(upload(h->init_b, weights[1], C));
  AP_TRY(upload(h->fc1_w, weights[2], static_cast<size_t>(Emid) * Ein));
  AP_TRY(upload(h->fc1_b, weights[3], Emid));
  AP_TRY(upload(h->fc2_w, weights[4], static_cast<size_t>(Eout) * Emid));
  AP_TRY(upload(h->fc2_b, weights[5], Eout));
  {
    std::vector<float> fw(static_cast<size_t>(N) * C * Eout), fb(static_cast<size_t>(N) * C);
    // packed fp32 GEMM operands, [K][N] with N contiguous
    const int N2 = 2 * C;
    std::vector<float> wd(static_cast<size_t>(N) * 3 * C * N2), bd(static_cast<size_t>(N) * N2);
    std::vector<float> wrs(static_cast<size_t>(N) * C * N2), brs(static_cast<size_t>(N) * N2);
    for (int n = 0; n < N; ++n) {
      const float* const* w = weights + 6 + 8 * n;
      std::memcpy(&fw[static_cast<size_t>(n) * C * Eout], w[0], sizeof(float) * C * Eout);
      std::memcpy(&fb[static_cast<size_t>(n) * C], w[1], sizeof(float) * C);
      float* wdn = &wd[static_cast<size_t>(n) * 3 * C * N2];
      for (int j = 0; j < N2; ++j) {
        const int t = j / 128, jj = j % 128;
        const int oc = jj < 64 ? 64 * t + jj : C + 64 * t + (jj - 64);  // tanh half | sigmoid half of the same channels
        bd[static_cast<size_t>(n) * N2 + j] = w[3][oc];
        for (int tap = 0; tap < 3; ++tap)
          for (int c = 0; c < C; ++c)
            wdn[static_cast<size_t>(tap * C + c) * N2 + j] = w[2][(static_cast<size_t>(oc) * C + c) * 3 + tap];
      }
      float* wrn = &wrs[static_cast<size_t>(n) * C * N2];
      for (int j = 0; j < N2; ++j) {
        const float* src = j < C ? w[4] + static_cast<size_t>(j) * C : w[6] + static_cast<size_t>(j - C) * C;
        brs[static_cast<size_t>(n) * N2 + j] = j < C ? w[5][j] : w[7][j - C];
        for (int c = 0; c < C; ++c) wrn[static_cast<size_t>(c) * N2 + j] = src[c];
      }
    }
    AP_TRY(upload(h->fct_w, fw));
    AP_TRY(upload(h->fct_b, fb));
    AP_TRY(upload(h->wd32, wd));
    AP_TRY(upload(h->bd32, bd));
    AP_TRY(upload(h->wrs32, wrs));
    AP_TRY(upload(h->brs32, brs));
    const float* const* tail = weights + 6 + 8 * N;
    const int Np = ((S + 127) / 128) * 128;
    std::vector<float> f1(static_cast<size_t>(S) * Np, 0.f);
    for (int o = 0; o < S; ++o)
      for (int c = 0; c < S; ++c) f1[static_cast<size_t>(c) * Np + o] = tail[0][static_cast<size_t>(o) * S + c];
    AP_TRY(upload(h->f1w32, f1));
    AP_TRY(upload(h->f1b, tail[1], S));
    AP_TRY(upload(h->f2w, tail[2], S));
    AP_TRY(upload(h->f2b, tail[3], 1));
  }
  {
    cudaError_t e = h->emb.alloc(sizeof(float) * Eout);
    if (e == cudaSuccess) e = h->ptab.alloc(sizeof(float) * (N + 1) * C);
    if (e == cudaSuccess) e = cudaMemset(h->ptab.p, 0, sizeof(float) * (N + 1) * C);
    if (e != cudaSuccess) {
      ap_diffwave_destroy(h);
      return fail(AP_ERR_CUDA, "ap_diffwave_create: %s", cudaGetErrorString(e));
    }
  }
  h->tc_capable = (C == 256 && S == 256);
  if (h->tc_capable) {
    AP_TRY(tc_net_create(&h->tc, *cfg, weights));
    h->mode = AP_MODE_BF16;
  }
#undef AP_TRY
  *out = h;
  return AP_OK;
}

extern "C" void ap_diffwave_destroy(ap_diffwave_t h) {
  if (!h) return;
  if (h->tc) tc_net_destroy(h->tc);
  delete h;
}

extern "C" int ap_diffwave_set_mode(ap_diffwave_t h, int mode) {
  AP_REQUIRE(h, "ap_diffwave_set_mode: null handle");
  AP_REQUIRE(mode == AP_MODE_BF16 || mode == AP_MODE_FP32 || mode == AP_MODE_FP16 || mode == AP_MODE_BF16X3,
             "ap_diffwave_set_mode: unknown mode %d", mode);
  if (mode != AP_MODE_FP32 && !h->tc_capable)
    return fail(AP_ERR_INVALID, "ap_diffwave_set_mode: the tensor-core kernels need res_channels == skip_channels == 256");
  if (mode != AP_MODE_FP32) tc_net_set_dtype(h->tc, mode == AP_MODE_FP16 ? 1 : (mode == AP_MODE_BF16X3 ? 2 : 0));
  if (mode != h->mode && (mode == AP_MODE_BF16X3 || h->mode == AP_MODE_BF16X3)) h->tc_chunk = 0;   // workspace layout changes
  h->mode = mode;
  return AP_OK;
}
extern "C" int ap_diffwave_get_mode(ap_diffwave_t h) { return h ? h->mode : AP_ERR_INVALID; }

static int reserve_ws(ap_diffwave_t h, int chunk, int L);
extern "C" int ap_diffwave_reserve(ap_diffwave_t h, int chunk, int L) {
  AP_REQUIRE(h && chunk > 0 && L > 0, "ap_diffwave_reserve: bad arguments");
  int rc = reserve_ws(h, chunk, L);
  if (rc == AP_OK) h->user_reserved = true;
  return rc;
}
static int reserve_ws(ap_diffwave_t h, int chunk, int L) {
  AP_CUDA(cudaSetDevice(h->device));
  if (h->mode != AP_MODE_FP32) {
    if (h->tc_chunk == chunk && h->tc_L == L) return AP_OK;
    h->tc_chunk = 0, h->tc_L = 0;            // the old buffers are released first: nothing may describe them if this fails
    int rc = tc_net_reserve(h->tc, chunk, L);
    if (rc != AP_OK) return rc;
    h->tc_chunk = chunk, h->tc_L = L;
    return AP_OK;
  }
  if (h->chunk == chunk && h->L == L) return AP_OK;
  h->chunk = 0, h->L = 0;
  const size_t n = static_cast<size_t>(chunk) * L * h->cfg.res_channels * sizeof(float);
  h->u0.release(), h->u1.release(), h->outb.release(), h->skip.release();
  cudaError_t e = h->u0.alloc(n);
  if (e == cudaSuccess) e = h->u1.alloc(n);
  if (e == cudaSuccess) e = h->outb.alloc(n);
  if (e == cudaSuccess) e = h->skip.alloc(n);
  if (e != cudaSuccess) {
    h->u0.release(), h->u1.release(), h->outb.release(), h->skip.release();
    (void)cudaGetLastError();
    return fail(AP_ERR_CUDA, "DiffWave fp32 workspace: %.2f GB for %d waveforms of length %d -> %s", 4.0 * n / 1e9, chunk, L,
                cudaGetErrorString(e));
  }
  h->chunk = chunk, h->L = L;
  return AP_OK;
}

static int grid_for(long long work_items, int threads) {
  long long b = ceil_div_ll(work_items, threads);
  const long long cap = static_cast<long long>(num_sms()) * 8;
  return static_cast<int>(b < cap ? (b > 0 ? b : 1) : cap);
}

static int step_embedding(ap_diffwave_t h, float t, cudaStream_t st) {
  const ap_wavenet_cfg& c = h->cfg;
  const size_t smem = sizeof(float) * (c.embed_dim_in + c.embed_dim_mid);
  embed_kernel<<<1, 512, smem, st>>>(t, h->freq.as<float>(), c.embed_dim_in, c.embed_dim_mid, c.embed_dim_out,
                                     h->fc1_w.as<float>(), h->fc1_b.as<float>(), h->fc2_w.as<float>(),
                                     h->fc2_b.as<float>(), h->emb.as<float>());
  AP_LAUNCH_CHECK();
  dim3 grid(ceil_div(c.res_channels, 8), c.num_res_layers);
  ptab_kernel<<<grid, 256, 0, st>>>(h->emb.as<float>(), c.embed_dim_out, c.res_channels, h->fct_w.as<float>(),
                                    h->fct_b.as<float>(), h->ptab.as<float>());
  AP_LAUNCH_CHECK();
  return AP_OK;
}

// fp32 FFMA path for one chunk (B <= h->chunk)
static int eps_fp32_chunk(ap_diffwave_t h, const float* x, float* eps, int B, int L, cudaStream_t st, int stop_after = -1) {
  const ap_wavenet_cfg& c = h->cfg;
  const int C = c.res_channels, N = c.num_res_layers, N2 = 2 * C;
  const long long M = static_cast<long long>(B) * L;
  AP_REQUIRE(M < (1ll << 31) / 4, "fp32 chunk too large");
  float* u[2] = {h->u0.as<float>(), h->u1.as<float>()};
  const float* ptab = h->ptab.as<float>();
  init_f32_kernel<<<grid_for(M * (C / 4), 256), 256, 0, st>>>(x, h->init_w.as<float>(), h->init_b.as<float>(), ptab, u[0], M, C);
  AP_LAUNCH_CHECK();
  AP_CUDA(cudaMemsetAsync(h->skip.p, 0, static_cast<size_t>(M) * C * sizeof(float), st));
  const float sqrt_half = static_cast<float>(std::sqrt(0.5));
  for (int n = 0; n < N; ++n) {
    const int d = 1 << (n % c.dilation_cycle);
    sgemm::Conv1dTaps al{u[n & 1], L, C, 3, d};
    GateEpi ge{h->outb.as<float>(), h->bd32.as<float>() + static_cast<size_t>(n) * N2, C};
    AP_CUDA(sgemm::launch(al, h->wd32.as<float>() + static_cast<size_t>(n) * 3 * C * N2, N2, 0, 1, static_cast<int>(M), N2,
                          3 * C, ge, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    sgemm::Conv1dTaps a2{h->outb.as<float>(), L, C, 1, 1};
    ResSkipEpi re{u[n & 1], u[(n + 1) & 1], h->skip.as<float>(), h->brs32.as<float>() + static_cast<size_t>(n) * N2,
                  ptab + static_cast<size_t>(n + 1) * C, C, sqrt_half};
    AP_CUDA(sgemm::launch(a2, h->wrs32.as<float>() + static_cast<size_t>(n) * C * N2, N2, 0, 1, static_cast<int>(M), N2, C,
                          re, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (n == stop_after) return AP_OK;
  }
  const int S = c.skip_channels, Np = ((S + 127) / 128) * 128;
  RowScaled ah{h->skip.as<float>(), S, static_cast<float>(std::sqrt(1.0 / N))};
  ReluEpi he{h->outb.as<float>(), h->f1b.as<float>(), S};
  AP_CUDA(sgemm::launch(ah, h->f1w32.as<float>(), Np, 0, 1, static_cast<int>(M), S, S, he, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  rowdot_kernel<<<grid_for(M * 32, 256), 256, 0, st>>>(h->outb.as<float>(), h->f2w.as<float>(), h->f2b.as<float>(), eps, M, S);
  AP_LAUNCH_CHECK();
  return AP_OK;
}

// Workspace for a batch of B waveforms of length L in the current mode: keeps an explicit ap_diffwave_reserve, otherwise
// (re)sizes the implicit one.  Returns the chunk (waveforms per pass) in *chunk_out.
static int ensure_workspace(ap_diffwave_t h, int B, int L, int* chunk_out) {
  const bool tc = h->mode != AP_MODE_FP32;
  int chunk = tc ? h->tc_chunk : h->chunk;
  const int curL = tc ? h->tc_L : h->L;
  // default chunk.  Workspace per position: fp32 4 tensors * 256 ch * 4 B = 4 KB; bf16 (N + 2) tensors * 256 ch * 2 B
  // = 19.4 KB for 36 layers (the gate output of EVERY layer is kept for k2_head); bf16x3 twice that (two planes).
  // bf16: 148 waveforms of 1 s = 18500 tiles = 125 full rounds over 74 CTA pairs (no ragged last wave) = 46 GB;
  // bf16x3: half as many waveforms in the same bytes.  On top of that shape heuristic the implicit workspace is bounded
  // by HALF of the device memory that is free right now (plus what this handle already holds) and by
  // AP_DIFFWAVE_WORKSPACE_GB if set, so a shared or smaller GPU gets a smaller chunk instead of an allocation failure.
  const long long budget_positions = tc ? (h->mode == AP_MODE_BF16X3 ? 74ll : 148ll) * 16000 : (1ll << 18);
  long long want = budget_positions / L;
  if (want > B) want = B;
  if (chunk == 0 || curL != L || (!h->user_reserved && chunk < want)) {
    const int N = h->cfg.num_res_layers;
    const double per_wave = tc ? (h->mode == AP_MODE_BF16X3 ? 2.0 : 1.0) * L * 512.0 * (N + 2) : L * 4096.0;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
      const double held = tc ? static_cast<double>(tc_net_workspace_bytes(h->tc))
                             : static_cast<double>(h->u0.bytes + h->u1.bytes + h->outb.bytes + h->skip.bytes);
      double budget = 0.5 * (static_cast<double>(free_b) + held);
      if (const char* env = std::getenv("AP_DIFFWAVE_WORKSPACE_GB")) {
        const double cap = std::atof(env) * 1e9;
        if (cap > 0 && cap < budget) budget = cap;
      }
      const long long fit = static_cast<long long>(budget / per_wave);
      if (want > fit) want = fit;
    }
    if (want < 1) want = 1;
    // an implicitly sized workspace grows with the batch (e.g. after a small backward pass); an explicit reserve is kept
    int rc = reserve_ws(h, static_cast<int>(want), L);
    if (rc != AP_OK) return rc;
    if (chunk == 0 || curL != L) h->user_reserved = false;
    chunk = static_cast<int>(want);
  }
  *chunk_out = chunk;
  return AP_OK;
}

extern "C" int ap_diffwave_eps(ap_diffwave_t h, const float* x, float t, float* eps, int B, int L, void* stream) {
  AP_REQUIRE(h && x && eps, "ap_diffwave_eps: null argument");
  AP_REQUIRE(B > 0 && L > 0, "ap_diffwave_eps: B and L must be positive (got %d, %d)", B, L);
  AP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool tc = h->mode != AP_MODE_FP32;
  int chunk = 0;
  int rc = ensure_workspace(h, B, L, &chunk);
  if (rc != AP_OK) return rc;
  rc = step_embedding(h, t, st);
  if (rc != AP_OK) return rc;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bn = B - b0 < chunk ? B - b0 : chunk;
    const float* xc = x + static_cast<size_t>(b0) * L;
    float* ec = eps + static_cast<size_t>(b0) * L;
    rc = tc ? tc_net_eps(h->tc, xc, h->ptab.as<float>(), ec, bn, L, st) : eps_fp32_chunk(h, xc, ec, bn, L, st);
    if (rc != AP_OK) return rc;
  }
  return AP_OK;
}

// Certification front end in one call: x0[b] = a * x_in[b] - b * eps_theta(x_in[b], t), x_in[b] = scale * (x + sigma * z[b]).
// Tensor-core modes with L % 4 == 0: the noisy copies are built inside the network's init kernel and x0 is formed in
// k2_head's epilogue (no separate smoothing-input / predict-x0 passes over HBM).  Otherwise (fp32 FFMA mode, ragged L) the
// same result is computed by the unfused sequence ap_smooth_inputs -> ap_diffwave_eps -> ap_predict_x0.
extern "C" int ap_diffwave_smooth_denoise(ap_diffwave_t h, const float* x, float sigma, float scale, const float* z,
                                          uint64_t seed, uint64_t offset, const uint64_t* offset_dev, float t,
                                          float sqrt_recip_ab, float sqrt_recipm1_ab, float* x0, int B, int L, void* stream) {
  AP_REQUIRE(h && x && x0, "ap_diffwave_smooth_denoise: null argument");
  AP_REQUIRE(B > 0 && L > 0, "ap_diffwave_smooth_denoise: B and L must be positive (got %d, %d)", B, L);
  AP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool fused = h->mode != AP_MODE_FP32 && L % 4 == 0;
  if (!fused) {
    AP_REQUIRE(!offset_dev, "ap_diffwave_smooth_denoise: a device-resident noise offset needs a tensor-core mode and L %% 4 == 0");
    const size_t nb = static_cast<size_t>(B) * L * sizeof(float);
    if (h->eps_buf.bytes < nb) AP_CUDA(h->eps_buf.alloc(nb));
    int rc = ap_smooth_inputs(x, sigma, scale, z, seed, offset, x0, B, L, stream);
    if (rc != AP_OK) return rc;
    rc = ap_diffwave_eps(h, x0, t, h->eps_buf.as<float>(), B, L, stream);
    if (rc != AP_OK) return rc;
    return predict_x0(x0, h->eps_buf.as<float>(), sqrt_recip_ab, sqrt_recipm1_ab, x0, static_cast<long long>(B) * L, st);
  }
  int chunk = 0;
  int rc = ensure_workspace(h, B, L, &chunk);
  if (rc != AP_OK) return rc;
  rc = step_embedding(h, t, st);
  if (rc != AP_OK) return rc;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bn = B - b0 < chunk ? B - b0 : chunk;
    const size_t e0 = static_cast<size_t>(b0) * L;   // a multiple of 4: element e of the call is lane e % 4 of block offset + e / 4
    SmoothSrc src{x, sigma, scale, z ? z + e0 : nullptr, seed, offset + e0 / 4, offset_dev, sqrt_recip_ab, sqrt_recipm1_ab};
    rc = tc_net_smooth_denoise(h->tc, src, h->ptab.as<float>(), x0 + e0, bn, L, st);
    if (rc != AP_OK) return rc;
  }
  return AP_OK;
}

__global__ void u64_add_kernel(unsigned long long* p, unsigned long long inc) { *p += inc; }
extern "C" int ap_u64_add(uint64_t* dev, uint64_t inc, void* stream) {
  AP_REQUIRE(dev, "ap_u64_add: null pointer");
  u64_add_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned long long*>(dev), inc);
  AP_LAUNCH_CHECK();
  return AP_OK;
}

static long long vjp_chunk(ap_diffwave_t h, int B, int L);

// Vector-Jacobian product of the network: g_x = (d eps_theta(x, t) / d x)^T g_eps (what autograd computes through the
// reference's WaveNet when an attack calls loss.backward(), robustness_eval/white_box_attack.py:438).  bf16 mode.
extern "C" int ap_diffwave_eps_vjp(ap_diffwave_t h, const float* x, float t, const float* g_eps, float* g_x, float* eps_out,
                                   int B, int L, void* stream) {
  AP_REQUIRE(h && x && g_eps && g_x, "ap_diffwave_eps_vjp: null argument");
  AP_REQUIRE(B > 0 && L > 0, "ap_diffwave_eps_vjp: B and L must be positive (got %d, %d)", B, L);
  if (h->mode != AP_MODE_BF16 && h->mode != AP_MODE_BF16X3)
    return fail(AP_ERR_STATE, "ap_diffwave_eps_vjp: the backward pass exists in AP_MODE_BF16 and AP_MODE_BF16X3 only");
  AP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long bchunk = vjp_chunk(h, B, L);
  if (h->tc_chunk < bchunk || h->tc_L != L) {
    int rc = reserve_ws(h, static_cast<int>(bchunk), L);
    if (rc != AP_OK) return rc;
    h->user_reserved = false;
  }
  const size_t nb = static_cast<size_t>(bchunk) * L * sizeof(float);
  if (h->eps_buf.bytes < nb) AP_CUDA(h->eps_buf.alloc(nb));
  int rc = step_embedding(h, t, st);
  if (rc != AP_OK) return rc;
  for (int b0 = 0; b0 < B; b0 += static_cast<int>(bchunk)) {
    const int bn = B - b0 < bchunk ? B - b0 : static_cast<int>(bchunk);
    const size_t off = static_cast<size_t>(b0) * L;
    rc = tc_net_vjp(h->tc, x + off, h->ptab.as<float>(), g_eps + off, g_x + off, eps_out ? eps_out + off : nullptr,
                    h->eps_buf.as<float>(), bn, L, static_cast<int>(bchunk), st);
    if (rc != AP_OK) return rc;
  }
  return AP_OK;
}

static long long vjp_chunk(ap_diffwave_t h, int B, int L) {   // sub-batch keeping the saved activations below ~24 GB
  const size_t per = tc_net_bwd_bytes_per_waveform(h->tc, L);
  long long bchunk = static_cast<long long>((24ull << 30) / per);
  if (bchunk < 1) bchunk = 1;
  return bchunk > B ? B : bchunk;
}

// Forward that keeps the backward's inputs: eps = eps_theta(x, t) and *token identifies the saved state.  A following
// ap_diffwave_eps_vjp_saved(token) skips the recomputation; the state is overwritten by the next saving forward (including the
// recomputation inside ap_diffwave_eps_vjp), so in a chain of evaluations only the last one can be reused -- which is the first
// one the backward pass visits.  Returns AP_ERR_STATE when B exceeds the backward sub-batch (use ap_diffwave_eps then).
extern "C" int ap_diffwave_eps_save(ap_diffwave_t h, const float* x, float t, float* eps, int B, int L, void* stream,
                                    unsigned long long* token) {
  AP_REQUIRE(h && x && eps && token, "ap_diffwave_eps_save: null argument");
  AP_REQUIRE(B > 0 && L > 0, "ap_diffwave_eps_save: B and L must be positive (got %d, %d)", B, L);
  *token = 0;
  if (h->mode != AP_MODE_BF16 && h->mode != AP_MODE_BF16X3)
    return fail(AP_ERR_STATE, "ap_diffwave_eps_save: AP_MODE_BF16 and AP_MODE_BF16X3 only");
  AP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long bchunk = vjp_chunk(h, B, L);
  if (B > bchunk) return fail(AP_ERR_STATE, "ap_diffwave_eps_save: batch %d exceeds the backward sub-batch %lld", B, bchunk);
  if (h->tc_chunk < B || h->tc_L != L) {
    int rc = reserve_ws(h, B, L);
    if (rc != AP_OK) return rc;
    h->user_reserved = false;
  }
  int rc = step_embedding(h, t, st);
  if (rc != AP_OK) return rc;
  return tc_net_eps_save(h->tc, x, h->ptab.as<float>(), eps, B, L, static_cast<int>(bchunk), st, token);
}

extern "C" int ap_diffwave_eps_vjp_saved(ap_diffwave_t h, unsigned long long token, const float* x, const float* g_eps, float* g_x,
                                         int B, int L, void* stream) {
  AP_REQUIRE(h && x && g_eps && g_x, "ap_diffwave_eps_vjp_saved: null argument");
  AP_REQUIRE(h->tc, "ap_diffwave_eps_vjp_saved: no tensor-core network");
  if (!tc_net_saved_state_is(h->tc, token, B, L)) return fail(AP_ERR_STATE, "ap_diffwave_eps_vjp_saved: the saved state is stale");
  AP_CUDA(cudaSetDevice(h->device));
  return tc_net_backward(h->tc, x, g_eps, g_x, B, L, static_cast<cudaStream_t>(stream));
}

extern "C" int ap_diffwave_purify_ddpm(ap_diffwave_t h, const float* x0, float* out, int t_star, const float* coef4,
                                       const float* z, uint64_t seed, uint64_t offset, int B, int L, void* stream) {
  AP_REQUIRE(h && x0 && out && coef4, "ap_diffwave_purify_ddpm: null argument");
  AP_REQUIRE(B > 0 && L > 0 && t_star >= 1, "ap_diffwave_purify_ddpm: bad sizes (B=%d L=%d t*=%d)", B, L, t_star);
  AP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = static_cast<long long>(B) * L;
  if (h->eps_buf.bytes < static_cast<size_t>(n) * sizeof(float)) AP_CUDA(h->eps_buf.alloc(static_cast<size_t>(n) * sizeof(float)));
  const uint64_t stride = ap_noise_offset_stride(B, L);
  int rc = diffuse(x0, coef4[0], coef4[1], z, seed, offset, out, n, st);  // diffwave_ddpm.py:49-73
  if (rc != AP_OK) return rc;
  for (int i = 0; i < t_star; ++i) {                                     // diffwave_ddpm.py:95-103
    const float* c = coef4 + 4 * (i + 1);
    rc = ap_diffwave_eps(h, out, c[3], h->eps_buf.as<float>(), B, L, stream);
    if (rc != AP_OK) return rc;
    const bool last = (i == t_star - 1);
    const float sigma = last ? 0.f : c[2];
    const float* zi = (z && !last) ? z + static_cast<size_t>(i + 1) * n : nullptr;
    rc = ddpm_step(out, h->eps_buf.as<float>(), c[0], c[1], sigma, zi, seed, offset + stride * (i + 1), n, st);
    if (rc != AP_OK) return rc;
  }
  return AP_OK;
}

// Debug / test hook: run the network up to and including residual layer `layer` in the current mode and return the next
// layer's input u_{layer+1} = h + fc_t(emb) and the gate output o_layer, both as fp32 (B, L, C).  B must fit one chunk.
extern "C" int ap_diffwave_debug_layer(ap_diffwave_t h, const float* x, float t, int layer, float* u_next, float* gate,
                                       int B, int L, void* stream) {
  AP_REQUIRE(h && x, "ap_diffwave_debug_layer: null argument");
  AP_REQUIRE(B > 0 && L > 0 && layer >= 0 && layer < h->cfg.num_res_layers, "ap_diffwave_debug_layer: bad arguments");
  AP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = reserve_ws(h, B, L);
  if (rc != AP_OK) return rc;
  h->user_reserved = false;
  rc = step_embedding(h, t, st);
  if (rc != AP_OK) return rc;
  if (h->mode != AP_MODE_FP32) return tc_net_debug_layer(h->tc, x, h->ptab.as<float>(), layer, u_next, gate, B, L, st);
  rc = eps_fp32_chunk(h, x, nullptr, B, L, st, layer);
  if (rc != AP_OK) return rc;
  const size_t bytes = static_cast<size_t>(B) * L * h->cfg.res_channels * sizeof(float);
  const float* un = ((layer + 1) & 1) ? h->u1.as<float>() : h->u0.as<float>();
  if (u_next) AP_CUDA(cudaMemcpyAsync(u_next, un, bytes, cudaMemcpyDeviceToDevice, st));
  if (gate) AP_CUDA(cudaMemcpyAsync(gate, h->outb.p, bytes, cudaMemcpyDeviceToDevice, st));
  return AP_OK;
}

extern "C" int ap_diffwave_profile(ap_diffwave_t h, int enable) {
  AP_REQUIRE(h, "ap_diffwave_profile: null handle");
  AP_REQUIRE(h->tc, "ap_diffwave_profile: only the tensor-core path is instrumented");
  tc_net_profile(h->tc, enable != 0);
  return AP_OK;
}
extern "C" int ap_diffwave_profile_read(ap_diffwave_t h, double* ms2, int* count2) {
  AP_REQUIRE(h && ms2 && count2, "ap_diffwave_profile_read: null argument");
  AP_REQUIRE(h->tc, "ap_diffwave_profile_read: only the tensor-core path is instrumented");
  AP_CUDA(cudaSetDevice(h->device));
  return tc_net_profile_read(h->tc, ms2, count2);
}

// development aid (AP_TC_DEBUG=1): per-CTA wait-cycle counters of the last k1_layer launch, 256 rows x 16 counters
extern "C" int ap_diffwave_debug_counters(ap_diffwave_t h, long long* host16x256) {
  AP_REQUIRE(h && h->tc && host16x256, "ap_diffwave_debug_counters: bad arguments");
  AP_CUDA(cudaSetDevice(h->device));
  AP_CUDA(cudaDeviceSynchronize());
  return tc_net_debug_counters(h->tc, host16x256);
}
