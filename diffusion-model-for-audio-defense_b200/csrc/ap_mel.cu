// K4: log-mel front end (replaces torchaudio MelSpectrogram + AmplitudeToDB as constructed at
// certified_robustness_eval.py:85-87 and kws_adaptive_attack_eval.py:74-76).
//   frames (implicit, centre padded)  x  windowed DFT basis [n_fft x (cos | -sin)]   -> power = re^2 + im^2   (GEMM, fp32)
//   power [rows x n_freq]  x  mel filterbank [n_freq x n_mels]  -> 10 log10(max(., 1e-10))
// The basis and the filterbank are computed once on the host in float64 and rounded to fp32.
#include <cmath>

#include "ap_common.cuh"
#include "ap_internal.h"
#include "ap_sgemm.cuh"

namespace ap {

// A operand: row m = (waveform b, frame f); k = sample n of the frame.  Sample index i = f*hop + n - n_fft/2 with
// zero ('constant') or 'reflect' padding, exactly torch.stft(center=True).
struct FrameLoader {
  const float* wav;
  int L, frames, hop, pad, reflect;
  __device__ __forceinline__ float at(const float* w, int i) const {
    if (i < 0) {
      if (!reflect) return 0.f;
      i = -i;
    } else if (i >= L) {
      if (!reflect) return 0.f;
      i = 2 * (L - 1) - i;
    }
    return (i >= 0 && i < L) ? w[i] : 0.f;
  }
  __device__ __forceinline__ float4 load4(int, int m, int k, int M, int K) const {
    if (m >= M || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
    const int b = m / frames, f = m - b * frames;
    const float* w = wav + static_cast<long long>(b) * L;
    const int i = f * hop + k - pad;
    if (i >= 0 && i + 3 < L) return make_float4(w[i], w[i + 1], w[i + 2], w[i + 3]);
    return make_float4(at(w, i), at(w, i + 1), at(w, i + 2), at(w, i + 3));
  }
};
// column tile t = [cos bins 64t..64t+63 | -sin bins 64t..64t+63]
struct PowerEpi {
  float* power;
  int ld, n_freq;
  __device__ __forceinline__ void store(int, int m, int n0, int tx, const float (&lo)[4], const float (&hi)[4], int) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kbin = (n0 >> 1) + tx * 4 + j;
      if (kbin < n_freq) power[static_cast<long long>(m) * ld + kbin] = fmaf(lo[j], lo[j], hi[j] * hi[j]);
    }
  }
};

// one warp per (b, f) row; lane = mel bin (looped for n_mels > 32); out[b][mel][f]
__global__ void __launch_bounds__(256) mel_db_kernel(const float* __restrict__ power, int ld, int n_freq,
                                                     const float* __restrict__ fb, int n_mels, int frames, long long rows,
                                                     float* __restrict__ spec) {
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long m = warp; m < rows; m += nwarps) {
    const float* pr = power + m * ld;
    const long long b = m / frames;
    const int f = static_cast<int>(m - b * frames);
    for (int mel = lane; mel < n_mels; mel += 32) {
      float acc = 0.f;
      for (int k = 0; k < n_freq; ++k) acc = fmaf(pr[k], fb[k * n_mels + mel], acc);
      spec[(b * n_mels + mel) * frames + f] = 10.0f * log10f(fmaxf(acc, 1e-10f));
    }
  }
}

// ---- backward (VJP wrt the waveform)
// forward epilogue that keeps the spectrum: reim[m][col] in the basis' column order (tile t = [re 64t.. | im 64t..])
struct ReImEpi {
  float* reim;
  int ld;   // = ncols
  __device__ __forceinline__ void store(int, int m, int n0, int tx, const float (&lo)[4], const float (&hi)[4], int) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      reim[static_cast<long long>(m) * ld + n0 + tx * 4 + j] = lo[j];
      reim[static_cast<long long>(m) * ld + n0 + 64 + tx * 4 + j] = hi[j];
    }
  }
};
// one warp per (b, f) row: mel power from the kept spectrum, g_mel = g_spec * (10 / ln 10) / melpow (0 where the clamp at
// 1e-10 is active), g_P[k] = sum_mel g_mel fb[k][mel], and in place reim <- 2 * (re, im) * g_P = d loss / d (re, im)
__global__ void __launch_bounds__(256) mel_bwd_kernel(float* __restrict__ reim, int ld, int n_freq, const float* __restrict__ fb,
                                                      int n_mels, int frames, long long rows, const float* __restrict__ g_spec) {
  extern __shared__ float sm[];   // per warp: n_mels gradients
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* gm = sm + wib * n_mels;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long m = warp; m < rows; m += nwarps) {
    float* pr = reim + m * ld;
    const long long b = m / frames;
    const int f = static_cast<int>(m - b * frames);
    for (int mel = lane; mel < n_mels; mel += 32) {
      float acc = 0.f;
      for (int k = 0; k < n_freq; ++k) {
        const int col = (k >> 6) * 128 + (k & 63);
        const float re = pr[col], im = pr[col + 64];
        acc = fmaf(fmaf(re, re, im * im), fb[k * n_mels + mel], acc);
      }
      gm[mel] = acc > 1e-10f ? g_spec[(b * n_mels + mel) * frames + f] * (4.342944819032518f / acc) : 0.f;
    }
    __syncwarp();
    for (int k = lane; k < n_freq; k += 32) {
      float gp = 0.f;
      for (int mel = 0; mel < n_mels; ++mel) gp = fmaf(gm[mel], fb[k * n_mels + mel], gp);
      const int col = (k >> 6) * 128 + (k & 63);
      pr[col] *= 2.f * gp;
      pr[col + 64] *= 2.f * gp;
    }
    __syncwarp();
  }
}
// A operand of the second GEMM: plain row-major rows (the padded columns of `reim` beyond n_freq hold zeros: basis columns
// there are zero, so the forward wrote zeros)
struct RowLoader {
  const float* a;
  int ld;
  __device__ __forceinline__ float4 load4(int, int m, int k, int M, int K) const {
    if (m >= M || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
    return *reinterpret_cast<const float4*>(a + static_cast<long long>(m) * ld + k);
  }
};
// g_frames[m][n] -> overlap-add into g_wav at i = f*hop + n - pad (reflect padding folds the mirrored samples back)
struct OverlapAddEpi {
  float* g_wav;
  int L, frames, hop, pad, reflect, n_fft;
  __device__ __forceinline__ void put(int m, int n, float v) const {
    if (n >= n_fft) return;
    const int b = m / frames, f = m - b * frames;
    int i = f * hop + n - pad;
    if (i < 0) {
      if (!reflect) return;
      i = -i;
    } else if (i >= L) {
      if (!reflect) return;
      i = 2 * (L - 1) - i;
    }
    if (i >= 0 && i < L) atomicAdd(g_wav + static_cast<long long>(b) * L + i, v);
  }
  __device__ __forceinline__ void store(int, int m, int n0, int tx, const float (&lo)[4], const float (&hi)[4], int) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      put(m, n0 + tx * 4 + j, lo[j]);
      put(m, n0 + 64 + tx * 4 + j, hi[j]);
    }
  }
};

}  // namespace ap

using namespace ap;

struct ap_mel_s {
  ap_mel_cfg cfg{};
  int device = 0;
  int n_freq = 0, ncols = 0, ld_power = 0;
  DevBuf basis, fb, power;
  long long power_rows = 0;
  // backward pass: transposed basis [ncols][n_fft padded to 128] and the kept spectrum
  std::vector<float> basis_host;
  DevBuf basis_t, reim;
  int ld_t = 0;
  long long reim_rows = 0;
  ap::MelTc* tc = nullptr;      // tensor-core forward (created on first eligible call)
  ~ap_mel_s() {
    if (tc) ap::mel_tc_destroy(tc);
  }
};

static double hz_to_mel(double f, bool slaney) {
  if (!slaney) return 2595.0 * std::log10(1.0 + f / 700.0);
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m, bool slaney) {
  if (!slaney) return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0);
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

extern "C" int ap_mel_create(ap_mel_t* out, const ap_mel_cfg* cfg, int device) {
  AP_REQUIRE(out && cfg, "ap_mel_create: null argument");
  *out = nullptr;
  AP_REQUIRE(cfg->n_fft >= 16 && cfg->n_fft % 4 == 0 && cfg->hop_length > 0 && cfg->n_mels > 0 && cfg->sample_rate > 0,
             "ap_mel_create: bad configuration (n_fft=%d hop=%d n_mels=%d)", cfg->n_fft, cfg->hop_length, cfg->n_mels);
  int rc = select_device(device);
  if (rc != AP_OK) return rc;
  auto* h = new ap_mel_s();
  h->cfg = *cfg;
  h->device = device;
  const int N = cfg->n_fft, nf = N / 2 + 1;
  h->n_freq = nf;
  const int tiles = (nf + 63) / 64;
  h->ncols = tiles * 128;
  h->ld_power = tiles * 64;
  const double PI = 3.14159265358979323846;
  std::vector<float> basis(static_cast<size_t>(N) * h->ncols, 0.f);
  for (int n = 0; n < N; ++n) {
    const double win = 0.5 - 0.5 * std::cos(2.0 * PI * n / N);   // periodic hann (torch.hann_window default)
    for (int k = 0; k < nf; ++k) {
      const long long nk = (static_cast<long long>(n) * k) % N;  // exact argument reduction
      const double ang = 2.0 * PI * static_cast<double>(nk) / N;
      const int t = k / 64, kk = k % 64;
      basis[static_cast<size_t>(n) * h->ncols + t * 128 + kk] = static_cast<float>(win * std::cos(ang));
      basis[static_cast<size_t>(n) * h->ncols + t * 128 + 64 + kk] = static_cast<float>(-win * std::sin(ang));
    }
  }
  // torchaudio.functional.melscale_fbanks(n_freqs, f_min=0, f_max=sr/2, n_mels, sr, norm, mel_scale)
  const bool slaney = cfg->slaney_scale != 0;
  const double f_max = cfg->sample_rate / 2.0;
  std::vector<double> f_pts(cfg->n_mels + 2);
  const double m_lo = hz_to_mel(0.0, slaney), m_hi = hz_to_mel(f_max, slaney);
  for (int i = 0; i < cfg->n_mels + 2; ++i) f_pts[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (cfg->n_mels + 1), slaney);
  std::vector<float> fb(static_cast<size_t>(nf) * cfg->n_mels);
  for (int k = 0; k < nf; ++k) {
    const double f = (cfg->sample_rate / 2) * static_cast<double>(k) / (nf - 1);
    for (int m = 0; m < cfg->n_mels; ++m) {
      const double down = (f - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
      const double up = (f_pts[m + 2] - f) / (f_pts[m + 2] - f_pts[m + 1]);
      double v = std::fmax(0.0, std::fmin(down, up));
      if (cfg->slaney_norm) v *= 2.0 / (f_pts[m + 2] - f_pts[m]);
      fb[static_cast<size_t>(k) * cfg->n_mels + m] = static_cast<float>(v);
    }
  }
  h->basis_host = basis;
  cudaError_t e = h->basis.upload(basis.data(), basis.size() * sizeof(float));
  if (e == cudaSuccess) e = h->fb.upload(fb.data(), fb.size() * sizeof(float));
  if (e != cudaSuccess) {
    delete h;
    return fail(AP_ERR_CUDA, "ap_mel_create: %s", cudaGetErrorString(e));
  }
  *out = h;
  return AP_OK;
}

extern "C" void ap_mel_destroy(ap_mel_t h) { delete h; }

extern "C" int ap_mel_frames(ap_mel_t h, int L) {
  AP_REQUIRE(h && L > 0, "ap_mel_frames: bad arguments");
  return 1 + L / h->cfg.hop_length;
}

extern "C" int ap_mel_db(ap_mel_t h, const float* wav, float* spec, int B, int L, void* stream) {
  AP_REQUIRE(h && wav && spec, "ap_mel_db: null argument");
  AP_REQUIRE(B > 0 && L > 0, "ap_mel_db: B and L must be positive (got %d, %d)", B, L);
  AP_REQUIRE(!h->cfg.reflect_pad || L > h->cfg.n_fft / 2, "ap_mel_db: reflect padding needs L > n_fft/2");
  AP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int frames = 1 + L / h->cfg.hop_length;
  const long long rows = static_cast<long long>(B) * frames;
  AP_REQUIRE(rows < (1ll << 30), "ap_mel_db: batch too large");
  if (mel_tc_eligible(h->cfg, L)) {     // tcgen05 DFT GEMM with the power / filterbank / dB tail fused into its epilogue
    if (!h->tc) {
      int rc = mel_tc_create(&h->tc, h->cfg);
      if (rc != AP_OK) return rc;
    }
    return mel_tc_forward(h->tc, wav, h->fb.as<float>(), spec, B, L, st);
  }
  if (rows > h->power_rows) {
    h->power_rows = 0;   // alloc() releases the old buffer first
    AP_CUDA(h->power.alloc(static_cast<size_t>(rows) * h->ld_power * sizeof(float)));
    h->power_rows = rows;
  }
  FrameLoader al{wav, L, frames, h->cfg.hop_length, h->cfg.n_fft / 2, h->cfg.reflect_pad};
  PowerEpi ep{h->power.as<float>(), h->ld_power, h->n_freq};
  AP_CUDA(sgemm::launch(al, h->basis.as<float>(), h->ncols, 0, 1, static_cast<int>(rows), h->ncols, h->cfg.n_fft, ep, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  long long blocks = ceil_div_ll(rows * 32, 256);
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  mel_db_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(h->power.as<float>(), h->ld_power, h->n_freq,
                                                               h->fb.as<float>(), h->cfg.n_mels, frames, rows, spec);
  AP_LAUNCH_CHECK();
  return AP_OK;
}

// g_wav = (d spec / d wav)^T g_spec: backward of power spectrogram -> mel -> 10 log10 (what autograd computes through
// torchaudio's MelSpectrogram + AmplitudeToDB).  wav, g_wav: device (B, L); g_spec: device (B, n_mels, frames).
extern "C" int ap_mel_vjp(ap_mel_t h, const float* wav, const float* g_spec, float* g_wav, int B, int L, void* stream) {
  AP_REQUIRE(h && wav && g_spec && g_wav, "ap_mel_vjp: null argument");
  AP_REQUIRE(B > 0 && L > 0, "ap_mel_vjp: B and L must be positive (got %d, %d)", B, L);
  AP_REQUIRE(!h->cfg.reflect_pad || L > h->cfg.n_fft / 2, "ap_mel_vjp: reflect padding needs L > n_fft/2");
  AP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int N = h->cfg.n_fft, frames = 1 + L / h->cfg.hop_length;
  const long long rows = static_cast<long long>(B) * frames;
  AP_REQUIRE(rows < (1ll << 30), "ap_mel_vjp: batch too large");
  if (!h->basis_t.p) {   // [ncols][ld_t]: basis transposed, N contiguous and padded to a multiple of 128 columns
    h->ld_t = ((N + 127) / 128) * 128;
    std::vector<float> bt(static_cast<size_t>(h->ncols) * h->ld_t, 0.f);
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < h->ncols; ++c) bt[static_cast<size_t>(c) * h->ld_t + n] = h->basis_host[static_cast<size_t>(n) * h->ncols + c];
    AP_CUDA(h->basis_t.upload(bt.data(), bt.size() * sizeof(float)));
  }
  if (rows > h->reim_rows) {
    h->reim_rows = 0;
    AP_CUDA(h->reim.alloc(static_cast<size_t>(rows) * h->ncols * sizeof(float)));
    h->reim_rows = rows;
  }
  FrameLoader al{wav, L, frames, h->cfg.hop_length, N / 2, h->cfg.reflect_pad};
  ReImEpi ep{h->reim.as<float>(), h->ncols};
  AP_CUDA(sgemm::launch(al, h->basis.as<float>(), h->ncols, 0, 1, static_cast<int>(rows), h->ncols, N, ep, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  long long blocks = ceil_div_ll(rows * 32, 256);
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  mel_bwd_kernel<<<static_cast<unsigned>(blocks), 256, 8 * h->cfg.n_mels * sizeof(float), st>>>(
      h->reim.as<float>(), h->ncols, h->n_freq, h->fb.as<float>(), h->cfg.n_mels, frames, rows, g_spec);
  AP_LAUNCH_CHECK();
  AP_CUDA(cudaMemsetAsync(g_wav, 0, static_cast<size_t>(B) * L * sizeof(float), st));
  RowLoader gl{h->reim.as<float>(), h->ncols};
  OverlapAddEpi oe{g_wav, L, frames, h->cfg.hop_length, N / 2, h->cfg.reflect_pad, N};
  AP_CUDA(sgemm::launch(gl, h->basis_t.as<float>(), h->ld_t, 0, 1, static_cast<int>(rows), h->ld_t, h->ncols, oe, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return AP_OK;
}
