// Internal (non-ABI) declarations shared between the translation units of libaudiopure_b200.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/audiopure.h"

namespace ap {

int num_sms();

// ap_update.cu
int diffuse(const float* x0, float a, float b, const float* z, uint64_t seed, uint64_t offset, float* xt, long long n,
            cudaStream_t st);
int ddpm_step(float* x, const float* eps, float c_eps, float sqrt_alpha, float sigma, const float* z, uint64_t seed,
              uint64_t offset, long long n, cudaStream_t st);
int sde_step(float* x, const float* eps, const ap_sde_coef& c, const float* z, uint64_t seed, uint64_t offset,
             long long n, cudaStream_t st);
int predict_x0(const float* xt, const float* eps, float a, float b, float* x0, long long n, cudaStream_t st);
int vote(const float* logits, int B, int K, long long* counts, int* pred, cudaStream_t st);

// Fused certification front end (RobustCertificate.smooth_predict + DiffWave.one_shot_denoise, certified_robust.py:44-54 and
// diffwave_ddpm.py:174-205): the network's first kernel builds the noisy copies x_in[b] = scale * (x + sigma * z[b]) of ONE
// input x (L) itself (and parks them in the output buffer), its last kernel turns eps into x0 = a * x_in - b * eps in place.
struct SmoothSrc {
  const float* x1;             // device (L): the input every row is a noisy copy of
  float sigma, scale;          // x_in = scale * (x1 + sigma * z)
  const float* z;              // device (B, L) host-generated N(0,1) noise, or null: Philox(seed) block offset + e / 4, lane e % 4
  uint64_t seed, offset;
  const uint64_t* offset_dev;  // optional device word added to `offset` (a captured CUDA graph replays with a moving offset)
  float a, b;                  // sqrt(1 / abar_t), sqrt(1 / abar_t - 1)
};

// ap_wavenet_tc.cu : the bf16 tensor-core (tcgen05 / TMEM / TMA) DiffWave network, C == S == 256 only.
//   weights: the ap_diffwave_create list (host fp32, weight-norm folded).
struct TcNet;
int tc_net_create(TcNet** out, const ap_wavenet_cfg& cfg, const float* const* weights);
void tc_net_destroy(TcNet* n);
void tc_net_set_dtype(TcNet* n, int dt);   // 0: bf16 operands, 1: fp16 operands (same kernels, same speed)
// (re)allocate the activation workspace for `chunk` waveforms of length L and encode the TMA tensor maps
int tc_net_reserve(TcNet* n, int chunk, int L);
size_t tc_net_workspace_bytes(const TcNet* n);
// eps[b, l] for b < B <= chunk.  ptab: device fp32 [num_layers + 1][256] step-embedding projections (row n = fc_t of
// layer n applied to the embedding; the extra last row is zero).  x, eps: device fp32 (B, L).
int tc_net_eps(TcNet* n, const float* x, const float* ptab, float* eps, int B, int L, cudaStream_t st);
// x0[b, :] for b < B <= chunk: the fused smoothing-input + one-shot-denoise form (x0 also serves as the x_in scratch)
int tc_net_smooth_denoise(TcNet* n, const SmoothSrc& src, const float* ptab, float* x0, int B, int L, cudaStream_t st);
// debug: run init + layers [0, layer]; returns u_{layer+1} and o_layer converted to fp32 (B, L, 256); either may be null
int tc_net_debug_layer(TcNet* n, const float* x, const float* ptab, int layer, float* u_next, float* gate, int B, int L,
                       cudaStream_t st);

// backward pass (bf16 mode): g_x = (d eps / d x)^T g_eps at (x, ptab) for B <= bchunk waveforms; the forward is recomputed
// with the activations the backward needs kept (tanh / sigmoid of every layer: ~1 KB per position per layer).
// eps_out may be null (eps_scratch, device fp32 (B, L), is then used for the recomputed eps).
int tc_net_vjp(TcNet* n, const float* x, const float* ptab, const float* g_eps, float* g_x, float* eps_out, float* eps_scratch,
               int B, int L, int bchunk, cudaStream_t st);
size_t tc_net_bwd_bytes_per_waveform(const TcNet* n, int L);
// the two halves of tc_net_vjp: a forward that keeps the backward's inputs (token = generation of the saved state) and the
// backward from that state
int tc_net_eps_save(TcNet* n, const float* x, const float* ptab, float* eps, int B, int L, int bchunk, cudaStream_t st,
                    unsigned long long* token);
bool tc_net_saved_state_is(const TcNet* n, unsigned long long token, int B, int L);
int tc_net_backward(TcNet* n, const float* x, const float* g_eps, float* g_x, int B, int L, cudaStream_t st);

// K4 on the tensor cores (ap_wavenet_tc.cu): windowed DFT as a tcgen05 GEMM (bf16 hi/lo split operands) with power -> mel ->
// dB fused into the epilogue.  Eligible: zero padding, 32 mels, hop % 64 == 0, n_fft a multiple (<= 8) of hop, n_fft/2 % 128 == 0,
// 32 frames per waveform (the SC09 front end: 2048 / 512 on 1 s clips).
struct MelTc;
bool mel_tc_eligible(const ap_mel_cfg& cfg, int L);
int mel_tc_create(MelTc** out, const ap_mel_cfg& cfg);
void mel_tc_destroy(MelTc* m);
// wav: device (B, L); fb: device [n_freq][32]; spec: device (B, 32, 32)
int mel_tc_forward(MelTc* m, const float* wav, const float* fb, float* spec, int B, int L, cudaStream_t st);

// per-launch CUDA-event timing of k1_layer ([0]) and k2_head ([1]); read synchronises on the recorded events
void tc_net_profile(TcNet* n, bool on);
int tc_net_debug_counters(TcNet* n, long long* host16x256);
int tc_net_profile_read(TcNet* n, double* ms, int* count);

}  // namespace ap
