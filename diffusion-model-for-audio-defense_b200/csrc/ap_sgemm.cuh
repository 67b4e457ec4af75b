// fp32 FFMA implicit-GEMM template (parity mode of the DiffWave network, mel DFT, classifier convolutions).
//   C[M x N] = A[M x K] * Bm[K x N]      A is produced by an ALoad functor (implicit im2col), Bm is a packed weight
// matrix with N contiguous and padded to a multiple of 128 columns; the epilogue functor receives, per output row,
// the thread's 4 "lo" columns (n0 + tx*4 + j) and 4 "hi" columns (n0 + 64 + tx*4 + j), which lets gate / power
// epilogues pair column j of the first half-tile with column j of the second.
// 128x128x16 tiles, 256 threads, 8x8 register tile per thread, register-staged double buffering.
#pragma once
#include <cuda_runtime.h>

namespace ap { namespace sgemm {

constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256, PAD = 4;

template <class ALoad, class Epi>
__global__ void __launch_bounds__(THREADS, 2)
kernel(ALoad aload, const float* __restrict__ Bm, int ldb, long long b_group_stride, int M, int N, int K, Epi epi) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM, z = blockIdx.z;
  const float* Bg = Bm + static_cast<long long>(z) * b_group_stride;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int idx = t + THREADS * j;
      const int row = idx >> 2, kq = idx & 3;
      ra[j] = aload.load4(z, m0 + row, k0 + kq * 4, M, K);
      const int krow = idx >> 5, nq = idx & 31;
      const int k = k0 + krow;
      rb[j] = (k < K) ? *reinterpret_cast<const float4*>(Bg + static_cast<long long>(k) * ldb + n0 + nq * 4)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int idx = t + THREADS * j;
      const int row = idx >> 2, kq = idx & 3;
      As[buf][kq * 4 + 0][row] = ra[j].x;
      As[buf][kq * 4 + 1][row] = ra[j].y;
      As[buf][kq * 4 + 2][row] = ra[j].z;
      As[buf][kq * 4 + 3][row] = ra[j].w;
      const int krow = idx >> 5, nq = idx & 31;
      *reinterpret_cast<float4*>(&Bs[buf][krow][nq * 4]) = rb[j];
    }
  };

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m < M) {
      const float lo[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
      const float hi[4] = {acc[i][4], acc[i][5], acc[i][6], acc[i][7]};
      epi.store(z, m, n0, tx, lo, hi, N);
    }
  }
}

// Narrow-N variant for N <= 16 (DenseNet's 48 -> 12 growth convolutions, the 1-channel data gradients of the stems): the
// 128-wide tile above would spend >= 7/8 of its FFMAs on padding columns.  128 x 16 x 16 tiles, one column and 8 rows per thread;
// the epilogue functor's per-element `put` is used.
template <class ALoad, class Epi>
__global__ void __launch_bounds__(THREADS, 2)
kernel_n16(ALoad aload, const float* __restrict__ Bm, int ldb, long long b_group_stride, int M, int N, int K, Epi epi) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][16];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * BM, z = blockIdx.z;
  const float* Bg = Bm + static_cast<long long>(z) * b_group_stride;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  float4 ra[2], rb = make_float4(0.f, 0.f, 0.f, 0.f);
  auto gload = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int idx = t + THREADS * j;
      ra[j] = aload.load4(z, m0 + (idx >> 2), k0 + (idx & 3) * 4, M, K);
    }
    if (t < 64) {
      const int k = k0 + (t >> 2);
      rb = (k < K) ? *reinterpret_cast<const float4*>(Bg + static_cast<long long>(k) * ldb + (t & 3) * 4)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int idx = t + THREADS * j;
      const int row = idx >> 2, kq = idx & 3;
      As[buf][kq * 4 + 0][row] = ra[j].x;
      As[buf][kq * 4 + 1][row] = ra[j].y;
      As[buf][kq * 4 + 2][row] = ra[j].z;
      As[buf][kq * 4 + 3][row] = ra[j].w;
    }
    if (t < 64) *reinterpret_cast<float4*>(&Bs[buf][t >> 2][(t & 3) * 4]) = rb;
  };
  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float b = Bs[buf][k][tx];
      acc[0] = fmaf(a0.x, b, acc[0]), acc[1] = fmaf(a0.y, b, acc[1]), acc[2] = fmaf(a0.z, b, acc[2]), acc[3] = fmaf(a0.w, b, acc[3]);
      acc[4] = fmaf(a1.x, b, acc[4]), acc[5] = fmaf(a1.y, b, acc[5]), acc[6] = fmaf(a1.z, b, acc[6]), acc[7] = fmaf(a1.w, b, acc[7]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  if (tx < N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      if (m < M) epi.put(z, m, tx, acc[i]);
    }
  }
}

template <class ALoad, class Epi>
inline cudaError_t launch_n16(const ALoad& aload, const float* Bm, int ldb, long long b_group_stride, int groups, int M, int N,
                              int K, const Epi& epi, cudaStream_t st) {
  dim3 grid(1, (M + BM - 1) / BM, groups);
  kernel_n16<ALoad, Epi><<<grid, THREADS, 0, st>>>(aload, Bm, ldb, b_group_stride, M, N, K, epi);
  return cudaGetLastError();
}

template <class ALoad, class Epi>
inline cudaError_t launch(const ALoad& aload, const float* Bm, int ldb, long long b_group_stride, int groups, int M, int N,
                          int K, const Epi& epi, cudaStream_t st) {
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, groups);
  kernel<ALoad, Epi><<<grid, THREADS, 0, st>>>(aload, Bm, ldb, b_group_stride, M, N, K, epi);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- A loaders
// Channels-last 1-D activations in[(b*L + l)*C + c]; k = tap*C + c reads position l + (tap - (taps-1)/2)*dilation,
// zero outside [0, L) (Conv1d padding = dilation*(k-1)/2, WaveNet.py:26).  C % 4 == 0.
struct Conv1dTaps {
  const float* in;
  int L, C, taps, dilation;
  __device__ __forceinline__ float4 load4(int, int m, int k, int M, int K) const {
    if (m >= M || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
    const int tap = k / C, c = k - tap * C;
    const int b = m / L, l = m - b * L;
    const int ls = l + (tap - (taps - 1) / 2) * dilation;
    if (ls < 0 || ls >= L) return make_float4(0.f, 0.f, 0.f, 0.f);
    return *reinterpret_cast<const float4*>(in + (static_cast<long long>(b) * L + ls) * C + c);
  }
};

}}  // namespace ap::sgemm
