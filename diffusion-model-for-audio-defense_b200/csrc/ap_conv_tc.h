// Tensor-core (TF32) convolution used by the classifiers; see ap_conv_tc.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "ap_common.cuh"

namespace ap {

// generic cuTensorMapEncodeTiled wrapper (SWIZZLE_128B, zero OOB fill); dims / box innermost first; strides_bytes has
// rank-1 entries (dims 1..rank-1).  Implemented in ap_wavenet_tc.cu.
int tma_encode(CUtensorMap* m, CUtensorMapDataType dtype, const void* ptr, int rank, const uint64_t* dims,
               const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);

struct ConvTcParams {
  int m_tiles, n_tiles, groups, taps_h, taps_w, kblocks, Cg, Ng, NT;
  int stride, pad, bh, bb, tiles_per_img;
  int relu, has_res, round_out;
  const float* bias;
};
struct ConvTcBinding {   // one convolution bound to concrete activation buffers
  CUtensorMap tmA, tmOut, tmRes;
  ConvTcParams p;
};
int conv_tc_n_tile(int Ng);   // output channels per tensor-core tile for a group width, 0 if none
bool conv_tc_supported(int Cin, int Cout, int groups, int H, int W, int kh, int kw, int stride, int pad);

struct ConvTc {
  int Cin = 0, Cout = 0, kh = 1, kw = 1, stride = 1, pad = 0, groups = 1, Cg = 0, Ng = 0, NT = 0, K = 0;
  DevBuf w, bias;
  CUtensorMap tmW;
  // w_folded: torch layout [Cout][Cin/groups][kh][kw] with BatchNorm already folded; bias_folded [Cout]
  int init(int cin, int cout, int kh, int kw, int stride, int pad, int groups, const float* w_folded, const float* bias_folded);
  int bind(ConvTcBinding* b, const float* in, int B, int H, int W, float* out, const float* residual, int relu, int round_out) const;
  int run(const ConvTcBinding& b, cudaStream_t st) const;
};

}  // namespace ap
