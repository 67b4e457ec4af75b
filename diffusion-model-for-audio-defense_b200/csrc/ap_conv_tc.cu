// K5: 2-D convolution (+ folded BatchNorm bias, optional residual add, ReLU) as a tcgen05 implicit GEMM in TF32 -- the
// precision class the reference's cuDNN path runs the classifier in on Ampere+ GPUs (torch.backends.cudnn.allow_tf32).
//
//   rows (M = 128)  : 128 output pixels = a (W_out x bh x bb) box of the NHWC output
//   K               : taps x input channels of the group, in 32-channel (128-byte) K-blocks; the A tile of one tap is ONE
//                     4-D TMA box of the NHWC input at the shifted coordinate -- zero padding is TMA out-of-bounds fill,
//                     stride-2 convolutions use the tensor map's element strides
//   N (64..256)     : output channels of the group, weights [Cout][K] K-major fp32 (pre-rounded to tf32)
// Accumulators are double-buffered in TMEM (2 x 256 columns), so the epilogue of tile i overlaps the MMAs of tile i+1.
// The epilogue stages 32-channel slabs in shared memory (SWIZZLE_128B) and writes them with TMA stores; the residual
// slab is prefetched into the same buffer by TMA and updated in place.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue.
#include <cuda.h>

#include "ap_common.cuh"
#include "ap_conv_tc.h"
#include "ap_internal.h"
#include "ap_ptx.cuh"

namespace ap {
namespace convtc {

using namespace ptx;

constexpr int A_BYTES = 128 * 128;             // [128 pixels][32 fp32]
constexpr int B_BYTES_MAX = 256 * 128;         // [<=256 out channels][32 fp32]
constexpr int STAGE_BYTES = A_BYTES + B_BYTES_MAX;
constexpr int NSTAGE = 3;
constexpr int STG_OFF = NSTAGE * STAGE_BYTES;  // 4 staging slabs of [128][32] fp32: [half-group][ping-pong]
constexpr int BAR_OFF = STG_OFF + 4 * A_BYTES;
constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;
constexpr int NTHREADS = 320;
enum { BAR_FULL = 0, BAR_EMPTY = 3, BAR_ACC_FULL = 6, BAR_ACC_EMPTY = 8, BAR_RES = 10, BAR_COUNT = 14 };

__global__ void __launch_bounds__(NTHREADS, 1)
k_conv(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
       const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes, const ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + BAR_OFF;
  auto bar = [&](int i) { return bars + 8u * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + BAR_OFF + 8 * BAR_COUNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) mbar_init(bar(BAR_FULL + i), 1), mbar_init(bar(BAR_EMPTY + i), 1);
    for (int i = 0; i < 2; ++i) mbar_init(bar(BAR_ACC_FULL + i), 1), mbar_init(bar(BAR_ACC_EMPTY + i), 8);
    for (int i = 0; i < 4; ++i) mbar_init(bar(BAR_RES + i), 1);
    fence_barrier_init();
    prefetch_tmap(&tmA), prefetch_tmap(&tmW), prefetch_tmap(&tmOut), prefetch_tmap(&tmRes);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_per_m = p.n_tiles * p.groups, total = p.m_tiles * n_per_m;
  const int nk = p.taps_h * p.taps_w * p.kblocks;
  const uint32_t stage_bytes = A_BYTES + static_cast<uint32_t>(p.NT) * 128u;

  auto decode = [&](int tile, int& g, int& nt, int& b0, int& oh0) {
    const int m = tile / n_per_m, rem = tile - m * n_per_m;
    g = rem / p.n_tiles, nt = rem - g * p.n_tiles;
    if (p.bb == 1) b0 = m / p.tiles_per_img, oh0 = (m - b0 * p.tiles_per_img) * p.bh;
    else b0 = m * p.bb, oh0 = 0;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        int g, nt, b0, oh0;
        decode(tile, g, nt, b0, oh0);
        for (int r = 0; r < p.taps_h; ++r)
          for (int s = 0; s < p.taps_w; ++s)
            for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
              const uint32_t st = it % NSTAGE, ph = (it / NSTAGE) & 1;
              mbar_wait(bar(BAR_EMPTY + st), ph ^ 1, 51);
              mbar_expect_tx(bar(BAR_FULL + st), stage_bytes);
              tma_load_4d(base + st * STAGE_BYTES, &tmA, bar(BAR_FULL + st), g * p.Cg + kb * 32, s - p.pad,
                          oh0 * p.stride + r - p.pad, b0);
              tma_load_2d(base + st * STAGE_BYTES + A_BYTES, &tmW, bar(BAR_FULL + st), (r * p.taps_w + s) * p.Cg + kb * 32,
                          g * p.Ng + nt * p.NT);
            }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // the whole warp runs the loop (uniform control flow keeps descriptors in uniform registers); one elected lane issues
    const uint32_t idesc = umma_idesc_tf32_f32(128, p.NT);
    uint32_t it = 0, ti = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++ti) {
      const uint32_t r = ti & 1;
      mbar_wait(bar(BAR_ACC_EMPTY + r), ((ti >> 1) & 1) ^ 1, 52);
      tc_fence_after();
      for (int kblk = 0; kblk < nk; ++kblk, ++it) {
        const uint32_t st = it % NSTAGE, ph = (it / NSTAGE) & 1;
        mbar_wait(bar(BAR_FULL + st), ph, 53);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = umma_desc_k_sw128(base + st * STAGE_BYTES), bd = umma_desc_k_sw128(base + st * STAGE_BYTES + A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32(tmem + r * 256, ad + 2 * k, bd + 2 * k, idesc, (kblk == 0 && k == 0) ? 0u : 1u);
          umma_commit(bar(BAR_EMPTY + st));
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(bar(BAR_ACC_FULL + r));
      __syncwarp();
    }
  } else {
    const int q = warp & 3, hs = (warp - 2) >> 2;
    const int htid = (threadIdx.x - 64) & 127;                    // thread index within the half-group
    const bool lead = htid == 0;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t row_off = row * 128, sw = row & 7;
    const int nslab = p.NT / 32;
    uint32_t ti = 0, sc = 0;                                      // sc: slabs this half-group has processed
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++ti) {
      int g, nt, b0, oh0;
      decode(tile, g, nt, b0, oh0);
      const uint32_t r = ti & 1;
      const int col_base = g * p.Ng + nt * p.NT;                  // first output channel of this tile
      mbar_wait(bar(BAR_ACC_FULL + r), (ti >> 1) & 1, 54);
      tc_fence_after();
      if (p.has_res && hs < nslab && lead) {                      // residual slab of my first slab -> buffer sc & 1
        bulk_wait_read<1>();
        mbar_expect_tx(bar(BAR_RES + hs * 2 + (sc & 1)), A_BYTES);
        tma_load_4d(base + STG_OFF + (hs * 2 + (sc & 1)) * A_BYTES, &tmRes, bar(BAR_RES + hs * 2 + (sc & 1)), col_base + hs * 32, 0,
                    oh0, b0);
      }
      for (int slab = hs; slab < nslab; slab += 2, ++sc) {
        const uint32_t buf = sc & 1, stg = base + STG_OFF + (hs * 2 + buf) * A_BYTES;
        const int col0 = col_base + slab * 32;
        if (p.has_res) {
          if (slab + 2 < nslab && lead) {                         // prefetch the next residual slab into the other buffer
            bulk_wait_read<0>();
            mbar_expect_tx(bar(BAR_RES + hs * 2 + (buf ^ 1)), A_BYTES);
            tma_load_4d(base + STG_OFF + (hs * 2 + (buf ^ 1)) * A_BYTES, &tmRes, bar(BAR_RES + hs * 2 + (buf ^ 1)), col0 + 64, 0, oh0, b0);
          }
          mbar_wait(bar(BAR_RES + hs * 2 + buf), (sc >> 1) & 1, 55);
        } else {
          if (lead) bulk_wait_read<1>();
          named_bar_sync(1 + hs, 128);
        }
        uint32_t acc[32];
        tmem_ld_32x32b_x32(lane_addr + r * 256 + slab * 32, acc);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t addr = stg + row_off + ((static_cast<uint32_t>(i) ^ sw) << 4);
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i * 4));
          float v[4] = {__uint_as_float(acc[i * 4 + 0]) + bv.x, __uint_as_float(acc[i * 4 + 1]) + bv.y,
                        __uint_as_float(acc[i * 4 + 2]) + bv.z, __uint_as_float(acc[i * 4 + 3]) + bv.w};
          if (p.has_res) {
            const uint4 rv = ld_shared_v4(addr);
            v[0] += __uint_as_float(rv.x), v[1] += __uint_as_float(rv.y), v[2] += __uint_as_float(rv.z), v[3] += __uint_as_float(rv.w);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (p.relu) v[e] = fmaxf(v[e], 0.f);
            if (p.round_out) v[e] = round_tf32(v[e]);
          }
          st_shared_v4(addr, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + hs, 128);
        if (lead) {
          tma_store_4d(&tmOut, stg, col0, 0, oh0, b0);
          bulk_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BAR_ACC_EMPTY + r));
    }
    if (lead) bulk_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace convtc

// ------------------------------------------------------------------------------------------------ host side
static inline float host_round_tf32(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  u = (u + 0x00000fffu + ((u >> 13) & 1u)) & 0xffffe000u;   // round to nearest even at 10 mantissa bits
  float y;
  std::memcpy(&y, &u, 4);
  return y;
}

// Output channels per tile (the UMMA N): 256 where it divides the group width, the whole group for 64 / 128, and 160 for
// WideResNet's 160 / 320 / 640-channel layers (any multiple of 32 up to 256 is a legal N for M = 128; the epilogue works in
// 32-column slabs).  0: no tensor-core tile for this width.
int conv_tc_n_tile(int Ng) {
  if (Ng > 0 && Ng % 256 == 0) return 256;
  if (Ng == 64 || Ng == 128) return Ng;
  if (Ng > 0 && Ng % 160 == 0) return 160;
  if (Ng > 0 && Ng % 128 == 0) return 128;         // 384 (the UNet's concatenated inputs, as outputs of the data-gradient twins)
  return 0;
}

bool conv_tc_supported(int Cin, int Cout, int groups, int H, int W, int kh, int kw, int stride, int pad) {
  if (groups <= 0 || Cin % groups || Cout % groups) return false;
  const int Cg = Cin / groups, Ng = Cout / groups;
  if (Cg % 32 || conv_tc_n_tile(Ng) == 0) return false;
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  if (Wo <= 0 || Ho <= 0 || Wo > 128 || 128 % Wo) return false;
  const int rows = 128 / Wo;                       // output rows per 128-pixel tile
  if (Ho >= rows) return Ho % rows == 0;
  return rows % Ho == 0;                           // several whole images per tile
}

int ConvTc::init(int cin, int cout, int kh_, int kw_, int stride_, int pad_, int groups_, const float* w_folded /*[Cout][Cg][kh][kw]*/,
                 const float* bias_folded) {
  Cin = cin, Cout = cout, kh = kh_, kw = kw_, stride = stride_, pad = pad_, groups = groups_;
  Cg = cin / groups, Ng = cout / groups, NT = conv_tc_n_tile(Ng), K = kh * kw * Cg;
  if (NT == 0 || Cg % 32) return fail(AP_ERR_INVALID, "ConvTc: no tensor-core tile for %d -> %d channels per group", Cg, Ng);
  std::vector<float> wp(static_cast<size_t>(cout) * K);
  for (int o = 0; o < cout; ++o)
    for (int c = 0; c < Cg; ++c)
      for (int r = 0; r < kh; ++r)
        for (int s = 0; s < kw; ++s)
          wp[static_cast<size_t>(o) * K + (r * kw + s) * Cg + c] = host_round_tf32(w_folded[((static_cast<size_t>(o) * Cg + c) * kh + r) * kw + s]);
  AP_CUDA(w.upload(wp.data(), wp.size() * sizeof(float)));
  AP_CUDA(bias.upload(bias_folded, sizeof(float) * cout));
  const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(cout)};
  const uint64_t strides[1] = {static_cast<uint64_t>(K) * 4};
  const uint32_t box[2] = {32, static_cast<uint32_t>(NT)}, es[2] = {1, 1};
  return tma_encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, w.p, 2, dims, strides, box, es);
}

// NHWC fp32 activation tensor map: box = (32 channels, bw, bh, bb) pixels, element strides (1, sw, sh, 1)
static int encode_act(CUtensorMap* m, const float* ptr, int B, int H, int W, int Cc, int bw, int bh, int bb, int stride) {
  const uint64_t dims[4] = {static_cast<uint64_t>(Cc), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(Cc) * 4, static_cast<uint64_t>(W) * Cc * 4, static_cast<uint64_t>(H) * W * Cc * 4};
  const uint32_t box[4] = {32, static_cast<uint32_t>(bw * stride), static_cast<uint32_t>(bh * stride), static_cast<uint32_t>(bb)};
  const uint32_t es[4] = {1, static_cast<uint32_t>(stride), static_cast<uint32_t>(stride), 1};
  return tma_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, ptr, 4, dims, strides, box, es);
}

int ConvTc::bind(ConvTcBinding* bnd, const float* in, int B, int H, int W, float* out, const float* residual, int relu,
                 int round_out) const {
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  const int rows = 128 / Wo;
  const int bh = Ho >= rows ? rows : Ho, bb = Ho >= rows ? 1 : rows / Ho;
  ConvTcParams& p = bnd->p;
  p.tiles_per_img = bb == 1 ? Ho / bh : 0;
  p.m_tiles = bb == 1 ? B * (Ho / bh) : (B + bb - 1) / bb;
  p.n_tiles = Ng / NT, p.groups = groups, p.taps_h = kh, p.taps_w = kw, p.kblocks = Cg / 32, p.Cg = Cg, p.Ng = Ng, p.NT = NT;
  p.stride = stride, p.pad = pad, p.bh = bh, p.bb = bb, p.relu = relu, p.has_res = residual != nullptr, p.round_out = round_out;
  p.bias = bias.as<float>();
  int rc = encode_act(&bnd->tmA, in, B, H, W, Cin, Wo, bh, bb, stride);
  if (rc == AP_OK) rc = encode_act(&bnd->tmOut, out, B, Ho, Wo, Cout, Wo, bh, bb, 1);
  if (rc == AP_OK) rc = encode_act(&bnd->tmRes, residual ? residual : out, B, Ho, Wo, Cout, Wo, bh, bb, 1);
  return rc;
}

int ConvTc::run(const ConvTcBinding& bnd, cudaStream_t st) const {
  static bool attr = false;
  if (!attr) {
    AP_CUDA(cudaFuncSetAttribute(convtc::k_conv, cudaFuncAttributeMaxDynamicSharedMemorySize, convtc::SMEM_BYTES));
    attr = true;
  }
  const int total = bnd.p.m_tiles * bnd.p.n_tiles * bnd.p.groups;
  const int grid = total < num_sms() ? total : num_sms();
  convtc::k_conv<<<grid, convtc::NTHREADS, convtc::SMEM_BYTES, st>>>(bnd.tmA, tmW, bnd.tmOut, bnd.tmRes, bnd.p);
  AP_LAUNCH_CHECK();
  return AP_OK;
}

}  // namespace ap
