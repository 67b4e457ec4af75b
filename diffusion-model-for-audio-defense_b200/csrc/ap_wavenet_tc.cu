// bf16 tensor-core DiffWave network for sm_100a: tcgen05.mma with TMEM accumulators, operands staged by TMA.
//
// Data layout in HBM (per chunk of waveforms):
//   U0/U1  bf16 [chunk][L][256]      layer input u_n = h_n + fc_t_n(emb)   (ping-pong; channels-last == K-major rows)
//   O      bf16 [N][chunk][L][256]   gate outputs o_n = tanh(.) * sigmoid(.) of every layer
//   weights bf16, K-major: Wd [N][2][256][768] (chunk j: rows 0..127 tanh ch 128j.., rows 128..255 sigmoid ch 128j..),
//                          Wr [N][256][256], Ws [N][256][256], Wf [256][256]
//
// Kernels
//   k1_layer   one residual block (WaveNet.py:75-97) per launch, persistent CTAs over 128-position tiles:
//              GEMM-1  a[128 x 512] = sum_{tap} U[l + (tap-1) d] . Wd      (K = 768, two N = 256 chunks, TMEM ping-pong)
//              epilogue-1  o = tanh(a_t + b) * sigmoid(a_s + b)  -> bf16 -> smem (A operand of GEMM-2) and TMA store to O
//              GEMM-2  r[128 x 256] = o . Wr                                (K = 256)
//              epilogue-2  u' = (u + r + b) * sqrt(.5) + p_next  -> bf16 -> TMA store
//   k2_head    the skip path of ALL layers and the head (WaveNet.py:133-135,160-162):
//              s[128 x 256] = sum_n O_n . Ws_n   (K = 36 * 256) ; y = relu((s + sum b) * sqrt(1/N) . Wf + b) ; eps = w2 . y + b2
//              (skip_total = sum_n skip_n is linear in o_n, so deferring it removes the fp32 skip read-modify-write
//               of 32.8 MB per layer per waveform from the layer kernel.)
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue.
#include <cmath>
#include <cstdlib>
#include <cuda.h>
#include <cuda_fp16.h>

#include "ap_common.cuh"
#include "ap_conv_tc.h"
#include "ap_internal.h"
#include "ap_ptx.cuh"

namespace ap {
namespace tc {

using namespace ptx;

constexpr int C = 256;                      // channels (res == skip)
constexpr int TILE_M = 128;                 // positions per tile (per CTA)
constexpr int A_BYTES = TILE_M * 128;       // [128 rows][64 bf16]  SWIZZLE_128B
constexpr int OUT_BYTES = 4 * A_BYTES;      // 4 K-blocks of [128][64] bf16 (gate output / staging)
constexpr int NTHREADS = 320;
constexpr int EPI_THREADS = 256;

// CG = 1: one CTA per tile, tcgen05 cta_group::1 (M = 128).
// CG = 2: a CTA pair (cluster of 2) works on two adjacent tiles with cta_group::2 (M = 256): each CTA stages its own 128
//         activation rows and HALF of every weight tile, so both the TMA fill and the MMA operand reads of shared memory
//         drop from 192 to 128 B/clk per SM (shared memory delivers 128 B/clk, which capped CG = 1 at ~63 % tensor duty).
// KIND 1 = k1_layer (biases live in registers: shared memory goes to a deeper TMA ring), KIND 2 = k2_head (1024 floats of
// bias / partial-dot scratch in shared memory).
template <int CG_, int KIND, int DT_ = 0> struct Geo {
  static constexpr int CG = CG_;
  static constexpr int DT = DT_;   // 0: bf16 operands, 1: fp16 operands
  static constexpr int B_ROWS = 256 / CG;
  static constexpr int B_BYTES = B_ROWS * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NSTAGE = CG == 1 ? 3 : (KIND == 1 ? 5 : 4);
  static constexpr int OUT_OFF = NSTAGE * STAGE_BYTES;
  static constexpr int BIAS_OFF = OUT_OFF + OUT_BYTES;
  static constexpr int BAR_OFF = BIAS_OFF + (KIND == 1 ? 0 : 4096);
  static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;   // + slack to align the base to 1024 B
  static constexpr uint32_t IDESC = DT_ == 0 ? umma_idesc_bf16_f32(128 * CG, 256) : umma_idesc_f16_f32(128 * CG, 256);
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
  static_assert(NSTAGE <= 5, "barrier table holds at most 5 stages");
};
enum { BAR_FULL = 0, BAR_EMPTY = 5, BAR_ACC_FULL = 10, BAR_ACC_EMPTY = 12, BAR_OUT_READY = 14, BAR_UC_FULL = 16, BAR_COUNT = 17 };

// Per-CTA view of the barrier array and the pair topology
template <class G> struct Ctx {
  static constexpr int CG = G::CG;
  uint32_t base, bars, lbars, rank;
  __device__ __forceinline__ uint32_t bar(int i) const { return bars + 8u * i; }    // local barrier (shared::cta)
  __device__ __forceinline__ uint32_t lbar(int i) const { return lbars + 8u * i; }  // leader's barrier (shared::cluster)
  __device__ __forceinline__ uint32_t stage_a(uint32_t s) const { return base + s * G::STAGE_BYTES; }
  __device__ __forceinline__ uint32_t stage_b(uint32_t s) const { return base + s * G::STAGE_BYTES + A_BYTES; }
  __device__ __forceinline__ uint32_t out_kb(int kb) const { return base + G::OUT_OFF + kb * A_BYTES; }
  // consumer -> MMA issuer signals (the issuer lives in the leader CTA)
  __device__ __forceinline__ void arrive_leader(int i) const {
    if (CG == 1) mbar_arrive(bar(i));
    else mbar_arrive_cluster(lbar(i));
  }
  // producer: arm the full barrier of stage s for `bytes` per CTA, after the slot was released
  __device__ __forceinline__ long long arm(uint32_t s, uint32_t ph, uint32_t bytes, int tag) const {
    const long long w = mbar_wait(bar(BAR_EMPTY + s), ph ^ 1, tag);
    if (CG == 1) mbar_expect_tx(bar(BAR_FULL + s), bytes);
    else if (rank == 0) mbar_expect_tx(bar(BAR_FULL + s), 2 * bytes);
    else mbar_arrive_cluster(lbar(BAR_FULL + s));
    return w;
  }
  __device__ __forceinline__ void load_a(uint32_t s, const CUtensorMap* m, int c0, int c1, int c2) const {
    if (CG == 1) tma_load_3d(stage_a(s), m, bar(BAR_FULL + s), c0, c1, c2);
    else tma_load_3d_pair(stage_a(s), m, lbar(BAR_FULL + s), c0, c1, c2);
  }
  // weight tile: this CTA's B_ROWS rows starting at row0 (+ rank * B_ROWS)
  __device__ __forceinline__ void load_b(uint32_t s, const CUtensorMap* m, int k0, int row0) const {
    if (CG == 1) tma_load_2d(stage_b(s), m, bar(BAR_FULL + s), k0, row0);
    else tma_load_2d_pair(stage_b(s), m, lbar(BAR_FULL + s), k0, row0 + static_cast<int>(rank) * G::B_ROWS);
  }
  // issue the 4 MMAs (K = 16 each) of one 64-wide K-block
  __device__ __forceinline__ void mma_kblock(uint32_t d_tmem, uint32_t a_smem, uint32_t b_smem, bool first) const {
    const uint64_t ad = umma_desc_k_sw128(a_smem), bd = umma_desc_k_sw128(b_smem);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (CG == 1) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, G::IDESC, (first && k == 0) ? 0u : 1u);
      else umma_bf16_pair(d_tmem, ad + 2 * k, bd + 2 * k, G::IDESC, (first && k == 0) ? 0u : 1u);
    }
  }
  __device__ __forceinline__ void commit(int i) const {   // MMA issuer -> barrier i of every CTA of the pair
    if (CG == 1) umma_commit(bar(i));
    else umma_commit_pair(bar(i), 3);
  }
};

// common prologue: barrier init, TMEM allocation, cluster handshake.  Returns the TMEM base address.
template <class G> __device__ __forceinline__ uint32_t tc_prologue(Ctx<G>& cx, uint8_t*& gen, int out_ready_count) {
  constexpr int CG = G::CG;
  extern __shared__ uint8_t smem_raw[];
  cx.base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  gen = smem_raw + (cx.base - smem_u32(smem_raw));
  cx.bars = cx.base + G::BAR_OFF;
  cx.rank = CG == 1 ? 0u : cluster_ctarank();
  cx.lbars = CG == 1 ? cx.bars : mapa_cluster(cx.bars, 0);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + G::BAR_OFF + 8 * BAR_COUNT);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < G::NSTAGE; ++i) mbar_init(cx.bar(BAR_FULL + i), CG), mbar_init(cx.bar(BAR_EMPTY + i), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(cx.bar(BAR_ACC_FULL + i), 1);
      mbar_init(cx.bar(BAR_ACC_EMPTY + i), 8 * CG);
      mbar_init(cx.bar(BAR_OUT_READY + i), out_ready_count * CG);
    }
    mbar_init(cx.bar(BAR_UC_FULL), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512), tmem_relinquish();
    else tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512), tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before any remote arrive / complete_tx
  tc_fence_after();
  return *tmem_slot;
}
template <int CG> __device__ __forceinline__ void tc_epilogue_teardown(uint32_t tmem) {
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();   // nobody exits (or frees TMEM) while the peer may still touch its smem / TMEM
  if ((threadIdx.x >> 5) == 1) {
    if (CG == 1) tmem_dealloc(tmem, 512);
    else tmem_dealloc_pair(tmem, 512);
  }
}
// tile enumeration: unit u = 0, 1, ... of this CTA -> (tile, valid)
template <int CG> struct Tiles {
  int n_tiles, first, stride;
  __device__ __forceinline__ Tiles(int n, uint32_t rank) : n_tiles(n) {
    if (CG == 1) first = blockIdx.x, stride = gridDim.x;
    else first = 2 * (blockIdx.x >> 1) + static_cast<int>(rank), stride = 2 * (gridDim.x >> 1);
  }
  __device__ __forceinline__ bool more(int tile) const {            // the PAIR still has work
    return CG == 1 ? tile < n_tiles : (tile & ~1) < n_tiles;
  }
};

struct K1Params {
  int n_tiles, tiles_per_sample, dilation, layer, chunk_alloc, last;
  const float* b_dil;   // [2][256] packed chunk order
  const float* b_res;   // [256]
  const float* p_next;  // [256]
  long long* dbg;       // optional [gridDim.x][16] wait-cycle counters (development aid), or null
};

template <int CG, int DT>
__global__ void __launch_bounds__(NTHREADS, 1)
k1_layer(const __grid_constant__ CUtensorMap tmUin, const __grid_constant__ CUtensorMap tmUout,
         const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmWd,
         const __grid_constant__ CUtensorMap tmWr, const K1Params p) {
  using G = Geo<CG, 1, DT>;
  Ctx<G> cx;
  uint8_t* gen;
  const uint32_t tmem = tc_prologue<G>(cx, gen, 1);
  if (threadIdx.x == 0)
    prefetch_tmap(&tmUin), prefetch_tmap(&tmUout), prefetch_tmap(&tmO), prefetch_tmap(&tmWd), prefetch_tmap(&tmWr);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Tiles<CG> tiles(p.n_tiles, cx.rank);
  const int oob_l0 = p.tiles_per_sample * TILE_M;   // a row coordinate past the end: TMA zero-fills the whole box

  if (warp == 0) {
    // ======================================================================================= TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      long long w_empty = 0;
      const long long t_start = clock64();
      for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride) {
        const bool valid = tile < p.n_tiles;
        const int b = valid ? tile / p.tiles_per_sample : 0;
        const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
        for (int j = 0; j < 2; ++j)
          for (int tap = 0; tap < 3; ++tap)
            for (int kb = 0; kb < 4; ++kb, ++it) {
              const uint32_t s = it % G::NSTAGE, ph = (it / G::NSTAGE) & 1;
              w_empty += cx.arm(s, ph, G::STAGE_BYTES, 1);
              cx.load_a(s, &tmUin, kb * 64, valid ? l0 + (tap - 1) * p.dilation : oob_l0, b);
              cx.load_b(s, &tmWd, tap * C + kb * 64, (p.layer * 2 + j) * 256);
            }
        if (!p.last)
          for (int kb = 0; kb < 4; ++kb, ++it) {
            const uint32_t s = it % G::NSTAGE, ph = (it / G::NSTAGE) & 1;
            w_empty += cx.arm(s, ph, G::B_BYTES, 2);
            cx.load_b(s, &tmWr, kb * 64, p.layer * 256);
          }
      }
      if (p.dbg) p.dbg[blockIdx.x * 16 + 0] = w_empty, p.dbg[blockIdx.x * 16 + 1] = clock64() - t_start;
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================================================================================= MMA issuer (leader CTA)
    if (lane == 0 && cx.rank == 0) {
      uint32_t it = 0, g = 0, ti = 0;
      long long w_full = 0, w_acc = 0, w_out = 0;
      const long long t_start = clock64();
      for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
        for (int j = 0; j < 2; ++j, ++g) {
          const uint32_t r = g & 1;
          w_acc += mbar_wait(cx.bar(BAR_ACC_EMPTY + r), ((g >> 1) & 1) ^ 1, 3);
          tc_fence_after();
          for (int kblk = 0; kblk < 12; ++kblk, ++it) {
            const uint32_t s = it % G::NSTAGE, ph = (it / G::NSTAGE) & 1;
            w_full += mbar_wait(cx.bar(BAR_FULL + s), ph, 4);
            tc_fence_after();
            cx.mma_kblock(tmem + r * 256, cx.stage_a(s), cx.stage_b(s), kblk == 0);
            cx.commit(BAR_EMPTY + s);
          }
          cx.commit(BAR_ACC_FULL + r);
        }
        if (!p.last) {
          const uint32_t r = g & 1;
          w_acc += mbar_wait(cx.bar(BAR_ACC_EMPTY + r), ((g >> 1) & 1) ^ 1, 5);
          for (int kb = 0; kb < 4; ++kb, ++it) {
            const uint32_t s = it % G::NSTAGE, ph = (it / G::NSTAGE) & 1;
            if (kb == 0) w_out += mbar_wait(cx.bar(BAR_OUT_READY + 0), ti & 1, 6);
            if (kb == 2) w_out += mbar_wait(cx.bar(BAR_OUT_READY + 1), ti & 1, 7);
            w_full += mbar_wait(cx.bar(BAR_FULL + s), ph, 8);
            tc_fence_after();
            cx.mma_kblock(tmem + r * 256, cx.out_kb(kb), cx.stage_b(s), kb == 0);
            cx.commit(BAR_EMPTY + s);
          }
          cx.commit(BAR_ACC_FULL + r);
          ++g;
        }
      }
      if (p.dbg) {
        long long* d = p.dbg + blockIdx.x * 16;
        d[2] = w_full, d[3] = w_acc, d[4] = w_out, d[5] = clock64() - t_start, d[6] = ti;
      }
    }
    __syncwarp();
  } else {
    // ======================================================================================= epilogue (8 warps)
    const int q = warp & 3, hsel = (warp - 2) >> 2, etid = threadIdx.x - 64;
    const int row = q * 32 + lane;                       // position within the tile == TMEM lane
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t row_off = row * 128, sw = row & 7;
    const float sqrt_half = 0.70710678118654752440f;
    // per-channel constants live in registers, one channel per lane, and are broadcast with warp shuffles:
    //   bd_t/bd_s[j][gq]: dilated-conv bias of tanh / sigmoid channel j*128 + hsel*64 + gq*32 + lane
    //   br/pn[gq]       : res-conv bias and next-layer step-embedding projection of channel hsel*128 + gq*32 + lane
    float bd_t[2][2], bd_s[2][2], br_r[4], pn_r[4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int gq = 0; gq < 2; ++gq) {
        bd_t[j][gq] = p.b_dil[j * 256 + hsel * 64 + gq * 32 + lane];
        bd_s[j][gq] = p.b_dil[j * 256 + 128 + hsel * 64 + gq * 32 + lane];
      }
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) br_r[gq] = p.b_res[hsel * 128 + gq * 32 + lane], pn_r[gq] = p.p_next[hsel * 128 + gq * 32 + lane];
    uint32_t g = 0, ti = 0;
    long long w_accfull = 0, w_uc = 0;
    const long long t_start = clock64();
    for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
      const bool valid = tile < p.n_tiles;
      const int b = valid ? tile / p.tiles_per_sample : 0;
      const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
      for (int j = 0; j < 2; ++j, ++g) {
        const uint32_t r = g & 1;
        w_accfull += mbar_wait(cx.bar(BAR_ACC_FULL + r), (g >> 1) & 1, 9);
        tc_fence_after();
        if (j == 0) {  // the staging region is about to be overwritten: previous TMA stores must have read it
          if (etid == 0) bulk_wait_read<0>();
          named_bar_sync(1, EPI_THREADS);
        }
        const uint32_t kb_base = cx.out_kb(2 * j + hsel) + row_off;
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
          const float bt_l = j == 0 ? bd_t[0][gq] : bd_t[1][gq], bs_l = j == 0 ? bd_s[0][gq] : bd_s[1][gq];
          uint32_t ta[32], sg[32];
          tmem_ld_32x32b_x32(lane_addr + r * 256 + hsel * 64 + gq * 32, ta);
          tmem_ld_32x32b_x32(lane_addr + r * 256 + 128 + hsel * 64 + gq * 32, sg);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c0 = i * 8 + 2 * e;
              const float a0 = __uint_as_float(ta[c0]) + __shfl_sync(0xffffffffu, bt_l, c0);
              const float a1 = __uint_as_float(ta[c0 + 1]) + __shfl_sync(0xffffffffu, bt_l, c0 + 1);
              const float s0 = __uint_as_float(sg[c0]) + __shfl_sync(0xffffffffu, bs_l, c0);
              const float s1 = __uint_as_float(sg[c0 + 1]) + __shfl_sync(0xffffffffu, bs_l, c0 + 1);
              pk[e] = pack2<DT>(tanh_approx(a0) * sigmoid_approx(s0), tanh_approx(a1) * sigmoid_approx(s1));
            }
            st_shared_v4(kb_base + (((gq * 4 + i) ^ sw) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + r);
        named_bar_sync(1, EPI_THREADS);
        if (etid == 0) {
          if (valid) {
            tma_store_3d(&tmO, cx.out_kb(2 * j), (2 * j) * 64, l0, p.layer * p.chunk_alloc + b);
            tma_store_3d(&tmO, cx.out_kb(2 * j + 1), (2 * j + 1) * 64, l0, p.layer * p.chunk_alloc + b);
            bulk_commit();
          }
          cx.arrive_leader(BAR_OUT_READY + j);
        }
      }
      if (!p.last) {
        const uint32_t r = g & 1;
        w_accfull += mbar_wait(cx.bar(BAR_ACC_FULL + r), (g >> 1) & 1, 10);   // GEMM-2 done: accumulators ready, `out` smem no longer read
        tc_fence_after();
        if (etid == 0) {
          bulk_wait_read<0>();                                    // the O stores have finished reading `out`
          mbar_expect_tx(cx.bar(BAR_UC_FULL), OUT_BYTES);
          for (int kb = 0; kb < 4; ++kb) tma_load_3d(cx.out_kb(kb), &tmUin, cx.bar(BAR_UC_FULL), kb * 64, l0, b);
        }
        w_uc += mbar_wait(cx.bar(BAR_UC_FULL), ti & 1, 11);
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const float br_l = br_r[gq], pn_l = pn_r[gq];
          uint32_t acc[32];
          tmem_ld_32x32b_x32(lane_addr + r * 256 + hsel * 128 + gq * 32, acc);
          tmem_ld_wait();
          const uint32_t kb_base = cx.out_kb(hsel * 2 + (gq >> 1)) + row_off;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t addr = kb_base + ((((gq & 1) * 4 + i) ^ sw) << 4);
            const uint4 uv = ld_shared_v4(addr);
            const uint32_t uw[4] = {uv.x, uv.y, uv.z, uv.w};
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c0 = i * 8 + 2 * e;
              const float h0 = (unpack_lo<DT>(uw[e]) + (__uint_as_float(acc[c0]) + __shfl_sync(0xffffffffu, br_l, c0))) * sqrt_half +
                               __shfl_sync(0xffffffffu, pn_l, c0);
              const float h1 = (unpack_hi<DT>(uw[e]) + (__uint_as_float(acc[c0 + 1]) + __shfl_sync(0xffffffffu, br_l, c0 + 1))) * sqrt_half +
                               __shfl_sync(0xffffffffu, pn_l, c0 + 1);
              pk[e] = pack2<DT>(h0, h1);
            }
            st_shared_v4(addr, make_uint4(pk[0], pk[1], pk[2], pk[3]));
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + r);
        named_bar_sync(1, EPI_THREADS);
        if (etid == 0 && valid) {
          for (int kb = 0; kb < 4; ++kb) tma_store_3d(&tmUout, cx.out_kb(kb), kb * 64, l0, b);
          bulk_commit();
        }
        ++g;
      }
    }
    if (etid == 0) bulk_wait_all<0>();
    if (p.dbg && etid == 0) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[7] = w_accfull, d[8] = w_uc, d[9] = clock64() - t_start;
    }
  }
  tc_epilogue_teardown<CG>(tmem);
}

// ------------------------------------------------------------------------------------------------ k2: skip sum + head
struct K2Params {
  int n_tiles, tiles_per_sample, L, num_layers, chunk_alloc;
  float scale;           // sqrt(1/N)
  const float* bskip;    // [256] sum over layers of the skip-conv biases
  const float* bf1;      // [256]
  const float* wf2;      // [256]
  const float* bf2;      // [1]
  float* eps;            // [B][L]
};
enum { BAR2_S_READY = BAR_OUT_READY };

template <int CG, int DT>
__global__ void __launch_bounds__(NTHREADS, 1)
k2_head(const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmWs,
        const __grid_constant__ CUtensorMap tmWf, const K2Params p) {
  using G = Geo<CG, 2, DT>;
  Ctx<G> cx;
  uint8_t* gen;
  const uint32_t tmem = tc_prologue<G>(cx, gen, 1);
  float* s_bias = reinterpret_cast<float*>(gen + G::BIAS_OFF);   // [0,256) bskip, [256,512) bf1, [512,768) wf2, [768,896) partial dots
  for (int i = threadIdx.x; i < 768; i += NTHREADS)
    s_bias[i] = i < 256 ? p.bskip[i] : (i < 512 ? p.bf1[i - 256] : p.wf2[i - 512]);
  if (threadIdx.x == 0) prefetch_tmap(&tmO), prefetch_tmap(&tmWs), prefetch_tmap(&tmWf);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = p.num_layers * 4;
  const Tiles<CG> tiles(p.n_tiles, cx.rank);
  const int oob_l0 = p.tiles_per_sample * TILE_M;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride) {
        const bool valid = tile < p.n_tiles;
        const int b = valid ? tile / p.tiles_per_sample : 0;
        const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
        for (int n = 0; n < p.num_layers; ++n)
          for (int kb = 0; kb < 4; ++kb, ++it) {
            const uint32_t s = it % G::NSTAGE, ph = (it / G::NSTAGE) & 1;
            cx.arm(s, ph, G::STAGE_BYTES, 21);
            cx.load_a(s, &tmO, kb * 64, l0, n * p.chunk_alloc + b);
            cx.load_b(s, &tmWs, kb * 64, n * 256);
          }
        for (int kb = 0; kb < 4; ++kb, ++it) {
          const uint32_t s = it % G::NSTAGE, ph = (it / G::NSTAGE) & 1;
          cx.arm(s, ph, G::B_BYTES, 22);
          cx.load_b(s, &tmWf, kb * 64, 0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && cx.rank == 0) {
      uint32_t it = 0, ti = 0;
      for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
        mbar_wait(cx.bar(BAR_ACC_EMPTY + 0), (ti & 1) ^ 1, 23);
        tc_fence_after();
        for (int kblk = 0; kblk < nkb; ++kblk, ++it) {
          const uint32_t s = it % G::NSTAGE, ph = (it / G::NSTAGE) & 1;
          mbar_wait(cx.bar(BAR_FULL + s), ph, 24);
          tc_fence_after();
          cx.mma_kblock(tmem, cx.stage_a(s), cx.stage_b(s), kblk == 0);
          cx.commit(BAR_EMPTY + s);
        }
        cx.commit(BAR_ACC_FULL + 0);
        mbar_wait(cx.bar(BAR_ACC_EMPTY + 1), (ti & 1) ^ 1, 25);
        mbar_wait(cx.bar(BAR2_S_READY), ti & 1, 26);
        tc_fence_after();
        for (int kb = 0; kb < 4; ++kb, ++it) {
          const uint32_t s = it % G::NSTAGE, ph = (it / G::NSTAGE) & 1;
          mbar_wait(cx.bar(BAR_FULL + s), ph, 27);
          tc_fence_after();
          cx.mma_kblock(tmem + 256, cx.out_kb(kb), cx.stage_b(s), kb == 0);
          cx.commit(BAR_EMPTY + s);
        }
        cx.commit(BAR_ACC_FULL + 1);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3, hsel = (warp - 2) >> 2, etid = threadIdx.x - 64;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t row_off = row * 128, sw = row & 7;
    float* s_part = s_bias + 768;
    uint32_t ti = 0;
    for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
      const bool valid = tile < p.n_tiles;
      const int b = valid ? tile / p.tiles_per_sample : 0;
      const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
      // ---- skip sum -> s (bf16, A operand of the head GEMM).  The previous tile's head MMAs finished reading the
      //      staging region before its ACC_FULL[1] fired, which this thread has already waited on.
      mbar_wait(cx.bar(BAR_ACC_FULL + 0), ti & 1, 28);
      tc_fence_after();
      const float* bs = s_bias + hsel * 128;
#pragma unroll 1
      for (int gq = 0; gq < 4; ++gq) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(lane_addr + hsel * 128 + gq * 32, acc);
        tmem_ld_wait();
        const uint32_t kb_base = cx.out_kb(hsel * 2 + (gq >> 1)) + row_off;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c0 = gq * 32 + i * 8 + 2 * e;
            pk[e] = pack2<DT>((__uint_as_float(acc[i * 8 + 2 * e]) + bs[c0]) * p.scale,
                                (__uint_as_float(acc[i * 8 + 2 * e + 1]) + bs[c0 + 1]) * p.scale);
          }
          st_shared_v4(kb_base + ((((gq & 1) * 4 + i) ^ sw) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + 0);
      named_bar_sync(1, EPI_THREADS);
      if (etid == 0) cx.arrive_leader(BAR2_S_READY);
      // ---- head: y = relu(acc + b) ; eps = w2 . y + b2
      mbar_wait(cx.bar(BAR_ACC_FULL + 1), ti & 1, 29);
      tc_fence_after();
      const float* bf = s_bias + 256 + hsel * 128;
      const float* w2 = s_bias + 512 + hsel * 128;
      float dot = 0.f;
#pragma unroll 1
      for (int gq = 0; gq < 4; ++gq) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(lane_addr + 256 + hsel * 128 + gq * 32, acc);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) dot = fmaf(fmaxf(__uint_as_float(acc[e]) + bf[gq * 32 + e], 0.f), w2[gq * 32 + e], dot);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + 1);
      if (hsel == 1) s_part[row] = dot;
      named_bar_sync(1, EPI_THREADS);
      if (hsel == 0 && valid && l0 + row < p.L) p.eps[static_cast<size_t>(b) * p.L + l0 + row] = dot + s_part[row] + p.bf2[0];
      named_bar_sync(1, EPI_THREADS);   // s_part is rewritten by the next tile
    }
  }
  tc_epilogue_teardown<CG>(tmem);
}

// ------------------------------------------------------------------------------------------------ init conv (bf16 out)
// u0[m][c] = bf16(max(w[c] x[m] + b[c], 0) + p0[c])      (WaveNet.py:147,13-19 and :84)
template <int DT>
__global__ void __launch_bounds__(256) init_h16_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, const float* __restrict__ p0,
                                                        uint4* __restrict__ u, long long M) {
  __shared__ float sw[C], sb[C], sp[C];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sw[i] = w[i], sb[i] = b[i], sp[i] = p0[i];
  __syncthreads();
  const long long total = M * (C / 8);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i >> 5;
    const int c = static_cast<int>(i & 31) * 8;
    const float xv = x[m];
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v0 = fmaxf(fmaf(sw[c + 2 * e], xv, sb[c + 2 * e]), 0.f) + sp[c + 2 * e];
      const float v1 = fmaxf(fmaf(sw[c + 2 * e + 1], xv, sb[c + 2 * e + 1]), 0.f) + sp[c + 2 * e + 1];
      pk[e] = pack2<DT>(v0, v1);
    }
    u[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// bf16 -> fp32 (debug dumps)
template <int DT>
__global__ void h16_to_f32_kernel(const uint16_t* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = unpack_lo<DT>(in[i]);
}

// ------------------------------------------------------------------------------------------------ self test kernel
// D[128 x 256] = A[128 x K] . B[256 x K]^T through the same TMA / UMMA / TMEM-load primitives (one CTA, one stage).
__global__ void __launch_bounds__(128, 1)
selftest_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ d,
                int K) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + Geo<1, 1>::STAGE_BYTES;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + Geo<1, 1>::STAGE_BYTES + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bars, 1), mbar_init(bars + 8, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  Ctx<Geo<1, 1>> cx;
  cx.base = base, cx.bars = bars, cx.lbars = bars, cx.rank = 0;
  if (threadIdx.x == 0) {
    for (int kb = 0; kb < K / 64; ++kb) {
      mbar_expect_tx(bars, Geo<1, 1>::STAGE_BYTES);
      tma_load_2d(base, &tmA, bars, kb * 64, 0);
      tma_load_2d(base + A_BYTES, &tmB, bars, kb * 64, 0);
      mbar_wait(bars, kb & 1, 40);
      tc_fence_after();
      cx.mma_kblock(tmem, base, base + A_BYTES, kb == 0);
      umma_commit(bars + 8);
      mbar_wait(bars + 8, kb & 1, 41);   // single stage: wait for the MMAs before refilling
    }
  }
  __syncthreads();
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < 256; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int e = 0; e < 32; ++e) d[row * 256 + c0 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// bf16 tensor, innermost dim contiguous; dims/box innermost first; SWIZZLE_128B (box[0] * 2 B == 128 B)
static int encode_bf16(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint32_t* box) {
  uint64_t strides[4];
  uint32_t es[5];
  uint64_t stride = 2;
  for (int i = 0; i < rank; ++i) {
    es[i] = 1;
    stride *= dims[i];
    if (i + 1 < rank) strides[i] = stride;
  }
  return tma_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, ptr, rank, dims, strides, box, es);
}

}  // namespace tc

int tma_encode(CUtensorMap* m, CUtensorMapDataType dtype, const void* ptr, int rank, const uint64_t* dims,
               const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  if (!fn) return fail(AP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5], gstride[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i], bx[i] = box[i], es[i] = elem_strides[i];
    if (i + 1 < rank) gstride[i] = strides_bytes[i];
  }
  CUresult r = fn(m, dtype, rank, const_cast<void*>(ptr), gdim, gstride, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return AP_OK;
}

// ================================================================================================ TcNet
struct TcNet {
  ap_wavenet_cfg cfg{};
  int N = 0;
  DevBuf wd, wr, ws, wf;                                   // bf16 operands
  DevBuf wd_h, wr_h, ws_h, wf_h;                           // fp16 operands (AP_MODE_FP16)
  int dt = 0;                                              // 0: bf16, 1: fp16
  DevBuf bd, br, bskip, bf1, wf2, bf2, init_w, init_b;     // fp32 vectors
  DevBuf u0, u1, o;
  int chunk = 0, L = 0;
  CUtensorMap tmU[2], tmO, tmWd, tmWr, tmWs, tmWf;   // weight maps: box of 256 rows (one CTA per tile)
  CUtensorMap tmWd2, tmWr2, tmWs2, tmWf2;           // weight maps: box of 128 rows (CTA pairs, each CTA stages half)
  CUtensorMap tmWd_h, tmWr_h, tmWs_h, tmWf_h, tmWd2_h, tmWr2_h, tmWs2_h, tmWf2_h;   // the same over the fp16 weights
  bool pair = true;                                 // cta_group::2 kernels (AP_TC_PAIR=0 selects the 1-CTA kernels)
  bool attr_set = false;
  // optional per-launch timing (bench.py roofline): CUDA events recorded on the launching stream around k1 / k2
  bool prof = false;
  DevBuf dbg;                       // wait-cycle counters of k1 launches of layer dbg_layer (AP_TC_DEBUG=1)
  int dbg_layer = 35;               // AP_TC_DEBUG_LAYER
  std::vector<cudaEvent_t> ev[2];   // [0] = k1 pairs, [1] = k2 pairs (start, stop interleaved)
  size_t ev_used[2] = {0, 0};
  ~TcNet() {
    for (auto& v : ev)
      for (cudaEvent_t e : v) cudaEventDestroy(e);
  }
};

static constexpr size_t kMaxProfPairs = 8192;
static cudaEvent_t prof_event(TcNet* n, int which) {
  if (!n->prof || n->ev_used[which] >= 2 * kMaxProfPairs) return nullptr;
  if (n->ev_used[which] == n->ev[which].size()) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    n->ev[which].push_back(e);
  }
  return n->ev[which][n->ev_used[which]++];
}
int tc_net_debug_counters(TcNet* n, long long* host16x256) {
  if (!n->dbg.p) return fail(AP_ERR_STATE, "set AP_TC_DEBUG=1 before creating the network");
  AP_CUDA(cudaMemcpy(host16x256, n->dbg.p, sizeof(long long) * 16 * 256, cudaMemcpyDeviceToHost));
  return AP_OK;
}
void tc_net_profile(TcNet* n, bool on) {
  n->prof = on;
  n->ev_used[0] = n->ev_used[1] = 0;
}
int tc_net_profile_read(TcNet* n, double* ms, int* count) {
  for (int w = 0; w < 2; ++w) {
    ms[w] = 0.0, count[w] = 0;
    for (size_t i = 0; i + 1 < n->ev_used[w]; i += 2) {
      AP_CUDA(cudaEventSynchronize(n->ev[w][i + 1]));
      float t = 0.f;
      AP_CUDA(cudaEventElapsedTime(&t, n->ev[w][i], n->ev[w][i + 1]));
      ms[w] += t, ++count[w];
    }
  }
  return AP_OK;
}

static int upload_bf16(DevBuf& d, const std::vector<uint16_t>& v) {
  AP_CUDA(d.upload(v.data(), v.size() * sizeof(uint16_t)));
  return AP_OK;
}
static int upload_f32(DevBuf& d, const std::vector<float>& v) {
  AP_CUDA(d.upload(v.data(), v.size() * sizeof(float)));
  return AP_OK;
}

int tc_net_create(TcNet** out, const ap_wavenet_cfg& cfg, const float* const* weights) {
  using namespace tc;
  *out = nullptr;
  if (cfg.res_channels != C || cfg.skip_channels != C) return fail(AP_ERR_INVALID, "tensor-core path needs 256 channels");
  TcNet* n = new TcNet();
  n->cfg = cfg;
  const int N = n->N = cfg.num_res_layers;
  std::vector<uint16_t> wd(static_cast<size_t>(N) * 512 * 768), wr(static_cast<size_t>(N) * C * C), ws(wr.size()),
      wf(static_cast<size_t>(C) * C);
  auto to_f16 = [](const std::vector<uint16_t>& bf, const std::vector<float>& src) {
    (void)bf;
    std::vector<uint16_t> out(src.size());
    for (size_t i = 0; i < src.size(); ++i) out[i] = __half_as_ushort(__float2half_rn(src[i]));
    return out;
  };
  std::vector<float> wd_f(wd.size()), wr_f(wr.size()), ws_f(ws.size()), wf_f(wf.size());
  std::vector<float> bd(static_cast<size_t>(N) * 512), br(static_cast<size_t>(N) * C), bskip(C, 0.f);
  std::vector<double> bsum(C, 0.0);
  for (int l = 0; l < N; ++l) {
    const float* const* w = weights + 6 + 8 * l;
    for (int j = 0; j < 2; ++j)
      for (int r = 0; r < 256; ++r) {
        const int oc = r < 128 ? 128 * j + r : C + 128 * j + (r - 128);
        bd[static_cast<size_t>(l) * 512 + j * 256 + r] = w[3][oc];
        uint16_t* dst = &wd[((static_cast<size_t>(l) * 2 + j) * 256 + r) * 768];
        for (int tap = 0; tap < 3; ++tap)
          for (int c = 0; c < C; ++c) {
            const float v = w[2][(static_cast<size_t>(oc) * C + c) * 3 + tap];
            dst[tap * C + c] = f32_to_bf16_rne(v);
            wd_f[((static_cast<size_t>(l) * 2 + j) * 256 + r) * 768 + tap * C + c] = v;
          }
      }
    for (int o = 0; o < C; ++o) {
      br[static_cast<size_t>(l) * C + o] = w[5][o];
      bsum[o] += w[7][o];
      for (int c = 0; c < C; ++c) {
        wr[(static_cast<size_t>(l) * C + o) * C + c] = f32_to_bf16_rne(w[4][static_cast<size_t>(o) * C + c]);
        ws[(static_cast<size_t>(l) * C + o) * C + c] = f32_to_bf16_rne(w[6][static_cast<size_t>(o) * C + c]);
        wr_f[(static_cast<size_t>(l) * C + o) * C + c] = w[4][static_cast<size_t>(o) * C + c];
        ws_f[(static_cast<size_t>(l) * C + o) * C + c] = w[6][static_cast<size_t>(o) * C + c];
      }
    }
  }
  for (int o = 0; o < C; ++o) bskip[o] = static_cast<float>(bsum[o]);
  const float* const* tail = weights + 6 + 8 * N;
  for (size_t i = 0; i < wf.size(); ++i) wf[i] = f32_to_bf16_rne(tail[0][i]), wf_f[i] = tail[0][i];
  int rc = AP_OK;
#define TRY(e) if (rc == AP_OK) rc = (e)
  TRY(upload_bf16(n->wd, wd));
  TRY(upload_bf16(n->wr, wr));
  TRY(upload_bf16(n->ws, ws));
  TRY(upload_bf16(n->wf, wf));
  TRY(upload_bf16(n->wd_h, to_f16(wd, wd_f)));
  TRY(upload_bf16(n->wr_h, to_f16(wr, wr_f)));
  TRY(upload_bf16(n->ws_h, to_f16(ws, ws_f)));
  TRY(upload_bf16(n->wf_h, to_f16(wf, wf_f)));
  TRY(upload_f32(n->bd, bd));
  TRY(upload_f32(n->br, br));
  TRY(upload_f32(n->bskip, bskip));
  TRY(upload_f32(n->bf1, std::vector<float>(tail[1], tail[1] + C)));
  TRY(upload_f32(n->wf2, std::vector<float>(tail[2], tail[2] + C)));
  TRY(upload_f32(n->bf2, std::vector<float>(tail[3], tail[3] + 1)));
  TRY(upload_f32(n->init_w, std::vector<float>(weights[0], weights[0] + C)));
  TRY(upload_f32(n->init_b, std::vector<float>(weights[1], weights[1] + C)));
  if (rc == AP_OK) {
    const uint64_t d1[2] = {768, static_cast<uint64_t>(N) * 512}, d2[2] = {256, static_cast<uint64_t>(N) * 256},
                   d3[2] = {256, 256};
    const uint32_t bw[2] = {64, 256};
    TRY(encode_bf16(&n->tmWd, n->wd.p, 2, d1, bw));
    TRY(encode_bf16(&n->tmWr, n->wr.p, 2, d2, bw));
    TRY(encode_bf16(&n->tmWs, n->ws.p, 2, d2, bw));
    TRY(encode_bf16(&n->tmWf, n->wf.p, 2, d3, bw));
    const uint32_t bw2[2] = {64, 128};
    TRY(encode_bf16(&n->tmWd2, n->wd.p, 2, d1, bw2));
    TRY(encode_bf16(&n->tmWr2, n->wr.p, 2, d2, bw2));
    TRY(encode_bf16(&n->tmWs2, n->ws.p, 2, d2, bw2));
    TRY(encode_bf16(&n->tmWf2, n->wf.p, 2, d3, bw2));
    TRY(encode_bf16(&n->tmWd_h, n->wd_h.p, 2, d1, bw));
    TRY(encode_bf16(&n->tmWr_h, n->wr_h.p, 2, d2, bw));
    TRY(encode_bf16(&n->tmWs_h, n->ws_h.p, 2, d2, bw));
    TRY(encode_bf16(&n->tmWf_h, n->wf_h.p, 2, d3, bw));
    TRY(encode_bf16(&n->tmWd2_h, n->wd_h.p, 2, d1, bw2));
    TRY(encode_bf16(&n->tmWr2_h, n->wr_h.p, 2, d2, bw2));
    TRY(encode_bf16(&n->tmWs2_h, n->ws_h.p, 2, d2, bw2));
    TRY(encode_bf16(&n->tmWf2_h, n->wf_h.p, 2, d3, bw2));
    const char* env = std::getenv("AP_TC_PAIR");
    if (env && env[0] == '0') n->pair = false;
    if (const char* e2 = std::getenv("AP_TC_DEBUG_LAYER")) n->dbg_layer = std::atoi(e2);
    env = std::getenv("AP_TC_DEBUG");
    if (env && env[0] == '1') {
      TRY((n->dbg.alloc(sizeof(long long) * 16 * 256) == cudaSuccess &&
           cudaMemset(n->dbg.p, 0, sizeof(long long) * 16 * 256) == cudaSuccess) ? AP_OK : fail(AP_ERR_CUDA, "dbg alloc"));
    }
  }
#undef TRY
  if (rc != AP_OK) {
    delete n;
    return rc;
  }
  *out = n;
  return AP_OK;
}

void tc_net_destroy(TcNet* n) { delete n; }
void tc_net_set_dtype(TcNet* n, int dt) { n->dt = dt ? 1 : 0; }

size_t tc_net_workspace_bytes(const TcNet* n) { return n->u0.bytes + n->u1.bytes + n->o.bytes; }

int tc_net_reserve(TcNet* n, int chunk, int L) {
  using namespace tc;
  const size_t per = static_cast<size_t>(chunk) * L * C * sizeof(uint16_t);
  AP_CUDA(n->u0.alloc(per));
  AP_CUDA(n->u1.alloc(per));
  AP_CUDA(n->o.alloc(per * n->N));
  n->chunk = chunk, n->L = L;
  const uint64_t du[3] = {256, static_cast<uint64_t>(L), static_cast<uint64_t>(chunk)};
  const uint64_t dO[3] = {256, static_cast<uint64_t>(L), static_cast<uint64_t>(chunk) * n->N};
  const uint32_t bx[3] = {64, 128, 1};
  int rc = encode_bf16(&n->tmU[0], n->u0.p, 3, du, bx);
  if (rc == AP_OK) rc = encode_bf16(&n->tmU[1], n->u1.p, 3, du, bx);
  if (rc == AP_OK) rc = encode_bf16(&n->tmO, n->o.p, 3, dO, bx);
  if (rc != AP_OK) return rc;
  if (!n->attr_set) {
    AP_CUDA(cudaFuncSetAttribute(k1_layer<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<1, 1>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k2_head<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<1, 2>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k1_layer<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 1>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k2_head<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 2>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k1_layer<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<1, 1>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k2_head<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<1, 2>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k1_layer<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 1>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k2_head<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 2>::SMEM_BYTES));
    n->attr_set = true;
  }
  return AP_OK;
}

// cluster-of-2 launch (CTA pairs); grid = 2 x min(#pairs, #SMs / 2)
static int pair_grid(int n_tiles) {
  const int pairs = (n_tiles + 1) / 2, cap = num_sms() / 2;
  return 2 * (pairs < cap ? pairs : cap);
}
template <class Kernel, class... Args>
static cudaError_t launch_pair(Kernel kernel, int grid, int smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(tc::NTHREADS), cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

static int tc_run_layers(TcNet* n, const float* x, const float* ptab, int B, int L, int layers, cudaStream_t st) {
  using namespace tc;
  const long long M = static_cast<long long>(B) * L;
  {
    long long blocks = ceil_div_ll(M * (C / 8), 256);
    const long long cap = static_cast<long long>(num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    if (n->dt == 0)
      init_h16_kernel<0><<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, n->init_w.as<float>(), n->init_b.as<float>(), ptab,
                                                                        n->u0.as<uint4>(), M);
    else
      init_h16_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, n->init_w.as<float>(), n->init_b.as<float>(), ptab,
                                                                        n->u0.as<uint4>(), M);
    AP_LAUNCH_CHECK();
  }
  const int tps = ceil_div(L, TILE_M), n_tiles = tps * B;
  const int grid = n_tiles < num_sms() ? n_tiles : num_sms();
  for (int l = 0; l < layers; ++l) {
    K1Params p;
    p.n_tiles = n_tiles, p.tiles_per_sample = tps, p.dilation = 1 << (l % n->cfg.dilation_cycle), p.layer = l;
    p.chunk_alloc = n->chunk, p.last = (l == n->N - 1);
    p.b_dil = n->bd.as<float>() + static_cast<size_t>(l) * 512;
    p.b_res = n->br.as<float>() + static_cast<size_t>(l) * C;
    p.p_next = ptab + static_cast<size_t>(l + 1) * C;
    p.dbg = (l == n->dbg_layer) ? n->dbg.as<long long>() : nullptr;
    cudaEvent_t e0 = prof_event(n, 0), e1 = e0 ? prof_event(n, 0) : nullptr;
    if (e1) cudaEventRecord(e0, st);
    const CUtensorMap &ui = n->tmU[l & 1], &uo = n->tmU[(l + 1) & 1];
    if (n->pair && n->dt == 0)
      AP_CUDA(launch_pair(k1_layer<2, 0>, pair_grid(n_tiles), Geo<2, 1>::SMEM_BYTES, st, ui, uo, n->tmO, n->tmWd2, n->tmWr2, p));
    else if (n->pair)
      AP_CUDA(launch_pair(k1_layer<2, 1>, pair_grid(n_tiles), Geo<2, 1>::SMEM_BYTES, st, ui, uo, n->tmO, n->tmWd2_h, n->tmWr2_h, p));
    else if (n->dt == 0)
      k1_layer<1, 0><<<grid, NTHREADS, Geo<1, 1>::SMEM_BYTES, st>>>(ui, uo, n->tmO, n->tmWd, n->tmWr, p);
    else
      k1_layer<1, 1><<<grid, NTHREADS, Geo<1, 1>::SMEM_BYTES, st>>>(ui, uo, n->tmO, n->tmWd_h, n->tmWr_h, p);
    if (e1) cudaEventRecord(e1, st);
    AP_LAUNCH_CHECK();
  }
  return AP_OK;
}

int tc_net_eps(TcNet* n, const float* x, const float* ptab, float* eps, int B, int L, cudaStream_t st) {
  using namespace tc;
  if (B > n->chunk || L != n->L) return fail(AP_ERR_STATE, "tc_net_eps: workspace reserved for chunk %d x L %d", n->chunk, n->L);
  int rc = tc_run_layers(n, x, ptab, B, L, n->N, st);
  if (rc != AP_OK) return rc;
  const int tps = ceil_div(L, TILE_M), n_tiles = tps * B;
  K2Params p;
  p.n_tiles = n_tiles, p.tiles_per_sample = tps, p.L = L, p.num_layers = n->N, p.chunk_alloc = n->chunk;
  p.scale = static_cast<float>(std::sqrt(1.0 / n->N));
  p.bskip = n->bskip.as<float>(), p.bf1 = n->bf1.as<float>(), p.wf2 = n->wf2.as<float>(), p.bf2 = n->bf2.as<float>();
  p.eps = eps;
  const int grid = n_tiles < num_sms() ? n_tiles : num_sms();
  cudaEvent_t e0 = prof_event(n, 1), e1 = e0 ? prof_event(n, 1) : nullptr;
  if (e1) cudaEventRecord(e0, st);
  if (n->pair && n->dt == 0)
    AP_CUDA(launch_pair(k2_head<2, 0>, pair_grid(n_tiles), Geo<2, 2>::SMEM_BYTES, st, n->tmO, n->tmWs2, n->tmWf2, p));
  else if (n->pair)
    AP_CUDA(launch_pair(k2_head<2, 1>, pair_grid(n_tiles), Geo<2, 2>::SMEM_BYTES, st, n->tmO, n->tmWs2_h, n->tmWf2_h, p));
  else if (n->dt == 0)
    k2_head<1, 0><<<grid, NTHREADS, Geo<1, 2>::SMEM_BYTES, st>>>(n->tmO, n->tmWs, n->tmWf, p);
  else
    k2_head<1, 1><<<grid, NTHREADS, Geo<1, 2>::SMEM_BYTES, st>>>(n->tmO, n->tmWs_h, n->tmWf_h, p);
  if (e1) cudaEventRecord(e1, st);
  AP_LAUNCH_CHECK();
  return AP_OK;
}

// debug: run init + layers [0, layer] and return u_{layer+1} and o_layer as fp32 (B, L, 256)
int tc_net_debug_layer(TcNet* n, const float* x, const float* ptab, int layer, float* u_next, float* gate, int B, int L,
                       cudaStream_t st) {
  using namespace tc;
  if (B > n->chunk || L != n->L) return fail(AP_ERR_STATE, "tc_net_debug_layer: workspace mismatch");
  int rc = tc_run_layers(n, x, ptab, B, L, layer + 1, st);
  if (rc != AP_OK) return rc;
  const long long cnt = static_cast<long long>(B) * L * C;
  const uint16_t* un = ((layer + 1) & 1) ? n->u1.as<uint16_t>() : n->u0.as<uint16_t>();
  const uint16_t* on = n->o.as<uint16_t>() + static_cast<size_t>(layer) * n->chunk * L * C;
  if (u_next) {
    if (n->dt == 0) h16_to_f32_kernel<0><<<num_sms() * 4, 256, 0, st>>>(un, u_next, cnt);
    else h16_to_f32_kernel<1><<<num_sms() * 4, 256, 0, st>>>(un, u_next, cnt);
    AP_LAUNCH_CHECK();
  }
  if (gate) {
    if (n->dt == 0) h16_to_f32_kernel<0><<<num_sms() * 4, 256, 0, st>>>(on, gate, cnt);
    else h16_to_f32_kernel<1><<<num_sms() * 4, 256, 0, st>>>(on, gate, cnt);
    AP_LAUNCH_CHECK();
  }
  return AP_OK;
}

}  // namespace ap

// ================================================================================================ C ABI: self test
extern "C" int ap_selftest_umma(const uint16_t* a_bf16, const uint16_t* b_bf16, float* d_out, int K, void* stream) {
  using namespace ap;
  using namespace ap::tc;
  AP_REQUIRE(a_bf16 && b_bf16 && d_out, "ap_selftest_umma: null pointer");
  AP_REQUIRE(K > 0 && K % 64 == 0, "ap_selftest_umma: K must be a positive multiple of 64 (got %d)", K);
  int dev = 0;
  AP_CUDA(cudaGetDevice(&dev));
  int rc = select_device(dev);
  if (rc != AP_OK) return rc;
  CUtensorMap ta, tb;
  const uint64_t da[2] = {static_cast<uint64_t>(K), 128}, db[2] = {static_cast<uint64_t>(K), 256};
  const uint32_t ba[2] = {64, 128}, bb[2] = {64, 256};
  rc = encode_bf16(&ta, a_bf16, 2, da, ba);
  if (rc == AP_OK) rc = encode_bf16(&tb, b_bf16, 2, db, bb);
  if (rc != AP_OK) return rc;
  const int smem = Geo<1, 1>::STAGE_BYTES + 128 + 1024;
  AP_CUDA(cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(ta, tb, d_out, K);
  AP_LAUNCH_CHECK();
  return AP_OK;
}
