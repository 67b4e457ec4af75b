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
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-3 idle, warps 4..11 =
// epilogue.  k1 moves registers from warp group 0 to the epilogue groups with setmaxnreg (96 / 200 per thread).
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <cuda.h>
#include <cuda_fp16.h>

#include "ap_common.cuh"
#include "ap_conv_tc.h"
#include "ap_internal.h"
#include "ap_philox.cuh"
#include "ap_ptx.cuh"

namespace ap {
namespace tc {

using namespace ptx;

constexpr int C = 256;                      // channels (res == skip)
// Gate variants (development switches, devtools/ab_compare.py).  MEASURED (r02, same box, sustained): evaluating the gate's
// exponentials on the FMA pipe (gate_poly2: one MUFU.RCP per channel instead of two MUFU.TANH) makes k1_layer SLOWER
// (1.72 -> 2.22 ms per 128 waveforms): a microbenchmark puts MUFU at 16 results/clk/SM, so the two tanh of a tile cost
// 4096 of its 14336 clk -- the epilogue is bound by its instruction count / dependency chains, not by the XU pipe, and the
// polynomial adds ~9 FMA-pipe + 6 ALU instructions per channel.  Kept for the record and for accuracy studies (3.7e-5 max
// abs error vs 7e-4 for MUFU.TANH).
#ifdef AP_GATE_POLY
constexpr bool kGatePoly = true;
#else
constexpr bool kGatePoly = false;
#endif
#ifdef AP_SPLIT_GATE_POLY
constexpr bool kSplitGatePoly = true;
#else
constexpr bool kSplitGatePoly = false;
#endif
constexpr int TILE_M = 128;                 // positions per tile (per CTA)
constexpr int A_BYTES = TILE_M * 128;       // [128 rows][64 bf16]  SWIZZLE_128B
constexpr int OUT_BYTES = 4 * A_BYTES;      // 4 K-blocks of [128][64] bf16 (gate output / skip-sum staging)
constexpr int NTHREADS = 384;               // warp group 0: TMA producer, MMA issuer, 2 idle warps; groups 1-2: epilogue
constexpr int EPI_WARP0 = 4;                // first epilogue warp (warp-group aligned: setmaxnreg works per warp group)
constexpr int EPI_THREADS = 256;

// CG = 1: one CTA per tile, tcgen05 cta_group::1 (M = 128).
// CG = 2: a CTA pair (cluster of 2) works on two adjacent tiles with cta_group::2 (M = 256): each CTA stages its own 128
//         activation rows and HALF of every weight tile, so both the TMA fill and the MMA operand reads of shared memory
//         drop from 192 to 128 B/clk per SM (shared memory delivers 128 B/clk, which capped CG = 1 at ~63 % tensor duty).
// KIND 1 = k1_layer, KIND 2 = k2_head.  Per-channel biases are kernel parameters (constant bank), so shared memory holds
// only the TMA ring (5 stages in pair mode), one 64 KB staging tile and 1 KB of scratch.
//
// DT = 2 ("bf16x3", AP_MODE_BF16X3): fp32-class arithmetic on the bf16 tensor cores.  Every activation and weight is kept as
// a pair of bf16 planes hi = bf16(v), lo = bf16(v - hi) (16 significand bits together) and every product is accumulated
// as hi*hi + lo*hi + hi*lo in fp32 (the dropped lo*lo term is below 2^-18 relative), i.e. three MMAs per K step.  The
// staging tile holds both planes (128 KB), which leaves a 3-stage ring.
template <int CG_, int KIND, int DT_ = 0> struct Geo {
  static constexpr int CG = CG_;
  static constexpr int DT = DT_;   // 0: bf16 operands, 1: fp16 operands, 2: bf16 hi/lo split operands
  static constexpr bool SPLIT = DT_ == 2;
  static constexpr int NCOMBO = SPLIT ? 3 : 1;   // (A plane, B plane) pairs per K block: (hi,hi) [, (lo,hi), (hi,lo)]
  static constexpr int B_ROWS = 256 / CG;
  static constexpr int B_BYTES = B_ROWS * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // bf16x3: k2_head stages both planes of the skip sum (128 KB, 3-stage ring); k1_split issues GEMM-2 in two K halves through
  // one 64 KB tile and keeps the 5-stage ring
  static constexpr bool WIDE_STAGING = SPLIT && KIND == 2;
  // k1_layer in pair mode moves the residual stream through shared memory with TMA (u sub-tiles in, u' sub-tiles out: two
  // 16 KB buffers) instead of 32-byte-per-thread global accesses; that costs one ring stage (4 instead of 5).
#ifdef AP_RES_DIRECT
  static constexpr bool RES_STAGED = false;
#else
  static constexpr bool RES_STAGED = KIND == 1 && !SPLIT && CG == 2;
#endif
#ifdef AP_K1_NSTAGE   // development aid: ring depth experiments
  static constexpr int NSTAGE = CG == 1 ? (SPLIT ? 2 : 3) : (WIDE_STAGING ? 3 : (RES_STAGED ? AP_K1_NSTAGE : 5));
#else
  static constexpr int NSTAGE = CG == 1 ? (SPLIT ? 2 : 3) : (WIDE_STAGING ? 3 : (RES_STAGED ? 4 : 5));
#endif
  static constexpr int OUT_OFF = NSTAGE * STAGE_BYTES;
  static constexpr int RES_OFF = OUT_OFF + (WIDE_STAGING || (SPLIT && CG == 1) ? 2 : 1) * OUT_BYTES;
  static constexpr int BIAS_OFF = RES_OFF + (RES_STAGED ? 2 * A_BYTES : 0);
  static constexpr int BAR_OFF = BIAS_OFF + 1024;   // k1: c2[256]; k2: 128 partial dots
  static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;   // + slack to align the base to 1024 B
  static constexpr uint32_t IDESC = DT_ == 1 ? umma_idesc_f16_f32(128 * CG, 256) : umma_idesc_bf16_f32(128 * CG, 256);
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
  static_assert(NSTAGE <= 5, "barrier table holds at most 5 stages");
};
enum { BAR_FULL = 0, BAR_EMPTY = 5, BAR_ACC_FULL = 10, BAR_ACC_EMPTY = 12, BAR_OUT_READY = 14, BAR_G2A_DONE = 16,
       BAR_RES_FULL = 17,   // [2] u sub-tile landed in residual buffer i (loader warp's TMA)
       BAR_RES_DONE = 19,   // [2] the 8 epilogue warps have overwritten buffer i with u' (-> storer warp)
       BAR_RES_FREE = 21,   // [2] the TMA store of buffer i has read it (-> loader warp)
       BAR_COUNT = 23 };

// position in the TMA ring: stage index and phase parity (no division in the producer / MMA-issue loops)
template <int NSTAGE> struct RingPos {
  uint32_t s = 0, ph = 0;
  __device__ __forceinline__ RingPos& operator++() {
    if (++s == NSTAGE) s = 0, ph ^= 1;
    return *this;
  }
};

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
  //           (wait_empty: every lane of the producer warp; arm: the elected lane, followed by its TMA loads)
  __device__ __forceinline__ long long wait_empty(uint32_t s, uint32_t ph, int tag) const {
    return mbar_wait(bar(BAR_EMPTY + s), ph ^ 1, tag);
  }
  __device__ __forceinline__ void arm(uint32_t s, uint32_t bytes) const {
    if (CG == 1) mbar_expect_tx(bar(BAR_FULL + s), bytes);
    else if (rank == 0) mbar_expect_tx(bar(BAR_FULL + s), 2 * bytes);
    else mbar_arrive_cluster(lbar(BAR_FULL + s));
  }
  __device__ __forceinline__ void load_a(uint32_t s, const CUtensorMap* m, int c0, int c1, int c2) const {
    if (CG == 1) tma_load_3d(stage_a(s), m, bar(BAR_FULL + s), c0, c1, c2);
    else tma_load_3d_pair(stage_a(s), m, lbar(BAR_FULL + s), c0, c1, c2);
  }
  // weight tile: this CTA's B_ROWS rows starting at row0 (+ rank * B_ROWS)
  __device__ __forceinline__ void load_b(uint32_t s, const CUtensorMap* m, int k0, int row0) const {
    if (CG == 1) tma_load_2d(stage_b(s), m, bar(BAR_FULL + s), k0, row0);
#ifdef AP_L2_HINTS   // weights: keep (1 MB per layer, re-read by every tile)
    else tma_load_2d_pair_hint(stage_b(s), m, lbar(BAR_FULL + s), k0, row0 + static_cast<int>(rank) * G::B_ROWS, l2_policy_evict_last());
#else
    else tma_load_2d_pair(stage_b(s), m, lbar(BAR_FULL + s), k0, row0 + static_cast<int>(rank) * G::B_ROWS);
#endif
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
    }
    for (int i = 0; i < 2; ++i) mbar_init(cx.bar(BAR_OUT_READY + i), out_ready_count * CG);
    mbar_init(cx.bar(BAR_G2A_DONE), 1);   // k1_split: first half of GEMM-2 done
    for (int i = 0; i < 2; ++i) {          // k1_layer's staged residual (per CTA, not per pair)
      mbar_init(cx.bar(BAR_RES_FULL + i), 1);
      mbar_init(cx.bar(BAR_RES_DONE + i), 8);
      mbar_init(cx.bar(BAR_RES_FREE + i), 1);
    }
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
  int n_tiles, tiles_per_sample, dilation, layer, chunk_alloc, last, L;
  const float* b_res;   // [256]
  const float* p_next;  // [256]
  const uint16_t* u_in; // [chunk][L][256] this layer's input (also read through tmUin as the GEMM-1 A operand)
  uint16_t* u_out;      // [chunk][L][256] next layer's input
  uint16_t* o_out;      // k1_split: [2 planes][N][chunk][L][256] gate outputs (written from registers)
  uint16_t* ts_out;     // SAVE kernels (backward pass): [N][ts_chunk][L][512] bf16, the gate's local derivatives
                        // d o / d a_t | d o / d a_s of every layer
  int ts_chunk;
  long long* dbg;       // optional [gridDim.x][16] wait-cycle counters (development aid), or null
  // hi/lo split mode (DT = 2): offset of the lo plane in the third coordinate of the U / O tensor maps and in the rows of
  // the Wd / Wr maps (0 otherwise); the lo plane of u_in / u_out starts u_plane * L * 256 elements after the hi plane
  int u_plane, o_plane, wd_plane, wr_plane;
  // dilated-conv biases in packed chunk order: [j][0..127] tanh bias of gate channel 128 j + c, [j][128..255] HALF the
  // sigmoid bias of the same channel (sigmoid(s) = 0.5 tanh(0.5 s) + 0.5) -- the SAVE kernels and k1_split; for k1_layer's
  // inference gate (gate_poly2) the host pre-scales them to 2 log2e b_t | -log2e b_s.  They live in the kernel-parameter
  // constant bank, so the bias add is a constant operand of the FADD / FFMA: no shared-memory or shuffle traffic in the epilogue.
  alignas(8) float bd[512];
};

// Development aid (AP_TC_DEBUG=1): SM-clock timestamps of the MMA jobs and the epilogue phases of tiles kTraceTile0.. of CTA 0,
// rows 200.. of the debug table ([tile][32 slots]); devtools/probe.py prints the reconstructed timeline.
constexpr uint32_t kTraceTile0 = 60, kTraceTiles = 3;
__device__ __forceinline__ void trace(const K1Params& p, uint32_t ti, int slot) {
  if (p.dbg && blockIdx.x == 0 && ti - kTraceTile0 < kTraceTiles) p.dbg[200 * 16 + (ti - kTraceTile0) * 32 + slot] = clock64();
}

// Order of the accumulation jobs of one CTA (pair), shared by the TMA producer, the MMA issuer and the epilogue warps:
//   G1c0(0) G1c1(0) | G1c0(1) G2(0) G1c1(1) | G1c0(2) G2(1) G1c1(2) | ... | G2(T-1)
// GEMM-2 of tile t is issued one chunk late, behind chunk 0 of tile t+1, so the tensor pipe never waits for the gate
// epilogue of chunk 1 (its only input dependency).  Job g accumulates into TMEM region g & 1; the gate epilogue frees its
// region as soon as the accumulators are in registers (before the MUFU work) and holds the packed gate values in
// registers until GEMM-2 of the previous tile has finished reading the 64 KB staging tile.  The residual epilogue does
// not touch shared memory: u is read and u' written with 32-byte-per-thread global accesses (full sectors), which
// takes its four 64 KB passes off the shared-memory crossbar (the resource that bounds this kernel).
template <class G, int HSEL, bool SAVE>
__device__ __forceinline__ void k1_epilogue(const Ctx<G>& cx, const K1Params& p, const uint32_t tmem,
                                            const Tiles<G::CG>& tiles, const CUtensorMap* tmO, const uint32_t c2_addr) {
  constexpr int DT = G::DT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q4 = warp & 3, etid = threadIdx.x - EPI_WARP0 * 32;
  const int row = q4 * 32 + lane;                       // position within the tile == TMEM lane
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q4 * 32) << 16);
  const uint32_t row_off = row * 128, sw = row & 7;
  const float sqrt_half = 0.70710678118654752440f;
  const int oob_l0 = p.tiles_per_sample * TILE_M;
  uint32_t g = 0;
  long long w_accfull = 0, w_g2 = 0, w_bulk = 0;
  const long long t_start = clock64();

  // ---- gate epilogue of chunk J: o = tanh(a_t + b) * sigmoid(a_s + b) -> bf16 -> staging (A operand of GEMM-2) + O
  auto gate = [&](auto jc, uint32_t ti, bool valid, int b, int l0) {
    constexpr int J = decltype(jc)::value;
    const uint32_t r = g & 1;
    w_accfull += mbar_wait(cx.bar(BAR_ACC_FULL + r), (g >> 1) & 1, 9);
    tc_fence_after();
    if (etid == 0) trace(p, ti, 8 + J * 8);
    uint32_t ta[2][32], sg[2][32];
#pragma unroll
    for (int gq = 0; gq < 2; ++gq) {
      tmem_ld_32x32b_x32(lane_addr + r * 256 + HSEL * 64 + gq * 32, ta[gq]);
      tmem_ld_32x32b_x32(lane_addr + r * 256 + 128 + HSEL * 64 + gq * 32, sg[gq]);
    }
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + r);   // accumulators are in registers: the region is free again
    if (etid == 0) trace(p, ti, 9 + J * 8);
    uint32_t pk[2][4][4];
    uint32_t tsk[2][4];                    // SAVE: packed tanh / sigmoid of the current 8 channels
#pragma unroll
    for (int gq = 0; gq < 2; ++gq)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c0 = i * 8 + 2 * e;
          const int cb = J * 256 + HSEL * 64 + gq * 32 + c0;
#ifdef AP_PROBE_NO_EPI   // roofline probe (wrong results): the epilogue keeps its protocol but does no arithmetic
          pk[gq][i][e] = ta[gq][c0] ^ sg[gq][c0 + 1];
          continue;
#endif
          if constexpr (!SAVE && kGatePoly) {
            // exponentials on the FMA pipe, one MUFU.RCP per channel (gate_poly2; p.bd holds 2 log2e b_t | -log2e b_s here)
            const float2 bt = reinterpret_cast<const float2*>(p.bd)[cb >> 1], bs = reinterpret_cast<const float2*>(p.bd)[(cb + 128) >> 1];
            const f32x2 x = fma2(pk2(__uint_as_float(ta[gq][c0]), __uint_as_float(ta[gq][c0 + 1])),
                                 pk2(2.885390081777927f, 2.885390081777927f), pk2(bt.x, bt.y));
            const f32x2 y = fma2(pk2(__uint_as_float(sg[gq][c0]), __uint_as_float(sg[gq][c0 + 1])),
                                 pk2(-1.4426950408889634f, -1.4426950408889634f), pk2(bs.x, bs.y));
            float o0, o1;
            up2(gate_poly2<3>(x, y), o0, o1);
            pk[gq][i][e] = pack2<DT>(o0, o1);
          } else {
            const float t0 = tanh_approx(__uint_as_float(ta[gq][c0]) + p.bd[cb]);
            const float t1 = tanh_approx(__uint_as_float(ta[gq][c0 + 1]) + p.bd[cb + 1]);
            const float s0 = tanh_approx(fmaf(__uint_as_float(sg[gq][c0]), 0.5f, p.bd[cb + 128]));
            const float s1 = tanh_approx(fmaf(__uint_as_float(sg[gq][c0 + 1]), 0.5f, p.bd[cb + 129]));
            pk[gq][i][e] = pack2<DT>(t0 * fmaf(s0, 0.5f, 0.5f), t1 * fmaf(s1, 0.5f, 0.5f));
            if constexpr (SAVE) {
              // the gate's local derivatives d o / d a_t = S (1 - T^2) and d o / d a_s = T S (1 - S), taken in fp32 here:
              // recomputing them from bf16-rounded T, S loses all precision where the gate saturates (1 - T^2 << 1)
              const float g0 = fmaf(s0, 0.5f, 0.5f), g1 = fmaf(s1, 0.5f, 0.5f);
              tsk[0][e] = pack2<DT>(g0 * fmaf(-t0, t0, 1.f), g1 * fmaf(-t1, t1, 1.f));
              tsk[1][e] = pack2<DT>(t0 * g0 * (1.f - g0), t1 * g1 * (1.f - g1));
            }
          }
          if constexpr (SAVE) {
            if (e == 3 && valid && l0 + row < p.L) {
              uint16_t* td = p.ts_out + ((static_cast<size_t>(p.layer) * p.ts_chunk + b) * p.L + l0 + row) * 512 + J * 128 +
                             HSEL * 64 + gq * 32 + i * 8;
              st_global_v4(td, make_uint4(tsk[0][0], tsk[0][1], tsk[0][2], tsk[0][3]));
              st_global_v4(td + 256, make_uint4(tsk[1][0], tsk[1][1], tsk[1][2], tsk[1][3]));
            }
          }
        }
    // the staging tile still holds o of the previous tile until its GEMM-2 (the NEXT job, g + 1) has completed
    if (etid == 0) trace(p, ti, 10 + J * 8);
    if (etid == 0) {                                       // and the O store of chunk J of the previous tile has read it
      if (J == 0 && !p.last && ti > 0) {
        w_g2 += mbar_wait(cx.bar(BAR_ACC_FULL + ((g + 1) & 1)), ((g + 1) >> 1) & 1, 12);
        tc_fence_after();
      }
      const long long t0 = clock64();
      bulk_wait_read<1>();
      w_bulk += clock64() - t0;
    }
    named_bar_sync(1, EPI_THREADS);
    if (etid == 0) trace(p, ti, 12 + J * 8);
    const uint32_t kb_base = cx.out_kb(2 * J + HSEL) + row_off;
#pragma unroll
    for (int gq = 0; gq < 2; ++gq)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        st_shared_v4(kb_base + (((gq * 4 + i) ^ sw) << 4), make_uint4(pk[gq][i][0], pk[gq][i][1], pk[gq][i][2], pk[gq][i][3]));
      }
    fence_proxy_async_smem();
    named_bar_sync(1, EPI_THREADS);
    if (etid == 0) {
      trace(p, ti, 13 + J * 8);
      if (valid) {
#ifdef AP_L2_HINTS   // gate outputs: not read again before k2_head
        tma_store_3d_hint(tmO, cx.out_kb(2 * J), (2 * J) * 64, l0, p.layer * p.chunk_alloc + b, l2_policy_evict_first());
        tma_store_3d_hint(tmO, cx.out_kb(2 * J + 1), (2 * J + 1) * 64, l0, p.layer * p.chunk_alloc + b, l2_policy_evict_first());
#else
        tma_store_3d(tmO, cx.out_kb(2 * J), (2 * J) * 64, l0, p.layer * p.chunk_alloc + b);
        tma_store_3d(tmO, cx.out_kb(2 * J + 1), (2 * J + 1) * 64, l0, p.layer * p.chunk_alloc + b);
#endif
      }
      trace(p, ti, 14 + J * 8);
      bulk_commit();                                       // always a group (possibly empty): wait_group counts stay uniform
      cx.arrive_leader(BAR_OUT_READY + J);
      trace(p, ti, 11 + J * 8);
    }
    ++g;
  };

  // ---- residual epilogue: u' = (u + r) * sqrt(.5) + (b_res * sqrt(.5) + p_next) -> bf16, global -> registers -> global
  uint32_t cur_ti = 0;
  uint32_t rcount = 0;   // residual sub-tiles consumed so far: buffer rcount & 1, barrier phase (rcount >> 1) & 1
  // ---- residual epilogue, staged (pair mode): u arrives in 16 KB sub-tiles [128 rows][64 ch] through TMA (loader warp), each
  //      thread rewrites its row's 32 channels in place, the storer warp sends the sub-tile to u_out with a TMA store.  No global
  //      accesses from the epilogue warps: the direct version's 32-byte-per-thread LDG / STG touched 32 cache lines per
  //      instruction (4096 LSU wavefronts per tile) and held up every other memory instruction of the SM, TMA issue included.
  auto residual_staged = [&]() {
    const uint32_t r = g & 1;
    if (etid == 0) trace(p, cur_ti, 24);
    w_accfull += mbar_wait(cx.bar(BAR_ACC_FULL + r), (g >> 1) & 1, 10);
    tc_fence_after();
    if (etid == 0) trace(p, cur_ti, 25);
    // all 128 accumulators of this thread's row into registers first: the TMEM region is free for the next MMA job at once,
    // however long the sub-tile round trips (store -> buffer free -> load) of passes 2 and 3 take
    uint32_t accs[4][32];
#ifndef AP_RES_LATE_RELEASE
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) tmem_ld_32x32b_x32(lane_addr + r * 256 + pass * 64 + HSEL * 32, accs[pass]);
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + r);
    if (etid == 0) trace(p, cur_ti, 26);
#endif
#pragma unroll
    for (int pass = 0; pass < 4; ++pass, ++rcount) {
      const uint32_t buf = rcount & 1;
      const uint32_t (&acc)[32] = accs[pass];
#ifdef AP_RES_LATE_RELEASE   // development aid: TMEM read pass by pass, released after the last pass (A/B in profiles/r02_ab_k1_variants.txt)
      tmem_ld_32x32b_x32(lane_addr + r * 256 + pass * 64 + HSEL * 32, accs[pass]);
      tmem_ld_wait();
      if (pass == 3) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + r);
      }
#endif
      mbar_wait(cx.bar(BAR_RES_FULL + buf), (rcount >> 1) & 1, 13);
      const uint32_t rb = cx.base + G::RES_OFF + buf * A_BYTES + row_off;
#ifdef AP_PROBE_NO_EPI
      if (acc[0] == 0x7fc12345u) st_shared_v4(rb, make_uint4(acc[1], acc[2], acc[3], acc[4]));
      if (true) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(cx.bar(BAR_RES_DONE + buf));
        continue;
      }
#endif
#pragma unroll
      for (int c = 0; c < 4; ++c) {                          // 8 channels = one 16-byte chunk of the swizzled row
        const uint32_t addr = rb + (((HSEL * 4 + c) ^ sw) << 4);
        const uint4 uv = ld_shared_v4(addr);
        const uint32_t uw[4] = {uv.x, uv.y, uv.z, uv.w};
        const int ch = pass * 64 + HSEL * 32 + c * 8;
        uint32_t pk[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint4 cv = ld_shared_v4(c2_addr + (ch + h * 4) * 4);   // warp-uniform address: broadcast
          const float cc[4] = {__uint_as_float(cv.x), __uint_as_float(cv.y), __uint_as_float(cv.z), __uint_as_float(cv.w)};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint32_t w2 = uw[h * 2 + e];
            const int a0 = c * 8 + h * 4 + e * 2;
            const float v0 = fmaf(unpack_lo<DT>(w2) + __uint_as_float(acc[a0]), sqrt_half, cc[e * 2]);
            const float v1 = fmaf(unpack_hi<DT>(w2) + __uint_as_float(acc[a0 + 1]), sqrt_half, cc[e * 2 + 1]);
            pk[h * 2 + e] = pack2<DT>(v0, v1);
          }
        }
        st_shared_v4(addr, make_uint4(pk[0], pk[1], pk[2], pk[3]));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(cx.bar(BAR_RES_DONE + buf));
    }
    if (etid == 0) trace(p, cur_ti, 27);
    ++g;
  };
  auto residual = [&](bool valid, int b, int l0) {
    if constexpr (G::RES_STAGED) {
      residual_staged();
      return;
    }
    const uint32_t r = g & 1;
    const bool live = valid && l0 + row < p.L;
    if (etid == 0) trace(p, cur_ti, 24);
    const size_t goff = (static_cast<size_t>(b) * p.L + (live ? l0 + row : 0)) * C + HSEL * 128;
    uint32_t uu[8][8];
    if (live) {
#pragma unroll
      for (int v = 0; v < 8; ++v) ld_global_v8(p.u_in + goff + v * 16, uu[v]);
    }
    if (etid == 0) trace(p, cur_ti, 31);
    w_accfull += mbar_wait(cx.bar(BAR_ACC_FULL + r), (g >> 1) & 1, 10);
    tc_fence_after();
    if (etid == 0) trace(p, cur_ti, 25);
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) {
      uint32_t acc[32];
      tmem_ld_32x32b_x32(lane_addr + r * 256 + HSEL * 128 + gq * 32, acc);
      tmem_ld_wait();
      if (gq == 3) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + r);
        if (etid == 0) trace(p, cur_ti, 26);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {                        // 16 channels = one 32-byte access (per plane)
        const int ch = HSEL * 128 + gq * 32 + i * 16;
        uint32_t pk[8];
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) {
          const uint4 cv = ld_shared_v4(c2_addr + (ch + e4 * 4) * 4);   // warp-uniform address: broadcast
          const float cc[4] = {__uint_as_float(cv.x), __uint_as_float(cv.y), __uint_as_float(cv.z), __uint_as_float(cv.w)};
          const int a0 = i * 16 + e4 * 4;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float u0 = unpack_lo<DT>(uu[gq * 2 + i][e4 * 2 + h]), u1 = unpack_hi<DT>(uu[gq * 2 + i][e4 * 2 + h]);
            const float v0 = fmaf(u0 + __uint_as_float(acc[a0 + 2 * h]), sqrt_half, cc[2 * h]);
            const float v1 = fmaf(u1 + __uint_as_float(acc[a0 + 2 * h + 1]), sqrt_half, cc[2 * h + 1]);
            pk[e4 * 2 + h] = pack2<DT>(v0, v1);
          }
        }
        if (live) st_global_v8(p.u_out + goff + (gq * 2 + i) * 16, pk);
      }
    }
    if (etid == 0) trace(p, cur_ti, 27);
    ++g;
  };

  uint32_t ti = 0;
  bool pvalid = false;
  int pb = 0, pl0 = 0;
  for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
    const bool valid = tile < p.n_tiles;
    const int b = valid ? tile / p.tiles_per_sample : 0;
    const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
    cur_ti = ti;
    gate(std::integral_constant<int, 0>{}, ti, valid, b, l0);
    if (!p.last && ti > 0) residual(pvalid, pb, pl0);
    gate(std::integral_constant<int, 1>{}, ti, valid, b, l0);
    pvalid = valid, pb = b, pl0 = l0;
  }
  if (!p.last && ti > 0) residual(pvalid, pb, pl0);
  if (etid == 0) bulk_wait_all<0>();
  if (p.dbg && etid == 0) {
    long long* d = p.dbg + blockIdx.x * 16;
    d[7] = w_accfull, d[8] = w_g2, d[9] = clock64() - t_start, d[10] = w_bulk;
  }
}

template <int CG, int DT, bool SAVE = false>
__global__ void __launch_bounds__(NTHREADS, 1)
k1_layer(const __grid_constant__ CUtensorMap tmUin, const __grid_constant__ CUtensorMap tmUout,
         const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmWd,
         const __grid_constant__ CUtensorMap tmWr, const __grid_constant__ K1Params p) {
  using G = Geo<CG, 1, DT>;
  static_assert(!G::SPLIT, "the bf16x3 mode has its own layer kernel (k1_split)");
  Ctx<G> cx;
  uint8_t* gen;
  const uint32_t tmem = tc_prologue<G>(cx, gen, 1);
  {  // c2[c] = b_res[c] * sqrt(.5) + p_next[c]: the per-channel constant of the residual update (WaveNet.py:84,97)
    float* s_c2 = reinterpret_cast<float*>(gen + G::BIAS_OFF);
    for (int i = threadIdx.x; i < C; i += NTHREADS) s_c2[i] = fmaf(p.b_res[i], 0.70710678118654752440f, p.p_next[i]);
  }
  if (threadIdx.x == 0)
    prefetch_tmap(&tmUin), prefetch_tmap(&tmUout), prefetch_tmap(&tmO), prefetch_tmap(&tmWd), prefetch_tmap(&tmWr);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Tiles<CG> tiles(p.n_tiles, cx.rank);
  const int oob_l0 = p.tiles_per_sample * TILE_M;   // a row coordinate past the end: TMA zero-fills the whole box

  if (warp < EPI_WARP0) setmaxnreg_dec<96>();
  if (warp == 0) {
    // ======================================================================================= TMA producer
    {   // whole warp, warp-uniform control flow; one elected lane arms the barrier and issues the loads
      RingPos<G::NSTAGE> it;
      uint32_t ti = 0;
      long long w_empty = 0;
      const long long t_start = clock64();
      // split mode: per K block the (A plane, B plane) pairs (hi,hi), (lo,hi), (hi,lo), each its own ring stage
      auto load_g1 = [&](int j, bool valid, int b, int l0) {
        for (int tap = 0; tap < 3; ++tap)
          for (int kb = 0; kb < 4; ++kb)
            for (int cmb = 0; cmb < G::NCOMBO; ++cmb, ++it) {
              const uint32_t s = it.s, ph = it.ph;
              w_empty += cx.wait_empty(s, ph, 1);
              if (elect_one()) {
#ifdef AP_PROBE_NO_TMA   // roofline probe (wrong results): no operand traffic, the MMAs run on whatever the ring holds
                if (cx.rank == 0) mbar_arrive(cx.bar(BAR_FULL + s)); else mbar_arrive_cluster(cx.lbar(BAR_FULL + s));
#else
                cx.arm(s, G::STAGE_BYTES);
                cx.load_a(s, &tmUin, kb * 64, valid ? l0 + (tap - 1) * p.dilation : oob_l0, b + (cmb == 1 ? p.u_plane : 0));
                cx.load_b(s, &tmWd, tap * C + kb * 64, (p.layer * 2 + j) * 256 + (cmb == 2 ? p.wd_plane : 0));
#endif
              }
              __syncwarp();
            }
      };
      auto load_wr = [&]() {
        for (int kb = 0; kb < 4; ++kb)
          for (int cmb = 0; cmb < G::NCOMBO; ++cmb, ++it) {
            const uint32_t s = it.s, ph = it.ph;
            w_empty += cx.wait_empty(s, ph, 2);
            if (elect_one()) {
#ifdef AP_PROBE_NO_TMA
              if (cx.rank == 0) mbar_arrive(cx.bar(BAR_FULL + s)); else mbar_arrive_cluster(cx.lbar(BAR_FULL + s));
#else
              cx.arm(s, G::B_BYTES);
              cx.load_b(s, &tmWr, kb * 64, p.layer * 256 + (cmb == 2 ? p.wr_plane : 0));
#endif
            }
            __syncwarp();
          }
      };
      for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
        const bool valid = tile < p.n_tiles;
        const int b = valid ? tile / p.tiles_per_sample : 0;
        const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
        load_g1(0, valid, b, l0);
        if (!p.last && ti > 0) load_wr();
        load_g1(1, valid, b, l0);
      }
      if (!p.last && ti > 0) load_wr();
      if (p.dbg && lane == 0) p.dbg[blockIdx.x * 16 + 0] = w_empty, p.dbg[blockIdx.x * 16 + 1] = clock64() - t_start;
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================================================================================= MMA issuer (leader CTA)
    // The whole warp runs the loop (warp-uniform control flow keeps the smem descriptors and barrier addresses in uniform
    // registers); one elected lane issues the tcgen05 instructions.  With a single divergent lane the compiler emitted
    // ~105 dependent SASS instructions per K block (ELECT + 5 R2UR per MMA): ~570 clk of issue for 512 clk of tensor work.
    if (cx.rank == 0) {
      RingPos<G::NSTAGE> it;
      uint32_t g = 0, ti = 0;
      long long w_full = 0, w_acc = 0, w_out = 0;
      const long long t_start = clock64();
      int tslot = 0;
      auto gemm1 = [&]() {
        const uint32_t r = g & 1;
        if (lane == 0) trace(p, ti, tslot);
        w_acc += mbar_wait(cx.bar(BAR_ACC_EMPTY + r), ((g >> 1) & 1) ^ 1, 3);
        tc_fence_after();
        if (lane == 0) trace(p, ti, tslot + 1);
        for (int kblk = 0; kblk < 12 * G::NCOMBO; ++kblk, ++it) {
          const uint32_t s = it.s, ph = it.ph;
          w_full += mbar_wait(cx.bar(BAR_FULL + s), ph, 4);
          tc_fence_after();
          if (elect_one()) {
            cx.mma_kblock(tmem + r * 256, cx.stage_a(s), cx.stage_b(s), kblk == 0);
            cx.commit(BAR_EMPTY + s);
          }
          __syncwarp();
        }
        if (elect_one()) cx.commit(BAR_ACC_FULL + r);
        __syncwarp();
        if (lane == 0) trace(p, ti, tslot + 2);
        ++g;
      };
      auto gemm2 = [&](uint32_t t_idx) {
        const uint32_t r = g & 1;
        if (lane == 0) trace(p, ti, 28);
        w_acc += mbar_wait(cx.bar(BAR_ACC_EMPTY + r), ((g >> 1) & 1) ^ 1, 5);
        tc_fence_after();
        if (lane == 0) trace(p, ti, 29);
        for (int kb = 0; kb < 4; ++kb) {
          if ((kb & 1) == 0) w_out += mbar_wait(cx.bar(BAR_OUT_READY + (kb >> 1)), t_idx & 1, 6);
          for (int cmb = 0; cmb < G::NCOMBO; ++cmb, ++it) {
            const uint32_t s = it.s, ph = it.ph;
            w_full += mbar_wait(cx.bar(BAR_FULL + s), ph, 8);
            tc_fence_after();
            if (elect_one()) {
              cx.mma_kblock(tmem + r * 256, cx.out_kb(kb + (cmb == 1 ? 4 : 0)), cx.stage_b(s), kb == 0 && cmb == 0);
              cx.commit(BAR_EMPTY + s);
            }
            __syncwarp();
          }
        }
        if (elect_one()) cx.commit(BAR_ACC_FULL + r);
        __syncwarp();
        if (lane == 0) trace(p, ti, 30);
        ++g;
      };
      for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
        tslot = 0;
        gemm1();
        if (!p.last && ti > 0) gemm2(ti - 1);
        tslot = 4;
        gemm1();
      }
      if (!p.last && ti > 0) gemm2(ti - 1);
      if (p.dbg && lane == 0) {
        long long* d = p.dbg + blockIdx.x * 16;
        d[2] = w_full, d[3] = w_acc, d[4] = w_out, d[5] = clock64() - t_start, d[6] = ti;
      }
    }
  } else if (warp == 2) {
    // ======================================================================================= residual loader (staged residual)
    // one 16 KB sub-tile [128 rows][64 channels] of this CTA's u tile per pass, into residual buffer (count & 1) once the storer
    // has released it; runs ahead of the epilogue by up to two sub-tiles
    if (G::RES_STAGED && !p.last) {
      uint32_t rcount = 0;
      for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride) {
        const bool valid = tile < p.n_tiles;
        const int b = valid ? tile / p.tiles_per_sample : 0;
        const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
        for (int pass = 0; pass < 4; ++pass, ++rcount) {
          const uint32_t buf = rcount & 1;
          mbar_wait(cx.bar(BAR_RES_FREE + buf), ((rcount >> 1) & 1) ^ 1, 14);
          if (elect_one()) {
            mbar_expect_tx(cx.bar(BAR_RES_FULL + buf), A_BYTES);
#ifdef AP_L2_HINTS   // last use of this part of u in the launch
            tma_load_3d_hint(cx.base + G::RES_OFF + buf * A_BYTES, &tmUin, cx.bar(BAR_RES_FULL + buf), pass * 64, l0, b, l2_policy_evict_first());
#else
            tma_load_3d(cx.base + G::RES_OFF + buf * A_BYTES, &tmUin, cx.bar(BAR_RES_FULL + buf), pass * 64, l0, b);
#endif
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 3) {
    // ======================================================================================= residual storer (staged residual)
    if (G::RES_STAGED && !p.last) {
      uint32_t rcount = 0;
      for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride) {
        const bool valid = tile < p.n_tiles;
        const int b = valid ? tile / p.tiles_per_sample : 0;
        const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
        for (int pass = 0; pass < 4; ++pass, ++rcount) {
          const uint32_t buf = rcount & 1;
          mbar_wait(cx.bar(BAR_RES_DONE + buf), (rcount >> 1) & 1, 15);
          if (elect_one()) {
#ifdef AP_L2_HINTS
            if (valid) tma_store_3d_hint(&tmUout, cx.base + G::RES_OFF + buf * A_BYTES, pass * 64, l0, b, l2_policy_evict_first());
#else
            if (valid) tma_store_3d(&tmUout, cx.base + G::RES_OFF + buf * A_BYTES, pass * 64, l0, b);   // rows >= L are clipped
#endif
            bulk_commit();
            bulk_wait_read<0>();                       // the store has read the buffer: the loader may refill it
            mbar_arrive(cx.bar(BAR_RES_FREE + buf));
          }
          __syncwarp();
        }
      }
      if (elect_one()) bulk_wait_all<0>();             // global writes of the last stores are complete before the CTA exits
      __syncwarp();
    }
  } else if (warp >= EPI_WARP0) {
    // ======================================================================================= epilogue (8 warps)
    const uint32_t c2_addr = cx.base + G::BIAS_OFF;
    setmaxnreg_inc<200>();
    if (((warp - EPI_WARP0) >> 2) == 0) k1_epilogue<G, 0, SAVE>(cx, p, tmem, tiles, &tmO, c2_addr);
    else k1_epilogue<G, 1, SAVE>(cx, p, tmem, tiles, &tmO, c2_addr);
  }
  tc_epilogue_teardown<CG>(tmem);
}

// ------------------------------------------------------------------------------------------------ k1 for the bf16x3 mode
// Same math as k1_layer with three MMAs per K step (Geo::NCOMBO) and both bf16 planes of every tensor.  Keeping o_hi AND o_lo
// of a whole tile staged for GEMM-2 would cost 128 KB of shared memory (a 3-stage ring); instead GEMM-2 is issued in two
// K halves -- G2a after the gate epilogue of chunk 0 (o channels 0..127), G2b after chunk 1 -- through one 64 KB staging
// tile (hi K-blocks in slots 0-1, lo in 2-3), which leaves the 5-stage ring of the bf16 kernel.  TMEM region X (columns
// 0..255) holds the GEMM-2 accumulator across both halves, region Y (256..511) every GEMM-1 chunk in turn (the gate epilogue
// frees it as soon as the accumulators are in registers).  Job order of the MMA warp per tile t:
//   G1c0(t) -> Y | G2b(t-1) -> X | G1c1(t) -> Y | G2a(t) -> X          epilogue: gate c0(t), residual(t-1), gate c1(t)
static_assert(BAR_COUNT * 8 + 8 <= 256, "barrier area");

template <class G, int HSEL, bool SAVE>
__device__ __forceinline__ void k1s_epilogue(const Ctx<G>& cx, const K1Params& p, const uint32_t tmem, const Tiles<G::CG>& tiles,
                                             const uint32_t c2_addr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q4 = warp & 3, etid = threadIdx.x - EPI_WARP0 * 32;
  const int row = q4 * 32 + lane;
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q4 * 32) << 16);
  const uint32_t row_off = row * 128, sw = row & 7;
  const float sqrt_half = 0.70710678118654752440f;
  const int oob_l0 = p.tiles_per_sample * TILE_M;
  const size_t lo_u = static_cast<size_t>(p.u_plane) * p.L * C, lo_o = static_cast<size_t>(p.o_plane) * p.L * C;
  uint32_t k = 0;   // gate jobs done (phase of the Y-region barriers)

  auto gate = [&](auto jc, uint32_t ti, bool valid, int b, int l0) {
    constexpr int J = decltype(jc)::value;
    mbar_wait(cx.bar(BAR_ACC_FULL + 1), k & 1, 70);
    tc_fence_after();
    uint32_t ta[2][32], sg[2][32];
#pragma unroll
    for (int gq = 0; gq < 2; ++gq) {
      tmem_ld_32x32b_x32(lane_addr + 256 + HSEL * 64 + gq * 32, ta[gq]);
      tmem_ld_32x32b_x32(lane_addr + 256 + 128 + HSEL * 64 + gq * 32, sg[gq]);
    }
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + 1);
    const bool live = valid && l0 + row < p.L;
    const size_t pos = static_cast<size_t>(b) * p.L + (live ? l0 + row : 0);
    uint32_t pk[2][4][4], pl[2][4][4], tsk[2][4];
#pragma unroll
    for (int gq = 0; gq < 2; ++gq)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c0 = i * 8 + 2 * e;
          const int cb = J * 256 + HSEL * 64 + gq * 32 + c0;
          // fp32-class gate (MUFU.TANH is good to ~2^-11 only); p.bd holds HALF the sigmoid bias
          const float a0 = __uint_as_float(ta[gq][c0]) + p.bd[cb], a1 = __uint_as_float(ta[gq][c0 + 1]) + p.bd[cb + 1];
          const float s0 = fmaf(2.f, p.bd[cb + 128], __uint_as_float(sg[gq][c0]));
          const float s1 = fmaf(2.f, p.bd[cb + 129], __uint_as_float(sg[gq][c0 + 1]));
          f32x2 o2;
          if constexpr (SAVE) {   // tanh and sigmoid separately: the backward pass keeps the gate's local derivatives
            const float t0 = tanh_exp(a0), t1 = tanh_exp(a1), g0 = sigmoid_exp(s0), g1 = sigmoid_exp(s1);
            o2 = pk2(t0 * g0, t1 * g1);
            tsk[0][e] = pack_bf16x2(g0 * fmaf(-t0, t0, 1.f), g1 * fmaf(-t1, t1, 1.f));
            tsk[1][e] = pack_bf16x2(t0 * g0 * (1.f - g0), t1 * g1 * (1.f - g1));
          } else if (kSplitGatePoly) {  // one quotient per channel; both exponentials on the FMA pipe (degree-5 polynomial: 1e-7)
            o2 = gate_poly2<5>(mul2(pk2(a0, a1), pk2(2.885390081777927f, 2.885390081777927f)),
                               mul2(pk2(s0, s1), pk2(-1.4426950408889634f, -1.4426950408889634f)));
          } else {                // one quotient per channel, two channels per instruction (packed fp32x2 arithmetic)
            o2 = gate_exp2(pk2(a0, a1), pk2(s0, s1));
          }
          split_bf16x2(o2, pk[gq][i][e], pl[gq][i][e]);
        }
        if constexpr (SAVE) {
          if (live) {
            uint16_t* td = p.ts_out + ((static_cast<size_t>(p.layer) * p.ts_chunk + b) * p.L + l0 + row) * 512 + J * 128 + HSEL * 64 +
                           gq * 32 + i * 8;
            st_global_v4(td, make_uint4(tsk[0][0], tsk[0][1], tsk[0][2], tsk[0][3]));
            st_global_v4(td + 256, make_uint4(tsk[1][0], tsk[1][1], tsk[1][2], tsk[1][3]));
          }
        }
      }
    // O (both planes) straight from registers: 4 x 32 bytes per thread and plane
    if (live) {
      uint16_t* od = p.o_out + (static_cast<size_t>(p.layer) * p.chunk_alloc * p.L + pos) * C + J * 128 + HSEL * 64;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const uint32_t (&q0)[4] = pk[v >> 1][(v & 1) * 2], (&q1)[4] = pk[v >> 1][(v & 1) * 2 + 1];
        const uint32_t h8[8] = {q0[0], q0[1], q0[2], q0[3], q1[0], q1[1], q1[2], q1[3]};
        st_global_v8(od + v * 16, h8);
        const uint32_t (&r0)[4] = pl[v >> 1][(v & 1) * 2], (&r1)[4] = pl[v >> 1][(v & 1) * 2 + 1];
        const uint32_t l8[8] = {r0[0], r0[1], r0[2], r0[3], r1[0], r1[1], r1[2], r1[3]};
        st_global_v8(od + lo_o + v * 16, l8);
      }
    }
    if (!p.last) {
      // the staging tile is free once the GEMM-2 half that read it has completed: G2b of the previous tile before chunk 0,
      // G2a of this tile before chunk 1
      if (J == 0) {
        if (ti > 0) mbar_wait(cx.bar(BAR_ACC_FULL + 0), (ti - 1) & 1, 71);
      } else {
        mbar_wait(cx.bar(BAR_G2A_DONE), ti & 1, 72);
      }
      tc_fence_after();
      const uint32_t hi_base = cx.out_kb(HSEL) + row_off, lo_base = cx.out_kb(2 + HSEL) + row_off;
#pragma unroll
      for (int gq = 0; gq < 2; ++gq)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          st_shared_v4(hi_base + (((gq * 4 + i) ^ sw) << 4), make_uint4(pk[gq][i][0], pk[gq][i][1], pk[gq][i][2], pk[gq][i][3]));
          st_shared_v4(lo_base + (((gq * 4 + i) ^ sw) << 4), make_uint4(pl[gq][i][0], pl[gq][i][1], pl[gq][i][2], pl[gq][i][3]));
        }
      fence_proxy_async_smem();
      named_bar_sync(1, EPI_THREADS);
      if (etid == 0) cx.arrive_leader(BAR_OUT_READY + J);
    }
    ++k;
  };

  auto residual = [&](uint32_t t_idx, bool valid, int b, int l0) {
    const bool live = valid && l0 + row < p.L;
    const size_t goff = (static_cast<size_t>(b) * p.L + (live ? l0 + row : 0)) * C + HSEL * 128;
    mbar_wait(cx.bar(BAR_ACC_FULL + 0), t_idx & 1, 73);
    tc_fence_after();
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) {
      uint32_t acc[32];
      tmem_ld_32x32b_x32(lane_addr + HSEL * 128 + gq * 32, acc);
      tmem_ld_wait();
      if (gq == 3) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + 0);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int ch = HSEL * 128 + gq * 32 + i * 16;
        uint32_t pk[8], pl[8], uh[8], ul[8];
        if (live) {
          ld_global_v8(p.u_in + goff + (gq * 2 + i) * 16, uh);
          ld_global_v8(p.u_in + lo_u + goff + (gq * 2 + i) * 16, ul);
        }
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) {
          const uint4 cv = ld_shared_v4(c2_addr + (ch + e4 * 4) * 4);
          const float cc[4] = {__uint_as_float(cv.x), __uint_as_float(cv.y), __uint_as_float(cv.z), __uint_as_float(cv.w)};
          const int a0 = i * 16 + e4 * 4;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t wh = uh[e4 * 2 + h], wl = ul[e4 * 2 + h];
            const f32x2 u2 = add2(pk2(bf16_lo(wh), bf16_hi(wh)), pk2(bf16_lo(wl), bf16_hi(wl)));
            const f32x2 r2 = pk2(__uint_as_float(acc[a0 + 2 * h]), __uint_as_float(acc[a0 + 2 * h + 1]));
            const f32x2 v2 = fma2(add2(u2, r2), pk2(sqrt_half, sqrt_half), pk2(cc[2 * h], cc[2 * h + 1]));
            split_bf16x2(v2, pk[e4 * 2 + h], pl[e4 * 2 + h]);
          }
        }
        if (live) {
          st_global_v8(p.u_out + goff + (gq * 2 + i) * 16, pk);
          st_global_v8(p.u_out + lo_u + goff + (gq * 2 + i) * 16, pl);
        }
      }
    }
  };

  uint32_t ti = 0;
  bool pvalid = false;
  int pb = 0, pl0 = 0;
  for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
    const bool valid = tile < p.n_tiles;
    const int b = valid ? tile / p.tiles_per_sample : 0;
    const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
    gate(std::integral_constant<int, 0>{}, ti, valid, b, l0);
    if (!p.last && ti > 0) residual(ti - 1, pvalid, pb, pl0);
    gate(std::integral_constant<int, 1>{}, ti, valid, b, l0);
    pvalid = valid, pb = b, pl0 = l0;
  }
  if (!p.last && ti > 0) residual(ti - 1, pvalid, pb, pl0);
}

template <bool SAVE>
__global__ void __launch_bounds__(NTHREADS, 1)
k1_split(const __grid_constant__ CUtensorMap tmUin, const __grid_constant__ CUtensorMap tmWd,
         const __grid_constant__ CUtensorMap tmWr, const __grid_constant__ K1Params p) {
  using G = Geo<2, 1, 2>;
  constexpr int CG = 2;
  Ctx<G> cx;
  uint8_t* gen;
  const uint32_t tmem = tc_prologue<G>(cx, gen, 1);
  {
    float* s_c2 = reinterpret_cast<float*>(gen + G::BIAS_OFF);
    for (int i = threadIdx.x; i < C; i += NTHREADS) s_c2[i] = fmaf(p.b_res[i], 0.70710678118654752440f, p.p_next[i]);
  }
  if (threadIdx.x == 0) prefetch_tmap(&tmUin), prefetch_tmap(&tmWd), prefetch_tmap(&tmWr);
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  const Tiles<CG> tiles(p.n_tiles, cx.rank);
  const int oob_l0 = p.tiles_per_sample * TILE_M;

  if (warp < EPI_WARP0) setmaxnreg_dec<96>();
  if (warp == 0) {
    // ======================================================================================= TMA producer
    RingPos<G::NSTAGE> it;
    uint32_t ti = 0;
    auto load_g1 = [&](int j, bool valid, int b, int l0) {
      for (int tap = 0; tap < 3; ++tap)
        for (int kb = 0; kb < 4; ++kb)
          for (int cmb = 0; cmb < 3; ++cmb, ++it) {
            const uint32_t s = it.s, ph = it.ph;
            cx.wait_empty(s, ph, 74);
            if (elect_one()) {
              cx.arm(s, G::STAGE_BYTES);
              cx.load_a(s, &tmUin, kb * 64, valid ? l0 + (tap - 1) * p.dilation : oob_l0, b + (cmb == 1 ? p.u_plane : 0));
              cx.load_b(s, &tmWd, tap * C + kb * 64, (p.layer * 2 + j) * 256 + (cmb == 2 ? p.wd_plane : 0));
            }
            __syncwarp();
          }
    };
    auto load_wr = [&](int pass) {   // res-conv weights of K-blocks 2 pass, 2 pass + 1
      for (int kbp = 0; kbp < 2; ++kbp)
        for (int cmb = 0; cmb < 3; ++cmb, ++it) {
          const uint32_t s = it.s, ph = it.ph;
          cx.wait_empty(s, ph, 75);
          if (elect_one()) {
            cx.arm(s, G::B_BYTES);
            cx.load_b(s, &tmWr, (pass * 2 + kbp) * 64, p.layer * 256 + (cmb == 2 ? p.wr_plane : 0));
          }
          __syncwarp();
        }
    };
    for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
      const bool valid = tile < p.n_tiles;
      const int b = valid ? tile / p.tiles_per_sample : 0;
      const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
      load_g1(0, valid, b, l0);
      if (!p.last && ti > 0) load_wr(1);
      load_g1(1, valid, b, l0);
      if (!p.last) load_wr(0);
    }
    if (!p.last && ti > 0) load_wr(1);
  } else if (warp == 1) {
    // ======================================================================================= MMA issuer (leader CTA)
    if (cx.rank == 0) {
      RingPos<G::NSTAGE> it;
      uint32_t k = 0, ti = 0;
      auto gemm1 = [&]() {
        mbar_wait(cx.bar(BAR_ACC_EMPTY + 1), (k & 1) ^ 1, 76);
        tc_fence_after();
        for (int kblk = 0; kblk < 36; ++kblk, ++it) {
          const uint32_t s = it.s, ph = it.ph;
          mbar_wait(cx.bar(BAR_FULL + s), ph, 77);
          tc_fence_after();
          if (elect_one()) {
            cx.mma_kblock(tmem + 256, cx.stage_a(s), cx.stage_b(s), kblk == 0);
            cx.commit(BAR_EMPTY + s);
          }
          __syncwarp();
        }
        if (elect_one()) cx.commit(BAR_ACC_FULL + 1);
        __syncwarp();
        ++k;
      };
      auto gemm2_half = [&](uint32_t t_idx, int pass) {
        if (pass == 0) {   // a new accumulation: region X must have been drained by the residual epilogue of tile t-1
          mbar_wait(cx.bar(BAR_ACC_EMPTY + 0), (t_idx & 1) ^ 1, 78);
        }
        mbar_wait(cx.bar(BAR_OUT_READY + pass), t_idx & 1, 79);
        tc_fence_after();
        for (int kbp = 0; kbp < 2; ++kbp)
          for (int cmb = 0; cmb < 3; ++cmb, ++it) {
            const uint32_t s = it.s, ph = it.ph;
            mbar_wait(cx.bar(BAR_FULL + s), ph, 80);
            tc_fence_after();
            if (elect_one()) {
              cx.mma_kblock(tmem, cx.out_kb((cmb == 1 ? 2 : 0) + kbp), cx.stage_b(s), pass == 0 && kbp == 0 && cmb == 0);
              cx.commit(BAR_EMPTY + s);
            }
            __syncwarp();
          }
        if (elect_one()) cx.commit(pass == 0 ? BAR_G2A_DONE : BAR_ACC_FULL + 0);
        __syncwarp();
      };
      for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
        gemm1();
        if (!p.last && ti > 0) gemm2_half(ti - 1, 1);
        gemm1();
        if (!p.last) gemm2_half(ti, 0);
      }
      if (!p.last && ti > 0) gemm2_half(ti - 1, 1);
    }
  } else if (warp >= EPI_WARP0) {
    setmaxnreg_inc<200>();
    const uint32_t c2_addr = cx.base + G::BIAS_OFF;
    if (((warp - EPI_WARP0) >> 2) == 0) k1s_epilogue<G, 0, SAVE>(cx, p, tmem, tiles, c2_addr);
    else k1s_epilogue<G, 1, SAVE>(cx, p, tmem, tiles, c2_addr);
  }
  tc_epilogue_teardown<CG>(tmem);
}

// ------------------------------------------------------------------------------------------------ k2: skip sum + head
struct K2Params {
  int n_tiles, tiles_per_sample, L, num_layers, chunk_alloc;
  float scale;           // sqrt(1/N)
  const float* bf2;      // [1]
  float* eps;            // [B][L]
  int o_plane, ws_plane, wf_plane;   // hi/lo split mode (DT = 2): offsets of the lo planes in the O / Ws / Wf tensor maps
  uint32_t* mask_out;                // SAVE kernels (backward pass): [B][L][8] bit c of the row = (head pre-activation c > 0)
  // fused one-shot denoise (certification): p.eps holds x_in on entry and receives x0 = x0_a * x_in - x0_b * eps
  int fuse_x0;
  float x0_a, x0_b;
  // per-channel vectors in the kernel-parameter constant bank (constant operands of the epilogue FMAs):
  float bskip_scaled[256];   // (sum over layers of the skip-conv biases) * sqrt(1/N)
  float bf1[256];            // final_conv.0 bias
  float wf2[256];            // final_conv.2 (256 -> 1) weight
};
enum { BAR2_S_READY = BAR_OUT_READY };

// epilogue of k2 for the column half HSEL (channels HSEL*128 .. +127)
template <class G, int HSEL, bool SAVE>
__device__ __forceinline__ void k2_epilogue(const Ctx<G>& cx, const K2Params& p, const uint32_t tmem,
                                            const Tiles<G::CG>& tiles, float* s_part) {
  constexpr int DT = G::DT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q4 = warp & 3, etid = threadIdx.x - EPI_WARP0 * 32;
  const int row = q4 * 32 + lane;
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q4 * 32) << 16);
  const uint32_t row_off = row * 128, sw = row & 7;
  const int oob_l0 = p.tiles_per_sample * TILE_M;
  uint32_t ti = 0;
  for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
    const bool valid = tile < p.n_tiles;
    const int b = valid ? tile / p.tiles_per_sample : 0;
    const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
    // ---- skip sum -> s (bf16, A operand of the head GEMM).  The previous tile's head MMAs finished reading the
    //      staging region before its ACC_FULL[1] fired, which this thread has already waited on.
    mbar_wait(cx.bar(BAR_ACC_FULL + 0), ti & 1, 28);
    tc_fence_after();
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) {
      uint32_t acc[32];
      tmem_ld_32x32b_x32(lane_addr + HSEL * 128 + gq * 32, acc);
      tmem_ld_wait();
      const uint32_t kb_base = cx.out_kb(HSEL * 2 + (gq >> 1)) + row_off;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t pk[4], pl[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c0 = HSEL * 128 + gq * 32 + i * 8 + 2 * e;
          const float v0 = fmaf(__uint_as_float(acc[i * 8 + 2 * e]), p.scale, p.bskip_scaled[c0]);
          const float v1 = fmaf(__uint_as_float(acc[i * 8 + 2 * e + 1]), p.scale, p.bskip_scaled[c0 + 1]);
          pk[e] = pack2<DT>(v0, v1);
          if constexpr (G::SPLIT) pl[e] = pack_bf16x2(v0 - bf16_lo(pk[e]), v1 - bf16_hi(pk[e]));
        }
        st_shared_v4(kb_base + ((((gq & 1) * 4 + i) ^ sw) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
        if constexpr (G::SPLIT)
          st_shared_v4(kb_base + 4 * A_BYTES + ((((gq & 1) * 4 + i) ^ sw) << 4), make_uint4(pl[0], pl[1], pl[2], pl[3]));
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
    float dot = 0.f;
    uint32_t mw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) {
      uint32_t acc[32];
      tmem_ld_32x32b_x32(lane_addr + 256 + HSEL * 128 + gq * 32, acc);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float pre = __uint_as_float(acc[e]) + p.bf1[HSEL * 128 + gq * 32 + e];
        dot = fmaf(fmaxf(pre, 0.f), p.wf2[HSEL * 128 + gq * 32 + e], dot);
        if constexpr (SAVE) mw[gq] |= (pre > 0.f ? 1u : 0u) << e;
      }
    }
    if constexpr (SAVE) {
      if (valid && l0 + row < p.L)
        st_global_v4(p.mask_out + (static_cast<size_t>(b) * p.L + l0 + row) * 8 + HSEL * 4, make_uint4(mw[0], mw[1], mw[2], mw[3]));
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + 1);
    if (HSEL == 1) s_part[row] = dot;
    named_bar_sync(1, EPI_THREADS);
    if (HSEL == 0 && valid && l0 + row < p.L) {
      float* dst = p.eps + static_cast<size_t>(b) * p.L + l0 + row;
      const float e = dot + s_part[row] + p.bf2[0];
      // _predict_x0_from_eps (diffwave_ddpm.py:195-205) in PredictX0Op's rounding order
      *dst = p.fuse_x0 ? __fsub_rn(__fmul_rn(p.x0_a, *dst), __fmul_rn(p.x0_b, e)) : e;
    }
    named_bar_sync(1, EPI_THREADS);   // s_part is rewritten by the next tile
  }
}

template <int CG, int DT, bool SAVE = false>
__global__ void __launch_bounds__(NTHREADS, 1)
k2_head(const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmWs,
        const __grid_constant__ CUtensorMap tmWf, const __grid_constant__ K2Params p) {
  using G = Geo<CG, 2, DT>;
  Ctx<G> cx;
  uint8_t* gen;
  const uint32_t tmem = tc_prologue<G>(cx, gen, 1);
  float* s_part = reinterpret_cast<float*>(gen + G::BIAS_OFF);   // [128] partial dots of the upper column half
  if (threadIdx.x == 0) prefetch_tmap(&tmO), prefetch_tmap(&tmWs), prefetch_tmap(&tmWf);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = p.num_layers * 4 * G::NCOMBO;
  const Tiles<CG> tiles(p.n_tiles, cx.rank);
  const int oob_l0 = p.tiles_per_sample * TILE_M;

  if (warp == 0) {
    {   // whole warp, one elected lane issues (see k1_layer)
      RingPos<G::NSTAGE> it;
      for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride) {
        const bool valid = tile < p.n_tiles;
        const int b = valid ? tile / p.tiles_per_sample : 0;
        const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
        for (int n = 0; n < p.num_layers; ++n)
          for (int kb = 0; kb < 4; ++kb)
            for (int cmb = 0; cmb < G::NCOMBO; ++cmb, ++it) {
              const uint32_t s = it.s, ph = it.ph;
              cx.wait_empty(s, ph, 21);
              if (elect_one()) {
                cx.arm(s, G::STAGE_BYTES);
                cx.load_a(s, &tmO, kb * 64, l0, n * p.chunk_alloc + b + (cmb == 1 ? p.o_plane : 0));
                cx.load_b(s, &tmWs, kb * 64, n * 256 + (cmb == 2 ? p.ws_plane : 0));
              }
              __syncwarp();
            }
        for (int kb = 0; kb < 4; ++kb)
          for (int cmb = 0; cmb < G::NCOMBO; ++cmb, ++it) {
            const uint32_t s = it.s, ph = it.ph;
            cx.wait_empty(s, ph, 22);
            if (elect_one()) {
              cx.arm(s, G::B_BYTES);
              cx.load_b(s, &tmWf, kb * 64, cmb == 2 ? p.wf_plane : 0);
            }
            __syncwarp();
          }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (cx.rank == 0) {   // whole warp, one elected lane issues (see k1_layer)
      RingPos<G::NSTAGE> it;
      uint32_t ti = 0;
      for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
        mbar_wait(cx.bar(BAR_ACC_EMPTY + 0), (ti & 1) ^ 1, 23);
        tc_fence_after();
        for (int kblk = 0; kblk < nkb; ++kblk, ++it) {
          const uint32_t s = it.s, ph = it.ph;
          mbar_wait(cx.bar(BAR_FULL + s), ph, 24);
          tc_fence_after();
          if (elect_one()) {
            cx.mma_kblock(tmem, cx.stage_a(s), cx.stage_b(s), kblk == 0);
            cx.commit(BAR_EMPTY + s);
          }
          __syncwarp();
        }
        if (elect_one()) cx.commit(BAR_ACC_FULL + 0);
        __syncwarp();
        mbar_wait(cx.bar(BAR_ACC_EMPTY + 1), (ti & 1) ^ 1, 25);
        mbar_wait(cx.bar(BAR2_S_READY), ti & 1, 26);
        tc_fence_after();
        for (int kb = 0; kb < 4; ++kb)
          for (int cmb = 0; cmb < G::NCOMBO; ++cmb, ++it) {
            const uint32_t s = it.s, ph = it.ph;
            mbar_wait(cx.bar(BAR_FULL + s), ph, 27);
            tc_fence_after();
            if (elect_one()) {
              cx.mma_kblock(tmem + 256, cx.out_kb(kb + (cmb == 1 ? 4 : 0)), cx.stage_b(s), kb == 0 && cmb == 0);
              cx.commit(BAR_EMPTY + s);
            }
            __syncwarp();
          }
        if (elect_one()) cx.commit(BAR_ACC_FULL + 1);
        __syncwarp();
      }
    }
  } else if (warp >= EPI_WARP0) {
    if (((warp - EPI_WARP0) >> 2) == 0) k2_epilogue<G, 0, SAVE>(cx, p, tmem, tiles, s_part);
    else k2_epilogue<G, 1, SAVE>(cx, p, tmem, tiles, s_part);
  }
  tc_epilogue_teardown<CG>(tmem);
}

// ------------------------------------------------------------------------------------------------ backward (VJP wrt x)
// g_x = (d eps / d x)^T g_eps for the bf16 network: the reference's DiffWave.forward is differentiable (white-box attacks
// call loss.backward() through the purifier, robustness_eval/white_box_attack.py:438); this is the same chain rule as
// autograd over WaveNet.py:75-97,120-135,164-172, as three streaming tcgen05 GEMMs per layer pass:
//   MODE 0 (head)      g_s  = (g_pre . Wf) * sqrt(1/N)                      g_pre = relu'(y) * w2 * g_eps   (K = 256)
//   MODE 1 (layer, 1)  g_o  = g_s . Ws_n + g_u' . Wr_n * sqrt(.5)           (K = 512; K = 256 for the last layer)
//                      g_a  = [ g_o * S (1 - T^2) | g_o * T S (1 - S) ]        (the two local derivatives are saved by k1<SAVE>)
//   MODE 2 (layer, 2)  g_u  = sum_tap g_a[l - (tap-1) d] . Wd_tap^T + sqrt(.5) g_u'   (K = 3 * 512: the transposed dilated conv)
// Gradients travel between the kernels as bf16 channels-last tensors (TMA sources of the next GEMM), accumulate in fp32.
struct KbParams {
  int n_tiles, tiles_per_sample, L, dilation, last;
  int nkb;             // K blocks per tap: 4 (head, last layer's MODE 1) or 8
  int b_row0;          // first row of this layer's weights in the B tensor map
  float scale;         // MODE 0: sqrt(1/N)
  const uint16_t* ts;  // MODE 1: [chunk][L][512] d o / d a_t | d o / d a_s of this layer's gate
  const uint16_t* g_next;  // MODE 2: [chunk][L][256] g_u of layer n+1 (unused when last)
  uint16_t* out;       // MODE 0: g_s [..][256]; MODE 1: g_a [..][512]; MODE 2: g_u [..][256]
};

template <class G, int MODE, int HSEL>
__device__ __forceinline__ void kb_epilogue(const Ctx<G>& cx, const KbParams& p, const uint32_t tmem, const Tiles<G::CG>& tiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q4 = warp & 3;
  const int row = q4 * 32 + lane;
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q4 * 32) << 16);
  const float sqrt_half = 0.70710678118654752440f;
  uint32_t ti = 0;
  for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
    const bool valid = tile < p.n_tiles;
    const int b = valid ? tile / p.tiles_per_sample : 0;
    const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : 0;
    const bool live = valid && l0 + row < p.L;
    const size_t pos = static_cast<size_t>(b) * p.L + (live ? l0 + row : 0);
    const uint32_t r = ti & 1;
    mbar_wait(cx.bar(BAR_ACC_FULL + r), (ti >> 1) & 1, 60);
    tc_fence_after();
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) {
      uint32_t acc[32];
      tmem_ld_32x32b_x32(lane_addr + r * 256 + HSEL * 128 + gq * 32, acc);
      tmem_ld_wait();
      if (gq == 3) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + r);
      }
      const int ch = HSEL * 128 + gq * 32;
#pragma unroll
      for (int i = 0; i < 2; ++i) {       // 16 channels = 32 bytes of bf16
        uint32_t o0[8], o1[8];
        if constexpr (MODE == 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            o0[e] = pack_bf16x2(__uint_as_float(acc[i * 16 + 2 * e]) * p.scale, __uint_as_float(acc[i * 16 + 2 * e + 1]) * p.scale);
          if (live) st_global_v8(p.out + pos * C + ch + i * 16, o0);
        } else if constexpr (MODE == 1) {
          uint32_t tv[8], sv[8];
          if (live) {
            ld_global_v8(p.ts + pos * 512 + ch + i * 16, tv);
            ld_global_v8(p.ts + pos * 512 + 256 + ch + i * 16, sv);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float g0 = __uint_as_float(acc[i * 16 + 2 * e]), g1 = __uint_as_float(acc[i * 16 + 2 * e + 1]);
            o0[e] = pack_bf16x2(g0 * bf16_lo(tv[e]), g1 * bf16_hi(tv[e]));          // d/d a_t
            o1[e] = pack_bf16x2(g0 * bf16_lo(sv[e]), g1 * bf16_hi(sv[e]));          // d/d a_s
          }
          if (live) {
            st_global_v8(p.out + pos * 512 + ch + i * 16, o0);
            st_global_v8(p.out + pos * 512 + 256 + ch + i * 16, o1);
          }
        } else {
          uint32_t gv[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          if (live && !p.last) ld_global_v8(p.g_next + pos * C + ch + i * 16, gv);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            o0[e] = pack_bf16x2(fmaf(bf16_lo(gv[e]), sqrt_half, __uint_as_float(acc[i * 16 + 2 * e])),
                                fmaf(bf16_hi(gv[e]), sqrt_half, __uint_as_float(acc[i * 16 + 2 * e + 1])));
          if (live) st_global_v8(p.out + pos * C + ch + i * 16, o0);
        }
      }
    }
  }
}

// tmA0: MODE 0 g_pre, MODE 1 g_s, MODE 2 g_a (512 channels); tmA1: MODE 1 g_u of layer n+1; tmB: the transposed weights
template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1)
k_bwd(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
      const __grid_constant__ CUtensorMap tmB, const __grid_constant__ KbParams p) {
  using G = Geo<2, 2, 0>;
  constexpr int CG = 2;
  Ctx<G> cx;
  uint8_t* gen;
  const uint32_t tmem = tc_prologue<G>(cx, gen, 1);
  if (threadIdx.x == 0) prefetch_tmap(&tmA0), prefetch_tmap(&tmA1), prefetch_tmap(&tmB);
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  const Tiles<CG> tiles(p.n_tiles, cx.rank);
  const int oob_l0 = p.tiles_per_sample * TILE_M;
  const int ntap = MODE == 2 ? 3 : 1;

  if (warp == 0) {
    RingPos<G::NSTAGE> it;
    for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride) {
      const bool valid = tile < p.n_tiles;
      const int b = valid ? tile / p.tiles_per_sample : 0;
      const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
      for (int tap = 0; tap < ntap; ++tap)
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          const uint32_t s = it.s, ph = it.ph;
          cx.wait_empty(s, ph, 61);
          if (elect_one()) {
            cx.arm(s, G::STAGE_BYTES);
            const int row = !valid ? oob_l0 : (MODE == 2 ? l0 - (tap - 1) * p.dilation : l0);
            if (MODE == 1 && kb >= 4) cx.load_a(s, &tmA1, (kb - 4) * 64, row, b);
            else cx.load_a(s, &tmA0, kb * 64, row, b);
            cx.load_b(s, &tmB, (tap * p.nkb + kb) * 64, p.b_row0);
          }
          __syncwarp();
        }
    }
  } else if (warp == 1) {
    if (cx.rank == 0) {
      RingPos<G::NSTAGE> it;
      uint32_t ti = 0;
      const int nk = ntap * p.nkb;
      for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
        const uint32_t r = ti & 1;
        mbar_wait(cx.bar(BAR_ACC_EMPTY + r), ((ti >> 1) & 1) ^ 1, 62);
        tc_fence_after();
        for (int k = 0; k < nk; ++k, ++it) {
          const uint32_t s = it.s, ph = it.ph;
          mbar_wait(cx.bar(BAR_FULL + s), ph, 63);
          tc_fence_after();
          if (elect_one()) {
            cx.mma_kblock(tmem + r * 256, cx.stage_a(s), cx.stage_b(s), k == 0);
            cx.commit(BAR_EMPTY + s);
          }
          __syncwarp();
        }
        if (elect_one()) cx.commit(BAR_ACC_FULL + r);
        __syncwarp();
      }
    }
  } else if (warp >= EPI_WARP0) {
    if (((warp - EPI_WARP0) >> 2) == 0) kb_epilogue<G, MODE, 0>(cx, p, tmem, tiles);
    else kb_epilogue<G, MODE, 1>(cx, p, tmem, tiles);
  }
  tc_epilogue_teardown<CG>(tmem);
}

// ---- MODE 2 of layer l fused with MODE 1 of layer l - 1 (the two are separate launches above because g_a is needed at l +- d;
// MODE 1 is pointwise in the position, so it can consume g_u(l) of the SAME tile straight from shared memory).  One launch per layer
// boundary with the structure of the forward's k1_layer: job A (K = 1536, TMEM region X) -> epilogue A adds sqrt(.5) g_u(l+1),
// writes g_u(l) (the next launch's residual term) and stages it as bf16 K-blocks -> job B (K = 512: g_s by TMA | staged g_u,
// region Y) -> epilogue B multiplies by the saved gate derivatives and writes g_a(l-1).  The MMA warp issues A(t), B(t-1), A(t+1),
// ...: the HBM-bound epilogue of one tile runs under the tensor-bound job of the next, instead of one after the other in two
// launches (0.83 ms per layer and 32 waveforms = HBM time + tensor time).
struct KfParams {
  int n_tiles, tiles_per_sample, L, dilation;
  int has_next;            // layer l < N - 1: g_u(l+1) exists
  int a_row0, b_row0;      // first rows of layer l in the WdT map and of layer l - 1 in the Wb map
  const uint16_t* ts;      // [chunk][L][512] gate derivatives of layer l - 1
  const uint16_t* g_next;  // [chunk][L][256] g_u(l+1)
  uint16_t* gu_out;        // [chunk][L][256] g_u(l)
  uint16_t* ga_out;        // [chunk][L][512] g_a(l-1)
  int ts_row0;             // k_bwd_fused_s: (l - 1) * chunk, first waveform of layer l - 1 in the derivative tensor map
};

template <class G, int HSEL>
__device__ __forceinline__ void kf_epilogue(const Ctx<G>& cx, const KfParams& p, const uint32_t tmem, const Tiles<G::CG>& tiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q4 = warp & 3, etid = threadIdx.x - EPI_WARP0 * 32;
  const int row = q4 * 32 + lane;
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q4 * 32) << 16);
  const uint32_t row_off = row * 128, sw = row & 7;
  const float sqrt_half = 0.70710678118654752440f;

  // epilogue B of unit t_idx: g_a(l-1) = acc_Y * (d o / d a_t | d o / d a_s)
  auto epi_b = [&](uint32_t t_idx, bool live, size_t pos) {
    mbar_wait(cx.bar(BAR_ACC_FULL + 1), t_idx & 1, 64);
    tc_fence_after();
    // the derivative rows of 64 channels (eight 32-byte loads) are requested together, ahead of the TMEM reads: one exposed
    // memory latency per half instead of one per 16 channels (the tcgen05.wait below is a compiler barrier for loads)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t tv[4][8], sv[4][8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int e = 0; e < 8; ++e) tv[j][e] = sv[j][e] = 0u;
        if (live) {
          ld_global_v8(p.ts + pos * 512 + HSEL * 128 + half * 64 + j * 16, tv[j]);
          ld_global_v8(p.ts + pos * 512 + 256 + HSEL * 128 + half * 64 + j * 16, sv[j]);
        }
      }
#pragma unroll
      for (int g2 = 0; g2 < 2; ++g2) {
        const int gq = half * 2 + g2;
        uint32_t acc[32];
        tmem_ld_32x32b_x32(lane_addr + 256 + HSEL * 128 + gq * 32, acc);
        tmem_ld_wait();
        if (gq == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + 1);
        }
        const int ch = HSEL * 128 + gq * 32;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          uint32_t o0[8], o1[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float g0 = __uint_as_float(acc[i * 16 + 2 * e]), g1 = __uint_as_float(acc[i * 16 + 2 * e + 1]);
            o0[e] = pack_bf16x2(g0 * bf16_lo(tv[g2 * 2 + i][e]), g1 * bf16_hi(tv[g2 * 2 + i][e]));
            o1[e] = pack_bf16x2(g0 * bf16_lo(sv[g2 * 2 + i][e]), g1 * bf16_hi(sv[g2 * 2 + i][e]));
          }
          if (live) {
            st_global_v8(p.ga_out + pos * 512 + ch + i * 16, o0);
            st_global_v8(p.ga_out + pos * 512 + 256 + ch + i * 16, o1);
          }
        }
      }
    }
  };

  uint32_t ti = 0;
  bool plive = false;
  size_t ppos = 0;
  for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
    const bool valid = tile < p.n_tiles;
    const int b = valid ? tile / p.tiles_per_sample : 0;
    const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : 0;
    const bool live = valid && l0 + row < p.L;
    const size_t pos = static_cast<size_t>(b) * p.L + (live ? l0 + row : 0);
    // ---- epilogue A: g_u(l) = acc_X + sqrt(.5) g_u(l+1); rows past the end of a waveform stage zeros.  The row of g_u(l+1)
    //      (eight 32-byte loads) is requested before the wait for the accumulator: its latency hides behind job A.
    uint32_t gn[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 8; ++e) gn[j][e] = 0u;
      if (live && p.has_next) ld_global_v8(p.g_next + pos * C + HSEL * 128 + j * 16, gn[j]);
    }
    mbar_wait(cx.bar(BAR_ACC_FULL + 0), ti & 1, 65);
    tc_fence_after();
    uint32_t pk[4][16];
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) {
      uint32_t acc[32];
      tmem_ld_32x32b_x32(lane_addr + HSEL * 128 + gq * 32, acc);
      tmem_ld_wait();
      if (gq == 3) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + 0);
      }
      const int ch = HSEL * 128 + gq * 32;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint32_t (&gv)[8] = gn[gq * 2 + i];
        uint32_t o0[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          o0[e] = live ? pack_bf16x2(fmaf(bf16_lo(gv[e]), sqrt_half, __uint_as_float(acc[i * 16 + 2 * e])),
                                     fmaf(bf16_hi(gv[e]), sqrt_half, __uint_as_float(acc[i * 16 + 2 * e + 1])))
                       : 0u;
          pk[gq][i * 8 + e] = o0[e];
        }
        if (live) st_global_v8(p.gu_out + pos * C + ch + i * 16, o0);
      }
    }
    // the staging tile is free once job B of the previous unit has completed
    if (ti > 0) {
      mbar_wait(cx.bar(BAR_ACC_FULL + 1), (ti - 1) & 1, 66);
      tc_fence_after();
    }
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) {
      const uint32_t kb_base = cx.out_kb(HSEL * 2 + (gq >> 1)) + row_off;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        st_shared_v4(kb_base + ((((gq & 1) * 4 + i) ^ sw) << 4), make_uint4(pk[gq][i * 4], pk[gq][i * 4 + 1], pk[gq][i * 4 + 2], pk[gq][i * 4 + 3]));
    }
    fence_proxy_async_smem();
    named_bar_sync(1, EPI_THREADS);
    if (etid == 0) cx.arrive_leader(BAR_OUT_READY);
    // ---- epilogue B of the previous unit (its job B was issued behind this unit's job A)
    if (ti > 0) epi_b(ti - 1, plive, ppos);
    plive = live, ppos = pos;
  }
  if (ti > 0) epi_b(ti - 1, plive, ppos);
}

// tmGa: g_a(l) (512 channels); tmGs: g_s; tmWdT / tmWb: the transposed weights
__global__ void __launch_bounds__(NTHREADS, 1)
k_bwd_fused(const __grid_constant__ CUtensorMap tmGa, const __grid_constant__ CUtensorMap tmGs, const __grid_constant__ CUtensorMap tmWdT,
            const __grid_constant__ CUtensorMap tmWb, const __grid_constant__ KfParams p) {
  using G = Geo<2, 2, 0>;
  constexpr int CG = 2;
  Ctx<G> cx;
  uint8_t* gen;
  const uint32_t tmem = tc_prologue<G>(cx, gen, 1);
  if (threadIdx.x == 0) prefetch_tmap(&tmGa), prefetch_tmap(&tmGs), prefetch_tmap(&tmWdT), prefetch_tmap(&tmWb);
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  const Tiles<CG> tiles(p.n_tiles, cx.rank);
  const int oob_l0 = p.tiles_per_sample * TILE_M;

  if (warp < EPI_WARP0) setmaxnreg_dec<96>();
  if (warp == 0) {
    RingPos<G::NSTAGE> it;
    auto load_job_b = [&](bool valid, int b, int l0) {
      for (int kb = 0; kb < 8; ++kb, ++it) {
        const uint32_t s = it.s, ph = it.ph;
        cx.wait_empty(s, ph, 67);
        if (elect_one()) {
          if (kb < 4) {      // g_s tile and the Ws rows
            cx.arm(s, G::STAGE_BYTES);
            cx.load_a(s, &tmGs, kb * 64, valid ? l0 : oob_l0, b);
          } else {           // Wr rows only: the A operand is the staged g_u
            cx.arm(s, G::B_BYTES);
          }
          cx.load_b(s, &tmWb, kb * 64, p.b_row0);
        }
        __syncwarp();
      }
    };
    uint32_t ti = 0;
    bool pvalid = false;
    int pb = 0, pl0 = 0;
    for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
      const bool valid = tile < p.n_tiles;
      const int b = valid ? tile / p.tiles_per_sample : 0;
      const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
      for (int tap = 0; tap < 3; ++tap)
        for (int kb = 0; kb < 8; ++kb, ++it) {
          const uint32_t s = it.s, ph = it.ph;
          cx.wait_empty(s, ph, 68);
          if (elect_one()) {
            cx.arm(s, G::STAGE_BYTES);
            cx.load_a(s, &tmGa, kb * 64, valid ? l0 - (tap - 1) * p.dilation : oob_l0, b);
            cx.load_b(s, &tmWdT, (tap * 8 + kb) * 64, p.a_row0);
          }
          __syncwarp();
        }
      if (ti > 0) load_job_b(pvalid, pb, pl0);
      pvalid = valid, pb = b, pl0 = l0;
    }
    if (ti > 0) load_job_b(pvalid, pb, pl0);
  } else if (warp == 1) {
    if (cx.rank == 0) {
      RingPos<G::NSTAGE> it;
      auto job_b = [&](uint32_t t_idx) {
        mbar_wait(cx.bar(BAR_ACC_EMPTY + 1), (t_idx & 1) ^ 1, 69);
        tc_fence_after();
        for (int kb = 0; kb < 8; ++kb, ++it) {
          const uint32_t s = it.s, ph = it.ph;
          if (kb == 4) {     // the staged g_u of this unit
            mbar_wait(cx.bar(BAR_OUT_READY), t_idx & 1, 58);
            tc_fence_after();
          }
          mbar_wait(cx.bar(BAR_FULL + s), ph, 59);
          tc_fence_after();
          if (elect_one()) {
            cx.mma_kblock(tmem + 256, kb < 4 ? cx.stage_a(s) : cx.out_kb(kb - 4), cx.stage_b(s), kb == 0);
            cx.commit(BAR_EMPTY + s);
          }
          __syncwarp();
        }
        if (elect_one()) cx.commit(BAR_ACC_FULL + 1);
        __syncwarp();
      };
      uint32_t ti = 0;
      for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
        mbar_wait(cx.bar(BAR_ACC_EMPTY + 0), (ti & 1) ^ 1, 57);
        tc_fence_after();
        for (int k = 0; k < 24; ++k, ++it) {
          const uint32_t s = it.s, ph = it.ph;
          mbar_wait(cx.bar(BAR_FULL + s), ph, 56);
          tc_fence_after();
          if (elect_one()) {
            cx.mma_kblock(tmem, cx.stage_a(s), cx.stage_b(s), k == 0);
            cx.commit(BAR_EMPTY + s);
          }
          __syncwarp();
        }
        if (elect_one()) cx.commit(BAR_ACC_FULL + 0);
        __syncwarp();
        if (ti > 0) job_b(ti - 1);
      }
      if (ti > 0) job_b(ti - 1);
    }
  } else if (warp >= EPI_WARP0) {
    setmaxnreg_inc<200>();
    if (((warp - EPI_WARP0) >> 2) == 0) kf_epilogue<G, 0>(cx, p, tmem, tiles);
    else kf_epilogue<G, 1>(cx, p, tmem, tiles);
  }
  tc_epilogue_teardown<CG>(tmem);
}

// ---- the same fused launch with every epilogue stream moved by TMA (the epilogue warps issue no global access).  Opt-in
// (AP_BWD_STAGED=1), kept as the record of an experiment: the 32-byte-per-thread loads / stores of kf_epilogue (96 sectors per position,
// one cache line per lane; ncu: l1tex throughput 75 % at boost clocks) looked like the limiter of the sustained loop, but this form
// runs in the same time: the pass is bound by tensor + memory energy under the power cap.  g_u(l+1) of the tile is TMA-loaded INTO the staging
// K-blocks once job B of the previous unit has read them, the epilogue adds the accumulator in place, and the staging tile is both
// job B's A operand and the source of the g_u(l) store; the saved gate derivatives come in as 16 KB sub-tiles [128 rows][64 ch]
// through four buffers, are multiplied in place and leave as g_a(l-1) (loader warp 2, storer warp 3, as in k1_layer).  Shared
// memory: 3-stage operand ring (96 KB) + staging (64 KB) + 4 stream buffers (64 KB).
struct GeoF {
  static constexpr int CG = 2, DT = 0, NCOMBO = 1;
  static constexpr bool SPLIT = false;
  static constexpr int B_ROWS = 128, B_BYTES = B_ROWS * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NSTAGE = 3;
  static constexpr int OUT_OFF = NSTAGE * STAGE_BYTES;
  static constexpr int RES_OFF = OUT_OFF + OUT_BYTES;          // four stream buffers of A_BYTES
  static constexpr int BIAS_OFF = RES_OFF + 4 * A_BYTES;
  static constexpr int BAR_OFF = BIAS_OFF;
  static constexpr int SMEM_BYTES = BAR_OFF + 512 + 1024;
  static constexpr uint32_t IDESC = umma_idesc_bf16_f32(256, 256);
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
};
enum { BARF_X_FULL = 24,      // [4] derivative sub-tile landed in stream buffer i
       BARF_X_DONE = 28,      // [4] the 8 epilogue warps have turned buffer i into g_a
       BARF_X_FREE = 32,      // [4] the store of buffer i has read it
       BARF_GN_FULL = 36,     // g_u(l+1) of the tile landed in the staging K-blocks
       BARF_GU_READY = 37,    // the staging tile holds g_u(l) (-> storer)
       BARF_GU_STORED = 38,   // the store of g_u(l) has read the staging tile (-> loader)
       BARF_COUNT = 39 };

template <int HSEL>
__device__ __forceinline__ void kfs_epilogue(const Ctx<GeoF>& cx, const KfParams& p, const uint32_t tmem, const Tiles<2>& tiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q4 = warp & 3, etid = threadIdx.x - EPI_WARP0 * 32;
  const int row = q4 * 32 + lane;
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q4 * 32) << 16);
  const uint32_t row_off = row * 128, sw = row & 7;
  const float sqrt_half = 0.70710678118654752440f;
  uint32_t xcount = 0;       // derivative sub-tiles consumed: buffer xcount & 3, phase (xcount >> 2) & 1

  auto epi_b = [&](uint32_t t_idx) {
    mbar_wait(cx.bar(BAR_ACC_FULL + 1), t_idx & 1, 64);
    tc_fence_after();
    uint32_t accs[4][32];
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) tmem_ld_32x32b_x32(lane_addr + 256 + pass * 64 + HSEL * 32, accs[pass]);
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + 1);
#pragma unroll
    for (int j = 0; j < 8; ++j, ++xcount) {          // plane j >> 2 (d/d a_t, d/d a_s), channels (j & 3) * 64 ...
      const uint32_t buf = xcount & 3;
      const uint32_t (&acc)[32] = accs[j & 3];
      mbar_wait(cx.bar(BARF_X_FULL + buf), (xcount >> 2) & 1, 52);
      const uint32_t rb = cx.base + GeoF::RES_OFF + buf * A_BYTES + row_off;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t addr = rb + (((HSEL * 4 + c) ^ sw) << 4);
        const uint4 dv = ld_shared_v4(addr);
        const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
        uint32_t pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          pk[k] = pack_bf16x2(__uint_as_float(acc[c * 8 + 2 * k]) * bf16_lo(dw[k]), __uint_as_float(acc[c * 8 + 2 * k + 1]) * bf16_hi(dw[k]));
        st_shared_v4(addr, make_uint4(pk[0], pk[1], pk[2], pk[3]));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(cx.bar(BARF_X_DONE + buf));
    }
  };

  uint32_t ti = 0;
  for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
    // ---- epilogue A: the accumulator to registers (region X is free at once), then g_u(l) = acc + sqrt(.5) g_u(l+1) in place in
    //      the staging K-blocks (rows past the end of a waveform are zero on both sides: TMA fills out-of-bounds rows with zeros)
    mbar_wait(cx.bar(BAR_ACC_FULL + 0), ti & 1, 65);
    tc_fence_after();
    uint32_t accs[4][32];
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) tmem_ld_32x32b_x32(lane_addr + pass * 64 + HSEL * 32, accs[pass]);
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + 0);
    mbar_wait(cx.bar(BARF_GN_FULL), ti & 1, 53);
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const uint32_t rb = cx.out_kb(pass) + row_off;
      const uint32_t (&acc)[32] = accs[pass];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t addr = rb + (((HSEL * 4 + c) ^ sw) << 4);
        const uint4 gv = ld_shared_v4(addr);
        const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
        uint32_t pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          pk[k] = pack_bf16x2(fmaf(bf16_lo(gw[k]), sqrt_half, __uint_as_float(acc[c * 8 + 2 * k])),
                              fmaf(bf16_hi(gw[k]), sqrt_half, __uint_as_float(acc[c * 8 + 2 * k + 1])));
        st_shared_v4(addr, make_uint4(pk[0], pk[1], pk[2], pk[3]));
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1, EPI_THREADS);
    if (etid == 0) {
      cx.arrive_leader(BAR_OUT_READY);
      mbar_arrive(cx.bar(BARF_GU_READY));
    }
    if (ti > 0) epi_b(ti - 1);
  }
  if (ti > 0) epi_b(ti - 1);
}

// tmGa: g_a(l); tmGs: g_s; tmWdT / tmWb: transposed weights; tmGn: g_u(l+1) (zeros for the top layer); tmGuOut: g_u(l);
// tmTs: saved gate derivatives [N chunk][L][512]; tmGaOut: g_a(l-1)
__global__ void __launch_bounds__(NTHREADS, 1)
k_bwd_fused_s(const __grid_constant__ CUtensorMap tmGa, const __grid_constant__ CUtensorMap tmGs, const __grid_constant__ CUtensorMap tmWdT,
              const __grid_constant__ CUtensorMap tmWb, const __grid_constant__ CUtensorMap tmGn, const __grid_constant__ CUtensorMap tmGuOut,
              const __grid_constant__ CUtensorMap tmTs, const __grid_constant__ CUtensorMap tmGaOut, const __grid_constant__ KfParams p) {
  using G = GeoF;
  constexpr int CG = 2;
  Ctx<G> cx;
  uint8_t* gen;
  const uint32_t tmem = tc_prologue<G>(cx, gen, 1);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(cx.bar(BARF_X_FULL + i), 1);
      mbar_init(cx.bar(BARF_X_DONE + i), 8);
      mbar_init(cx.bar(BARF_X_FREE + i), 1);
    }
    mbar_init(cx.bar(BARF_GN_FULL), 1);
    mbar_init(cx.bar(BARF_GU_READY), 1);
    mbar_init(cx.bar(BARF_GU_STORED), 1);
    fence_barrier_init();
    prefetch_tmap(&tmGa), prefetch_tmap(&tmGs), prefetch_tmap(&tmWdT), prefetch_tmap(&tmWb);
    prefetch_tmap(&tmGn), prefetch_tmap(&tmGuOut), prefetch_tmap(&tmTs), prefetch_tmap(&tmGaOut);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Tiles<CG> tiles(p.n_tiles, cx.rank);
  const int oob_l0 = p.tiles_per_sample * TILE_M;

  if (warp < EPI_WARP0) setmaxnreg_dec<96>();
  if (warp == 0) {
    RingPos<G::NSTAGE> it;
    auto load_job_b = [&](bool valid, int b, int l0) {
      for (int kb = 0; kb < 8; ++kb, ++it) {
        const uint32_t s = it.s, ph = it.ph;
        cx.wait_empty(s, ph, 67);
        if (elect_one()) {
          if (kb < 4) {
            cx.arm(s, G::STAGE_BYTES);
            cx.load_a(s, &tmGs, kb * 64, valid ? l0 : oob_l0, b);
          } else {
            cx.arm(s, G::B_BYTES);
          }
          cx.load_b(s, &tmWb, kb * 64, p.b_row0);
        }
        __syncwarp();
      }
    };
    uint32_t ti = 0;
    bool pvalid = false;
    int pb = 0, pl0 = 0;
    for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
      const bool valid = tile < p.n_tiles;
      const int b = valid ? tile / p.tiles_per_sample : 0;
      const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
      for (int tap = 0; tap < 3; ++tap)
        for (int kb = 0; kb < 8; ++kb, ++it) {
          const uint32_t s = it.s, ph = it.ph;
          cx.wait_empty(s, ph, 68);
          if (elect_one()) {
            cx.arm(s, G::STAGE_BYTES);
            cx.load_a(s, &tmGa, kb * 64, valid ? l0 - (tap - 1) * p.dilation : oob_l0, b);
            cx.load_b(s, &tmWdT, (tap * 8 + kb) * 64, p.a_row0);
          }
          __syncwarp();
        }
      if (ti > 0) load_job_b(pvalid, pb, pl0);
      pvalid = valid, pb = b, pl0 = l0;
    }
    if (ti > 0) load_job_b(pvalid, pb, pl0);
  } else if (warp == 1) {
    if (cx.rank == 0) {
      RingPos<G::NSTAGE> it;
      auto job_b = [&](uint32_t t_idx) {
        mbar_wait(cx.bar(BAR_ACC_EMPTY + 1), (t_idx & 1) ^ 1, 69);
        tc_fence_after();
        for (int kb = 0; kb < 8; ++kb, ++it) {
          const uint32_t s = it.s, ph = it.ph;
          if (kb == 4) {
            mbar_wait(cx.bar(BAR_OUT_READY), t_idx & 1, 58);
            tc_fence_after();
          }
          mbar_wait(cx.bar(BAR_FULL + s), ph, 59);
          tc_fence_after();
          if (elect_one()) {
            cx.mma_kblock(tmem + 256, kb < 4 ? cx.stage_a(s) : cx.out_kb(kb - 4), cx.stage_b(s), kb == 0);
            cx.commit(BAR_EMPTY + s);
          }
          __syncwarp();
        }
        if (elect_one()) cx.commit(BAR_ACC_FULL + 1);
        __syncwarp();
      };
      uint32_t ti = 0;
      for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
        mbar_wait(cx.bar(BAR_ACC_EMPTY + 0), (ti & 1) ^ 1, 57);
        tc_fence_after();
        for (int k = 0; k < 24; ++k, ++it) {
          const uint32_t s = it.s, ph = it.ph;
          mbar_wait(cx.bar(BAR_FULL + s), ph, 56);
          tc_fence_after();
          if (elect_one()) {
            cx.mma_kblock(tmem, cx.stage_a(s), cx.stage_b(s), k == 0);
            cx.commit(BAR_EMPTY + s);
          }
          __syncwarp();
        }
        if (elect_one()) cx.commit(BAR_ACC_FULL + 0);
        __syncwarp();
        if (ti > 0) job_b(ti - 1);
      }
      if (ti > 0) job_b(ti - 1);
    }
  } else if (warp == 2) {
    // ======================================================================================= stream loader
    uint32_t xcount = 0, ti = 0;
    bool pvalid = false;
    int pb = 0, pl0 = 0;
    auto load_ts = [&](int j, bool valid, int b, int l0) {
      const uint32_t buf = xcount & 3;
      mbar_wait(cx.bar(BARF_X_FREE + buf), ((xcount >> 2) & 1) ^ 1, 50);
      if (lane == 0) {
        mbar_expect_tx(cx.bar(BARF_X_FULL + buf), A_BYTES);
        tma_load_3d(cx.base + G::RES_OFF + buf * A_BYTES, &tmTs, cx.bar(BARF_X_FULL + buf), (j >> 2) * 256 + (j & 3) * 64,
                    valid ? l0 : oob_l0, p.ts_row0 + b);
      }
      __syncwarp();
      ++xcount;
    };
    for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
      const bool valid = tile < p.n_tiles;
      const int b = valid ? tile / p.tiles_per_sample : 0;
      const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
      if (ti > 0)
        for (int j = 0; j < 4; ++j) load_ts(j, pvalid, pb, pl0);
      if (ti > 0) {          // the staging tile: job B of the previous unit has read it, and so has the store of its g_u
        mbar_wait(cx.bar(BAR_ACC_FULL + 1), (ti - 1) & 1, 51);
        mbar_wait(cx.bar(BARF_GU_STORED), (ti - 1) & 1, 49);
      }
      if (lane == 0) {
        mbar_expect_tx(cx.bar(BARF_GN_FULL), OUT_BYTES);
        for (int pass = 0; pass < 4; ++pass) tma_load_3d(cx.out_kb(pass), &tmGn, cx.bar(BARF_GN_FULL), pass * 64, valid ? l0 : oob_l0, b);
      }
      __syncwarp();
      if (ti > 0)
        for (int j = 4; j < 8; ++j) load_ts(j, pvalid, pb, pl0);
      pvalid = valid, pb = b, pl0 = l0;
    }
    if (ti > 0)
      for (int j = 0; j < 8; ++j) load_ts(j, pvalid, pb, pl0);
  } else if (warp == 3) {
    // ======================================================================================= stream storer
    uint32_t xcount = 0, ti = 0;
    bool pvalid = false;
    int pb = 0, pl0 = 0;
    auto store_ga = [&](int j, bool valid, int b, int l0) {
      const uint32_t buf = xcount & 3;
      mbar_wait(cx.bar(BARF_X_DONE + buf), (xcount >> 2) & 1, 48);
      if (lane == 0) {
        if (valid) tma_store_3d(&tmGaOut, cx.base + G::RES_OFF + buf * A_BYTES, (j >> 2) * 256 + (j & 3) * 64, l0, b);
        bulk_commit();
        bulk_wait_read<0>();
        mbar_arrive(cx.bar(BARF_X_FREE + buf));
      }
      __syncwarp();
      ++xcount;
    };
    for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride, ++ti) {
      const bool valid = tile < p.n_tiles;
      const int b = valid ? tile / p.tiles_per_sample : 0;
      const int l0 = valid ? (tile - b * p.tiles_per_sample) * TILE_M : oob_l0;
      mbar_wait(cx.bar(BARF_GU_READY), ti & 1, 47);
      if (lane == 0) {
        if (valid)
          for (int pass = 0; pass < 4; ++pass) tma_store_3d(&tmGuOut, cx.out_kb(pass), pass * 64, l0, b);
        bulk_commit();
        bulk_wait_read<0>();
        mbar_arrive(cx.bar(BARF_GU_STORED));
      }
      __syncwarp();
      if (ti > 0)
        for (int j = 0; j < 8; ++j) store_ga(j, pvalid, pb, pl0);
      pvalid = valid, pb = b, pl0 = l0;
    }
    if (ti > 0)
      for (int j = 0; j < 8; ++j) store_ga(j, pvalid, pb, pl0);
    if (lane == 0) bulk_wait_all<0>();
    __syncwarp();
  } else if (warp >= EPI_WARP0) {
    setmaxnreg_inc<200>();
    if (((warp - EPI_WARP0) >> 2) == 0) kfs_epilogue<0>(cx, p, tmem, tiles);
    else kfs_epilogue<1>(cx, p, tmem, tiles);
  }
  tc_epilogue_teardown<CG>(tmem);
}

// g_pre[m][c] = mask(m, c) ? g_eps[m] * w2[c] : 0   (bf16; backward of eps = w2 . relu(pre) + b2, WaveNet.py:161-162)
__global__ void __launch_bounds__(256) gpre_kernel(const float* __restrict__ g_eps, const uint32_t* __restrict__ mask,
                                                    const float* __restrict__ w2, uint4* __restrict__ g_pre, long long M) {
  __shared__ float sw[C];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sw[i] = w2[i];
  __syncthreads();
  const long long total = M * (C / 8);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i >> 5;
    const int c = static_cast<int>(i & 31) * 8;
    const float g = g_eps[m];
    const uint32_t bits = mask[m * 8 + (c >> 5)] >> (c & 31);
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      pk[e] = pack_bf16x2((bits >> (2 * e)) & 1u ? g * sw[c + 2 * e] : 0.f, (bits >> (2 * e + 1)) & 1u ? g * sw[c + 2 * e + 1] : 0.f);
    g_pre[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// g_x[m] = sum_c g_u0[m][c] * w[c] * [w[c] x[m] + b[c] > 0]   (backward of u0 = relu(w x + b) + p0, WaveNet.py:147,13-19);
// one warp per position, 8 channels per lane
__global__ void __launch_bounds__(256) gx_kernel(const uint4* __restrict__ g_u0, const float* __restrict__ x,
                                                  const float* __restrict__ w, const float* __restrict__ b,
                                                  float* __restrict__ g_x, long long M) {
  const int lane = threadIdx.x & 31;
  float wv[8], bv[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) wv[e] = w[lane * 8 + e], bv[e] = b[lane * 8 + e];
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long m = warp0; m < M; m += nwarps) {
    const float xv = x[m];
    const uint4 g = g_u0[m * 32 + lane];
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (fmaf(wv[2 * e], xv, bv[2 * e]) > 0.f) acc = fmaf(bf16_lo(gw[e]), wv[2 * e], acc);
      if (fmaf(wv[2 * e + 1], xv, bv[2 * e + 1]) > 0.f) acc = fmaf(bf16_hi(gw[e]), wv[2 * e + 1], acc);
    }
    for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) g_x[m] = acc;
  }
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

// Certification front end fused into the init conv: a block walks 1024 consecutive positions e = b * L + l of the (B, L) batch.
// Phase 1: thread i draws the Philox block of positions [e0 + 4 i, +4) -- or reads the caller's noise -- and forms
//          x_in = scale * (x1[l] + sigma * z) in SmoothOp's rounding order (ap_update.cu); x_in goes to shared memory and to
//          `xin_out` (the x0 output buffer, which k2_head's epilogue reads back and overwrites).
// Phase 2: the 1024 x 256 channels-last tile of u0 = bf16(relu(w x_in + b) + p0) (hi / lo planes in split mode).
// Requires (B * L) % 4 == 0 handling of the tail by element count; the Philox block of element e is offset + e / 4.
template <int DT>
__global__ void __launch_bounds__(256) init_smooth_kernel(const SmoothSrc src, const float* __restrict__ w, const float* __restrict__ b,
                                                           const float* __restrict__ p0, uint4* __restrict__ u, long long lo_off,
                                                           float* __restrict__ xin_out, long long M, int L) {
  __shared__ float sw[C], sb[C], sp[C];
  __shared__ __align__(16) float sx[1024];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sw[i] = w[i], sb[i] = b[i], sp[i] = p0[i];
  const uint64_t off = src.offset + (src.offset_dev ? *src.offset_dev : 0ull);
  for (long long e0 = static_cast<long long>(blockIdx.x) * 1024; e0 < M; e0 += static_cast<long long>(gridDim.x) * 1024) {
    __syncthreads();                                       // sw/sb/sp visible; previous tile's sx consumed
    {
      const long long e = e0 + 4 * threadIdx.x;
      if (e < M) {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        const int cnt = M - e >= 4 ? 4 : static_cast<int>(M - e);
        if (src.z) {
          for (int j = 0; j < cnt; ++j) z[j] = src.z[e + j];
        } else {
          normal4(off + static_cast<uint64_t>(e >> 2), src.seed, z);
        }
        float xi[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xv = j < cnt ? src.x1[(e + j) % L] : 0.f;
          xi[j] = __fmul_rn(src.scale, __fadd_rn(xv, __fmul_rn(src.sigma, z[j])));
          sx[4 * threadIdx.x + j] = xi[j];
          if (j < cnt) xin_out[e + j] = xi[j];
        }
      }
    }
    __syncthreads();
    const long long npos = M - e0 < 1024 ? M - e0 : 1024;
    for (long long i = threadIdx.x; i < npos * 32; i += blockDim.x) {
      const int c = static_cast<int>(i & 31) * 8;
      const float xv = sx[i >> 5];
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float v0, v1;
        if (DT == 2) {   // the roundings of init_split_kernel
          v0 = __fadd_rn(fmaxf(__fadd_rn(__fmul_rn(sw[c + 2 * k], xv), sb[c + 2 * k]), 0.f), sp[c + 2 * k]);
          v1 = __fadd_rn(fmaxf(__fadd_rn(__fmul_rn(sw[c + 2 * k + 1], xv), sb[c + 2 * k + 1]), 0.f), sp[c + 2 * k + 1]);
          ph[k] = pack_bf16x2(v0, v1);
          pl[k] = pack_bf16x2(v0 - bf16_lo(ph[k]), v1 - bf16_hi(ph[k]));
        } else {         // the roundings of init_h16_kernel
          v0 = fmaxf(fmaf(sw[c + 2 * k], xv, sb[c + 2 * k]), 0.f) + sp[c + 2 * k];
          v1 = fmaxf(fmaf(sw[c + 2 * k + 1], xv, sb[c + 2 * k + 1]), 0.f) + sp[c + 2 * k + 1];
          ph[k] = pack2<DT == 1 ? 1 : 0>(v0, v1);
        }
      }
      u[e0 * 32 + i] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
      if (DT == 2) u[lo_off + e0 * 32 + i] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
  }
}

// split mode: u0 as hi / lo bf16 planes (`lo_off` uint4 elements apart)
__global__ void __launch_bounds__(256) init_split_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ b, const float* __restrict__ p0,
                                                          uint4* __restrict__ u, long long lo_off, long long M) {
  __shared__ float sw[C], sb[C], sp[C];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sw[i] = w[i], sb[i] = b[i], sp[i] = p0[i];
  __syncthreads();
  const long long total = M * (C / 8);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i >> 5;
    const int c = static_cast<int>(i & 31) * 8;
    const float xv = x[m];
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // the same roundings as the fp32 path (init_f32_kernel): mul, add, relu, add
      const float v0 = __fadd_rn(fmaxf(__fadd_rn(__fmul_rn(sw[c + 2 * e], xv), sb[c + 2 * e]), 0.f), sp[c + 2 * e]);
      const float v1 = __fadd_rn(fmaxf(__fadd_rn(__fmul_rn(sw[c + 2 * e + 1], xv), sb[c + 2 * e + 1]), 0.f), sp[c + 2 * e + 1]);
      ph[e] = pack_bf16x2(v0, v1);
      pl[e] = pack_bf16x2(v0 - bf16_lo(ph[e]), v1 - bf16_hi(ph[e]));
    }
    u[i] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    u[lo_off + i] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

// hi + lo planes -> fp32 (debug dumps, split mode)
__global__ void split_to_f32_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, float* __restrict__ out,
                                    long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = bf16_lo(hi[i]) + bf16_lo(lo[i]);
}

// bf16 -> fp32 (debug dumps)
template <int DT>
__global__ void h16_to_f32_kernel(const uint16_t* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = unpack_lo<DT>(in[i]);
}


// ------------------------------------------------------------------------------------------------ K4: log-mel on the tensor cores
// The windowed-DFT contraction of the log-mel front end (torchaudio MelSpectrogram + AmplitudeToDB, built at
// certified_robustness_eval.py:85-87) as a tcgen05 GEMM with fp32-class accuracy and the whole tail fused into its epilogue:
//     frames[(b, f)][n] . basis[n][cos k | -sin k]  ->  power = re^2 + im^2  ->  mel filterbank  ->  10 log10(max(., 1e-10))
// so that the power spectrum never reaches HBM (the FFMA version wrote and re-read 262 KB per waveform).
//  * A operand without im2col: with hop H and n_fft = Q H, frame f is the concatenation of Q consecutive H-sample blocks of the
//    zero-padded waveform, frame[f][q H + r] = Xp[f + q][r].  One 3-D TMA box (64 r, 32 frames, 4 waveforms) of Xp at block
//    offset q is therefore the [128 rows][64 k] A tile of K-block (q, r0) -- the same trick as the dilated taps of k1_layer.
//  * fp32-class accuracy on bf16 tensor cores: Xp and the basis are kept as bf16 hi / lo planes and every product is
//    accumulated as hi*hi + lo*hi + hi*lo (three MMAs per K step, as in AP_MODE_BF16X3): 2e-3 dB needs more than tf32's 10 bits.
//  * N tile t = [cos bins 128 t .. +127 | -sin bins 128 t .. +127]; the sine row of bin 0 is identically zero, so it carries
//    the Nyquist bin (n_fft / 2) instead: n_freq - 1 = 1024 bins = 8 tiles with nothing padded.
//  * CTA pairs (cta_group::2, M = 256 = 8 waveforms x 32 frames); a pair walks the N tiles of its rows while every epilogue
//    thread keeps the 32 mel sums of its row in registers: thread = row (TMEM lane), two column halves on the two warp groups.
struct MelGeo {
  static constexpr int CG = 2, DT = 0;
  static constexpr int B_ROWS = 128, B_BYTES = B_ROWS * 128, STAGE_BYTES = A_BYTES + B_BYTES, NSTAGE = 5;
  static constexpr int OUT_OFF = NSTAGE * STAGE_BYTES;     // filterbank rows of the current N tile: [129][32] fp32 (row 128: Nyquist)
  static constexpr int RED_OFF = OUT_OFF + 129 * 32 * 4 + 128;   // mel sums of the upper column half: [128 rows][33] fp32
  static constexpr int BIAS_OFF = RED_OFF + 128 * 33 * 4;
  static constexpr int BAR_OFF = (BIAS_OFF + 255) / 256 * 256;
  static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;
  static constexpr uint32_t IDESC = umma_idesc_bf16_f32(256, 256);
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
};
struct MelParams {
  int n_tiles;        // CTA tiles of 128 rows = 4 waveforms x 32 frames
  int B;              // waveforms with an output
  int x_plane;        // offset of the lo plane in the third coordinate of the Xp tensor map (= padded batch)
  int basis_plane;    // offset of the lo plane in the rows of the basis tensor map (= NT * 256)
  int NT, KB, hop64;  // N tiles (n_freq - 1) / 128, K blocks n_fft / 64, K blocks per hop
  const float* fb;    // [n_freq][32]
  float* spec;        // [B][32 mels][32 frames]
};

__global__ void __launch_bounds__(NTHREADS, 1)
k_mel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ MelParams p) {
  using G = MelGeo;
  Ctx<G> cx;
  uint8_t* gen;
  const uint32_t tmem = tc_prologue<G>(cx, gen, 1);
  if (threadIdx.x == 0) prefetch_tmap(&tmX), prefetch_tmap(&tmB);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Tiles<2> tiles(p.n_tiles, cx.rank);
  const int stages_per_ntile = p.KB * 3;

  if (warp == 0) {
    // ======================================================================================= TMA producer
    RingPos<G::NSTAGE> it;
    for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride) {
      const int b0 = tile < p.n_tiles ? tile * 4 : 2 * p.x_plane;     // past the end: the whole box is zero-filled
      for (int nt = 0; nt < p.NT; ++nt)
        for (int kb = 0; kb < p.KB; ++kb)
          for (int cmb = 0; cmb < 3; ++cmb, ++it) {    // (hi, hi), (lo, hi), (hi, lo)
            const uint32_t s = it.s, ph = it.ph;
            cx.wait_empty(s, ph, 90);
            if (elect_one()) {
              cx.arm(s, G::STAGE_BYTES);
              cx.load_a(s, &tmX, (kb % p.hop64) * 64, kb / p.hop64, b0 + (cmb == 1 && tile < p.n_tiles ? p.x_plane : 0));
              cx.load_b(s, &tmB, kb * 64, nt * 256 + (cmb == 2 ? p.basis_plane : 0));
            }
            __syncwarp();
          }
    }
  } else if (warp == 1) {
    // ======================================================================================= MMA issuer (leader CTA)
    if (cx.rank == 0) {
      RingPos<G::NSTAGE> it;
      uint32_t g = 0;
      for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride)
        for (int nt = 0; nt < p.NT; ++nt, ++g) {
          const uint32_t r = g & 1;
          mbar_wait(cx.bar(BAR_ACC_EMPTY + r), ((g >> 1) & 1) ^ 1, 91);
          tc_fence_after();
          for (int k = 0; k < stages_per_ntile; ++k, ++it) {
            const uint32_t s = it.s, ph = it.ph;
            mbar_wait(cx.bar(BAR_FULL + s), ph, 92);
            tc_fence_after();
            if (elect_one()) {
              cx.mma_kblock(tmem + r * 256, cx.stage_a(s), cx.stage_b(s), k == 0);
              cx.commit(BAR_EMPTY + s);
            }
            __syncwarp();
          }
          if (elect_one()) cx.commit(BAR_ACC_FULL + r);
          __syncwarp();
        }
    }
  } else if (warp >= EPI_WARP0) {
    // ======================================================================================= epilogue (8 warps)
    const int hsel = (warp - EPI_WARP0) >> 2, q4 = warp & 3, etid = threadIdx.x - EPI_WARP0 * 32;
    const int row = q4 * 32 + lane;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q4 * 32) << 16);
    const uint32_t fb_addr = cx.base + G::OUT_OFF;
    float* fb_s = reinterpret_cast<float*>(gen + G::OUT_OFF);
    float* red = reinterpret_cast<float*>(gen + G::RED_OFF);
    uint32_t g = 0;
    for (int tile = tiles.first; tiles.more(tile); tile += tiles.stride) {
      float mel[32];
#pragma unroll
      for (int m = 0; m < 32; ++m) mel[m] = 0.f;
      for (int nt = 0; nt < p.NT; ++nt, ++g) {
        const uint32_t r = g & 1;
        named_bar_sync(1, EPI_THREADS);                    // the previous N tile's filterbank rows are no longer read
        {
          const float4* src = reinterpret_cast<const float4*>(p.fb + static_cast<size_t>(nt) * 128 * 32);
          float4* dst = reinterpret_cast<float4*>(fb_s);
          for (int i = etid; i < 128 * 8; i += EPI_THREADS) dst[i] = src[i];
          if (nt == 0 && etid < 8)                          // the Nyquist bin's row (tile 0 carries it in the sine slot of bin 0)
            dst[128 * 8 + etid] = reinterpret_cast<const float4*>(p.fb + static_cast<size_t>(p.NT) * 128 * 32)[etid];
        }
        named_bar_sync(1, EPI_THREADS);
        mbar_wait(cx.bar(BAR_ACC_FULL + r), (g >> 1) & 1, 93);
        tc_fence_after();
#pragma unroll
        for (int grp = 0; grp < 2; ++grp) {
          uint32_t re[32], im[32];
          tmem_ld_32x32b_x32(lane_addr + r * 256 + hsel * 64 + grp * 32, re);
          tmem_ld_32x32b_x32(lane_addr + r * 256 + 128 + hsel * 64 + grp * 32, im);
          tmem_ld_wait();
          if (grp == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) cx.arrive_leader(BAR_ACC_EMPTY + r);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float c = __uint_as_float(re[j]), s = __uint_as_float(im[j]);
            const int bin = hsel * 64 + grp * 32 + j;
            const bool nyq = (nt == 0) & (bin == 0);
            const float pw = nyq ? c * c : fmaf(c, c, s * s);
            const uint32_t fr = fb_addr + bin * 128;
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              const uint4 w = ld_shared_v4(fr + v * 16);    // warp-uniform address: broadcast
              mel[4 * v + 0] = fmaf(pw, __uint_as_float(w.x), mel[4 * v + 0]);
              mel[4 * v + 1] = fmaf(pw, __uint_as_float(w.y), mel[4 * v + 1]);
              mel[4 * v + 2] = fmaf(pw, __uint_as_float(w.z), mel[4 * v + 2]);
              mel[4 * v + 3] = fmaf(pw, __uint_as_float(w.w), mel[4 * v + 3]);
            }
            if (nyq) {
              const float pn = s * s;
#pragma unroll
              for (int m = 0; m < 32; ++m) mel[m] = fmaf(pn, fb_s[128 * 32 + m], mel[m]);
            }
          }
        }
      }
      // combine the two column halves, dB, store spec[b][mel][frame] (a warp = the 32 frames of one waveform: coalesced)
      named_bar_sync(1, EPI_THREADS);
      if (hsel == 1) {
#pragma unroll
        for (int m = 0; m < 32; ++m) red[row * 33 + m] = mel[m];
      }
      named_bar_sync(1, EPI_THREADS);
      if (hsel == 0 && tile < p.n_tiles) {
        const int b = tile * 4 + q4;
        if (b < p.B) {
#pragma unroll
          for (int m = 0; m < 32; ++m) {
            const float v = mel[m] + red[row * 33 + m];
            p.spec[(static_cast<size_t>(b) * 32 + m) * 32 + lane] = 10.0f * log10f(fmaxf(v, 1e-10f));
          }
        }
      }
    }
  }
  tc_epilogue_teardown<2>(tmem);
}

// zero-padded hop-sized blocks of the waveform as bf16 hi / lo planes: xp[plane][b][blk][r] = split(wav[b][(blk * hop + r) - pad])
__global__ void __launch_bounds__(256) mel_prep_kernel(const float* __restrict__ wav, uint32_t* __restrict__ xp, long long plane_words,
                                                        int B, int Bpad, int L, int hop, int nblk, int pad) {
  const long long per_b = static_cast<long long>(nblk) * hop / 2, total = per_b * Bpad;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / per_b);
    const long long e = (i - b * per_b) * 2 - pad;          // sample index of the first element of the pair
    float v0 = 0.f, v1 = 0.f;
    if (b < B) {
      const float* w = wav + static_cast<long long>(b) * L;
      if (e >= 0 && e < L) v0 = w[e];
      if (e + 1 >= 0 && e + 1 < L) v1 = w[e + 1];
    }
    const uint32_t hi = pack_bf16x2(v0, v1);
    xp[i] = hi;
    xp[plane_words + i] = pack_bf16x2(v0 - bf16_lo(hi), v1 - bf16_hi(hi));
  }
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
  DevBuf wd_s, wr_s, ws_s, wf_s;                           // bf16 hi plane followed by the lo plane (AP_MODE_BF16X3)
  CUtensorMap tmWd_s, tmWr_s, tmWs_s, tmWf_s;              // over both planes, box of 128 rows (CTA pairs only)
  int dt = 0;                                              // 0: bf16, 1: fp16, 2: bf16 hi/lo split (bf16x3)
  bool ws_split = false;                                   // layout of the reserved workspace (two planes per tensor)
  // backward pass (bf16 mode): transposed weights, saved activations and gradient buffers for bchunk waveforms
  DevBuf wb, wdt, wft;                                     // [N][256][512], [N][256][1536], [256][256] bf16, K-major
  CUtensorMap tmWb, tmWdT, tmWfT;
  DevBuf ts, mask, g_pre, g_s, g_a, g_a2, g_u[2];     // g_a / g_a2: the fused backward launches read one and write the other
  CUtensorMap tmGpre, tmGs, tmGa, tmGa2, tmGu[2], tmTs;
  int bchunk = 0, bL = 0;
  bool bwd_attr = false;
  unsigned long long save_gen = 0;   // generation of the saved forward state (tokens of ap_diffwave_eps_save)
  int save_B = 0, save_dt = 0;
  DevBuf br, bf2, init_w, init_b, wf2_dev;                 // fp32 vectors
  std::vector<float> bskip_host, bf1_host, wf2_host;       // k2's per-channel vectors (kernel params)
  std::vector<float> bd_host;                              // [N][512] dilated-conv biases in k1's packed order (kernel params)
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
        bd[static_cast<size_t>(l) * 512 + j * 256 + r] = r < 128 ? w[3][oc] : 0.5f * w[3][oc];   // sigmoid half: see K1Params
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
  auto to_split = [](const std::vector<float>& src) {   // hi plane, then lo = bf16(v - hi)
    std::vector<uint16_t> out(2 * src.size());
    for (size_t i = 0; i < src.size(); ++i) {
      const uint16_t hi = f32_to_bf16_rne(src[i]);
      uint32_t hb = static_cast<uint32_t>(hi) << 16;
      float hf;
      std::memcpy(&hf, &hb, 4);
      out[i] = hi;
      out[src.size() + i] = f32_to_bf16_rne(src[i] - hf);
    }
    return out;
  };
  TRY(upload_bf16(n->wd_s, to_split(wd_f)));
  TRY(upload_bf16(n->wr_s, to_split(wr_f)));
  TRY(upload_bf16(n->ws_s, to_split(ws_f)));
  TRY(upload_bf16(n->wf_s, to_split(wf_f)));
  {  // backward operands (see k_bwd): Wb[n][o][k] = Ws_n[k][o] (k < 256) | Wr_n[k-256][o] * sqrt(.5);
     // WdT[n][c][tap*512 + oc] = Wd_n[oc][c][tap];  WfT[c][o] = Wf[o][c]
    std::vector<uint16_t> wb(static_cast<size_t>(N) * C * 512), wdt(static_cast<size_t>(N) * C * 1536), wft(static_cast<size_t>(C) * C);
    const float sh = static_cast<float>(std::sqrt(0.5));
    for (int l = 0; l < N; ++l) {
      const float* const* w = weights + 6 + 8 * l;
      for (int o = 0; o < C; ++o)
        for (int k = 0; k < C; ++k) {
          wb[(static_cast<size_t>(l) * C + o) * 512 + k] = f32_to_bf16_rne(w[6][static_cast<size_t>(k) * C + o]);
          wb[(static_cast<size_t>(l) * C + o) * 512 + 256 + k] = f32_to_bf16_rne(w[4][static_cast<size_t>(k) * C + o] * sh);
        }
      for (int c = 0; c < C; ++c)
        for (int tap = 0; tap < 3; ++tap)
          for (int oc = 0; oc < 512; ++oc)
            wdt[(static_cast<size_t>(l) * C + c) * 1536 + tap * 512 + oc] = f32_to_bf16_rne(w[2][(static_cast<size_t>(oc) * C + c) * 3 + tap]);
    }
    for (int c = 0; c < C; ++c)
      for (int o = 0; o < C; ++o) wft[static_cast<size_t>(c) * C + o] = f32_to_bf16_rne(tail[0][static_cast<size_t>(o) * C + c]);
    TRY(upload_bf16(n->wb, wb));
    TRY(upload_bf16(n->wdt, wdt));
    TRY(upload_bf16(n->wft, wft));
  }
  n->bd_host = bd;
  TRY(upload_f32(n->br, br));
  n->bskip_host = bskip;
  n->bf1_host.assign(tail[1], tail[1] + C);
  n->wf2_host.assign(tail[2], tail[2] + C);
  TRY(upload_f32(n->wf2_dev, n->wf2_host));
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
    const uint64_t s1[2] = {768, static_cast<uint64_t>(N) * 1024}, s2[2] = {256, static_cast<uint64_t>(N) * 512}, s3[2] = {256, 512};
    TRY(encode_bf16(&n->tmWd_s, n->wd_s.p, 2, s1, bw2));
    TRY(encode_bf16(&n->tmWr_s, n->wr_s.p, 2, s2, bw2));
    TRY(encode_bf16(&n->tmWs_s, n->ws_s.p, 2, s2, bw2));
    TRY(encode_bf16(&n->tmWf_s, n->wf_s.p, 2, s3, bw2));
    const uint64_t b1[2] = {512, static_cast<uint64_t>(N) * 256}, b2[2] = {1536, static_cast<uint64_t>(N) * 256};
    TRY(encode_bf16(&n->tmWb, n->wb.p, 2, b1, bw2));
    TRY(encode_bf16(&n->tmWdT, n->wdt.p, 2, b2, bw2));
    TRY(encode_bf16(&n->tmWfT, n->wft.p, 2, d3, bw2));
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
void tc_net_set_dtype(TcNet* n, int dt) { n->dt = dt; }

size_t tc_net_workspace_bytes(const TcNet* n) { return n->u0.bytes + n->u1.bytes + n->o.bytes; }

int tc_net_reserve(TcNet* n, int chunk, int L) {
  using namespace tc;
  const int planes = n->dt == 2 ? 2 : 1;   // split mode: hi plane [chunk] followed by the lo plane [chunk]
  const size_t per = static_cast<size_t>(planes) * chunk * L * C * sizeof(uint16_t);
  // the old buffers go first (two workspaces do not fit): from here on nothing -- sizes, tensor maps, the saved backward
  // state -- may describe them, so a failed allocation leaves an EMPTY workspace rather than dangling pointers
  n->chunk = 0, n->L = 0, n->save_B = 0;
  std::memset(n->tmU, 0, sizeof(n->tmU)), std::memset(&n->tmO, 0, sizeof(n->tmO));
  n->u0.release(), n->u1.release(), n->o.release();
  cudaError_t e = n->u0.alloc(per);
  if (e == cudaSuccess) e = n->u1.alloc(per);
  if (e == cudaSuccess) e = n->o.alloc(per * n->N);
  if (e != cudaSuccess) {
    n->u0.release(), n->u1.release(), n->o.release();
    (void)cudaGetLastError();
    return fail(AP_ERR_CUDA, "DiffWave workspace: %.2f GB for %d waveforms of length %d (%d layers, %s) -> %s; lower it with "
                "ap_diffwave_reserve or AP_DIFFWAVE_WORKSPACE_GB", per * (n->N + 2.0) / 1e9, chunk, L, n->N,
                planes == 2 ? "bf16x3" : "bf16", cudaGetErrorString(e));
  }
  const uint64_t du[3] = {256, static_cast<uint64_t>(L), static_cast<uint64_t>(chunk) * planes};
  const uint64_t dO[3] = {256, static_cast<uint64_t>(L), static_cast<uint64_t>(chunk) * n->N * planes};
  const uint32_t bx[3] = {64, 128, 1};
  int rc = encode_bf16(&n->tmU[0], n->u0.p, 3, du, bx);
  if (rc == AP_OK) rc = encode_bf16(&n->tmU[1], n->u1.p, 3, du, bx);
  if (rc == AP_OK) rc = encode_bf16(&n->tmO, n->o.p, 3, dO, bx);
  if (rc != AP_OK) {
    n->u0.release(), n->u1.release(), n->o.release();
    return rc;
  }
  n->chunk = chunk, n->L = L, n->ws_split = planes == 2;   // committed only now: allocations and encodes succeeded
  if (!n->attr_set) {
    AP_CUDA(cudaFuncSetAttribute(k1_layer<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<1, 1>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k2_head<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<1, 2>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k1_layer<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 1>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k2_head<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 2>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k1_layer<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<1, 1>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k2_head<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<1, 2>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k1_layer<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 1>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k2_head<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 2>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k1_split<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 1, 2>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k2_head<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 2, 2>::SMEM_BYTES));
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

static int tc_run_layers(TcNet* n, const float* x, const float* ptab, int B, int L, int layers, cudaStream_t st,
                         bool save = false, const SmoothSrc* smooth = nullptr, float* xin_out = nullptr) {
  using namespace tc;
  const long long M = static_cast<long long>(B) * L;
  if (smooth) {
    long long blocks = ceil_div_ll(M, 1024);
    const long long cap = static_cast<long long>(num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    const long long lo_off = static_cast<long long>(n->chunk) * L * (C / 8);
    const unsigned g = static_cast<unsigned>(blocks);
    if (n->dt == 2)
      init_smooth_kernel<2><<<g, 256, 0, st>>>(*smooth, n->init_w.as<float>(), n->init_b.as<float>(), ptab, n->u0.as<uint4>(), lo_off, xin_out, M, L);
    else if (n->dt == 0)
      init_smooth_kernel<0><<<g, 256, 0, st>>>(*smooth, n->init_w.as<float>(), n->init_b.as<float>(), ptab, n->u0.as<uint4>(), 0, xin_out, M, L);
    else
      init_smooth_kernel<1><<<g, 256, 0, st>>>(*smooth, n->init_w.as<float>(), n->init_b.as<float>(), ptab, n->u0.as<uint4>(), 0, xin_out, M, L);
    AP_LAUNCH_CHECK();
  } else {
    long long blocks = ceil_div_ll(M * (C / 8), 256);
    const long long cap = static_cast<long long>(num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    if (n->dt == 2)
      init_split_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, n->init_w.as<float>(), n->init_b.as<float>(), ptab,
                                                                       n->u0.as<uint4>(),
                                                                       static_cast<long long>(n->chunk) * L * (C / 8), M);
    else if (n->dt == 0)
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
    p.chunk_alloc = n->chunk, p.last = (l == n->N - 1), p.L = L;
    p.u_in = (l & 1) ? n->u1.as<uint16_t>() : n->u0.as<uint16_t>();
    p.u_out = (l & 1) ? n->u0.as<uint16_t>() : n->u1.as<uint16_t>();
    p.o_out = n->o.as<uint16_t>();
    std::memcpy(p.bd, n->bd_host.data() + static_cast<size_t>(l) * 512, sizeof(p.bd));
    if (!save && n->dt != 2 && tc::kGatePoly) {
      // k1_layer's gate evaluates 2^(2 log2e (a_t + b_t)) and 2^(-log2e (a_s + b_s)) on the FMA pipe (gate_poly2): the biases
      // travel pre-scaled, so that scale and bias are ONE packed FFMA.  bd_host holds b_t and HALF of b_s.
      for (int j = 0; j < 2; ++j)
        for (int c = 0; c < 128; ++c) {
          p.bd[j * 256 + c] *= 2.885390081777927f;
          p.bd[j * 256 + 128 + c] *= -2.f * 1.4426950408889634f;
        }
    }
    p.b_res = n->br.as<float>() + static_cast<size_t>(l) * C;
    p.p_next = ptab + static_cast<size_t>(l + 1) * C;
    p.dbg = (l == n->dbg_layer) ? n->dbg.as<long long>() : nullptr;
    p.ts_out = save ? n->ts.as<uint16_t>() : nullptr, p.ts_chunk = n->bchunk;
    p.u_plane = p.o_plane = p.wd_plane = p.wr_plane = 0;
    if (n->dt == 2) p.u_plane = n->chunk, p.o_plane = n->N * n->chunk, p.wd_plane = n->N * 512, p.wr_plane = n->N * 256;
    cudaEvent_t e0 = prof_event(n, 0), e1 = e0 ? prof_event(n, 0) : nullptr;
    if (e1) cudaEventRecord(e0, st);
    const CUtensorMap& ui = n->tmU[l & 1];
    const CUtensorMap& uo = n->tmU[(l + 1) & 1];
    if (save && n->dt == 2)
      AP_CUDA(launch_pair(k1_split<true>, pair_grid(n_tiles), Geo<2, 1, 2>::SMEM_BYTES, st, ui, n->tmWd_s, n->tmWr_s, p));
    else if (save)
      AP_CUDA(launch_pair(k1_layer<2, 0, true>, pair_grid(n_tiles), Geo<2, 1>::SMEM_BYTES, st, ui, uo, n->tmO, n->tmWd2, n->tmWr2, p));
    else if (n->dt == 2)
      AP_CUDA(launch_pair(k1_split<false>, pair_grid(n_tiles), Geo<2, 1, 2>::SMEM_BYTES, st, ui, n->tmWd_s, n->tmWr_s, p));
    else if (n->pair && n->dt == 0)
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

static int tc_net_eps_impl(TcNet* n, const float* x, const float* ptab, float* eps, int B, int L, cudaStream_t st, bool save,
                           const SmoothSrc* smooth = nullptr);
int tc_net_eps(TcNet* n, const float* x, const float* ptab, float* eps, int B, int L, cudaStream_t st) {
  return tc_net_eps_impl(n, x, ptab, eps, B, L, st, false);
}
int tc_net_smooth_denoise(TcNet* n, const SmoothSrc& src, const float* ptab, float* x0, int B, int L, cudaStream_t st) {
  return tc_net_eps_impl(n, nullptr, ptab, x0, B, L, st, false, &src);
}
static int tc_net_eps_impl(TcNet* n, const float* x, const float* ptab, float* eps, int B, int L, cudaStream_t st, bool save,
                           const SmoothSrc* smooth) {
  using namespace tc;
  if (B > n->chunk || L != n->L || n->ws_split != (n->dt == 2))
    return fail(AP_ERR_STATE, "tc_net_eps: workspace reserved for chunk %d x L %d (%s layout)", n->chunk, n->L,
                n->ws_split ? "split" : "single-plane");
  int rc = tc_run_layers(n, x, ptab, B, L, n->N, st, save, smooth, eps);
  if (rc != AP_OK) return rc;
  const int tps = ceil_div(L, TILE_M), n_tiles = tps * B;
  K2Params p;
  p.mask_out = save ? n->mask.as<uint32_t>() : nullptr;
  p.fuse_x0 = smooth != nullptr, p.x0_a = smooth ? smooth->a : 0.f, p.x0_b = smooth ? smooth->b : 0.f;
  p.n_tiles = n_tiles, p.tiles_per_sample = tps, p.L = L, p.num_layers = n->N, p.chunk_alloc = n->chunk;
  p.scale = static_cast<float>(std::sqrt(1.0 / n->N));
  p.bf2 = n->bf2.as<float>();
  p.o_plane = p.ws_plane = p.wf_plane = 0;
  if (n->dt == 2) p.o_plane = n->N * n->chunk, p.ws_plane = n->N * 256, p.wf_plane = 256;
  for (int c = 0; c < C; ++c) p.bskip_scaled[c] = n->bskip_host[c] * p.scale, p.bf1[c] = n->bf1_host[c], p.wf2[c] = n->wf2_host[c];
  p.eps = eps;
  const int grid = n_tiles < num_sms() ? n_tiles : num_sms();
  cudaEvent_t e0 = prof_event(n, 1), e1 = e0 ? prof_event(n, 1) : nullptr;
  if (e1) cudaEventRecord(e0, st);
  if (save && n->dt == 2)
    AP_CUDA(launch_pair(k2_head<2, 2, true>, pair_grid(n_tiles), Geo<2, 2, 2>::SMEM_BYTES, st, n->tmO, n->tmWs_s, n->tmWf_s, p));
  else if (save)
    AP_CUDA(launch_pair(k2_head<2, 0, true>, pair_grid(n_tiles), Geo<2, 2>::SMEM_BYTES, st, n->tmO, n->tmWs2, n->tmWf2, p));
  else if (n->dt == 2)
    AP_CUDA(launch_pair(k2_head<2, 2>, pair_grid(n_tiles), Geo<2, 2, 2>::SMEM_BYTES, st, n->tmO, n->tmWs_s, n->tmWf_s, p));
  else if (n->pair && n->dt == 0)
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

// ---- backward: g_x = (d eps / d x)^T g_eps at (x, ptab); also returns eps when eps_out != nullptr.  bf16 pair mode only.
static int tc_bwd_reserve(TcNet* n, int chunk, int L) {
  using namespace tc;
  if (n->bchunk == chunk && n->bL == L) return AP_OK;
  const size_t pos = static_cast<size_t>(chunk) * L;
  DevBuf* bufs[8] = {&n->ts, &n->mask, &n->g_pre, &n->g_s, &n->g_a, &n->g_a2, &n->g_u[0], &n->g_u[1]};
  const size_t sizes[8] = {pos * 512 * 2 * n->N, pos * 8 * 4, pos * 256 * 2, pos * 256 * 2, pos * 512 * 2, pos * 512 * 2, pos * 256 * 2,
                           pos * 256 * 2};
  n->bchunk = 0, n->bL = 0;                 // nothing describes the released buffers if an allocation below fails
  n->save_B = 0, ++n->save_gen;
  for (DevBuf* d : bufs) d->release();
  for (int i = 0; i < 8; ++i) {
    cudaError_t e = bufs[i]->alloc(sizes[i]);
    if (e != cudaSuccess) {
      for (DevBuf* d : bufs) d->release();
      (void)cudaGetLastError();
      size_t total = 0;
      for (size_t s : sizes) total += s;
      return fail(AP_ERR_CUDA, "DiffWave backward workspace: %.2f GB for %d waveforms of length %d -> %s", total / 1e9, chunk, L,
                  cudaGetErrorString(e));
    }
  }
  const uint64_t d256[3] = {256, static_cast<uint64_t>(L), static_cast<uint64_t>(chunk)};
  const uint64_t d512[3] = {512, static_cast<uint64_t>(L), static_cast<uint64_t>(chunk)};
  const uint32_t bx[3] = {64, 128, 1};
  int rc = encode_bf16(&n->tmGpre, n->g_pre.p, 3, d256, bx);
  if (rc == AP_OK) rc = encode_bf16(&n->tmGs, n->g_s.p, 3, d256, bx);
  if (rc == AP_OK) rc = encode_bf16(&n->tmGa, n->g_a.p, 3, d512, bx);
  if (rc == AP_OK) rc = encode_bf16(&n->tmGa2, n->g_a2.p, 3, d512, bx);
  const uint64_t dts[3] = {512, static_cast<uint64_t>(L), static_cast<uint64_t>(chunk) * n->N};
  if (rc == AP_OK) rc = encode_bf16(&n->tmTs, n->ts.p, 3, dts, bx);
  if (rc == AP_OK) rc = encode_bf16(&n->tmGu[0], n->g_u[0].p, 3, d256, bx);
  if (rc == AP_OK) rc = encode_bf16(&n->tmGu[1], n->g_u[1].p, 3, d256, bx);
  if (rc != AP_OK) {
    for (DevBuf* d : bufs) d->release();
    return rc;
  }
  if (!n->bwd_attr) {
    AP_CUDA(cudaFuncSetAttribute(k_bwd<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 2>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k_bwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 2>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k_bwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 2>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k_bwd_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 2>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k_bwd_fused_s, cudaFuncAttributeMaxDynamicSharedMemorySize, GeoF::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k1_layer<2, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 1>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k2_head<2, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 2>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k1_split<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 1, 2>::SMEM_BYTES));
    AP_CUDA(cudaFuncSetAttribute(k2_head<2, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2, 2, 2>::SMEM_BYTES));
    n->bwd_attr = true;
  }
  n->bchunk = chunk, n->bL = L;
  return AP_OK;
}

size_t tc_net_bwd_bytes_per_waveform(const TcNet* n, int L) {
  return static_cast<size_t>(L) * (static_cast<size_t>(n->N) * 1024 + 32 + 512 + 512 + 1024 + 1024 + 1024);
}

// forward that keeps what the backward needs (the gate's local derivatives of every layer, the head's ReLU mask); returns a
// token identifying the saved state (any later saving forward, re-reservation or mode change invalidates it)
int tc_net_eps_save(TcNet* n, const float* x, const float* ptab, float* eps, int B, int L, int bchunk, cudaStream_t st,
                    unsigned long long* token) {
  using namespace tc;
  if (n->dt == 1) return fail(AP_ERR_STATE, "backward pass: bf16 / bf16x3 modes only");
  if (B > n->chunk || L != n->L || n->ws_split != (n->dt == 2)) return fail(AP_ERR_STATE, "tc_net_eps_save: forward workspace mismatch");
  if (B > bchunk) return fail(AP_ERR_STATE, "tc_net_eps_save: batch %d exceeds the backward chunk %d", B, bchunk);
  int rc = tc_bwd_reserve(n, bchunk, L);
  if (rc != AP_OK) return rc;
  n->save_B = 0;
  rc = tc_net_eps_impl(n, x, ptab, eps, B, L, st, true);
  if (rc != AP_OK) return rc;
  n->save_B = B, n->save_dt = n->dt;
  if (token) *token = ++n->save_gen;
  else ++n->save_gen;
  return AP_OK;
}
bool tc_net_saved_state_is(const TcNet* n, unsigned long long token, int B, int L) {
  return token != 0 && token == n->save_gen && n->save_B == B && n->bL == L && n->save_dt == n->dt;
}

// backward from the saved state of the last tc_net_eps_save (same x)
int tc_net_backward(TcNet* n, const float* x, const float* g_eps, float* g_x, int B, int L, cudaStream_t st) {
  using namespace tc;
  if (n->save_B != B || n->bL != L) return fail(AP_ERR_STATE, "tc_net_backward: no saved forward state for this batch");
  const long long M = static_cast<long long>(B) * L;
  const int tps = ceil_div(L, TILE_M), n_tiles = tps * B, grid = pair_grid(n_tiles), smem = Geo<2, 2>::SMEM_BYTES;
  {
    long long blocks = ceil_div_ll(M * (C / 8), 256);
    const long long cap = static_cast<long long>(num_sms()) * 8;
    gpre_kernel<<<static_cast<unsigned>(blocks < cap ? blocks : cap), 256, 0, st>>>(g_eps, n->mask.as<uint32_t>(), n->wf2_dev.as<float>(),
                                                                                    n->g_pre.as<uint4>(), M);
    AP_LAUNCH_CHECK();
  }
  KbParams p{};
  p.n_tiles = n_tiles, p.tiles_per_sample = tps, p.L = L, p.dilation = 1, p.last = 0;
  p.nkb = 4, p.b_row0 = 0, p.scale = static_cast<float>(std::sqrt(1.0 / n->N));
  p.ts = nullptr, p.g_next = nullptr, p.out = n->g_s.as<uint16_t>();
  AP_CUDA(launch_pair(k_bwd<0>, grid, smem, st, n->tmGpre, n->tmGpre, n->tmWfT, p));
  AP_LAUNCH_CHECK();
  static const bool unfused = std::getenv("AP_BWD_UNFUSED") != nullptr;      // development aid: the two-launch form of every layer
  if (unfused) {
    for (int l = n->N - 1; l >= 0; --l) {
      const bool last = l == n->N - 1;
      // g_u of layer l+1 lives in g_u[(l+1) & 1]; this layer writes g_u[l & 1]
      p.last = last, p.nkb = last ? 4 : 8, p.b_row0 = l * 256;
      p.ts = n->ts.as<uint16_t>() + static_cast<size_t>(l) * n->bchunk * L * 512;
      p.out = n->g_a.as<uint16_t>();
      AP_CUDA(launch_pair(k_bwd<1>, grid, smem, st, n->tmGs, n->tmGu[(l + 1) & 1], n->tmWb, p));
      AP_LAUNCH_CHECK();
      p.nkb = 8, p.dilation = 1 << (l % n->cfg.dilation_cycle);
      p.g_next = n->g_u[(l + 1) & 1].as<uint16_t>(), p.out = n->g_u[l & 1].as<uint16_t>();
      AP_CUDA(launch_pair(k_bwd<2>, grid, smem, st, n->tmGa, n->tmGa, n->tmWdT, p));
      AP_LAUNCH_CHECK();
    }
  } else {
    // MODE 1 of the last layer alone, then one fused launch per layer boundary (MODE 2 of l + MODE 1 of l - 1), then MODE 2 of
    // layer 0 alone.  g_a(l) lives in ga[l & 1].
    DevBuf* ga[2] = {&n->g_a, &n->g_a2};
    const CUtensorMap* tga[2] = {&n->tmGa, &n->tmGa2};
    const int top = n->N - 1;
    static const bool staged = std::getenv("AP_BWD_STAGED") != nullptr;
    if (staged)        // the top layer has no g_u(l+1): the fused launches read zeros in its place
      AP_CUDA(cudaMemsetAsync(n->g_u[(top + 1) & 1].p, 0, static_cast<size_t>(M) * 512, st));
    p.last = 1, p.nkb = 4, p.b_row0 = top * 256;
    p.ts = n->ts.as<uint16_t>() + static_cast<size_t>(top) * n->bchunk * L * 512;
    p.out = ga[top & 1]->as<uint16_t>();
    AP_CUDA(launch_pair(k_bwd<1>, grid, smem, st, n->tmGs, n->tmGu[0], n->tmWb, p));
    AP_LAUNCH_CHECK();
    for (int l = top; l >= 1; --l) {
      KfParams f{};
      f.n_tiles = n_tiles, f.tiles_per_sample = tps, f.L = L, f.dilation = 1 << (l % n->cfg.dilation_cycle);
      f.has_next = l < top, f.a_row0 = l * 256, f.b_row0 = (l - 1) * 256;
      f.ts = n->ts.as<uint16_t>() + static_cast<size_t>(l - 1) * n->bchunk * L * 512;
      f.g_next = n->g_u[(l + 1) & 1].as<uint16_t>(), f.gu_out = n->g_u[l & 1].as<uint16_t>();
      f.ga_out = ga[(l - 1) & 1]->as<uint16_t>();
      f.ts_row0 = (l - 1) * n->bchunk;
      if (staged)
        AP_CUDA(launch_pair(k_bwd_fused_s, grid, GeoF::SMEM_BYTES, st, *tga[l & 1], n->tmGs, n->tmWdT, n->tmWb, n->tmGu[(l + 1) & 1],
                            n->tmGu[l & 1], n->tmTs, *tga[(l - 1) & 1], f));
      else
        AP_CUDA(launch_pair(k_bwd_fused, grid, smem, st, *tga[l & 1], n->tmGs, n->tmWdT, n->tmWb, f));
      AP_LAUNCH_CHECK();
    }
    p.last = top == 0, p.nkb = 8, p.dilation = 1, p.b_row0 = 0;
    p.g_next = n->g_u[1].as<uint16_t>(), p.out = n->g_u[0].as<uint16_t>();
    AP_CUDA(launch_pair(k_bwd<2>, grid, smem, st, *tga[0], *tga[0], n->tmWdT, p));
    AP_LAUNCH_CHECK();
  }
  {
    long long blocks = ceil_div_ll(M * 32, 256);
    const long long cap = static_cast<long long>(num_sms()) * 16;
    gx_kernel<<<static_cast<unsigned>(blocks < cap ? blocks : cap), 256, 0, st>>>(n->g_u[0].as<uint4>(), x, n->init_w.as<float>(),
                                                                                  n->init_b.as<float>(), g_x, M);
    AP_LAUNCH_CHECK();
  }
  return AP_OK;
}

int tc_net_vjp(TcNet* n, const float* x, const float* ptab, const float* g_eps, float* g_x, float* eps_out, float* eps_scratch,
               int B, int L, int bchunk, cudaStream_t st) {
  int rc = tc_net_eps_save(n, x, ptab, eps_out ? eps_out : eps_scratch, B, L, bchunk, st, nullptr);
  if (rc != AP_OK) return rc;
  return tc_net_backward(n, x, g_eps, g_x, B, L, st);
}

// debug: run init + layers [0, layer] and return u_{layer+1} and o_layer as fp32 (B, L, 256)
int tc_net_debug_layer(TcNet* n, const float* x, const float* ptab, int layer, float* u_next, float* gate, int B, int L,
                       cudaStream_t st) {
  using namespace tc;
  if (B > n->chunk || L != n->L || n->ws_split != (n->dt == 2)) return fail(AP_ERR_STATE, "tc_net_debug_layer: workspace mismatch");
  int rc = tc_run_layers(n, x, ptab, B, L, layer + 1, st);
  if (rc != AP_OK) return rc;
  const long long cnt = static_cast<long long>(B) * L * C;
  const uint16_t* un = ((layer + 1) & 1) ? n->u1.as<uint16_t>() : n->u0.as<uint16_t>();
  const uint16_t* on = n->o.as<uint16_t>() + static_cast<size_t>(layer) * n->chunk * L * C;
  if (n->dt == 2) {
    const size_t plane = static_cast<size_t>(n->chunk) * L * C;
    if (u_next) {
      split_to_f32_kernel<<<num_sms() * 4, 256, 0, st>>>(un, un + plane, u_next, cnt);
      AP_LAUNCH_CHECK();
    }
    if (gate) {
      split_to_f32_kernel<<<num_sms() * 4, 256, 0, st>>>(on, on + plane * n->N, gate, cnt);
      AP_LAUNCH_CHECK();
    }
    return AP_OK;
  }
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


// ================================================================================================ K4 on the tensor cores (host)
struct MelTc {
  int n_fft = 0, hop = 0, n_freq = 0, nblk = 0, NT = 0;
  DevBuf basis, xp;           // bf16 hi plane followed by the lo plane
  CUtensorMap tmB{}, tmX{};
  int bpad = 0, L = 0;        // what xp / tmX are sized for
  bool attr = false;
};
void mel_tc_destroy(MelTc* m) { delete m; }

bool mel_tc_eligible(const ap_mel_cfg& c, int L) {
  if (std::getenv("AP_MEL_FFMA")) return false;             // development aid: force the FFMA path
  return !c.reflect_pad && c.n_mels == 32 && c.hop_length % 64 == 0 && c.n_fft % c.hop_length == 0 && c.n_fft / c.hop_length <= 8 &&
         (c.n_fft / 2) % 128 == 0 && 1 + L / c.hop_length == 32;
}

// basis_rows(k, n): double-precision windowed DFT rows; the caller passes the window
int mel_tc_create(MelTc** out, const ap_mel_cfg& c) {
  using namespace tc;
  *out = nullptr;
  auto* m = new MelTc();
  m->n_fft = c.n_fft, m->hop = c.hop_length, m->n_freq = c.n_fft / 2 + 1, m->NT = (c.n_fft / 2) / 128;
  const int N = c.n_fft, rows = m->NT * 256;
  const double PI = 3.14159265358979323846;
  std::vector<uint16_t> b(static_cast<size_t>(2) * rows * N);
  for (int t = 0; t < m->NT; ++t)
    for (int j = 0; j < 256; ++j) {
      const int k = t * 128 + (j & 127);
      const bool sine = j >= 128, nyq = sine && k == 0;     // the (all-zero) sine row of bin 0 carries the Nyquist bin
      uint16_t* hi = b.data() + static_cast<size_t>(t * 256 + j) * N;
      uint16_t* lo = hi + static_cast<size_t>(rows) * N;
      for (int n = 0; n < N; ++n) {
        const double win = 0.5 - 0.5 * std::cos(2.0 * PI * n / N);          // periodic hann (torch.hann_window default)
        const long long nk = (static_cast<long long>(n) * (nyq ? N / 2 : k)) % N;
        const double ang = 2.0 * PI * static_cast<double>(nk) / N;
        const float v = static_cast<float>(nyq ? win * std::cos(ang) : (sine ? -win * std::sin(ang) : win * std::cos(ang)));
        hi[n] = f32_to_bf16_rne(v);
        uint32_t hb = static_cast<uint32_t>(hi[n]) << 16;
        float hf;
        std::memcpy(&hf, &hb, 4);
        lo[n] = f32_to_bf16_rne(v - hf);
      }
    }
  cudaError_t e = m->basis.upload(b.data(), b.size() * sizeof(uint16_t));
  if (e != cudaSuccess) {
    delete m;
    return fail(AP_ERR_CUDA, "mel basis upload: %s", cudaGetErrorString(e));
  }
  const uint64_t db[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(2 * rows)};
  const uint32_t bb[2] = {64, 128};
  int rc = encode_bf16(&m->tmB, m->basis.p, 2, db, bb);
  if (rc != AP_OK) {
    delete m;
    return rc;
  }
  *out = m;
  return AP_OK;
}

int mel_tc_forward(MelTc* m, const float* wav, const float* fb, float* spec, int B, int L, cudaStream_t st) {
  using namespace tc;
  const int frames = 1 + L / m->hop, Q = m->n_fft / m->hop;
  if (frames != 32) return fail(AP_ERR_STATE, "mel_tc_forward: 32 frames per waveform only");
  const int bpad = (B + 7) / 8 * 8, nblk = frames + Q - 1;
  if (bpad > m->bpad || L != m->L) {
    m->bpad = 0;
    const size_t words = static_cast<size_t>(bpad) * nblk * m->hop / 2;
    AP_CUDA(m->xp.alloc(2 * words * sizeof(uint32_t)));
    const uint64_t dx[3] = {static_cast<uint64_t>(m->hop), static_cast<uint64_t>(nblk), static_cast<uint64_t>(2 * bpad)};
    const uint32_t bx[3] = {64, 32, 4};
    int rc = encode_bf16(&m->tmX, m->xp.p, 3, dx, bx);
    if (rc != AP_OK) return rc;
    m->bpad = bpad, m->L = L, m->nblk = nblk;
  }
  if (!m->attr) {
    AP_CUDA(cudaFuncSetAttribute(k_mel, cudaFuncAttributeMaxDynamicSharedMemorySize, MelGeo::SMEM_BYTES));
    m->attr = true;
  }
  // the tensor map spans m->bpad waveforms; the lo plane starts m->bpad entries into its third dimension
  const long long plane_words = static_cast<long long>(m->bpad) * nblk * m->hop / 2;
  {
    long long blocks = ceil_div_ll(plane_words, 256);
    const long long cap = static_cast<long long>(num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    mel_prep_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(wav, m->xp.as<uint32_t>(), plane_words, B, m->bpad, L, m->hop, nblk,
                                                                   m->n_fft / 2);
    AP_LAUNCH_CHECK();
  }
  MelParams p;
  p.n_tiles = (B + 3) / 4, p.B = B, p.x_plane = m->bpad, p.basis_plane = m->NT * 256;
  p.NT = m->NT, p.KB = m->n_fft / 64, p.hop64 = m->hop / 64, p.fb = fb, p.spec = spec;
  AP_CUDA(launch_pair(k_mel, pair_grid(p.n_tiles), MelGeo::SMEM_BYTES, st, m->tmX, m->tmB, p));
  AP_LAUNCH_CHECK();
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
