// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM).
// Written for this project; instruction spellings checked against the PTX ISA as used by CUTLASS 4.x.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace ap { namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trapped launch (cudaErrorLaunchFailure), never a hung GPU.
// Returns the number of cycles spent waiting.  try_wait may itself suspend the thread for a while before it reports
// success, so the default (fast) build under-reports: it reads the clock only on the slow path, which keeps the wait to
// two instructions in the MMA-issue loop.  Build with -DAP_TC_TIMED_WAITS (devtools/build_variant.sh) for exact counters.
__device__ __forceinline__ long long mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
#ifndef AP_TC_TIMED_WAITS
  if (mbar_try_wait(bar, parity)) return 0;
#endif
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      printf("[audiopure] mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
  return clock64() - t0;
}

// ------------------------------------------------------------------ proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// TMA stores (smem -> global, bulk async group); OOB parts of the box are clipped
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// ---- the same copies with an L2 eviction-priority hint (experiment: -DAP_L2_HINTS, see DESIGN.md "What limits k1 now")
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d_pair_hint(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the N most recent bulk groups of this thread have finished READING their shared-memory source
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// named barrier among `nthreads` threads (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// 32-byte global accesses (one full sector per thread), not allocated in L1
__device__ __forceinline__ void ld_global_v8(const void* p, uint32_t (&r)[8]) {
  asm volatile("ld.global.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&r)[8]) {
  asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void st_global_v4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// two fp32 -> packed bf16x2 (lo in bits [0,16), hi in bits [16,32)), round to nearest even
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
// fp16 variants (AP_MODE_FP16): saturating conversion, so an out-of-range activation becomes +-65504 instead of inf
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ float f16_lo(uint32_t v) { return __half2float(__ushort_as_half(static_cast<unsigned short>(v & 0xffffu))); }
__device__ __forceinline__ float f16_hi(uint32_t v) { return __half2float(__ushort_as_half(static_cast<unsigned short>(v >> 16))); }
// DT = 0: bf16, DT = 1: fp16
template <int DT> __device__ __forceinline__ uint32_t pack2(float lo, float hi) { return DT == 1 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); }
template <int DT> __device__ __forceinline__ float unpack_lo(uint32_t v) { return DT == 1 ? f16_lo(v) : bf16_lo(v); }
template <int DT> __device__ __forceinline__ float unpack_hi(uint32_t v) { return DT == 1 ? f16_hi(v) : bf16_hi(v); }

// ------------------------------------------------------------------ TMEM allocation (one full warp executes these)
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ------------------------------------------------------------------ CTA pairs (cluster of 2, tcgen05 cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// arrive on an mbarrier given by a shared::cluster address (possibly in the peer CTA)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in THIS CTA's shared memory, the transaction bytes are counted on the mbarrier at
// `bar_cluster` (a shared::cluster address, normally the leader CTA's barrier)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T with M = 256 split over the pair; issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all MMAs issued so far completed) on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

// ------------------------------------------------------------------ UMMA descriptors
// Shared-memory operand descriptor, K-major, SWIZZLE_128B canonical layout:
//   rows of 128 B (64 bf16) at a 128 B pitch, 8-row swizzle atoms 1024 B apart (SBO), 16-B chunk index XOR (row & 7).
//   bits [0,14) addr>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=64) |
//   [46,48) version=1 (Blackwell) | [61,64) layout=2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32, both operands K-major.
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
// same with fp16 operands (a_format = b_format = 0)
__host__ __device__ constexpr uint32_t umma_idesc_f16_f32(int M, int N) {
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Instruction descriptor, kind::tf32: fp32 operands read as tf32, fp32 accumulate, both operands K-major (K = 8 per MMA).
__host__ __device__ constexpr uint32_t umma_idesc_tf32_f32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// fp32 -> tf32 (round to nearest, ties away), result kept in a 32-bit container
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// Arrive on an mbarrier once every MMA issued so far by this thread has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ------------------------------------------------------------------ TMEM -> registers (warp-collective; lane i gets TMEM lane base+i)
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ register reallocation between warp groups
// every warp of the (4-warp, aligned) warp group executes the same instruction; N is a multiple of 8 in [24, 256]
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ------------------------------------------------------------------ math
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_approx(float x) { return fmaf(tanh_approx(0.5f * x), 0.5f, 0.5f); }
// fp32-class variants (absolute error ~1e-7): tanh(x) = 1 - 2 / (1 + e^{2x}), sigmoid(x) = 1 / (1 + e^{-x}).
// Saturation is exact: e^{2x} -> inf gives 1, e^{2x} -> 0 gives -1 (x clamped so that 1 + e^{..} stays below 2^126,
// the range in which the fast division is accurate).
__device__ __forceinline__ float tanh_exp(float x) {
  const float e = __expf(2.f * fminf(fmaxf(x, -40.f), 40.f));
  return 1.f - __fdividef(2.f, 1.f + e);
}
// packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2 on sm_100): two independent fp32 operations per instruction
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void up2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// v -> (bf16x2 hi, bf16x2 lo) with lo = bf16(v - float(hi)) for a pair of values
__device__ __forceinline__ void split_bf16x2(f32x2 v, uint32_t& hi, uint32_t& lo) {
  float v0, v1;
  up2(v, v0, v1);
  hi = pack_bf16x2(v0, v1);
  float r0, r1;
  up2(fma2(pk2(bf16_lo(hi), bf16_hi(hi)), pk2(-1.f, -1.f), v), r0, r1);
  lo = pack_bf16x2(r0, r1);
}
// the gate of two channels at once: tanh(a) * sigmoid(s) = (E - 1) / ((E + 1)(1 + F)), E = e^{2a}, F = e^{-s} (see gate_exp)
__device__ __forceinline__ f32x2 gate_exp2(f32x2 a, f32x2 s) {
  float a0, a1, x0, x1, y0, y1, E0, E1, F0, F1, r0, r1;
  up2(a, a0, a1);
  up2(mul2(pk2(fminf(a0, 20.f), fminf(a1, 20.f)), pk2(2.885390081777927f, 2.885390081777927f)), x0, x1);
  up2(mul2(s, pk2(-1.4426950408889634f, -1.4426950408889634f)), y0, y1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(E0) : "f"(x0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(E1) : "f"(x1));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(F0) : "f"(y0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(F1) : "f"(y1));
  const f32x2 E = pk2(E0, E1), one = pk2(1.f, 1.f);
  up2(mul2(add2(E, one), add2(pk2(F0, F1), one)), x0, x1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(x0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(x1));
  return mul2(add2(E, pk2(-1.f, -1.f)), pk2(r0, r1));
}
// tanh(a) * sigmoid(s) with one reciprocal: (E - 1) / ((E + 1)(1 + F)), E = e^{2a}, F = e^{-s}.  a is clamped above so
// that E + 1 stays finite; E -> 0 gives -1 / (1 + F), F -> inf gives 0.  Absolute error ~2e-7 (checked against float64).
__device__ __forceinline__ float gate_exp(float a, float s) {
  float E, F, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(E) : "f"(fminf(a, 20.f) * 2.885390081777927f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(F) : "f"(s * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((E + 1.f) * (1.f + F)));
  return (E - 1.f) * r;
}
// ---- the gate WITHOUT transcendental-pipe exponentials.  ncu on k1_layer: the XU pipe (MUFU.TANH x 2 per gate channel plus the
// F2FP packs) is 97.5 % busy -- the layer kernel is bound by it, not by the tensor pipe (79 %).  2^x is therefore evaluated on
// the FMA / ALU pipes: x = j + f with j = round(x), |f| <= 0.5 (magic-number rounding), 2^f by a minimax polynomial of degree
// DEG in packed fp32x2 arithmetic, 2^j by adding j into the exponent field.  Relative error 7.5e-5 (DEG 3), 2.6e-6 (DEG 4),
// 7.5e-8 (DEG 5).  One MUFU.RCP per channel remains:
//     tanh(a) sigmoid(s) = (E - 1) / ((E + 1)(1 + F)),  E = 2^(2 log2e a),  F = 2^(-log2e s)
// max abs error of the gate vs float64: 3.7e-5 (DEG 3; MUFU.TANH's 2^-11 gives 7e-4), 1.3e-6 (DEG 4), < 1e-7 (DEG 5).
template <int DEG> __host__ __device__ constexpr float ex2_coef(int k) {
  return DEG == 3 ? (k == 0 ? 0.99992807f : k == 1 ? 0.69326099f : k == 2 ? 0.24261112f : 0.055171638f)
       : DEG == 4 ? (k == 0 ? 0.99999926f : k == 1 ? 0.69312181f : k == 2 ? 0.24024745f : k == 3 ? 0.055917860f : 0.0095700983f)
                  : (k == 0 ? 1.00000007f : k == 1 ? 0.69314697f : k == 2 ? 0.24022120f : k == 3 ? 0.055507133f
                     : k == 4 ? 0.0096755413f : 0.0013276468f);
}
// 2^x for two values; x is clamped to [-64, 64] (the caller's quotient saturates long before)
template <int DEG> __device__ __forceinline__ f32x2 ex2_poly2(f32x2 x) {
  float x0, x1;
  up2(x, x0, x1);
  x = pk2(fminf(fmaxf(x0, -64.f), 64.f), fminf(fmaxf(x1, -64.f), 64.f));
  const float M = 12582912.f;                                       // 1.5 * 2^23: x + M has round(x) in its low mantissa bits
  const f32x2 r = add2(x, pk2(M, M));
  const f32x2 f = fma2(add2(r, pk2(-M, -M)), pk2(-1.f, -1.f), x);   // x - round(x)
  f32x2 p = pk2(ex2_coef<DEG>(DEG), ex2_coef<DEG>(DEG));
#pragma unroll
  for (int k = DEG - 1; k >= 0; --k) p = fma2(p, f, pk2(ex2_coef<DEG>(k), ex2_coef<DEG>(k)));
  float p0, p1, r0, r1;
  up2(p, p0, p1);
  up2(r, r0, r1);
  return pk2(__int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23)),
             __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23)));
}
// the gate of two channels: x = 2 log2e (a_t + b_t), y = -log2e (a_s + b_s) (the caller folds scale and bias into one FFMA2)
template <int DEG> __device__ __forceinline__ f32x2 gate_poly2(f32x2 x, f32x2 y) {
  const f32x2 E = ex2_poly2<DEG>(x), F = ex2_poly2<DEG>(y), one = pk2(1.f, 1.f);
  float d0, d1, r0, r1;
  up2(mul2(add2(E, one), add2(F, one)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
  return mul2(add2(E, pk2(-1.f, -1.f)), pk2(r0, r1));
}
__device__ __forceinline__ float sigmoid_exp(float x) {
  return __fdividef(1.f, 1.f + __expf(-fminf(fmaxf(x, -80.f), 80.f)));
}

}}  // namespace ap::ptx
