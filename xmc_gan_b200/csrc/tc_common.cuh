// tc_common.cuh — sm_100a building blocks for the tcgen05 kernels: mbarrier, TMA, TMEM
// allocation / load / store, UMMA shared-memory + instruction descriptors and the MMA issue.
// Everything is inline PTX (no CUTLASS); descriptor bit layouts follow the PTX ISA
// "tcgen05 matrix descriptor" / "instruction descriptor" tables.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace xmc {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU box.  On timeout the CTA-wide abort flag and
// the global error word are set; every later wait of the CTA returns at once and the kernel runs to
// completion with garbage results.  The error word is read ON THE DEVICE by the kernels that follow
// (xmc_infonce_loss turns the loss into NaN, xmc_normalize_transpose_backward the gradients), so a
// timed-out step can never pass for a good one; no host synchronisation is needed for that.
struct WaitCtx {
  volatile int* abort_flag;   // shared memory
  int* err;                   // global (workspace)
};
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// try_wait itself may suspend the thread for an implementation-defined time, so the bound is on
// elapsed wall time, not on the number of polls.  4 s: two orders of magnitude above the longest
// kernel of the path (so time-slicing or a debugger does not trip it), short enough for a test box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, const WaitCtx& w, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_ns();
  for (uint32_t spin = 0;; ++spin) {
    if (*w.abort_flag) return;
    if (mbar_try_wait(bar, parity)) return;
    if ((spin & 63) == 63 && global_ns() - t0 > 4000000000ull) {
      *w.abort_flag = 1;
      if (w.err) atomicExch(w.err, code + 1000 * (int)blockIdx.x);
      return;
    }
  }
}

// ---- TMA ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// CTA pair (cta_group::2): this CTA's box into ITS shared memory, the bytes completed on the mbarrier at the same offset in
// the pair's LEADER (even rank: bit 24 of a shared::cluster address selects the CTA of the pair)
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// arrive on the mbarrier at this offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// thread-block cluster: rank of this CTA, and a full arrive + wait of every thread of every CTA
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// warm the L2 with a tile that will be loaded later (no smem destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// smem box -> global tile (bulk async-group completion); out-of-bounds rows of the box are clipped
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// smem box += into a global fp32 tile (element-wise add performed by the TMA unit / L2)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
// 16-byte asynchronous copy global -> shared (no registers); src_bytes < 16 zero-fills the rest of the chunk
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// the executing thread's earlier cp.async copies arrive on `bar` when they have landed (counts as one pending arrival)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / TMA reading smem)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 1-D bulk reduce-add (fp32) smem -> global
__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
               ::"l"(reinterpret_cast<uint64_t>(gdst)), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {      // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// pair-wide allocation: the same warp of BOTH CTAs of the pair executes these (same column count, same dst offset)
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp gets lane (base_lane + t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ---- descriptors ----------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B canonical layouts (16-byte units in the fields):
//   K-major : rows of 64 bf16 (128 B), 8-row groups SBO apart; LBO unused (1).
//   MN-major: 64 contiguous MN elements (128 B) per K row, next 64-element MN block LBO apart,
//             8-K-row groups SBO apart.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) |
         (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32) |
         (1ull << 46) |      // descriptor version (Blackwell)
         (2ull << 61);       // SWIZZLE_128B
}
// MN-major operand of 32-bit elements (kind::tf32): layout type SWIZZLE_128B_BASE32B — rows of 128 bytes (32 MN elements),
// 4-K-row groups SBO apart, next 32-element MN block LBO apart; inside a group the 32-BYTE chunk index is XORed with
// (row mod 4) (address bits [5,7) ^= bits [7,9)).
__device__ __forceinline__ uint64_t smem_desc_base32b(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) |
         (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32) |
         (1ull << 46) |
         (1ull << 61);       // SWIZZLE_128B_BASE32B
}
// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32, M x N, per-operand major-ness.
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// kind::tf32: fp32 storage read as tf32 (10-bit mantissa, K = 8 per instruction) -> fp32
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// A descriptor is a plain 64-bit value: only its low 14 bits (start address >> 4) change between
// k-steps / blocks / stages, so stepping it is an add of (byte offset >> 4).
typedef uint64_t Desc;
__device__ __forceinline__ Desc make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return smem_desc(saddr, lbo_bytes, sbo_bytes);
}

// The whole MMA-issuing (and TMA-producing) role runs inside ONE `if (elect_one())` of a warp whose
// index is provably uniform (warp_index() below).  ptxas then keeps descriptors, TMEM addresses and
// barrier addresses in uniform registers and emits back-to-back UTCHMMA with ~2 uniform-datapath
// instructions between them (per-call elect.sync cost ~18 SASS instructions per MMA and left the
// tensor pipe waiting on the issuing warp).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ int warp_index() { return __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0); }

// D[tmem] (+)= A[tmem] * B[smem]          (single issuing thread)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, Desc b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b), "r"(idesc), "r"(accumulate ? 1u : 0u)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]          (single issuing thread)
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, Desc a, Desc b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate ? 1u : 0u)
      : "memory");
}
__device__ __forceinline__ void mma_ss_tf32(uint32_t d_tmem, Desc a, Desc b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate ? 1u : 0u)
      : "memory");
}
// All previously issued MMAs of this thread arrive on `bar` when complete (implies
// fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// cta_group::2: ONE instruction of the pair's leader multiplies over both SMs — D[256 x N]: rows 0-127 in the leader's tensor
// memory, rows 128-255 in the peer's (same column address); A: each CTA's own 128 rows, B: N/2 rows from each CTA, both
// at the descriptor's offsets in either CTA's shared memory.  idesc carries M = 256.
__device__ __forceinline__ void mma_ss_pair(uint32_t d_tmem, Desc a, Desc b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate ? 1u : 0u)
      : "memory");
}
__device__ __forceinline__ void mma_ss_tf32_pair(uint32_t d_tmem, Desc a, Desc b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate ? 1u : 0u)
      : "memory");
}
// completion of the leader's earlier pair MMAs, delivered to the mbarrier at this offset in both CTAs of the pair
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// lane L returns sum over the warp of v[L] (31 shuffles instead of 160)
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float keep = hi ? v[k + off] : v[k];
      const float send = hi ? v[k] : v[k + off];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);   // .x (low 16 bits) = lo
  return *reinterpret_cast<uint32_t*>(&v);
}

constexpr int kStgPitch = 20;    // words per staged row (16 + 4): rows stay 16-byte aligned, the row-per-lane writes are conflict-free
// Epilogue stores.  A thread owns one row of the accumulator tile (TMEM lane), so storing from registers sends each
// 16-byte piece of a warp instruction to a different row: 32 half-used sectors (measured: the epilogues took as long
// as loads + MMAs together).  Instead the warp parks its [32 rows x W words] block in shared memory, one row per lane, and
// re-reads it as (row, 16-byte chunk): every instruction then covers whole 64- or 128-byte row segments.
template <typename F>
__device__ __forceinline__ void warp_rows_out(uint32_t* stg, int lane, const uint32_t* w, F&& emit) {
  constexpr int W = 16, CH = W / 4, RPI = 32 / CH;           // 16 words = four 16-byte chunks per row; 8 rows per warp instruction
  __syncwarp();                                              // the previous block has been read
#pragma unroll
  for (int q = 0; q < CH; ++q)
    *reinterpret_cast<uint4*>(stg + lane * kStgPitch + 4 * q) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 32 / RPI; ++i) {
    const int r = i * RPI + lane / CH, ch = lane % CH;
    emit(r, 4 * ch, *reinterpret_cast<const uint4*>(stg + r * kStgPitch + 4 * ch));
  }
}


}  // namespace tc
}  // namespace xmc
