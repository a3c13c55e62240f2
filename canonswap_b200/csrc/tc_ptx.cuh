// PTX building blocks shared by the tcgen05 convolution kernels (sm_100a): mbarrier, TMA tiled loads,
// tcgen05 fences / commit / ld / st, the K-major SWIZZLE_128B UMMA descriptor and the per-stage MMA issue block.
#pragma once
#include "ctx.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace cs {
namespace tc {

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded spin: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  unsigned long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if ((spin & 1023u) == 1023u) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();       // 4 s
    }
  }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
// cta_group::2 variants: issued by both CTAs of a pair, the transaction bytes land on the LEADER CTA's barrier
// (shared::cluster address with the peer bit cleared, as CUTLASS' SM100_TMA_2SM_LOAD does).
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3,
                                                int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 format: version 1 at bits 46-47,
// layout type 2 at bits 61-63, stride byte offset = 1024 B between 8-row groups). The low word is the
// 16-byte-granular start address, so a K advance of 32 B (one 16-element bf16 K step) is +2 and the
// lo half of a [hi x32 | lo x32] row (+64 B) is +4.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  constexpr uint64_t HI = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29))) << 32;
  return HI | (uint64_t)((saddr & 0x3FFFFu) >> 4);
}

// All the MMAs of one pipeline stage (one 32-channel K block) + the commit that frees the stage, issued
// by one elected lane of a converged warp in a single asm block (the issue thread is the critical path
// of the kernel: no divergence bookkeeping, no per-MMA descriptor rebuild).
//   hi*hi -> [d_main] (first MMA accumulates iff acc_main), lo*hi and hi*lo -> [d_corr] (iff acc_corr).
// (ilh / ihl are the instruction descriptors of the lo*hi / hi*lo MMAs: same as hi*hi, both halves are bf16.)
#define CS_MMA_HEAD                                   \
  "{\n\t"                                             \
  ".reg .pred pe, pm, pc, pt;\n\t"                    \
  ".reg .b64 a2, a4, a6, b2, b4, b6;\n\t"             \
  ".reg .b32 ilh, ihl;\n\t"                           \
  "mov.b32 ilh, %4;\n\t"                              \
  "mov.b32 ihl, %4;\n\t"                              \
  "elect.sync _|pe, 0xffffffff;\n\t"                  \
  "setp.ne.b32 pm, %5, 0;\n\t"                        \
  "setp.ne.b32 pc, %6, 0;\n\t"                        \
  "setp.eq.b32 pt, %4, %4;\n\t"                       \
  "add.s64 a2, %2, 2;\n\t"                            \
  "add.s64 a4, %2, 4;\n\t"                            \
  "add.s64 a6, %2, 6;\n\t"                            \
  "add.s64 b2, %3, 2;\n\t"                            \
  "add.s64 b4, %3, 4;\n\t"                            \
  "add.s64 b6, %3, 6;\n\t"
#define CS_MMA(D, A, B, I, P) "@pe tcgen05.mma.cta_group::1.kind::f16 [" D "], " A ", " B ", " I ", " P ";\n\t"
#define CS_MMA_TAIL "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t}"
// pair variants: one MMA spans both CTAs (M = 256), the commit arrives on the barrier of BOTH CTAs
#define CS_MMA2(D, A, B, I, P) "@pe tcgen05.mma.cta_group::2.kind::f16 [" D "], " A ", " B ", " I ", " P ";\n\t"
#define CS_MMA2_TAIL                                                                                         \
  "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%7], mk;\n\t}"
#define CS_MMA2_HEAD CS_MMA_HEAD ".reg .b16 mk;\n\tmov.b16 mk, 3;\n\t"
#define CS_MMA_OPS                                                                                                     \
  ::"r"(d_main), "r"(d_corr), "l"(ad), "l"(bd), "r"(idesc), "r"(acc_main), "r"(acc_corr), "r"(bar) : "memory"

template <int NPASS, int KSTEPS, int CTAS = 1>
__device__ __forceinline__ void mma_stage(uint32_t d_main, uint32_t d_corr, uint64_t ad, uint64_t bd, uint32_t idesc,
                                          uint32_t acc_main, uint32_t acc_corr, uint32_t bar) {
  if constexpr (CTAS == 2) {
    static_assert(NPASS == 3, "pair MMAs are built for the 3-pass path only");
    if constexpr (KSTEPS == 2) {
      asm volatile(CS_MMA2_HEAD CS_MMA2("%0", "%2", "%3", "%4", "pm") CS_MMA2("%0", "a2", "b2", "%4", "pt")
                   CS_MMA2("%1", "a4", "%3", "ilh", "pc") CS_MMA2("%1", "a6", "b2", "ilh", "pt")
                   CS_MMA2("%1", "%2", "b4", "ihl", "pt") CS_MMA2("%1", "a2", "b6", "ihl", "pt") CS_MMA2_TAIL CS_MMA_OPS);
    } else {
      asm volatile(CS_MMA2_HEAD CS_MMA2("%0", "%2", "%3", "%4", "pm") CS_MMA2("%1", "a4", "%3", "ilh", "pc") CS_MMA2("%1", "%2", "b4", "ihl", "pt")
                       CS_MMA2_TAIL CS_MMA_OPS);
    }
  } else if constexpr (NPASS == 3 && KSTEPS == 2) {
    asm volatile(CS_MMA_HEAD CS_MMA("%0", "%2", "%3", "%4", "pm") CS_MMA("%0", "a2", "b2", "%4", "pt")      // hi*hi
                 CS_MMA("%1", "a4", "%3", "ilh", "pc") CS_MMA("%1", "a6", "b2", "ilh", "pt")                  // lo*hi
                 CS_MMA("%1", "%2", "b4", "ihl", "pt") CS_MMA("%1", "a2", "b6", "ihl", "pt") CS_MMA_TAIL CS_MMA_OPS);
  } else if constexpr (NPASS == 3 && KSTEPS == 1) {
    asm volatile(CS_MMA_HEAD CS_MMA("%0", "%2", "%3", "%4", "pm") CS_MMA("%1", "a4", "%3", "ilh", "pc") CS_MMA("%1", "%2", "b4", "ihl", "pt")
                     CS_MMA_TAIL CS_MMA_OPS);
  } else if constexpr (NPASS == 2 && KSTEPS == 2) {
    asm volatile(CS_MMA_HEAD CS_MMA("%0", "%2", "%3", "%4", "pm") CS_MMA("%0", "a2", "b2", "%4", "pt") CS_MMA("%1", "a4", "%3", "ilh", "pc")
                     CS_MMA("%1", "a6", "b2", "ilh", "pt") CS_MMA_TAIL CS_MMA_OPS);
  } else if constexpr (NPASS == 2 && KSTEPS == 1) {
    asm volatile(CS_MMA_HEAD CS_MMA("%0", "%2", "%3", "%4", "pm") CS_MMA("%1", "a4", "%3", "ilh", "pc") CS_MMA_TAIL CS_MMA_OPS);
  } else if constexpr (NPASS == 1 && KSTEPS == 2) {
    asm volatile(CS_MMA_HEAD CS_MMA("%0", "%2", "%3", "%4", "pm") CS_MMA("%0", "a2", "b2", "%4", "pt") CS_MMA_TAIL CS_MMA_OPS);
  } else {
    asm volatile(CS_MMA_HEAD CS_MMA("%0", "%2", "%3", "%4", "pm") CS_MMA_TAIL CS_MMA_OPS);
  }
}

// the same 6 MMAs (3 passes x 2 K-steps, one CTA) without the commit: several operand windows of one pipeline stage
// (conv3s_tc.cu: the three kh taps of a halo tile) are issued back to back and the last one frees the stage
__device__ __forceinline__ void mma_stage32_nocommit(uint32_t d_main, uint32_t d_corr, uint64_t ad, uint64_t bd, uint32_t idesc,
                                                     uint32_t acc_main, uint32_t acc_corr) {
  const uint32_t bar = 0;
  asm volatile(CS_MMA_HEAD CS_MMA("%0", "%2", "%3", "%4", "pm") CS_MMA("%0", "a2", "b2", "%4", "pt")
               CS_MMA("%1", "a4", "%3", "ilh", "pc") CS_MMA("%1", "a6", "b2", "ilh", "pt")
               CS_MMA("%1", "%2", "b4", "ihl", "pt") CS_MMA("%1", "a2", "b6", "ihl", "pt") "}" CS_MMA_OPS);
}

template <int NPASS, int CTAS = 1>
__device__ __forceinline__ void mma_stage_k(int ksteps, uint32_t d_main, uint32_t d_corr, uint64_t ad, uint64_t bd,
                                            uint32_t idesc, uint32_t acc_main, uint32_t acc_corr, uint32_t bar) {
  if (ksteps == 2) mma_stage<NPASS, 2, CTAS>(d_main, d_corr, ad, bd, idesc, acc_main, acc_corr, bar);
  else mma_stage<NPASS, 1, CTAS>(d_main, d_corr, ad, bd, idesc, acc_main, acc_corr, bar);
}


__device__ __forceinline__ void tc_st16_zero(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int TC_THREADS = 192;                             // 2 role warps + 4 epilogue warps
constexpr int TC_THREADS_MAX = 320;                         // ... or 8 epilogue warps (conv_tc.cu, wide tiles)
constexpr int A_TILE_BYTES = 128 * 128;
constexpr int STG_LD = 36;                                  // floats per staged row (32 + pad, 16-B aligned)
constexpr int STG_BYTES = 4 * 32 * STG_LD * 4;              // one 32x32 fp32 staging tile per epilogue warp

// ------------------------------------------------------------------------------------------
// Accumulator plan + position-dependent truncation pre-compensation
// ------------------------------------------------------------------------------------------
// The tensor core adds every MMA into the fp32 TMEM accumulator with truncation toward zero: event t loses ~kappa * acc_t,
// so an element loses kappa * sum_t acc_t = kappa * sum_s rem(s) * m_s, where m_s is the contribution of K step s and
// rem(s) the number of truncation events from its entry to the end of the chain.  That is a LINEAR functional of the
// products whose coefficients depend only on the issue order -- so the weights of K step s are packed pre-multiplied by
// (1 + kappa * rem(s)) and the loss cancels to first order per element, at no run-time cost (measured on B200,
// tools/poscomp_probe.py: per-conv rms error 2.5 - 2.8x below the constant epilogue factor it replaces; kappa = 3.3e-8).
// The plan (sets, chunk) therefore belongs to the packed weights: pack_tc fixes it, conv_tc launches with it.
struct TcPlan {
  int nsets = 1;      // TMEM accumulator sets of BN columns per buffer (set 0 = corrections when nsets > 1 and npass > 1)
  int chunk = 1 << 30;// K iterations (32-channel blocks) per hi*hi set
  int nacc = 1;       // accumulator buffers (2 = epilogue of tile i overlaps the MMAs of tile i+1)
  int npass = 3;
  bool thin = false;   // two co-resident CTAs per SM, half of the TMEM columns each
};
// hi*hi events issued in iterations [0, i) of a tap-major / block-minor walk (the last block of a tap may hold one K step)
__host__ __device__ inline int tc_events_upto(int i, int nblk, int last_ksteps) { return 2 * i - (i / nblk) * (2 - last_ksteps); }
// truncation events from the entry of (iteration it, K step ks) into its accumulator to the end of that accumulator's chain
__host__ __device__ inline int tc_remaining_events(int nsets, int chunk, int npass, int niter, int nblk, int last_ksteps, int it, int ks) {
  const bool corr = npass > 1 && nsets > 1;
  if (nsets == 1) return npass * (tc_events_upto(niter, nblk, last_ksteps) - tc_events_upto(it, nblk, last_ksteps)) - ks;
  (void)corr;
  int end = (it / chunk + 1) * chunk;
  if (end > niter) end = niter;
  return tc_events_upto(end, nblk, last_ksteps) - tc_events_upto(it, nblk, last_ksteps) - ks;
}
TcPlan tc_make_plan(int BN, int niter, int npass, int single_chain, int max_sets, bool double_buffer, int single_main);   // conv_tc.cu

PFN_cuTensorMapEncodeTiled_v12000 encode_fn();             // conv_tc.cu
int pick_box(int dim, int cap, int* log2out);               // conv_tc.cu

}  // namespace tc
}  // namespace cs
