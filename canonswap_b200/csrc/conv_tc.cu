// tcgen05 implicit-GEMM convolution for sm_100a (the dense-contraction hot path: every conv2d / conv3d
// of the five networks with Cin >= 16 and Cout >= 8).
//
//   D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * W[cout, tap, cin]
//
//   M tile  = 128 output pixels = one 4-D spatial box (bb x bd x bh x bw) of the channels-last input,
//             so the A tile of filter tap (kd,kh,kw) is the SAME box shifted by the tap: one TMA tiled
//             load per stage, the zero padding of the convolution is TMA out-of-bounds zero fill.
//   N tile  = BN <= 256 output channels, K = taps x Cin walked in 32-channel blocks.
//   operand = split fp16: every fp32 value v is stored as hi = fp16(v), lo = fp16(v - hi); a 32-channel
//             block is the 128-byte row [hi x32 | lo x32], which is exactly one SWIZZLE_128B row, so a
//             stage is ONE A box (128 rows) and ONE B box (BN rows) and the MMA descriptors of the
//             hi / lo halves are 64-byte K-advances inside the swizzle atom.
//   math    = 3 x tcgen05.mma.kind::f16 (fp16 x fp16 -> fp32 in TMEM) per 16-channel K step:
//             hi*hi + hi*lo + lo*hi  (~2^-22 relative, for the 1e-3 fp32 parity bar of BASELINE.json);
//             `npass` 2 / 1 drop the correction terms (measurement only).
//   roles   = warp 0: TMA producer, warp 1: TMEM alloc + MMA issuer, warps 2-5 (2-9 on wide tiles): epilogue
//             (tcgen05.ld -> bias / activation / residual / per-pixel multiplier -> fp32 channels-last and / or the next
//             conv's operand); persistent tile loop, double-buffered TMEM, tcgen05 pair mode: DESIGN.md section 4.2.
//   the 16-bit storage type of the operands is spelled __nv_bfloat16 in the code (a 16-bit slot); the bits are fp16.
#include "tc_ptx.cuh"
#include <mutex>

namespace cs {

using namespace tc;

namespace {

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
struct ConvTcK {
  int B, D, H, W;                  // output geometry (the tensor map carries the input extents)
  int lbw, lbh, lbd, lbb;          // log2 of the box extents (product 128)
  int ntw, nth, ntd;               // tiles per dimension (batch tiles = gridDim.x / (ntw*nth*ntd))
  int KD, KH, KW, PD, PH, PW;
  int nblk;                        // 32-channel blocks of the (padded) input
  int last_ksteps;                 // 16-channel K steps in the last block (1 or 2)
  int rowA;                        // bf16 elements per pixel of the operand = nblk * 64
  int BN, Cout, stages, npass;
  int tcols, nsets, chunk;         // TMEM columns, accumulator sets of BN columns, K iterations per hi*hi set
  int nacc;                        // accumulator buffers (of nsets * BN columns): 2 = the epilogue of tile i overlaps the MMAs of tile i+1
  int m_units, n_tiles;            // persistent tile loop: units of CTAS consecutive M tiles x N tiles (unit u -> m = u % m_units, n = u / m_units)
  float acc_scale;                 // compensation of the tensor core's round-toward-zero accumulation (see host code)
  float kappa;                     // > 0: the weights carry the position-dependent pre-compensation; the epilogue takes back what it
  int Din;                         //      assumed for filter taps that fall into the zero padding (no products, no truncation)
  float out_scale;                 // 1 / ConvW::wmul
  int zrows;                       // > 0: depth-dependent weights, B rows of depth slice d start at d * zrows
  const float* bias; int act; float slope;
  const float* res; long rb, rd, rh, rw;
  const float* mult;
  float* y; long yb, yd, yh, yw;               // may be null when only the operand is emitted
  int vec4;
  int plain;                       // no bias / activation / residual / multiplier / emission: the rows are copied out as they are
  // optional: also write act(v * escale[n] + eshift[n]) as the split-fp16 operand of the next conv (dense, output geometry)
  __nv_bfloat16* emit; int erow; const float* escale; const float* eshift; int eact; float eslope; float emul;
  // SPADE epilogue (see Epilogue::sp_x)
  const float* sp_x; const float* sp_mean; const float* sp_rstd; int sp_C, sp_xs, sp_Hx, sp_Wx;
  // phase mode (conv of a nearest-upsampled input computed on the low-resolution operand, see Epilogue::phase_shift):
  // N tile p = output phase (a, b) = (p >> ph_s, p & (2^ph_s - 1)); only the taps in tapmask[p] are walked; the tile's
  // pixels land at (oh * 2^ph_s + a, ow * 2^ph_s + b) of the output.
  int ph_s; unsigned tapmask[16];                 // bit t = filter tap t (kd-major, up to 27 taps) is walked by this phase
  int ph_tpp;                                     // N tiles per phase (a phase produces ph_tpp * BN output channels)
};


struct TileOrg { int w0, h0, d0, b0, n0, nrow0, ph_a, ph_b, chan0; uint32_t tmask; };
// origin of M tile `m_tile` (a 4-D spatial box) and N tile `n_idx`
__device__ __forceinline__ TileOrg tile_origin(const ConvTcK& k, int m_tile, int n_idx) {
  TileOrg o;
  int t = m_tile;
  const int tw = t % k.ntw; t /= k.ntw;
  const int th = t % k.nth; t /= k.nth;
  const int td = t % k.ntd; const int tb = t / k.ntd;
  o.w0 = tw << k.lbw; o.h0 = th << k.lbh; o.d0 = td << k.lbd; o.b0 = tb << k.lbb;
  o.n0 = n_idx * k.BN;
  o.nrow0 = o.n0 + o.d0 * k.zrows;               // first B row of this tile
  const int phase = k.ph_s ? n_idx / k.ph_tpp : 0;
  o.tmask = k.ph_s ? k.tapmask[phase] : 0xFFFFFFFFu;
  o.ph_a = k.ph_s ? (phase >> k.ph_s) : 0;
  o.ph_b = k.ph_s ? (phase & ((1 << k.ph_s) - 1)) : 0;
  o.chan0 = k.ph_s ? phase * k.ph_tpp * k.BN : 0;  // every phase produces output channels 0 .. ph_tpp * BN - 1
  return o;
}

// RES / EMIT: compile-time epilogue variants (residual add, operand emission) -- the epilogue is not overlapped
// with the main loop, so the plain variant must not carry the registers of the fused ones.
// CTAS = 2: the CTAs of a 2-CTA cluster (two consecutive M tiles, same N tile) run as a tcgen05 pair: each loads its own
// A tile and HALF of the B tile, the leader issues cta_group::2 MMAs (M = 256) that read A / B from both CTAs' shared
// memory and accumulate into both CTAs' TMEM.  Halving the B bytes written and read per CTA takes the kernel off the
// shared-memory bandwidth limit that bounds the 1-CTA form at N = 256 (3 MMAs per operand load).
// BCOMP: border take-back of the truncation pre-compensation (multi-tap convs whose weights carry it); a template parameter
// so that the 1x1 / Winograd GEMM instantiations, whose epilogue paces the kernel, carry none of its code.
// sigmoid / GELU out of line: one copy of the expf / erff code per kernel instead of one per unrolled call site
__device__ __noinline__ float4 act_rare4(float4 v, int kind) {
  return make_float4(apply_act(v.x, kind, 0.f), apply_act(v.y, kind, 0.f), apply_act(v.z, kind, 0.f), apply_act(v.w, kind, 0.f));
}

template <bool RES, bool EMIT, int CTAS, bool SPADE = false, bool BCOMP = false>
__global__ void __launch_bounds__(TC_THREADS_MAX) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB, ConvTcK k) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint32_t cta_rank = 0;
  if constexpr (CTAS == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const uint32_t brows = (uint32_t)k.BN / CTAS;                       // B rows held by this CTA
  const uint32_t stage_bytes = A_TILE_BYTES + brows * 128u;
  const uint32_t stg = base + (uint32_t)k.stages * stage_bytes;       // epilogue staging
  const int egroups = ((int)(blockDim.x >> 5) - 2) >> 2;              // epilogue warp groups (4 warps each): 1 or 2
  const uint32_t bars = stg + (uint32_t)egroups * STG_BYTES;          // full[stages], empty[stages], tmem_full[2], tmem_empty[2], tmem slot
  const uint32_t tmem_full = bars + 16u * k.stages;
  const uint32_t tmem_empty = tmem_full + 16u;
  const uint32_t tmem_slot = tmem_empty + 16u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (uint32_t)k.tcols;
  // Accumulator sets.  The tensor core adds every MMA into the fp32 TMEM accumulator with truncation, so
  // the error of one accumulator grows linearly with its chain length (~2^-24 |acc| per MMA, measured).
  // The chain is therefore split over `nsets` accumulators of BN columns that the epilogue sums in
  // fp32 registers: set 0 takes the small correction products (lo*hi, hi*lo), sets 1.. take hi*hi of
  // consecutive K ranges of `chunk` iterations.
  const int corr = (k.npass > 1 && k.nsets > 1) ? 1 : 0;
  const uint32_t acc_cols = (uint32_t)(k.nsets * k.BN);
  const int taps = k.KD * k.KH * k.KW;
  // persistent tile loop: cluster c (one CTA, or a tcgen05 pair) walks units c, c + #clusters, ...
  const int unit0 = (int)(blockIdx.x / CTAS), unit_step = (int)(gridDim.x / CTAS), units = k.m_units * k.n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < k.stages; ++s) { mbar_init(bars + 8u * s, 1); mbar_init(bars + 8u * (k.stages + s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full + 8u * b, 1);
      mbar_init(tmem_empty + 8u * b, (uint32_t)(CTAS * ((int)(blockDim.x >> 5) - 2)));   // every epilogue warp of the cluster
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // Pair mode: tcgen05.alloc.cta_group::2 is a two-CTA protocol -- the leader CTA finds the columns for both SMs and hands
  // the result to its peer through a mailbox in the peer's (reserved) shared memory.  Like any access to a peer's shared
  // memory it must not start before the peer CTA is known to be running: without this barrier a peer that started late
  // (two concurrent lanes competing for registers) missed the message and spun in its alloc forever.
  if constexpr (CTAS == 2) cluster_sync_all();
  if (warp == 1) {
    if constexpr (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  if constexpr (CTAS == 2) {
    if (warp == 1) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  if (warp == 0) {
    // ===== TMA producer (one lane): runs ahead across tiles, bounded only by the free pipeline stages =====
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int u = unit0; u < units; u += unit_step) {
        const TileOrg o = tile_origin(k, (u % k.m_units) * CTAS + (int)cta_rank, u / k.m_units);
        for (int tap = 0; tap < taps; ++tap) {
          if (!((o.tmask >> (tap & 31)) & 1u)) continue;
          const int kw = tap % k.KW; const int r = tap / k.KW; const int kh = r % k.KH; const int kd = r / k.KH;
          const int cw = o.w0 + kw - k.PW, ch = o.h0 + kh - k.PH, cd = o.d0 + kd - k.PD;
          const int kcol = tap * k.rowA;
          for (int blk = 0; blk < k.nblk; ++blk) {
            const uint32_t fb = bars + 8u * s;
            mbar_wait(fb + 8u * k.stages, ph ^ 1u);
            const uint32_t sa = base + (uint32_t)s * stage_bytes;
            if constexpr (CTAS == 2) {
              if (cta_rank == 0) mbar_expect_tx(fb, 2u * stage_bytes);          // both CTAs' bytes land on the leader's barrier
              tma_load_5d_2sm(sa, &tmA, fb, blk * 64, cw, ch, cd, o.b0);
              tma_load_2d_2sm(sa + A_TILE_BYTES, &tmB, fb, kcol + blk * 64, o.nrow0 + (int)(cta_rank * brows));
            } else {
              mbar_expect_tx(fb, stage_bytes);
              tma_load_5d(sa, &tmA, fb, blk * 64, cw, ch, cd, o.b0);
              tma_load_2d(sa + A_TILE_BYTES, &tmB, fb, kcol + blk * 64, o.nrow0);
            }
            if (++s == k.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1 && (CTAS == 1 || cta_rank == 0)) {
    // ===== MMA issuer: the whole warp walks the loop converged, one elected lane issues =====
    // instruction descriptor: D fp32, A/B fp16 (IDESC_AB_FMT), both K-major, N = BN, M = 128 per CTA
    const uint32_t idesc = (1u << 4) | IDESC_AB_FMT | ((uint32_t)(k.BN >> 3) << 17) | (((128u * CTAS) >> 4) << 24);
    int s = 0; uint32_t ph = 0;
    int it = 0;
    for (int u = unit0; u < units; u += unit_step, ++it) {
      const uint32_t tmask = k.ph_s ? k.tapmask[(u / k.m_units) / k.ph_tpp] : 0xFFFFFFFFu;
      const int buf = k.nacc == 2 ? (it & 1) : 0, use = k.nacc == 2 ? (it >> 1) : it;
      if (use > 0) {                                       // the epilogue warps (of both CTAs) have drained this buffer
        mbar_wait(tmem_empty + 8u * buf, (uint32_t)((use - 1) & 1));
        tc_fence_after();
      }
      const uint32_t d_corr0 = tmem_base + (uint32_t)buf * acc_cols;
      uint32_t d_main = d_corr0 + (uint32_t)(corr * k.BN);
      int in_set = 0;
      uint32_t acc_corr = 0;
      for (int tap = 0; tap < taps; ++tap) {
        if (!((tmask >> (tap & 31)) & 1u)) continue;
        for (int blk = 0; blk < k.nblk; ++blk) {
          const uint32_t fb = bars + 8u * s;
          mbar_wait(fb, ph);
          tc_fence_after();
          const uint32_t sa = base + (uint32_t)s * stage_bytes;
          const uint64_t ad = umma_desc(sa), bd = umma_desc(sa + A_TILE_BYTES);
          const int ksteps = (blk == k.nblk - 1) ? k.last_ksteps : 2;
          const uint32_t acc_main = in_set > 0 ? 1u : 0u;
          const uint32_t eb = fb + 8u * k.stages;
          if constexpr (CTAS == 2) {
            mma_stage_k<3, 2>(ksteps, d_main, corr ? d_corr0 : d_main, ad, bd, idesc, acc_main, corr ? acc_corr : 1u, eb);
          } else if (corr && k.npass == 3) {
            mma_stage_k<3>(ksteps, d_main, d_corr0, ad, bd, idesc, acc_main, acc_corr, eb);
          } else if (corr) {
            mma_stage_k<2>(ksteps, d_main, d_corr0, ad, bd, idesc, acc_main, acc_corr, eb);
          } else if (k.npass == 3) {                       // single accumulator: corrections follow hi*hi in place
            mma_stage_k<3>(ksteps, d_main, d_main, ad, bd, idesc, acc_main, 1u, eb);
          } else if (k.npass == 2) {
            mma_stage_k<2>(ksteps, d_main, d_main, ad, bd, idesc, acc_main, 1u, eb);
          } else {
            mma_stage_k<1>(ksteps, d_main, d_main, ad, bd, idesc, acc_main, 1u, eb);
          }
          acc_corr = 1u;
          if (++in_set == k.chunk) { in_set = 0; d_main += (uint32_t)k.BN; }
          if (++s == k.stages) { s = 0; ph ^= 1u; }
        }
      }
      const uint32_t tf = tmem_full + 8u * buf;
      if constexpr (CTAS == 2) {
        asm volatile(
            "{\n\t.reg .pred pe;\n\t.reg .b16 mk;\n\tmov.b16 mk, 3;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], mk;\n\t}"
            ::"r"(tf) : "memory");
      } else {
        asm volatile(
            "{\n\t.reg .pred pe;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
            ::"r"(tf) : "memory");
      }
    }
  } else if (warp >= 2) {
    // ===== epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 =====
    // phase 1: each lane pulls its pixel row (32 columns, all accumulator sets summed) into a padded smem tile;
    // phase 2: the warp walks the tile 4 rows x 128 B at a time so global stores / residual loads are coalesced.
    // with 2 groups (wide tiles) the groups take alternate 32-column chunks: the epilogue is not overlapped with the main
    // loop, so its throughput is on the critical path of every tile
    const int q = warp & 3;
    const int eg = (warp - 2) >> 2;
    float* tile = reinterpret_cast<float*>(smem_raw + (stg - raw)) + (eg * 4 + q) * 32 * STG_LD;
    // row mapping of the coalesced phase: 4 rows x 8 lanes (x 8 passes); SPADE: 8 rows x 4 lanes (x 4 passes)
    constexpr int RSTEP = SPADE ? 8 : 4;
    const int sub = SPADE ? (lane >> 2) : (lane >> 3), c4 = (lane & 7) * 4;
    const bool act_rare = !act_is_leaky(k.act);              // sigmoid / GELU: two layers of the whole path
    const float aslope = leaky_slope(k.act, k.slope), easlope = leaky_slope(k.eact, k.eslope);
    int it = 0;
    for (int u = unit0; u < units; u += unit_step, ++it) {
    const TileOrg o = tile_origin(k, (u % k.m_units) * CTAS + (int)cta_rank, u / k.m_units);
    const int w0 = o.w0, h0 = o.h0, d0 = o.d0, b0 = o.b0, n0 = o.n0, ph_a = o.ph_a, ph_b = o.ph_b, chan0 = o.chan0;
    long yoff[8], roff[RES ? 8 : 1], epix[EMIT ? 8 : 1], xoff[SPADE ? 8 : 1];
    int sbase[SPADE ? 8 : 1];
    float mu[8];
    uint32_t vmask = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int r = q * 32 + sub + RSTEP * i;
      const int ow = w0 + (r & ((1 << k.lbw) - 1)); r >>= k.lbw;
      const int oh = h0 + (r & ((1 << k.lbh) - 1)); r >>= k.lbh;
      const int od = d0 + (r & ((1 << k.lbd) - 1)); r >>= k.lbd;
      const int ob = b0 + r;
      const bool valid = (SPADE ? i < 4 : true) && ow < k.W && oh < k.H && od < k.D && ob < k.B;
      if (valid) vmask |= 1u << i;
      const int oh2 = (oh << k.ph_s) + ph_a, ow2 = (ow << k.ph_s) + ph_b;        // output position (phase mode: upsampled grid)
      yoff[i] = ob * k.yb + od * k.yd + oh2 * k.yh + ow2 * k.yw;
      if constexpr (RES) roff[i] = ob * k.rb + od * k.rd + oh2 * k.rh + ow2 * k.rw;
      if constexpr (EMIT) epix[i] = (((long)ob * k.D + od) * (k.H << k.ph_s) + oh2) * (k.W << k.ph_s) + ow2;
      if constexpr (SPADE) {
        xoff[i] = (((long)ob * k.sp_Hx + (oh >> k.sp_xs)) * k.sp_Wx + (ow >> k.sp_xs)) * k.sp_C;
        sbase[i] = ob * k.sp_C;
      }
      mu[i] = (k.mult && valid) ? k.mult[(((long)ob * k.D + od) * k.H + oh) * k.W + ow] : 1.f;
    }
    // Border rows: the packed pre-compensation assumes that every K step is followed by the full chain of truncation
    // events, but taps in the zero padding add exact zeros (no truncation).  Depth padding removes a SUFFIX of the tap
    // order (kd is the slowest index): a uniform over-compensation of kappa * (zero events of the set), taken back exactly;
    // h / w padding removes interleaved taps: kappa * fraction * (real events) / 2 on average.
    int b_it0 = 0x7fffffff, b_lead = 0; float b_frac = 0.f;
    constexpr bool bcomp = BCOMP;
    if constexpr (BCOMP) {
      int r = q * 32 + lane;
      const int ow = w0 + (r & ((1 << k.lbw) - 1)); r >>= k.lbw;
      const int oh = h0 + (r & ((1 << k.lbh) - 1)); r >>= k.lbh;
      const int od = d0 + (r & ((1 << k.lbd) - 1));
      int vh = 0, vw = 0;
      for (int kh = 0; kh < k.KH; ++kh) vh += (oh + kh - k.PH >= 0 && oh + kh - k.PH < k.H) ? 1 : 0;
      for (int kw = 0; kw < k.KW; ++kw) vw += (ow + kw - k.PW >= 0 && ow + kw - k.PW < k.W) ? 1 : 0;
      b_frac = 1.f - (float)(vh * vw) / (float)(k.KH * k.KW);
      if (k.KD > 1 && k.PD > 0) {
        const int kd0 = k.Din - od + k.PD;                   // first depth tap beyond the last input slice
        if (kd0 < k.KD) b_it0 = (kd0 < 0 ? 0 : kd0) * k.KH * k.KW * k.nblk;
        const int lead = k.PD - od;                          // depth taps before the first input slice
        if (lead > 0) b_lead = lead * k.KH * k.KW * k.nblk;
      }
    }
    const int niter_all = taps * k.nblk;
    auto set_factor = [&](int set_idx) -> float {            // set_idx: hi*hi set (0-based); single accumulator: the whole chain
      const int a = k.nsets == 1 ? 0 : set_idx * k.chunk;
      int e = k.nsets == 1 ? niter_all : a + k.chunk;
      if (e > niter_all) e = niter_all;
      const int mul = k.nsets == 1 ? k.npass : 1;
      const int z0 = b_it0 > a ? (b_it0 < e ? b_it0 : e) : a;         // zero suffix [z0, e)
      const int r0 = b_lead > a ? (b_lead < e ? b_lead : e) : a;      // real iterations [r0, z0)
      const int zev = (tc_events_upto(e, k.nblk, k.last_ksteps) - tc_events_upto(z0, k.nblk, k.last_ksteps)) * mul;
      const int rev = z0 > r0 ? (tc_events_upto(z0, k.nblk, k.last_ksteps) - tc_events_upto(r0, k.nblk, k.last_ksteps)) * mul : 0;
      return 1.f - k.kappa * ((float)zev + 0.5f * b_frac * (float)rev);
    };
    if (RES && eg == 0) {                                   // pull the residual tile towards L2 while the MMAs run
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (!((vmask >> i) & 1u)) continue;
        for (int c = c4 * 8; c < k.BN && n0 + c < k.Cout; c += 256)       // one 128-B line per lane, 8 lanes per row
          asm volatile("prefetch.global.L2 [%0];" ::"l"(k.res + roff[i] + n0 + c));
      }
    }
    // SPADE: the input x of every chunk this warp will emit does not depend on the accumulators: it is gathered into
    // registers here, i.e. while the MMAs of the tile are still running (<= 4 chunks per warp: 2 epilogue groups for
    // BN > 64, checked by the host).  Only the x loads are issued here -- 16 independent 16-byte loads in flight; the
    // (L1-resident) statistics are read and x_hat = (x - mean) * rstd is formed per row after the accumulator wait
    // (ncu: with the three dependent loads per row in this place the epilogue warps spent half their time stalled on them).
    float4 xh[SPADE ? 4 : 1][SPADE ? 4 : 1];
    if constexpr (SPADE) {
      const int cc = (lane & 3) * 4;
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int c0 = (eg + ci * egroups) * 32;
        const bool live = c0 < k.BN && n0 + c0 < k.Cout;
        const int ch = ((n0 + c0) >> 1) + cc;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          xh[ci][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (live && ((vmask >> i) & 1u)) xh[ci][i] = *reinterpret_cast<const float4*>(k.sp_x + xoff[i] + ch);
        }
      }
    }
    const int buf = k.nacc == 2 ? (it & 1) : 0, use = k.nacc == 2 ? (it >> 1) : it;
    mbar_wait(tmem_full + 8u * buf, (uint32_t)(use & 1));
    tc_fence_after();
    const uint32_t trow = tmem_base + (uint32_t)buf * acc_cols + ((uint32_t)(q * 32) << 16);
    int cidx = -1;
    for (int c0 = eg * 32; c0 < k.BN; c0 += 32 * egroups) {
      ++cidx;
      if (n0 + c0 >= k.Cout) break;                         // warp-uniform
      uint32_t vv[2][16];                                   // both halves of the chunk in flight before one wait
#pragma unroll
      for (int half = 0; half < 2; ++half)
        if (c0 + 16 * half < k.BN) tc_ld16(trow + (uint32_t)(c0 + 16 * half) + (uint32_t)(corr * k.BN), vv[half]);
      tc_ld_wait();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (c0 + 16 * half < k.BN) {                        // warp-uniform (BN is a multiple of 16)
          uint32_t* v = vv[half];
          const uint32_t tcol = trow + (uint32_t)(c0 + 16 * half);
          if constexpr (BCOMP) {
            const float f0 = set_factor(0);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * f0);
          }
          for (int st = corr + 1; st < k.nsets; ++st) {     // hi*hi sets in K order ...
            uint32_t u[16];
            tc_ld16(tcol + (uint32_t)(st * k.BN), u);
            tc_ld_wait();
            float fs = 1.f;
            if constexpr (BCOMP) fs = set_factor(st - corr);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(fmaf(__uint_as_float(u[j]), fs, __uint_as_float(v[j])));
          }
          if (corr) {                                       // ... then the small correction terms
            uint32_t u[16];
            tc_ld16(tcol, u);
            tc_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
              v[j] = __float_as_uint(fmaf(__uint_as_float(v[j]), k.acc_scale, __uint_as_float(u[j])) * k.out_scale);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * k.acc_scale * k.out_scale);
          }
          float4* dst = reinterpret_cast<float4*>(tile + lane * STG_LD + 16 * half);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                 __uint_as_float(v[4 * j + 3]));
        }
      }
      __syncwarp();
      if constexpr (SPADE) {
        // chunk = [gamma of 16 channels | beta of the same 16 channels]; lane -> (row of 8, 4 of the 16 channels)
        const int cc = (lane & 3) * 4;
        const int ncol = n0 + c0;
        const int ch = (ncol >> 1) + cc;
        float4 bg = make_float4(0.f, 0.f, 0.f, 0.f), bb = bg;
        if (k.bias) {
          bg = __ldg(reinterpret_cast<const float4*>(k.bias + ncol + cc));
          bb = __ldg(reinterpret_cast<const float4*>(k.bias + ncol + 16 + cc));
        }
        float4 xv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          xv[i] = xh[0][i];
#pragma unroll
          for (int ci = 1; ci < 4; ++ci) if (ci == cidx) xv[i] = xh[ci][i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!((vmask >> i) & 1u)) continue;
          const float* trp = tile + (sub + 8 * i) * STG_LD;
          const float4 mn = __ldg(reinterpret_cast<const float4*>(k.sp_mean + sbase[i] + ch));
          const float4 rs = __ldg(reinterpret_cast<const float4*>(k.sp_rstd + sbase[i] + ch));
          const float4 g = *reinterpret_cast<const float4*>(trp + cc);
          const float4 bt = *reinterpret_cast<const float4*>(trp + 16 + cc);
          const float4 xn = make_float4((xv[i].x - mn.x) * rs.x, (xv[i].y - mn.y) * rs.y, (xv[i].z - mn.z) * rs.z, (xv[i].w - mn.w) * rs.w);
          float v0 = xn.x * (1.f + (g.x + bg.x)) + (bt.x + bb.x);
          float v1 = xn.y * (1.f + (g.y + bg.y)) + (bt.y + bb.y);
          float v2 = xn.z * (1.f + (g.z + bg.z)) + (bt.z + bb.z);
          float v3 = xn.w * (1.f + (g.w + bg.w)) + (bt.w + bb.w);
          v0 = apply_leaky(v0, easlope); v1 = apply_leaky(v1, easlope);
          v2 = apply_leaky(v2, easlope); v3 = apply_leaky(v3, easlope);
          if (k.emit) {
            uint2 hv, lv;
            split_operand4(v0 * k.emul, v1 * k.emul, v2 * k.emul, v3 * k.emul, hv, lv);
            __nv_bfloat16* ep = k.emit + epix[i] * k.erow + (ch >> 5) * 64 + (ch & 31);
            *reinterpret_cast<uint2*>(ep) = hv;
            *reinterpret_cast<uint2*>(ep + 32) = lv;
          } else {                                           // fp32 modulated activation [.., sp_C] (input of a Winograd conv)
            *reinterpret_cast<float4*>(k.y + yoff[i] + ch) = make_float4(v0, v1, v2, v3);
          }
        }
        __syncwarp();
        continue;
      }
      if constexpr (!RES && !EMIT) {
        // plain GEMM output (the Winograd GEMMs, whose tile period is set by these warps: ncu showed them 91 % busy with ~70
        // instructions per row quad in the generic path below): 32 full columns -> one LDS.128 + one STG.128 per row
        if (k.plain && c0 + 32 <= k.BN && n0 + c0 + 32 <= k.Cout) {     // warp-uniform
          const int ncp = n0 + c0 + c4 - chan0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (!((vmask >> i) & 1u)) continue;
            *reinterpret_cast<float4*>(k.y + yoff[i] + ncp) = *reinterpret_cast<const float4*>(tile + (sub + 4 * i) * STG_LD + c4);
          }
          __syncwarp();
          continue;
        }
      }
      const int n = n0 + c0 + c4;
      if (c0 + c4 < k.BN && n < k.Cout) {
        float bz[4] = {0.f, 0.f, 0.f, 0.f};
        if (k.bias) {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (n + j < k.Cout) bz[j] = __ldg(k.bias + n + j);
        }
        const int nc = n - chan0;                           // output channel (phase mode: every N tile is channels 0 .. BN-1)
        const bool full4 = k.vec4 && (n + 3 < k.Cout);
        float4 rr4[RES ? 8 : 1];
        if constexpr (RES) {
          if (full4) {                                      // all residual loads in flight before the first use
#pragma unroll
            for (int i = 0; i < 8; ++i)
              rr4[i] = ((vmask >> i) & 1u) ? *reinterpret_cast<const float4*>(k.res + roff[i] + nc) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        float es[4] = {1.f, 1.f, 1.f, 1.f}, eb[4] = {0.f, 0.f, 0.f, 0.f};
        if (EMIT && k.escale) {                             // emission needs Cout % 32 == 0: n .. n+3 are valid
          const float4 s4 = __ldg(reinterpret_cast<const float4*>(k.escale + nc));
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(k.eshift + nc));
          es[0] = s4.x; es[1] = s4.y; es[2] = s4.z; es[3] = s4.w;
          eb[0] = b4.x; eb[1] = b4.y; eb[2] = b4.z; eb[3] = b4.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (!((vmask >> i) & 1u)) continue;
          const float4 a = *reinterpret_cast<const float4*>(tile + (sub + 4 * i) * STG_LD + c4);
          float o[4] = {a.x + bz[0], a.y + bz[1], a.z + bz[2], a.w + bz[3]};
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = apply_leaky(o[j], aslope);
          if (act_rare) {                                    // warp-uniform; one out-of-line copy of the expf / erff code
            const float4 r = act_rare4(make_float4(a.x + bz[0], a.y + bz[1], a.z + bz[2], a.w + bz[3]), k.act);
            o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
          }
          if constexpr (RES) {
            if (full4) {
              o[0] += rr4[i].x; o[1] += rr4[i].y; o[2] += rr4[i].z; o[3] += rr4[i].w;
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) if (n + j < k.Cout) o[j] += k.res[roff[i] + nc + j];
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] *= mu[i];
          if (k.y) {
            float* yp = k.y + yoff[i] + nc;
            if (full4) {
              *reinterpret_cast<float4*>(yp) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) if (n + j < k.Cout) yp[j] = o[j];
            }
          }
          if constexpr (EMIT) {                             // the next conv's split-fp16 operand, transform fused
            float e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] = apply_leaky(fmaf(o[j], es[j], eb[j]), easlope) * k.emul;
            uint2 hv, lv;
            split_operand4(e[0], e[1], e[2], e[3], hv, lv);
            __nv_bfloat16* ep = k.emit + epix[i] * k.erow + (nc >> 5) * 64 + (nc & 31);
            *reinterpret_cast<uint2*>(ep) = hv;
            *reinterpret_cast<uint2*>(ep + 32) = lv;
          }
        }
      }
      __syncwarp();
    }
    // this warp's TMEM reads of the tile are complete (every tcgen05.ld above was waited for): hand the accumulator
    // buffer back to the MMA issuer -- the leader CTA's, which also writes the peer's TMEM in pair mode
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if constexpr (CTAS == 2)
        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"((tmem_empty + 8u * buf) & PEER_BIT_MASK) : "memory");
      else
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty + 8u * buf) : "memory");
    }
    }  // tile loop
  }

  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();   // the peer may still read this CTA's smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CTAS == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// weight packing: w32 [tap][Cin][Cout] fp32 -> [Cout_p][tap][nblk][hi 32 | lo 32] bf16
// ------------------------------------------------------------------------------------------
struct PackPlan {                   // issue order of the kernel that will consume the packed rows (see TcPlan, tc_ptx.cuh)
  int nsets, chunk, npass, last_ksteps;
  float kappa;                      // pre-compensation per truncation event (0 = none)
  int ph_s, rows_per_phase;         // phase mode: rows [p * rows_per_phase, ..) belong to phase p, which walks the taps in tapmask[p]
  unsigned tapmask[16];
};

__global__ void __launch_bounds__(256) pack_tc_kernel(const float* __restrict__ w32, __nv_bfloat16* __restrict__ out, int taps,
                                                      int Cin, int Cout, int Cout_p, int nblk, float wmul, PackPlan pp) {
  const long rowlen = (long)taps * nblk * 64;
  const long total = (long)Cout_p * taps * nblk * 32;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int j = (int)(i & 31); long r = i >> 5;
    int blk = (int)(r % nblk); r /= nblk;
    int tap = (int)(r % taps); int co = (int)(r / taps);
    int ci = blk * 32 + j;
    float v = (co < Cout && ci < Cin) ? w32[((long)tap * Cin + ci) * Cout + co] * wmul : 0.f;
    if (pp.kappa != 0.f) {
      int it = tap * nblk + blk, niter = taps * nblk;
      if (pp.ph_s) {                                          // ordinal of the tap among the taps this phase walks
        const unsigned m = pp.tapmask[co / pp.rows_per_phase];
        it = __popc(m & ((1u << tap) - 1u)) * nblk + blk;
        niter = __popc(m) * nblk;
      }
      const int rem = tc_remaining_events(pp.nsets, pp.chunk, pp.npass, niter, nblk, pp.last_ksteps, it, j >> 4);
      v *= 1.0f + pp.kappa * (float)rem;
    }
    __nv_bfloat16 hi, lo;
    split_operand(v, hi, lo);
    long o = (long)co * rowlen + ((long)tap * nblk + blk) * 64 + j;
    out[o] = hi;
    out[o + 32] = lo;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

int pick_bn(int Cout) {
  int c16 = round_up(Cout, 16);
  if (c16 <= 256) return c16;
  // split into equal tiles of <= 256 columns
  int nt = (c16 + 255) / 256;
  return round_up((c16 + nt - 1) / nt, 16);
}

bool g_attr_set[64] = {};
constexpr int MAX_DYN_SMEM = 220 * 1024;

// phase-mode tap masks: output row Y = y*f + a reads input rows floor((Y + dy) / f): a = 0 -> {y-1: dy=-1, y: dy=0,+1};
// a = f-1 -> {y: dy=-1,0, y+1: dy=+1}; else only row y.  The packed weights hold the per-phase tap sums; taps outside the
// mask are zero and skipped.
// (a kernel with KD = 3 -- the hourglass decoder convs, whose input is upsampled in (h, w) only -- walks the same in-plane
//  taps at every depth tap: the 9-bit mask is replicated per kd)
void phase_tapmasks(int ps, int KD, unsigned* out) {
  const int f = 1 << ps;
  for (int a = 0; a < f; ++a)
    for (int b = 0; b < f; ++b) {
      unsigned rows = 2u | (a == 0 ? 1u : 0u) | (a == f - 1 ? 4u : 0u), cols = 2u | (b == 0 ? 1u : 0u) | (b == f - 1 ? 4u : 0u);
      unsigned m = 0;
      for (int ty = 0; ty < 3; ++ty)
        for (int tx = 0; tx < 3; ++tx) if (((rows >> ty) & 1u) && ((cols >> tx) & 1u)) m |= 1u << (ty * 3 + tx);
      unsigned all = 0;
      for (int kd = 0; kd < KD; ++kd) all |= m << (9 * kd);
      out[a * f + b] = all;
    }
}

}  // namespace

namespace tc {
// Accumulator sets: chains of <= ~256 MMAs per TMEM accumulator.  A short K (whole chain <= single_chain MMAs) runs in ONE
// accumulator; otherwise one set for the correction products + hi*hi sets of <= ~256 MMAs.  If two such buffers fit into the
// TMEM columns the accumulators are double-buffered: the epilogue of tile i overlaps the MMAs of tile i+1.  Thin N tiles
// (BN <= 64) run two CTAs per SM with half of the TMEM each.
// niter: K iterations a tile walks (phase form: the ACTIVE taps only); single_main: one hi*hi set whatever the chain length
// (phase form with a non-uniform number of taps per phase)
TcPlan tc_make_plan(int BN, int niter, int npass, int single_chain, int max_sets, bool double_buffer, int single_main) {
  TcPlan p;
  p.npass = npass;
  const int steps_main = niter * 2;
  // thin N tiles (short MMAs: the per-stage issue overhead dominates) run as TWO co-resident CTAs per SM with half of the
  // TMEM each -- unless the chain is so long that it needs the whole TMEM for accumulator sets
  const bool thin = BN <= 64 && (steps_main + 2) / 3 <= 256;
  p.thin = thin;
  int want = (steps_main * npass <= single_chain) ? 1 : 1 + (steps_main + 255) / 256;
  if (npass == 1 && want > 1) want -= 1;
  if (max_sets > 0 && want > max_sets) want = max_sets;
  const int tmem_budget = thin ? 256 : 512;
  while (want > 1 && BN * want > tmem_budget) --want;
  p.nacc = (double_buffer && 2 * BN * want <= tmem_budget) ? 2 : 1;
  const int per_buf = tmem_budget / p.nacc;
  int nsets = want == 1 ? 1 : per_buf / BN;               // spare columns shorten the chains further
  if (nsets > 16) nsets = 16;
  if (max_sets > 0 && nsets > max_sets) nsets = max_sets;
  if (nsets < 1) nsets = 1;
  const int corr = (npass > 1 && nsets > 1) ? 1 : 0;
  int nmain = nsets - corr;
  if (nmain > niter) nmain = niter;
  const int chunk = (niter + nmain - 1) / nmain;          // K iterations per hi*hi set
  nmain = (niter + chunk - 1) / chunk;                    // sets actually written
  if (single_main) nmain = 1;
  p.nsets = corr + nmain;
  p.chunk = single_main ? (1 << 30) : chunk;
  return p;
}
}  // namespace tc

namespace tc {
PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  if (!fn) throw Error(CS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  return fn;
}

// power-of-two box extent p <= cap covering `dim` with the fewest padded elements (ties: larger p)
int pick_box(int dim, int cap, int* log2out) {
  int best = 1, bestl = 0;
  long best_pad = dim;
  for (int l = 1, p = 2; p <= cap; ++l, p <<= 1) {
    long pad = (long)((dim + p - 1) / p) * p;
    if (pad <= best_pad) { best = p; bestl = l; best_pad = pad; }
    if (p >= dim) break;
  }
  *log2out = bestl;
  return best;
}

}  // namespace tc

bool conv_tc_supported(const ConvW& w, const Act& out) {
  (void)out;
  return w.wtc != nullptr;
}

// (Cin < 16, e.g. the RGB input of the first conv, is zero-padded to one 16-channel K step)
bool conv_tc_shape_ok(int Cin, int Cout) { return Cin >= 1 && Cout >= 1; }

Opd conv_tc_alloc_operand(Arena& A, const ConvW& w, const Act& out) {
  Opd o;
  o.B = out.B; o.D = out.D; o.H = out.H; o.W = out.W;
  o.nblk = (w.Cin + 31) / 32;
  o.amul = w.amul;
  o.p = A.bf16((size_t)out.B * out.D * out.H * out.W * o.nblk * 64);
  return o;
}

void pack_tc(cs_ctx* ctx, ConvW& w, cudaStream_t stream) {
  if (!conv_tc_shape_ok(w.Cin, w.Cout) || !w.w32) return;
  const int nblk = (w.Cin + 31) / 32;
  int BN = pick_bn(w.Cout);
  // Accumulator-set policy.  The hi*hi products of a tile accumulate in 512 / BN - 1 TMEM sets (one more holds the small
  // correction products); every MMA added into a set truncates it (round toward zero), so the error grows with the chain
  // length per set (the first-order loss is pre-compensated in the packed weights; what remains grows with the chain too:
  // tests/test_gpu_configs.py).  Halve the N tile (down to tc_bn_min = 128) until a chain is at most `chain_max` MMAs:
  // BN = 256 leaves one hi*hi set (a 3x3 conv over 512 channels would chain 288 MMAs), BN = 128 three (96 each).  (BN = 64
  // -- seven sets -- was measured for the deep hourglass levels, K up to 27 x 1024: their chains drop from 576 to 247, but
  // the narrow tile re-reads the A operand twice as often: +1.8 ms per step for no visible end-to-end gain.)
  {
    const int chain_max = ctx->tc_chain_max > 0 ? ctx->tc_chain_max : 256;
    const int steps_main = w.taps() * nblk * 2;
    while (BN > ctx->tc_bn_min && (steps_main + (512 / BN - 1) - 1) / (512 / BN - 1) > chain_max) BN = round_up(BN / 2, 16);
  }
  if (ctx->tc_bn_max > 0 && BN > ctx->tc_bn_max) BN = ctx->tc_bn_max;
  int rows_per_phase = 0;
  if (w.phase_shift > 0) {                                  // N tiles never straddle output phases: Cout = 4^ps * (tiles per phase) * BN rows
    rows_per_phase = w.Cout >> (2 * w.phase_shift);
    BN = rows_per_phase;
    while (BN > 256) BN /= 2;
    CS_REQUIRE(BN % 16 == 0 && rows_per_phase % BN == 0 && w.KH == 3 && w.KW == 3 && (w.KD == 1 || (w.KD == 3 && w.phase_shift == 1)), CS_ERR_WEIGHTS,
               "pack_tc: bad phase-form conv");
  }
  const int Cout_p = round_up(w.Cout, BN);
  const size_t n = (size_t)Cout_p * w.taps() * nblk * 64;
  if (!w.wtc) w.wtc = static_cast<__nv_bfloat16*>(ctx->dmalloc(n * sizeof(__nv_bfloat16)));
  w.nblk = nblk; w.BN = BN; w.Cout_p = Cout_p;
  // the accumulator plan is fixed here: the packed rows carry the truncation pre-compensation of this issue order
  const int npass = ctx->tc_passes >= 1 && ctx->tc_passes <= 3 ? ctx->tc_passes : 3;
  // phase form, x2: every phase walks 2 x 2 of the 3 x 3 in-plane taps (uniform: ordinary accumulator sets over the active
  // iterations); x4: 4, 6 or 9 taps depending on the phase -> one hi*hi set
  const int niter_plan = w.phase_shift == 1 ? w.KD * 4 * nblk : w.taps() * nblk;
  const TcPlan plan = tc_make_plan(BN, niter_plan, npass, ctx->tc_single_chain, ctx->tc_sets, ctx->tc_dbuf != 0, w.phase_shift >= 2 ? 1 : 0);
  w.plan_nsets = plan.nsets; w.plan_chunk = plan.chunk; w.plan_nacc = plan.nacc; w.plan_npass = npass; w.plan_thin = plan.thin;
  w.plan_kappa = (float)ctx->tc_poscomp * 1e-10f;
  PackPlan pp{};
  pp.nsets = plan.nsets; pp.chunk = plan.chunk; pp.npass = npass; pp.kappa = w.plan_kappa;
  pp.last_ksteps = ((w.Cin - (nblk - 1) * 32) + 15) / 16;
  pp.ph_s = w.phase_shift; pp.rows_per_phase = rows_per_phase;
  if (w.phase_shift) phase_tapmasks(w.phase_shift, w.KD, pp.tapmask);
  long total = (long)Cout_p * w.taps() * nblk * 32;
  long blocks = (total + 255) / 256; if (blocks > 148L * 16) blocks = 148L * 16;
  pack_tc_kernel<<<(unsigned)blocks, 256, 0, stream>>>(w.w32, w.wtc, w.taps(), w.Cin, w.Cout, Cout_p, nblk, w.wmul, pp);
  check_launch("pack_tc");
}

void conv_tc(const Launcher& L, const Opd& x, const ConvW& w, const ConvGeom& g, const Epilogue& e, Act y) {
  L.count();
  if (L.dry) return;
  CS_REQUIRE(w.wtc != nullptr, CS_ERR_WEIGHTS, "conv_tc: tcgen05 operand not packed");
  // stride-1 'same' convolutions, or a kernel spanning the full depth without depth padding (Do = 1)
  const bool same_d = g.Do == x.D && g.PD == w.KD / 2;
  const bool full_d = g.Do == 1 && w.KD == x.D && g.PD == 0;
  const int ps = e.phase_shift;
  CS_REQUIRE((same_d || full_d) && g.Ho == (x.H << ps) && g.Wo == (x.W << ps) && g.PH == w.KH / 2 && g.PW == w.KW / 2,
             CS_ERR_INVALID, "conv_tc: unsupported geometry");
  if (ps) {
    CS_REQUIRE(ps <= 2 && (w.KD == 1 || (w.KD == 3 && ps == 1)) && w.KH == 3 && w.KW == 3 && y.C == (w.Cout >> (2 * ps)) && y.C % w.BN == 0 &&
                   w.zrows == 0 && !e.residual && !e.mult && !e.sp_x, CS_ERR_INVALID, "conv_tc: bad phase-mode conv");
  } else {
    CS_REQUIRE(y.C == w.Cout || (e.sp_x && !e.emit && y.C == e.sp_C), CS_ERR_INVALID, "conv_tc: channel mismatch");
  }
  CS_REQUIRE(x.nblk == w.nblk, CS_ERR_INVALID, "conv_tc: channel mismatch");

  operand_absmax(L, x, w.id);
  ConvTcK k{};
  k.emul = e.emit_mul;
  k.B = x.B; k.D = g.Do; k.H = x.H; k.W = x.W;
  int cap = 128;
  int bw = pick_box(x.W, cap, &k.lbw); cap /= bw;
  int bh = pick_box(x.H, cap, &k.lbh); cap /= bh;
  int bd = 1; k.lbd = 0;                                   // depth-dependent weights: one depth slice per tile
  if (w.zrows == 0) { bd = pick_box(g.Do, cap, &k.lbd); cap /= bd; }
  int bb = cap; k.lbb = 0; while ((1 << k.lbb) < bb) ++k.lbb;
  k.ntw = (x.W + bw - 1) / bw; k.nth = (x.H + bh - 1) / bh; k.ntd = (g.Do + bd - 1) / bd;
  const int ntb = (x.B + bb - 1) / bb;
  k.KD = w.KD; k.KH = w.KH; k.KW = w.KW; k.PD = g.PD; k.PH = g.PH; k.PW = g.PW;
  k.nblk = w.nblk;
  k.last_ksteps = ((w.Cin - (w.nblk - 1) * 32) + 15) / 16;
  k.rowA = w.nblk * 64;
  k.BN = w.BN; k.Cout = w.Cout; k.zrows = w.zrows;
  k.ph_s = ps;
  k.ph_tpp = ps ? y.C / w.BN : 1;
  if (ps) phase_tapmasks(ps, w.KD, k.tapmask);
  k.npass = w.plan_nsets > 0 ? w.plan_npass : (L.npass >= 1 && L.npass <= 3 ? L.npass : 3);
  k.bias = w.bias; k.act = e.act; k.slope = e.slope;
  k.res = e.residual; k.rb = e.rs_b; k.rd = e.rs_d; k.rh = e.rs_h; k.rw = e.rs_w;
  k.mult = e.mult;
  k.y = y.p; k.yb = y.sb; k.yd = y.sd; k.yh = y.sh; k.yw = y.sw;
  if (e.sp_x) {
    // output: the split operand of the consumer conv (e.emit, y.p null) or the fp32 activation itself (y = [.., sp_C])
    CS_REQUIRE(!e.residual && !e.mult && w.Cout == 2 * e.sp_C && e.sp_C % 32 == 0 && g.Do == 1 &&
                   ((e.emit && y.p == nullptr && e.emit_nblk * 32 >= e.sp_C) || (!e.emit && y.p != nullptr && y.C == e.sp_C)),
               CS_ERR_INVALID, "conv_tc: bad SPADE epilogue");
    k.emit = e.emit; k.erow = e.emit_nblk * 64; k.eact = e.emit_act; k.eslope = e.emit_slope;
    k.sp_x = e.sp_x; k.sp_mean = e.sp_mean; k.sp_rstd = e.sp_rstd; k.sp_C = e.sp_C; k.sp_xs = e.sp_xshift;
    k.sp_Hx = e.sp_Hx; k.sp_Wx = e.sp_Wx;
  } else if (e.emit) {
    CS_REQUIRE(y.C % 32 == 0 && e.emit_nblk * 32 >= y.C, CS_ERR_INVALID, "conv_tc: operand emission needs Cout % 32 == 0");
    k.emit = e.emit; k.erow = e.emit_nblk * 64; k.escale = e.emit_scale; k.eshift = e.emit_shift; k.eact = e.emit_act;
    k.eslope = e.emit_slope;
  }
  CS_REQUIRE(act_is_leaky(e.emit_act), CS_ERR_INVALID, "conv_tc: the fused operand activation must be none / relu / leaky relu");
  auto al4 = [](long v) { return (v & 3) == 0; };
  k.vec4 = al4(y.sb) && al4(y.sd) && al4(y.sh) && al4(y.sw) && ((uintptr_t)y.p % 16 == 0) &&
           (!e.residual || (al4(e.rs_b) && al4(e.rs_d) && al4(e.rs_h) && al4(e.rs_w) && ((uintptr_t)e.residual % 16 == 0)));

  k.plain = k.vec4 && y.p && !w.bias && e.act == ACT_NONE && !e.residual && !e.mult && !e.emit && !e.sp_x;

  const unsigned m_tiles = (unsigned)(k.ntw * k.nth * k.ntd * ntb);
  // tcgen05 pair mode (cta_group::2): wide N tiles, an even number of M tiles, weights shared by the pair.  The kernel is
  // persistent, so the cluster set-up is paid once per launch and short-K convs profit too (halved B traffic per CTA).
  const int niter = w.taps() * w.nblk;
  // (depth-dependent weights: the two M tiles of a pair must lie in the same depth slice, i.e. an even number of tiles per slice)
  const bool pair = L.pair && k.BN >= 128 && (m_tiles % 2 == 0) && (w.zrows == 0 || (k.ntw * k.nth) % 2 == 0) && k.npass == 3 &&
                    niter >= L.pair_min_iter && ps == 0;
  const int stage_bytes = A_TILE_BYTES + (pair ? k.BN / 2 : k.BN) * 128;
  // thin N tiles (short MMAs: the per-stage issue overhead dominates) run as TWO co-resident CTAs per SM, each with half of the
  // shared memory and of the TMEM columns
  const bool thin = w.plan_nsets > 0 ? w.plan_thin : (k.BN <= 64 && (niter * 2 + 2) / 3 <= 256);
  // accumulator plan: the one the weights were packed for (pack_tc); hand-packed weights (no plan) get the default policy
  {
    TcPlan plan;
    if (w.plan_nsets > 0) { plan.nsets = w.plan_nsets; plan.chunk = w.plan_chunk; plan.nacc = w.plan_nacc; plan.npass = w.plan_npass; plan.thin = w.plan_thin; }
    else plan = tc_make_plan(k.BN, niter, k.npass, L.single_chain, L.max_sets, L.double_buffer, ps ? 1 : 0);
    CS_REQUIRE(!ps || w.phase_shift == ps, CS_ERR_INVALID, "conv_tc: phase-mode launch of a conv not packed in phase form");
    k.nsets = plan.nsets; k.chunk = plan.chunk; k.nacc = plan.nacc;
    const int corr = (k.npass > 1 && k.nsets > 1) ? 1 : 0;
    int tcols = 32;
    while (tcols < k.nacc * k.nsets * k.BN) tcols <<= 1;
    k.tcols = tcols;
    // The tensor core truncates (rounds toward zero) when it adds an MMA into the fp32 accumulator.  Weights packed with the
    // position-dependent pre-compensation (ConvW::plan_kappa, tc_ptx.cuh) need no epilogue factor; otherwise the epilogue
    // scales the hi*hi sum by 1 + c * chain (the mean loss of a chain, CS_OPT_TC_COMP).
    const int per_set = k.chunk < niter ? k.chunk : niter;
    const int chain = per_set * 2 * (corr ? 1 : k.npass);
    k.acc_scale = w.plan_kappa != 0.f ? 1.0f : 1.0f + L.acc_comp * 1e-10f * (float)chain;
    k.kappa = w.plan_kappa; k.Din = x.D;
    k.out_scale = 1.0f / (w.wmul * x.amul);
  }
  const int egroups = k.BN > 64 ? 2 : 1;                   // 8 epilogue warps on wide tiles
  CS_REQUIRE(!k.sp_x || (k.BN + 32 * egroups - 1) / (32 * egroups) <= 4, CS_ERR_INVALID, "conv_tc: SPADE epilogue holds <= 4 chunks per warp");
  const int stg_bytes = egroups * STG_BYTES;
  int stages = ((thin ? 100 * 1024 : MAX_DYN_SMEM) - 2048 - stg_bytes) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 1) stages = 1;
  k.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + stg_bytes + 1024 + 16 * stages + 64;
  // persistent grid: one CTA (or pair) per SM walks the tiles round-robin, N tile-major so concurrent CTAs share the weights
  const int ctas = pair ? 2 : 1;
  k.m_units = (int)m_tiles / ctas;
  k.n_tiles = (w.zrows > 0 ? w.zrows : w.Cout_p) / k.BN;
  const int units = k.m_units * k.n_tiles;
  static int n_sm = 0;
  if (!n_sm) { int dev0 = 0; CS_CUDA(cudaGetDevice(&dev0)); CS_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev0)); }

  // tensor maps
  auto enc = encode_fn();
  CUtensorMap tmA, tmB;
  {
    const cuuint64_t pix = (cuuint64_t)x.row() * 2;          // a block range of a wider operand keeps the wider pixel stride
    cuuint64_t dims[5] = {(cuuint64_t)k.rowA, (cuuint64_t)x.W, (cuuint64_t)x.H, (cuuint64_t)x.D, (cuuint64_t)x.B};
    cuuint64_t strides[4] = {pix, pix * x.W, pix * x.W * x.H, pix * x.W * x.H * x.D};
    cuuint32_t box[5] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, (cuuint32_t)bb};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x.p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CS_REQUIRE(r == CUDA_SUCCESS, CS_ERR_CUDA, "conv_tc: cuTensorMapEncodeTiled(A) failed");
  }
  {
    const cuuint64_t rowlen = (cuuint64_t)w.taps() * k.rowA;
    cuuint64_t dims[2] = {rowlen, (cuuint64_t)w.Cout_p};
    cuuint64_t strides[1] = {rowlen * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(pair ? k.BN / 2 : k.BN)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w.wtc, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CS_REQUIRE(r == CUDA_SUCCESS, CS_ERR_CUDA, "conv_tc: cuTensorMapEncodeTiled(B) failed");
  }
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, ConvTcK);
  // [residual][emit][pair][border take-back]; the SPADE epilogue has its own four
  static const KernelFn fns[2][2][2][2] = {
      {{{conv_tc_kernel<false, false, 1, false, false>, conv_tc_kernel<false, false, 1, false, true>},
        {conv_tc_kernel<false, false, 2, false, false>, conv_tc_kernel<false, false, 2, false, true>}},
       {{conv_tc_kernel<false, true, 1, false, false>, conv_tc_kernel<false, true, 1, false, true>},
        {conv_tc_kernel<false, true, 2, false, false>, conv_tc_kernel<false, true, 2, false, true>}}},
      {{{conv_tc_kernel<true, false, 1, false, false>, conv_tc_kernel<true, false, 1, false, true>},
        {conv_tc_kernel<true, false, 2, false, false>, conv_tc_kernel<true, false, 2, false, true>}},
       {{conv_tc_kernel<true, true, 1, false, false>, conv_tc_kernel<true, true, 1, false, true>},
        {conv_tc_kernel<true, true, 2, false, false>, conv_tc_kernel<true, true, 2, false, true>}}}};
  // (no border take-back in the SPADE epilogue: its K is short -- at most kappa * 0.56 * 108 = 2e-6 on corner pixels -- and the
  //  epilogue, which already holds the normalised activation in registers, paces these kernels: +9 % measured with it)
  static const KernelFn fns_spade[2] = {conv_tc_kernel<false, true, 1, true, false>, conv_tc_kernel<false, true, 2, true, false>};
  if (!g_attr_set[dev & 63]) {
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        for (int c = 0; c < 2; ++c)
          for (int d = 0; d < 2; ++d)
            CS_CUDA(cudaFuncSetAttribute(fns[a][b][c][d], cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM + 4096));
    for (int c = 0; c < 2; ++c)
      CS_CUDA(cudaFuncSetAttribute(fns_spade[c], cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM + 4096));
    g_attr_set[dev & 63] = true;
  }
  if (w.zrows > 0) CS_REQUIRE(bd == 1 && w.zrows % k.BN == 0, CS_ERR_INVALID, "conv_tc: depth-dependent weights need one depth per tile");
  const long M = (long)x.B * g.Do * g.Ho * g.Wo;          // phase mode: the algorithmic conv runs on the upsampled grid
  char desc[120];
  snprintf(desc, sizeof(desc), "tc M=%ld Cin=%d Cout=%d k=%dx%dx%d BN=%d st=%d sets=%d acc=%d %s%s%s%s%s tiles=%dx%d", M, w.Cin, y.C, w.KD,
           w.KH, w.KW, k.BN, stages, k.nsets, k.nacc, pair ? "pair " : "", k.res ? "res " : "", k.emit ? "emit " : "", k.sp_x ? "spade " : "",
           ps ? "phase " : "", (int)m_tiles, k.n_tiles);
  ProfScope pscope(L, PK_CONV_TC, e.alg_flops > 0.0 ? e.alg_flops : 2.0 * (double)M * (e.sp_x ? w.Cout : y.C) * w.Cin * w.taps(), 0.0, desc);
  const bool has_res = k.res != nullptr, has_emit = k.emit != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(64 + 128 * egroups); cfg.dynamicSmemBytes = smem; cfg.stream = L.stream;
  cudaLaunchAttribute attr[1];
  if (pair) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  const int bcomp = (k.kappa != 0.f && w.taps() > 1 && ps == 0) ? 1 : 0;
  KernelFn fn = fns[has_res][has_emit][pair ? 1 : 0][bcomp];
  if (k.sp_x) fn = fns_spade[pair ? 1 : 0];
  int clusters = thin ? 2 * n_sm : n_sm;
  if (pair) {
    // co-resident 2-CTA clusters (GPC boundaries can strand an SM): a persistent grid must not exceed one wave
    static std::map<std::pair<const void*, size_t>, int> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_pair(reinterpret_cast<const void*>(fn), smem);
    auto itc = cache.find(key);
    if (itc == cache.end()) {
      cfg.gridDim = dim3((unsigned)n_sm / 2 * 2);
      int nc = 0;
      CS_CUDA(cudaOccupancyMaxActiveClusters(&nc, fn, &cfg));
      if (nc < 1) nc = 1;
      if (nc > n_sm / 2) nc = n_sm / 2;
      itc = cache.emplace(key, nc).first;
    }
    clusters = itc->second;
  }
  if (clusters > units) clusters = units;
  cfg.gridDim = dim3((unsigned)(clusters * ctas));
  CS_CUDA(cudaLaunchKernelEx(&cfg, fn, tmA, tmB, k));
  check_launch("conv_tc");
}

}  // namespace cs
