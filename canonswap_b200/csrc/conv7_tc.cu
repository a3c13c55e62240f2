// The 7x7x7 mask convolution of DenseMotionNetwork (reference dense_motion.py:18,88: Conv3d(142 -> 22, k=7, p=3)
// on the [B,142,16,h,w] hourglass output) -- the largest single op of the path (140.5 GFLOP per pass) and the
// worst shaped one for an implicit GEMM (N = 22).  Dedicated tcgen05 kernel, "depth-stacked, kh-split":
//
//   * one CTA = 128 (h,w) pixels x 8 output depths x 22 channels, for ONE filter row kh.
//     For an input slice z and a filter tap (kh,kw) the A tile (128 pixels x 32 channels of slice z, shifted by
//     the tap; TMA zero-fills the h/w padding) contributes to every output depth d = z-3 .. z+3 at once:
//     the B tile stacks the 7 depth taps kd = z-d+3 along N (7 x 24 = 168 columns, clipped to the CTA's 8 depths),
//     so N is 32..176 instead of 22, every A tile is loaded once for 7 depth taps, and TMEM holds the
//     8 x 24 accumulator columns of the CTA (+ the same again for the split-fp16 correction products).
//     A depth slot is 24 columns (22 channels + 2 pad), not 32: 25 % fewer MMA columns.  UMMA needs N % 16 == 0, so
//     an odd number of slots is rounded up by 8 columns; those 8 extra B rows / D columns are harmless by
//     construction: the used slots end either at the last slot of the filter position (followed by 8 zero rows
//     in the packed weights) or at the last depth of the CTA's group (the extra columns are scratch TMEM).
//   * the 7 filter rows kh go to 7 different CTAs that write 7 partial logit tensors, summed (+ bias) by a small
//     fp32 kernel.  That keeps every TMEM accumulation chain at 7 z x 7 kw x 9 K-steps = 441 MMAs (the tensor
//     core accumulates with truncation: error grows with the chain length, see conv_tc.cu) and gives the
//     grid 7x more CTAs (3584 at B = 8) for 148 SMs.
//
// Operand format, pipeline roles and MMA issue are those of conv_tc.cu.
#include "tc_ptx.cuh"

namespace cs {

using namespace tc;

namespace {

constexpr int C7_COUT_P = 24;                    // 22 -> 24 columns per output depth
constexpr int C7_GROUP = 8;                      // output depths per CTA
constexpr int C7_BROWS = 7 * C7_COUT_P + 8;      // 176 B rows per filter position (kh,kw): 7 depth slots + 8 zero rows
// CTAS = 1: 5 stages of A (16 KB) + B (22 KB).  CTAS = 2 (tcgen05 pair, two pixel tiles share the weights): each CTA holds
// half of the B rows, 8 stages of 16 + 11 KB.  The kernel is paced by the bytes it keeps in flight (a slot is refilled only
// after its MMAs have completed), so in pair mode the epilogue's staging tile does not get its own
// shared memory: it reuses stage 0, which is idle once the last MMA has completed (tmem_full), and the 18 KB buy an 8th stage.
template <int CTAS> struct C7Cfg {
  static constexpr int BROWS = C7_BROWS / CTAS;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + BROWS * 128;
  static constexpr int STAGES = CTAS == 2 ? 8 : 5;
  static constexpr bool STG_IN_STAGE0 = CTAS == 2;
  static constexpr int SMEM = STAGES * STAGE_BYTES + (STG_IN_STAGE0 ? 0 : STG_BYTES) + 1024 + 16 * STAGES + 32;
  static_assert(STG_BYTES <= STAGE_BYTES, "the staging tile must fit into one stage");
};

struct Conv7K {
  int B, H, W;                     // D = 16
  int lbw, lbh, ntw, nth;
  int nblk, last_ksteps, Cout;
  float* parts;                    // [7][B,16,H,W,ldo]
  long part_stride;                // elements between partial tensors
  int ldo;                         // channel stride of a partial (24)
  float acc_scale;                 // constant round-toward-zero compensation of the hi*hi chain (weights without pre-compensation)
  float out_scale;                 // 1 / ConvW::wmul
  float kappa;                     // > 0: the packed weights carry the position-dependent pre-compensation (tc_ptx.cuh); the epilogue
  int ev_slice;                    //      takes back what was assumed for taps in the zero padding. ev_slice = hi*hi events per input slice
};

template <int CTAS>
__global__ void __launch_bounds__(TC_THREADS) conv7_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB, Conv7K k) {
  constexpr int C7_STAGES = C7Cfg<CTAS>::STAGES;
  constexpr int C7_STAGE_BYTES = C7Cfg<CTAS>::STAGE_BYTES;
  uint32_t cta_rank = 0;
  if constexpr (CTAS == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t stg = C7Cfg<CTAS>::STG_IN_STAGE0 ? base : base + (uint32_t)C7_STAGES * C7_STAGE_BYTES;
  const uint32_t bars = base + (uint32_t)C7_STAGES * C7_STAGE_BYTES + (C7Cfg<CTAS>::STG_IN_STAGE0 ? 0u : (uint32_t)STG_BYTES);
  const uint32_t tmem_full = bars + 16u * C7_STAGES;
  const uint32_t tmem_slot = tmem_full + 8u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int t = blockIdx.x;
  const int tw = t % k.ntw; t /= k.ntw;
  const int th = t % k.nth; const int b = t / k.nth;
  const int w0 = tw << k.lbw, h0 = th << k.lbh;
  const int g = blockIdx.y;                      // output depths [8g, 8g+8)
  const int kh = blockIdx.z;
  const int zlo = max(0, C7_GROUP * g - 3), zhi = min(15, C7_GROUP * g + C7_GROUP - 1 + 3);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < C7_STAGES; ++s) { mbar_init(bars + 8u * s, 1); mbar_init(bars + 8u * (C7_STAGES + s), 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // pair mode: the two-CTA allocation protocol writes to the peer's shared memory -- the peer must be running (conv_tc.cu)
  if constexpr (CTAS == 2) cluster_sync_all();
  if (warp == 1) {
    if constexpr (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  // zero the accumulators: every MMA below accumulates (a slice touches a sliding window of depth columns)
  if (warp >= 2) {
    const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int c = 0; c < 512; c += 16) tc_st16_zero(trow + (uint32_t)c);
    tc_st_wait();
  }
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();    // barriers + zeroed accumulators of BOTH CTAs
  tc_fence_after();
  if constexpr (CTAS == 2) {
    if (warp == 1) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int z = zlo; z <= zhi; ++z) {
        const int jlo = max(0, C7_GROUP * g - z + 3);          // first depth tap slot whose output is in the group
        const int nd = min(z + 3, C7_GROUP * g + C7_GROUP - 1) - max(z - 3, C7_GROUP * g) + 1;   // output depths of this slice
        for (int kw = 0; kw < 7; ++kw) {
          // pair mode: this CTA supplies rows [rank * N/2, (rank+1) * N/2) of the N = round16(24 * nd) row B operand
          const int npad = (nd * C7_COUT_P + 15) & ~15;
          const int brow = (kh * 7 + kw) * C7_BROWS + jlo * C7_COUT_P + (CTAS == 2 ? (int)cta_rank * (npad / 2) : 0);
          for (int blk = 0; blk < k.nblk; ++blk) {
            const uint32_t fb = bars + 8u * s;
            mbar_wait(fb + 8u * C7_STAGES, ph ^ 1u);
            const uint32_t sa = base + (uint32_t)s * C7_STAGE_BYTES;
            if constexpr (CTAS == 2) {
              if (cta_rank == 0) mbar_expect_tx(fb, 2u * C7_STAGE_BYTES);
              tma_load_5d_2sm(sa, &tmA, fb, blk * 64, w0 + kw - 3, h0 + kh - 3, z, b);
              tma_load_2d_2sm(sa + A_TILE_BYTES, &tmB, fb, blk * 64, brow);
            } else {
              mbar_expect_tx(fb, C7_STAGE_BYTES);
              tma_load_5d(sa, &tmA, fb, blk * 64, w0 + kw - 3, h0 + kh - 3, z, b);
              tma_load_2d(sa + A_TILE_BYTES, &tmB, fb, blk * 64, brow);
            }
            if (++s == C7_STAGES) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1 && (CTAS == 1 || cta_rank == 0)) {
    // ===== MMA issuer (converged warp, elected lane inside the asm block; the pair's leader only) =====
    int s = 0; uint32_t ph = 0;
    for (int z = zlo; z <= zhi; ++z) {
      const int dlo = max(z - 3, C7_GROUP * g), dhi = min(z + 3, C7_GROUP * g + C7_GROUP - 1);
      const uint32_t N = ((uint32_t)(dhi - dlo + 1) * C7_COUT_P + 15u) & ~15u;
      const uint32_t idesc = (1u << 4) | IDESC_AB_FMT | ((N >> 3) << 17) | (((128u * CTAS) >> 4) << 24);
      const uint32_t d_main = tmem_base + (uint32_t)((dlo - C7_GROUP * g) * C7_COUT_P);
      const uint32_t d_corr = d_main + 256u;
      for (int kw = 0; kw < 7; ++kw) {
        for (int blk = 0; blk < k.nblk; ++blk) {
          const uint32_t fb = bars + 8u * s;
          mbar_wait(fb, ph);
          tc_fence_after();
          const uint32_t sa = base + (uint32_t)s * C7_STAGE_BYTES;
          const int ksteps = (blk == k.nblk - 1) ? k.last_ksteps : 2;
          mma_stage_k<3, CTAS>(ksteps, d_main, d_corr, umma_desc(sa), umma_desc(sa + A_TILE_BYTES), idesc, 1u, 1u,
                               fb + 8u * C7_STAGES);
          if (++s == C7_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
    if constexpr (CTAS == 2) {
      asm volatile(
          "{\n\t.reg .pred pe;\n\t.reg .b16 mk;\n\tmov.b16 mk, 3;\n\t"
          "elect.sync _|pe, 0xffffffff;\n\t"
          "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], mk;\n\t}"
          ::"r"(tmem_full) : "memory");
    } else {
      asm volatile(
          "{\n\t.reg .pred pe;\n\t"
          "elect.sync _|pe, 0xffffffff;\n\t"
          "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
          ::"r"(tmem_full) : "memory");
    }
  } else if (warp >= 2) {
    // ===== epilogue: per output depth, 32 columns (main + corr) -> smem tile -> coalesced partial rows =====
    const int q = warp & 3;
    float* tile = reinterpret_cast<float*>(smem_raw + (stg - raw)) + q * 32 * STG_LD;
    const int sub = lane >> 3, c4 = (lane & 7) * 4;
    long poff[8];
    uint32_t vmask = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int r = q * 32 + sub + 4 * i;
      const int ow = w0 + (r & ((1 << k.lbw) - 1)); r >>= k.lbw;
      const int oh = h0 + r;
      if (ow < k.W && oh < k.H) vmask |= 1u << i;
      poff[i] = ((long)oh * k.W + ow) * k.ldo;
    }
    float* pbase = k.parts + (long)kh * k.part_stride + (long)b * 16 * k.H * k.W * k.ldo;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    // pre-compensated weights: output depths 13..15 have no slices z > 15 (a suffix of the chain: exact take-back), columns
    // within 3 pixels of the left / right edge lose interleaved kw taps (kappa * fraction * real events / 2 on average)
    float fw = 0.f;
    if (k.kappa != 0.f) {
      const int my_ow = w0 + ((q * 32 + lane) & ((1 << k.lbw) - 1));
      int vw = 0;
      for (int kw = 0; kw < 7; ++kw) vw += (my_ow + kw - 3 >= 0 && my_ow + kw - 3 < k.W) ? 1 : 0;
      fw = 0.5f * k.kappa * (1.f - (float)vw * (1.f / 7.f));
    }
    for (int dl = 0; dl < C7_GROUP; ++dl) {
      float msc = k.acc_scale;
      if (k.kappa != 0.f) {
        const int d = C7_GROUP * g + dl;
        const int nz = min(15, d + 3) - max(0, d - 3) + 1;
        const int zmiss = max(0, d + 3 - 15);
        msc = 1.f - k.kappa * (float)(zmiss * k.ev_slice) - fw * (float)(nz * k.ev_slice);
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[16], u[16];
        // (the second half reads 8 valid columns + 8 of the next depth slot, which are not stored: c4 < ldo below)
        tc_ld16(trow + (uint32_t)(dl * C7_COUT_P + 16 * half), v);
        tc_ld16(trow + (uint32_t)(256 + dl * C7_COUT_P + 16 * half), u);
        tc_ld_wait();
        float4* dst = reinterpret_cast<float4*>(tile + lane * STG_LD + 16 * half);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_float4(fmaf(__uint_as_float(v[4 * j]), msc, __uint_as_float(u[4 * j])) * k.out_scale,
                               fmaf(__uint_as_float(v[4 * j + 1]), msc, __uint_as_float(u[4 * j + 1])) * k.out_scale,
                               fmaf(__uint_as_float(v[4 * j + 2]), msc, __uint_as_float(u[4 * j + 2])) * k.out_scale,
                               fmaf(__uint_as_float(v[4 * j + 3]), msc, __uint_as_float(u[4 * j + 3])) * k.out_scale);
      }
      __syncwarp();
      if (c4 < k.ldo) {
        float* pd = pbase + (long)(C7_GROUP * g + dl) * k.H * k.W * k.ldo + c4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (!((vmask >> i) & 1u)) continue;
          *reinterpret_cast<float4*>(pd + poff[i]) = *reinterpret_cast<const float4*>(tile + (sub + 4 * i) * STG_LD + c4);
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CTAS == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// w32 [tap = (kd*7+kh)*7+kw][Cin][Cout] fp32 -> rows ((kh*7+kw)*7 + j)*32 + co, j <-> kd = 6 - j (ascending output
// depth d = z - 3 + j), columns [blk][hi 32 | lo 32]
// kappa: truncation pre-compensation per event.  Issue order of conv7_tc_kernel for one (kh): z ascending, kw, blk, then
// hh k0, hh k1 into the main accumulator (corrections go to their own): K step (kd, kw, blk, ks) of the chain of an output
// depth is followed by (6-kd) slices x ev_slice + (6-kw) x ev_kw + the rest of its kw group.
__global__ void __launch_bounds__(256) pack_conv7_kernel(const float* __restrict__ w32, __nv_bfloat16* __restrict__ out, int Cin,
                                                         int Cout, int nblk, float wmul, float kappa) {
  const int last_ks = ((Cin - (nblk - 1) * 32) + 15) / 16;
  const int ev_kw = 2 * (nblk - 1) + last_ks, ev_slice = 7 * ev_kw;
  const long total = 49L * 7 * C7_COUT_P * nblk * 32;         // the 8 pad rows per filter position stay zero (memset)
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int e = (int)(i & 31); long r = i >> 5;
    const int blk = (int)(r % nblk); r /= nblk;
    const int co = (int)(r % C7_COUT_P); r /= C7_COUT_P;
    const int j = (int)(r % 7); const int khw = (int)(r / 7);       // khw = kh*7 + kw
    const int kd = 6 - j, ci = blk * 32 + e;
    const int tap = kd * 49 + khw;
    const int kw = khw % 7;
    const int rem = (6 - kd) * ev_slice + (6 - kw) * ev_kw + (ev_kw - (2 * blk + (e >> 4)));
    const float v = (co < Cout && ci < Cin) ? w32[((long)tap * Cin + ci) * Cout + co] * wmul * (1.0f + kappa * (float)rem) : 0.f;
    __nv_bfloat16 hi, lo;
    split_operand(v, hi, lo);
    const long row = (long)khw * C7_BROWS + j * C7_COUT_P + co;
    const long o = row * (nblk * 64L) + blk * 64 + e;
    out[o] = hi;
    out[o + 32] = lo;
  }
}

// logits[i] = sum_p parts[p][i] + bias[channel]
__global__ void __launch_bounds__(256) sum_parts_kernel(const float4* __restrict__ parts, long stride4, int nparts,
                                                        const float* __restrict__ bias, int Cout, int ldo4,
                                                        float4* __restrict__ out, long n4) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 a = __ldg(parts + i);
    for (int p = 1; p < nparts; ++p) {
      const float4 v = __ldg(parts + p * stride4 + i);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    const int c = (int)(i % ldo4) * 4;
    if (bias) {
      if (c < Cout) a.x += bias[c];
      if (c + 1 < Cout) a.y += bias[c + 1];
      if (c + 2 < Cout) a.z += bias[c + 2];
      if (c + 3 < Cout) a.w += bias[c + 3];
    }
    out[i] = a;
  }
}

bool g_attr7[64] = {};

}  // namespace

bool conv7_supported(const ConvW& w, const Act& out) {
  return w.w7 != nullptr && out.D == 16 && (long)out.H * out.W >= 128 && out.sw % 4 == 0 && out.sw >= w.Cout && out.sw <= C7_COUT_P &&
         out.sh == (long)out.W * out.sw && out.sd == (long)out.H * out.sh && out.sb == 16 * out.sd;
}

size_t conv7_scratch_floats(const Act& out) { return (size_t)7 * out.B * 16 * out.H * out.W * out.sw; }

void pack_conv7(cs_ctx* ctx, ConvW& w) {
  if (!(w.KD == 7 && w.KH == 7 && w.KW == 7 && w.Cout <= C7_COUT_P && w.Cin >= 16 && w.w32)) return;
  const int nblk = (w.Cin + 31) / 32;
  const size_t n = (size_t)49 * C7_BROWS * nblk * 64;
  if (!w.w7) w.w7 = static_cast<__nv_bfloat16*>(ctx->dmalloc(n * sizeof(__nv_bfloat16)));
  CS_CUDA(cudaMemset(w.w7, 0, n * sizeof(__nv_bfloat16)));
  w.w7_kappa = (float)ctx->tc_poscomp * 1e-10f;
  pack_conv7_kernel<<<148 * 8, 256>>>(w.w32, w.w7, w.Cin, w.Cout, nblk, w.wmul, w.w7_kappa);
  check_launch("pack_conv7");
}

// x: split-fp16 operand [B,16,H,W,nblk*64]; out: logits [B,16,H,W,ldo] (dense, ldo = out.sw >= Cout);
// scratch: conv7_scratch_floats(out) floats
void conv7_tc(const Launcher& L, const Opd& x, const ConvW& w, Act out, float* scratch) {
  L.count(); L.count();
  if (L.dry) return;
  CS_REQUIRE(conv7_supported(w, out) && x.D == 16 && x.nblk == (w.Cin + 31) / 32 && x.B == out.B && x.H == out.H &&
                 x.W == out.W, CS_ERR_INVALID, "conv7_tc: unsupported geometry");
  Conv7K k{};
  k.B = x.B; k.H = x.H; k.W = x.W;
  int cap = 128;
  const int bw = pick_box(x.W, cap, &k.lbw); cap /= bw;
  const int bh = cap; k.lbh = 0; while ((1 << k.lbh) < bh) ++k.lbh;
  k.ntw = (x.W + bw - 1) / bw; k.nth = (x.H + bh - 1) / bh;
  k.nblk = x.nblk;
  k.last_ksteps = ((w.Cin - (x.nblk - 1) * 32) + 15) / 16;
  k.Cout = w.Cout;
  k.ldo = (int)out.sw;
  k.parts = scratch;
  k.part_stride = (long)x.B * 16 * x.H * x.W * k.ldo;
  k.out_scale = 1.0f / (w.wmul * x.amul);
  operand_absmax(L, x, w.id);
  k.ev_slice = 7 * (2 * (x.nblk - 1) + k.last_ksteps);
  k.kappa = w.w7_kappa;
  k.acc_scale = w.w7_kappa != 0.f ? 1.0f : 1.0f + L.acc_comp * 1e-10f * (float)(7 * k.ev_slice);   // chain: 7 z x 7 kw x K steps

  const unsigned gx = (unsigned)(k.ntw * k.nth * x.B);
  const bool pair = L.pair && (gx % 2 == 0);
  auto enc = encode_fn();
  CUtensorMap tmA, tmB;
  const int rowA = x.nblk * 64;
  {
    const cuuint64_t pix = (cuuint64_t)rowA * 2;
    cuuint64_t dims[5] = {(cuuint64_t)rowA, (cuuint64_t)x.W, (cuuint64_t)x.H, 16, (cuuint64_t)x.B};
    cuuint64_t strides[4] = {pix, pix * x.W, pix * x.W * x.H, pix * x.W * x.H * 16};
    cuuint32_t box[5] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x.p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CS_REQUIRE(r == CUDA_SUCCESS, CS_ERR_CUDA, "conv7_tc: cuTensorMapEncodeTiled(A) failed");
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)rowA, (cuuint64_t)(49 * C7_BROWS)};
    cuuint64_t strides[1] = {(cuuint64_t)rowA * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(pair ? C7_BROWS / 2 : C7_BROWS)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w.w7, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CS_REQUIRE(r == CUDA_SUCCESS, CS_ERR_CUDA, "conv7_tc: cuTensorMapEncodeTiled(B) failed");
  }
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (!g_attr7[dev & 63]) {
    CS_CUDA(cudaFuncSetAttribute(conv7_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, C7Cfg<1>::SMEM));
    CS_CUDA(cudaFuncSetAttribute(conv7_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, C7Cfg<2>::SMEM));
    g_attr7[dev & 63] = true;
  }
  const long M = (long)x.B * 16 * x.H * x.W;
  {
    ProfScope ps(L, PK_CONV_TC, 2.0 * (double)M * w.Cout * w.Cin * 343.0, 0.0, "conv7");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(gx, 16 / C7_GROUP, 7); cfg.blockDim = dim3(TC_THREADS); cfg.stream = L.stream;
    cfg.dynamicSmemBytes = pair ? C7Cfg<2>::SMEM : C7Cfg<1>::SMEM;
    cudaLaunchAttribute attr[1];
    if (pair) {
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      CS_CUDA(cudaLaunchKernelEx(&cfg, conv7_tc_kernel<2>, tmA, tmB, k));
    } else {
      CS_CUDA(cudaLaunchKernelEx(&cfg, conv7_tc_kernel<1>, tmA, tmB, k));
    }
  }
  {
    const long n4 = M * k.ldo / 4;
    ProfScope ps(L, PK_OTHER, 0.0, (double)n4 * 16.0 * 8.0, "sum_parts");
    long blocks = (n4 + 255) / 256; if (blocks > 148L * 16) blocks = 148L * 16;
    sum_parts_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(reinterpret_cast<const float4*>(scratch), k.part_stride / 4, 7, w.bias,
                                                           w.Cout, k.ldo / 4, reinterpret_cast<float4*>(out.p), n4);
    check_launch("sum_parts");
  }
}

}  // namespace cs
