// The 32 -> 32 channel 3x3x3 convolutions of the 32x16 feature volume (ResBlock3d / ResBlock3D_stage3_leak, reference
// util.py:80-102,515-544: 36 of them per frame).  As a plain implicit GEMM they have N = 32 and re-read every input tile
// from L2 once per filter tap (27x): the generic kernel runs them at the L2 bandwidth limit.  Dedicated tcgen05 kernel,
// "depth-stacked, weights resident, persistent":
//
//   * one tile = 128 (h,w) pixels x 8 output depths x 32 channels = 256 fp32 accumulator columns; the TMEM holds TWO such
//     buffers, so the epilogue of tile i (bias / activation / residual / operand emission / statistics, ~25 % of a tile)
//     overlaps the MMAs of tile i+1.  One persistent CTA per SM walks the tiles round-robin (round 1: one 16-depth tile per
//     CTA, the whole TMEM, epilogue exposed, 1.73 waves).
//   * for an input slice z and an in-plane tap (kh,kw) the A tile (128 pixels x 32 channels, shifted by the tap, TMA
//     zero-fills the h/w padding) feeds the output depths d = z-1, z, z+1 at once: the B tile stacks the depth taps
//     kd = z-d+1 along N (3 x 32 = 96 columns landing at TMEM column 32*(z-1)), so the MMAs are 3x wider.  A tile of 8
//     depths walks 10 input slices (one halo slice on either side, N = 32 there).
//   * halo tiles: the pixel tile is 8 (w) x 16 (h), so a shift by one image row is a shift by 8 operand rows = exactly one
//     1024-byte SWIZZLE_128B atom.  One TMA box of 8 x 18 pixels (18 KB) therefore serves the three kh taps of a (z, kw):
//     their A descriptors start 0 / 1024 / 2048 bytes into the stage.
//   * the complete weight set (9 in-plane taps x 96 rows x 128 B = 108 KB) is loaded into shared memory once per CTA;
//     the pipeline only streams A tiles.
//   * accumulation chain per output column: 3 slices x 9 taps x 2 K-steps x 3 passes = 162 MMAs in one accumulator.  The
//     packed weights carry the position-dependent pre-compensation of the tensor core's accumulate truncation for exactly
//     this issue order (tc_ptx.cuh): K step (kd,kh,kw,ks) is followed by (2-kd)*54 + (2-kw)*18 + (2-kh)*6 + 6-ks events.
//     Depth 15 has no slice z = 16: its chain ends 54 events early, which the epilogue undoes with one factor.
//   * optional per-tile partial sums (sum x, sum x^2 per channel) of the fp32 output for the GroupNorm / InstanceNorm that
//     follows (reference util.py:521-523): written per tile, reduced in a fixed order by stats_from_tiles -- no separate
//     read of the tensor, deterministic, independent of the batch size.
//
// Operand format, MMA issue and the fused epilogue (bias, activation, residual, fp32 store and / or emission of the next
// conv's operand) are those of conv_tc.cu.
#include "tc_ptx.cuh"

namespace cs {

using namespace tc;

namespace {

constexpr int C3_BTAP_BYTES = 96 * 128;                   // one in-plane tap: 3 depth taps x 32 couts x [hi 32 | lo 32]
constexpr int C3_B_BYTES = 9 * C3_BTAP_BYTES;             // 108 KB resident weights
constexpr int C3_HALO_BYTES = 18 * 8 * 128;               // 8 (w) x 18 (h) pixels x [hi 32 | lo 32]
constexpr int C3_STAGES = 4;                              // 4 x 18 KB in flight (18 MMAs per stage)
constexpr int C3_EGROUPS = 2;                             // 8 epilogue warps: the groups take alternate output depths
constexpr int C3_DG = 8;                                  // output depths per tile
constexpr int C3_RED_BYTES = 2 * 8 * 64 * 4;              // statistics: [parity][epilogue warp][32 channels x (s1, s2)]
constexpr int C3_SMEM = C3_B_BYTES + C3_STAGES * C3_HALO_BYTES + C3_EGROUPS * STG_BYTES + C3_RED_BYTES + 1024 + 16 * C3_STAGES + 96;
constexpr int C3_THREADS = 64 + 128 * C3_EGROUPS;

struct Conv3sK {
  int B, H, W;                     // D = 16, C = 32
  int ntw, nth, tiles;             // tiles of 8 (w) x 16 (h) pixels x 8 depths: tile = ((b * nth + th) * ntw + tw) * 2 + g
  const float* bias; int act; float slope;
  const float* res; long rb, rd, rh, rw;
  float* y; long yb, yd, yh, yw;
  __nv_bfloat16* emit; const float* escale; const float* eshift; int eact; float eslope; float emul;
  float out_scale;                 // 1 / ConvW::wmul (x the constant truncation compensation when the weights carry none)
  float kappa;                     // > 0: pre-compensated weights; the epilogue takes back what was assumed for taps in the zero padding:
                                   // depth 15 ends 54 events early (exact), h / w border rows lose interleaved taps (kappa * fraction * events / 2)
  float* stats;                    // optional [tiles][32][2] partial sums of the fp32 output
};

template <bool RES, bool EMIT, bool STATS>
__global__ void __launch_bounds__(C3_THREADS) conv3s_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB, Conv3sK k) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bres = base;                                        // resident weights
  const uint32_t abase = base + C3_B_BYTES;                          // A stages
  const uint32_t stg = abase + (uint32_t)C3_STAGES * C3_HALO_BYTES;
  const uint32_t red = stg + C3_EGROUPS * STG_BYTES;
  const uint32_t bars = red + C3_RED_BYTES;                          // full[S], empty[S], tmem_full[2], tmem_empty[2], wready, slot
  const uint32_t tmem_full = bars + 16u * C3_STAGES;
  const uint32_t tmem_empty = tmem_full + 16u;
  const uint32_t wready = tmem_empty + 16u;
  const uint32_t tmem_slot = wready + 8u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < C3_STAGES; ++s) { mbar_init(bars + 8u * s, 1); mbar_init(bars + 8u * (C3_STAGES + s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full + 8u * b, 1); mbar_init(tmem_empty + 8u * b, 4 * C3_EGROUPS); }
    mbar_init(wready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  // zero both accumulator buffers: every MMA accumulates (a slice touches a sliding window of depth columns); after that
  // the epilogue re-zeroes a buffer before it hands it back
  if (warp >= 2) {
    const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int c = ((warp - 2) >> 2) * 16; c < 512; c += 16 * C3_EGROUPS) tc_st16_zero(trow + (uint32_t)c);
    tc_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ===== TMA producer: the weights once, then the A tiles of every tile of this CTA =====
    if (lane == 0) {
      mbar_expect_tx(wready, C3_B_BYTES);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(bres + (uint32_t)tap * C3_BTAP_BYTES, &tmB, wready, 0, tap * 96);
      int s = 0; uint32_t ph = 0;
      for (int t = blockIdx.x; t < k.tiles; t += gridDim.x) {
        const int g = t & 1; int r = t >> 1;
        const int tw = r % k.ntw; r /= k.ntw;
        const int th = r % k.nth; const int b = r / k.nth;
        const int w0 = tw << 3, h0 = th << 4;
        const int zlo = max(0, C3_DG * g - 1), zhi = min(15, C3_DG * g + C3_DG);
        for (int z = zlo; z <= zhi; ++z) {
          for (int kw = 0; kw < 3; ++kw) {
            const uint32_t fb = bars + 8u * s;
            mbar_wait(fb + 8u * C3_STAGES, ph ^ 1u);
            mbar_expect_tx(fb, C3_HALO_BYTES);
            tma_load_5d(abase + (uint32_t)s * C3_HALO_BYTES, &tmA, fb, 0, w0 + kw - 1, h0 - 1, z, b);
            if (++s == C3_STAGES) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (converged warp, elected lane inside the asm block) =====
    mbar_wait(wready, 0);
    int s = 0; uint32_t ph = 0;
    int it = 0;
    for (int t = blockIdx.x; t < k.tiles; t += gridDim.x, ++it) {
      const int g = t & 1;
      const int buf = it & 1, use = it >> 1;
      if (use > 0) {                                       // the epilogue warps have drained and re-zeroed this buffer
        mbar_wait(tmem_empty + 8u * buf, (uint32_t)((use - 1) & 1));
        tc_fence_after();
      }
      const int d0 = C3_DG * g;
      const int zlo = max(0, d0 - 1), zhi = min(15, d0 + C3_DG);
      for (int z = zlo; z <= zhi; ++z) {
        const int dlo = max(z - 1, d0), dhi = min(z + 1, d0 + C3_DG - 1);
        const uint32_t N = (uint32_t)(dhi - dlo + 1) * 32u;
        const uint32_t idesc = (1u << 4) | IDESC_AB_FMT | ((N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t d_acc = tmem_base + (uint32_t)(buf * 256 + (dlo - d0) * 32);
        const uint32_t brow = (uint32_t)(dlo - (z - 1)) * 32u * 128u;  // skip the depth-tap slots of outputs outside the tile
        for (int kw = 0; kw < 3; ++kw) {
          const uint32_t fb = bars + 8u * s;
          mbar_wait(fb, ph);
          tc_fence_after();
          const uint32_t sa = abase + (uint32_t)s * C3_HALO_BYTES;
          // kh = 0, 1, 2: the 128-row window of the halo tile starting kh image rows (= kh swizzle atoms) down
          mma_stage32_nocommit(d_acc, d_acc, umma_desc(sa), umma_desc(bres + (uint32_t)kw * C3_BTAP_BYTES + brow), idesc, 1u, 1u);
          mma_stage32_nocommit(d_acc, d_acc, umma_desc(sa + 1024u), umma_desc(bres + (uint32_t)(3 + kw) * C3_BTAP_BYTES + brow), idesc, 1u, 1u);
          mma_stage<3, 2>(d_acc, d_acc, umma_desc(sa + 2048u), umma_desc(bres + (uint32_t)(6 + kw) * C3_BTAP_BYTES + brow), idesc, 1u, 1u,
                          fb + 8u * C3_STAGES);
          if (++s == C3_STAGES) { s = 0; ph ^= 1u; }
        }
      }
      asm volatile(
          "{\n\t.reg .pred pe;\n\t"
          "elect.sync _|pe, 0xffffffff;\n\t"
          "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
          ::"r"(tmem_full + 8u * buf) : "memory");
    }
  } else {
    // ===== epilogue: per output depth 32 columns -> smem tile -> coalesced rows =====
    const int q = warp & 3, eg = (warp - 2) >> 2;
    float* tile = reinterpret_cast<float*>(smem_raw + (stg - raw)) + (eg * 4 + q) * 32 * STG_LD;
    float* redf = reinterpret_cast<float*>(smem_raw + (red - raw));
    const int sub = lane >> 3, c4 = (lane & 7) * 4;
    float bz[4] = {0.f, 0.f, 0.f, 0.f};
    if (k.bias) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(k.bias + c4));
      bz[0] = t4.x; bz[1] = t4.y; bz[2] = t4.z; bz[3] = t4.w;
    }
    float es[4] = {1.f, 1.f, 1.f, 1.f}, eb[4] = {0.f, 0.f, 0.f, 0.f};
    if (EMIT && k.escale) {
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(k.escale + c4));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(k.eshift + c4));
      es[0] = s4.x; es[1] = s4.y; es[2] = s4.z; es[3] = s4.w;
      eb[0] = b4.x; eb[1] = b4.y; eb[2] = b4.z; eb[3] = b4.w;
    }
    const long dstride_e = (long)k.H * k.W;
    const float aslope = leaky_slope(k.act, k.slope), easlope = leaky_slope(k.eact, k.eslope);
    const uint32_t trow0 = tmem_base + ((uint32_t)(q * 32) << 16);
    int it = 0;
    for (int t = blockIdx.x; t < k.tiles; t += gridDim.x, ++it) {
      const int g = t & 1; int r0 = t >> 1;
      const int tw = r0 % k.ntw; r0 /= k.ntw;
      const int th = r0 % k.nth; const int b = r0 / k.nth;
      const int w0 = tw << 3, h0 = th << 4, d0 = C3_DG * g;
      long yoff[8], roff[RES ? 8 : 1], epix[EMIT ? 8 : 1];
      float bfr[8];                                          // kappa / 2 x fraction of the 9 in-plane taps in the padding
      uint32_t vmask = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int r = q * 32 + sub + 4 * i;
        const int ow = w0 + (r & 7); r >>= 3;
        const int oh = h0 + r;
        if (ow < k.W && oh < k.H) vmask |= 1u << i;
        const int vh = 3 - (oh == 0) - (oh == k.H - 1), vw = 3 - (ow == 0) - (ow == k.W - 1);
        bfr[i] = 0.5f * k.kappa * (1.f - (float)(vh * vw) * (1.f / 9.f));
        yoff[i] = b * k.yb + oh * k.yh + ow * k.yw + c4;
        if constexpr (RES) roff[i] = b * k.rb + oh * k.rh + ow * k.rw + c4;
        if constexpr (EMIT) epix[i] = (((long)b * 16) * k.H + oh) * k.W + ow;
      }
      if constexpr (RES) {                                  // pull this warp's residual rows towards L2 while the MMAs run
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (!((vmask >> i) & 1u)) continue;
          for (int d = d0 + eg; d < d0 + C3_DG; d += C3_EGROUPS)
            if ((lane & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(k.res + roff[i] + d * k.rd));
        }
      }
      const int buf = it & 1, use = it >> 1;
      mbar_wait(tmem_full + 8u * buf, (uint32_t)(use & 1));
      tc_fence_after();
      const uint32_t trow = trow0 + (uint32_t)(buf * 256);
      float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
      for (int dl = eg; dl < C3_DG; dl += C3_EGROUPS) {
        const int d = d0 + dl;
        const float ev_real = (d == 0 || d == 15) ? 108.f : 162.f;     // events of the slices that exist
        const float zsc = d == 15 ? 1.f - 54.f * k.kappa : 1.f;
        {
          uint32_t v[32];
          tc_ld16(trow + (uint32_t)(dl * 32), v);
          tc_ld16(trow + (uint32_t)(dl * 32 + 16), v + 16);
          tc_ld_wait();
          float4* dst = reinterpret_cast<float4*>(tile + lane * STG_LD);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                 __uint_as_float(v[4 * j + 3]));
        }
        // this warp's columns of depth dl are in registers / smem: clear them for the tile after next
        tc_st16_zero(trow + (uint32_t)(dl * 32));
        tc_st16_zero(trow + (uint32_t)(dl * 32 + 16));
        __syncwarp();
        float4 rr4[RES ? 8 : 1];
        if constexpr (RES) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            rr4[i] = ((vmask >> i) & 1u) ? *reinterpret_cast<const float4*>(k.res + roff[i] + d * k.rd) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (!((vmask >> i) & 1u)) continue;
          const float4 a = *reinterpret_cast<const float4*>(tile + (sub + 4 * i) * STG_LD + c4);
          const float osc = k.out_scale * (zsc - bfr[i] * ev_real);
          float o[4] = {fmaf(a.x, osc, bz[0]), fmaf(a.y, osc, bz[1]), fmaf(a.z, osc, bz[2]), fmaf(a.w, osc, bz[3])};
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = apply_leaky(o[j], aslope);
          if constexpr (RES) { o[0] += rr4[i].x; o[1] += rr4[i].y; o[2] += rr4[i].z; o[3] += rr4[i].w; }
          if constexpr (STATS) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { s1[j] += o[j]; s2[j] = fmaf(o[j], o[j], s2[j]); }
          }
          if (k.y) *reinterpret_cast<float4*>(k.y + yoff[i] + d * k.yd) = make_float4(o[0], o[1], o[2], o[3]);
          if constexpr (EMIT) {
            float e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] = apply_leaky(fmaf(o[j], es[j], eb[j]), easlope) * k.emul;
            uint2 hv, lv;
            split_operand4(e[0], e[1], e[2], e[3], hv, lv);
            __nv_bfloat16* ep = k.emit + (epix[i] + d * dstride_e) * 64 + c4;
            *reinterpret_cast<uint2*>(ep) = hv;
            *reinterpret_cast<uint2*>(ep + 32) = lv;
          }
        }
        __syncwarp();
      }
      // hand the (re-zeroed) accumulator buffer back to the MMA issuer
      tc_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty + 8u * buf) : "memory");
      if constexpr (STATS) {
        // per-tile partial sums: lanes with the same channel quad (xor 8, 16) -> the 8 epilogue warps (fixed order) -> global
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 8); s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 8);
          s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16); s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
        }
        float* rp = redf + ((it & 1) * 8 + (warp - 2)) * 64;
        if (lane < 8) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { rp[(c4 + j) * 2] = s1[j]; rp[(c4 + j) * 2 + 1] = s2[j]; }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");      // the 8 epilogue warps
        if (warp == 2) {
          const float* r0p = redf + (it & 1) * 8 * 64;
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int wq = 0; wq < 8; ++wq) { a0 += r0p[wq * 64 + lane * 2]; a1 += r0p[wq * 64 + lane * 2 + 1]; }
          *reinterpret_cast<float2*>(k.stats + ((long)t * 32 + lane) * 2) = make_float2(a0, a1);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// per-(b,c) mean / rstd from the per-tile partial sums of conv3s_tc (fixed summation order, fp64)
__global__ void stats_from_tiles_kernel(const float* __restrict__ part, int tiles_per_sample, float* __restrict__ mean,
                                        float* __restrict__ rstd, int n /*B*32*/, double inv_count, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = i >> 5, c = i & 31;
  double s1 = 0.0, s2 = 0.0;
  const float* p = part + ((long)b * tiles_per_sample * 32 + c) * 2;
  for (int t = 0; t < tiles_per_sample; ++t) { s1 += (double)p[(long)t * 64]; s2 += (double)p[(long)t * 64 + 1]; }
  const double m = s1 * inv_count;
  double var = s2 * inv_count - m * m;
  if (var < 0.0) var = 0.0;
  mean[i] = (float)m;
  rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

// w32 [tap = (kd*3+kh)*3+kw][32][32] fp32 -> rows ((kh*3+kw)*3 + j)*32 + co with kd = 2 - j (ascending output depth
// d = z - 1 + j), 64 columns [hi 32 | lo 32]
// kappa: truncation pre-compensation per event; K step (kd, kh, kw, ks) is followed by (2-kd)*54 + (2-kw)*18 + (2-kh)*6 + 6-ks
// events of its accumulator (issue order of conv3s_tc_kernel: z ascending, kw, kh, then hh k0, hh k1, lh, lh, hl, hl)
__global__ void __launch_bounds__(256) pack_conv3s_kernel(const float* __restrict__ w32, __nv_bfloat16* __restrict__ out, float wmul,
                                                          float kappa) {
  const int total = 9 * 3 * 32 * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i & 31; int r = i >> 5;
    const int co = r & 31; r >>= 5;
    const int j = r % 3; const int khw = r / 3;
    const int kd = 2 - j, kh = khw / 3, kw = khw % 3;
    const int tap = kd * 9 + khw;
    const int rem = (2 - kd) * 54 + (2 - kw) * 18 + (2 - kh) * 6 + 6 - (ci >> 4);
    const float v = w32[((long)tap * 32 + ci) * 32 + co] * wmul * (1.0f + kappa * (float)rem);
    __nv_bfloat16 hi, lo;
    split_operand(v, hi, lo);
    const long o = (((long)khw * 3 + j) * 32 + co) * 64 + ci;
    out[o] = hi;
    out[o + 32] = lo;
  }
}

bool g_attr3[64] = {};

}  // namespace

bool conv3s_supported(const ConvW& w, int H, int W) {
  return w.w3s != nullptr && (long)H * W >= 128;
}

void pack_conv3s(cs_ctx* ctx, ConvW& w) {
  if (!(w.KD == 3 && w.KH == 3 && w.KW == 3 && w.Cin == 32 && w.Cout == 32 && w.w32)) return;
  if (!w.w3s) w.w3s = static_cast<__nv_bfloat16*>(ctx->dmalloc((size_t)9 * 96 * 64 * sizeof(__nv_bfloat16)));
  w.w3s_kappa = (float)ctx->tc_poscomp * 1e-10f;
  pack_conv3s_kernel<<<108, 256>>>(w.w32, w.w3s, w.wmul, w.w3s_kappa);
  check_launch("pack_conv3s");
}

// x: split-fp16 operand [B,16,H,W,64] of a 32-channel volume; y (fp32, may have a null pointer when only the operand is
// emitted) / residual: channels-last with generic strides and channel stride 1 (the [B,h,w,16,32] volume view).
// stats_part != null: also writes conv3s_stats_floats() per-tile partial sums of the fp32 output (see conv3s_stats).
size_t conv3s_stats_floats(int B, int H, int W) { return (size_t)B * ((H + 15) / 16) * ((W + 7) / 8) * 2 * 64; }

void conv3s_tc(const Launcher& L, const Opd& x, const ConvW& w, const Epilogue& e, Act y, float* stats_part) {
  L.count();
  if (L.dry) return;
  CS_REQUIRE(conv3s_supported(w, x.H, x.W) && x.D == 16 && x.nblk == 1 && y.D == 16 && y.C == 32 && y.H == x.H && y.W == x.W &&
                 y.B == x.B, CS_ERR_INVALID, "conv3s_tc: unsupported geometry");
  auto al4 = [](long v) { return (v & 3) == 0; };
  CS_REQUIRE(al4(y.sb) && al4(y.sd) && al4(y.sh) && al4(y.sw) && ((uintptr_t)y.p % 16 == 0), CS_ERR_INVALID,
             "conv3s_tc: output must be 16-byte aligned");
  CS_REQUIRE(!e.mult, CS_ERR_INVALID, "conv3s_tc: per-pixel multiplier not supported");
  CS_REQUIRE(act_is_leaky(e.act) && act_is_leaky(e.emit_act), CS_ERR_INVALID, "conv3s_tc: activation must be none / relu / leaky relu");
  Conv3sK k{};
  k.B = x.B; k.H = x.H; k.W = x.W;
  k.ntw = (x.W + 7) / 8; k.nth = (x.H + 15) / 16;
  k.tiles = k.ntw * k.nth * x.B * 2;
  k.bias = w.bias; k.act = e.act; k.slope = e.slope;
  k.res = e.residual; k.rb = e.rs_b; k.rd = e.rs_d; k.rh = e.rs_h; k.rw = e.rs_w;
  if (e.residual)
    CS_REQUIRE(al4(e.rs_b) && al4(e.rs_d) && al4(e.rs_h) && al4(e.rs_w) && ((uintptr_t)e.residual % 16 == 0), CS_ERR_INVALID,
               "conv3s_tc: residual must be 16-byte aligned");
  k.y = y.p; k.yb = y.sb; k.yd = y.sd; k.yh = y.sh; k.yw = y.sw;
  if (e.emit) {
    CS_REQUIRE(e.emit_nblk == 1, CS_ERR_INVALID, "conv3s_tc: emitted operand must have 32 channels");
    k.emit = e.emit; k.escale = e.emit_scale; k.eshift = e.emit_shift; k.eact = e.emit_act; k.eslope = e.emit_slope;
  }
  // weights packed with the position-dependent pre-compensation need no epilogue factor except at depth 15 (54 events early);
  // otherwise the constant factor of the mean loss of a 162-MMA chain (CS_OPT_TC_COMP)
  k.kappa = w.w3s_kappa;
  k.out_scale = (w.w3s_kappa != 0.f ? 1.0f : 1.0f + L.acc_comp * 1e-10f * 162.f) / (w.wmul * x.amul);
  k.emul = e.emit_mul;
  operand_absmax(L, x, w.id);
  k.stats = stats_part;

  auto enc = encode_fn();
  CUtensorMap tmA, tmB;
  {
    const cuuint64_t pix = 64 * 2;
    cuuint64_t dims[5] = {64, (cuuint64_t)x.W, (cuuint64_t)x.H, 16, (cuuint64_t)x.B};
    cuuint64_t strides[4] = {pix, pix * x.W, pix * x.W * x.H, pix * x.W * x.H * 16};
    cuuint32_t box[5] = {64, 8, 18, 1, 1};                 // halo tile: 8 (w) x 18 (h) pixels
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x.p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CS_REQUIRE(r == CUDA_SUCCESS, CS_ERR_CUDA, "conv3s_tc: cuTensorMapEncodeTiled(A) failed");
  }
  {
    cuuint64_t dims[2] = {64, 9 * 96};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 96};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w.w3s, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CS_REQUIRE(r == CUDA_SUCCESS, CS_ERR_CUDA, "conv3s_tc: cuTensorMapEncodeTiled(B) failed");
  }
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, Conv3sK);
  static const KernelFn fns[2][2][2] = {{{conv3s_tc_kernel<false, false, false>, conv3s_tc_kernel<false, false, true>},
                                         {conv3s_tc_kernel<false, true, false>, conv3s_tc_kernel<false, true, true>}},
                                        {{conv3s_tc_kernel<true, false, false>, conv3s_tc_kernel<true, false, true>},
                                         {conv3s_tc_kernel<true, true, false>, conv3s_tc_kernel<true, true, true>}}};
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  static int n_sm = 0;
  if (!g_attr3[dev & 63]) {
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        for (int c = 0; c < 2; ++c) CS_CUDA(cudaFuncSetAttribute(fns[a][b][c], cudaFuncAttributeMaxDynamicSharedMemorySize, C3_SMEM));
    CS_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    g_attr3[dev & 63] = true;
  }
  const long M = (long)x.B * 16 * x.H * x.W;
  ProfScope ps(L, PK_CONV_TC, 2.0 * (double)M * 32 * 32 * 27, 0.0, "conv3s");
  const int grid = k.tiles < n_sm ? k.tiles : n_sm;        // persistent: one CTA per SM walks the tiles round-robin
  fns[k.res != nullptr][k.emit != nullptr][k.stats != nullptr]<<<grid, C3_THREADS, C3_SMEM, L.stream>>>(tmA, tmB, k);
  check_launch("conv3s_tc");
}

// mean / rstd per (b, c) of the tensor whose per-tile partial sums a conv3s_tc launch wrote (GroupNorm(32,32) == per-channel
// instance norm over D*H*W, reference util.py:521-523): fixed summation order, fp64
void conv3s_stats(const Launcher& L, const float* stats_part, int B, int H, int W, float* mean, float* rstd, float eps) {
  L.count();
  if (L.dry) return;
  const int tps = ((H + 15) / 16) * ((W + 7) / 8) * 2;
  const int n = B * 32;
  stats_from_tiles_kernel<<<(n + 127) / 128, 128, 0, L.stream>>>(stats_part, tps, mean, rstd, n, 1.0 / ((double)16 * H * W), eps);
  check_launch("stats_from_tiles");
}

}  // namespace cs
