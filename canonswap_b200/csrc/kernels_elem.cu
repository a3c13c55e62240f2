// Element-wise / reduction / layout kernels of the generator hot path (HBM-bound, sm_100a).
//   prep            gather (+nearest-up / 2x2 avg-pool / concat) -> normalise -> SPADE modulate -> (+add) -> act
//                   -> fp32 channels-last and/or split-fp16 operand planes for the tcgen05 conv
//   instance_stats  per-(b,c) mean / rstd (InstanceNorm2d, GroupNorm(32,32)), reference util.py:286,521
//   adaptive_blend  mask*out_mod + (1-mask)*out_std, reference adaptive_modulate.py:186
//   nchw<->cl       layout shims at the per-stage C-ABI boundary
//   ingest / emit   u8 HWC -> fp32 (can_swap_e2e.py:147-163);  sigmoid + PixelShuffle(2) + u8 (spade_generator.py:56-57, can_swap_e2e.py:314-322)
#include "common.cuh"

namespace cs {

// ------------------------------------------------------------------------------------------
// prep
// ------------------------------------------------------------------------------------------
struct PrepK {
  const float* s0; long s0b, s0d, s0h, s0w; int C0;
  const float* s1; long s1b, s1d, s1h, s1w; int C1;
  int upshift, pool2, norm, act; float slope;
  const float *scale, *shift, *mean, *rstd, *gb, *add; long ab, ad, ah, aw;
  int B, D, H, W;           // output geometry
  int Cl;                   // logical channels C0 + C1
  int Cout;                 // channels written (>= Cl, pad written as zero)
  // outputs
  float* o32; long ob, od, oh, ow;
  __nv_bfloat16* opl; long prow;   // split-fp16 operand: dense pixels, prow = nblk*64 elements per pixel
  float amul;                      // Opd::amul
};

__global__ void __launch_bounds__(256) prep_kernel(PrepK k) {
  long total = (long)k.B * k.D * k.H * k.W * k.Cout;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    // 32-bit divisions (total < 2^32: host check); a 64-bit division costs ~100 instructions
    const unsigned ui = (unsigned)idx, upix = ui / (unsigned)k.Cout;
    const int c = (int)(ui - upix * (unsigned)k.Cout);
    const long pix = upix;
    const unsigned t1 = upix / (unsigned)k.W, t2 = t1 / (unsigned)k.H;
    const int w = (int)(upix - t1 * (unsigned)k.W), h = (int)(t1 - t2 * (unsigned)k.H);
    const int b = (int)(t2 / (unsigned)k.D), d = (int)(t2 - (unsigned)b * (unsigned)k.D);
    float v = 0.f;
    if (c < k.Cl) {
      if (c < k.C0) {
        if (k.pool2) {
          const float* q = k.s0 + b * k.s0b + d * k.s0d + (2 * h) * k.s0h + (2 * w) * k.s0w + c;
          v = 0.25f * ((q[0] + q[k.s0w]) + (q[k.s0h] + q[k.s0h + k.s0w]));
        } else {
          v = k.s0[b * k.s0b + d * k.s0d + (long)(h >> k.upshift) * k.s0h + (long)(w >> k.upshift) * k.s0w + c];
        }
        if (k.norm == NORM_AFFINE_C) {
          v = v * k.scale[c] + k.shift[c];
        } else if (k.norm == NORM_STATS_BC) {
          v = (v - k.mean[b * k.C0 + c]) * k.rstd[b * k.C0 + c];
          if (k.scale) v = v * k.scale[c] + k.shift[c];
        }
        if (k.gb) {
          const float* g = k.gb + pix * (2L * k.C0) + (c >> 4) * 32 + (c & 15);   // [gamma x16 | beta x16] chunks
          v = v * (1.f + g[0]) + g[16];
        }
      } else {
        v = k.s1[b * k.s1b + d * k.s1d + h * k.s1h + w * k.s1w + (c - k.C0)];
      }
      if (k.add) v += k.add[b * k.ab + d * k.ad + h * k.ah + w * k.aw + c];
      v = apply_act(v, k.act, k.slope);
    }
    if (k.o32 && c < k.Cl) k.o32[b * k.ob + d * k.od + h * k.oh + w * k.ow + c] = v;
    if (k.opl) {
      long o = pix * k.prow + (c >> 5) * 64 + (c & 31);
      split_operand(v * k.amul, k.opl[o], k.opl[o + 32]);
    }
  }
}

static PrepK make_prepk(const Prep& p) {
  PrepK k{};
  k.s0 = p.src0.p; k.s0b = p.src0.sb; k.s0d = p.src0.sd; k.s0h = p.src0.sh; k.s0w = p.src0.sw; k.C0 = p.src0.C;
  k.s1 = p.src1.p; k.s1b = p.src1.sb; k.s1d = p.src1.sd; k.s1h = p.src1.sh; k.s1w = p.src1.sw;
  k.C1 = p.src1.p ? p.src1.C : 0;
  k.upshift = p.upshift; k.pool2 = p.pool2; k.norm = p.norm; k.act = p.act; k.slope = p.slope;
  k.scale = p.scale; k.shift = p.shift; k.mean = p.mean; k.rstd = p.rstd; k.gb = p.gb;
  k.add = p.add.p; k.ab = p.add.sb; k.ad = p.add.sd; k.ah = p.add.sh; k.aw = p.add.sw;
  k.Cl = k.C0 + k.C1;
  return k;
}

// 4 channels per thread: float4 loads, float4 / 8-byte bf16x4 stores. Requires every channel count,
// stride and base pointer involved to be a multiple of 4 elements (checked by prep_vec_ok).
__global__ void __launch_bounds__(256) prep_kernel_v4(PrepK k) {
  const int C4 = k.Cout >> 2;
  const long total = (long)k.B * k.D * k.H * k.W * C4;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const unsigned ui = (unsigned)idx, pix = ui / (unsigned)C4;      // 32-bit divisions (total < 2^32: host check)
    const int c = (int)(ui - pix * (unsigned)C4) * 4;
    const unsigned t1 = pix / (unsigned)k.W, t2 = t1 / (unsigned)k.H;
    const int w = (int)(pix - t1 * (unsigned)k.W), h = (int)(t1 - t2 * (unsigned)k.H);
    const int b = (int)(t2 / (unsigned)k.D), d = (int)(t2 - (unsigned)b * (unsigned)k.D);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < k.Cl) {
      if (c < k.C0) {
        if (k.pool2) {
          const float* q = k.s0 + b * k.s0b + d * k.s0d + (2 * h) * k.s0h + (2 * w) * k.s0w + c;
          const float4 a0 = *reinterpret_cast<const float4*>(q), a1 = *reinterpret_cast<const float4*>(q + k.s0w);
          const float4 a2 = *reinterpret_cast<const float4*>(q + k.s0h), a3 = *reinterpret_cast<const float4*>(q + k.s0h + k.s0w);
          v.x = 0.25f * ((a0.x + a1.x) + (a2.x + a3.x)); v.y = 0.25f * ((a0.y + a1.y) + (a2.y + a3.y));
          v.z = 0.25f * ((a0.z + a1.z) + (a2.z + a3.z)); v.w = 0.25f * ((a0.w + a1.w) + (a2.w + a3.w));
        } else {
          v = *reinterpret_cast<const float4*>(k.s0 + b * k.s0b + d * k.s0d + (long)(h >> k.upshift) * k.s0h +
                                               (long)(w >> k.upshift) * k.s0w + c);
        }
        if (k.norm == NORM_AFFINE_C) {
          const float4 sc = *reinterpret_cast<const float4*>(k.scale + c), sh = *reinterpret_cast<const float4*>(k.shift + c);
          v.x = v.x * sc.x + sh.x; v.y = v.y * sc.y + sh.y; v.z = v.z * sc.z + sh.z; v.w = v.w * sc.w + sh.w;
        } else if (k.norm == NORM_STATS_BC) {
          const float4 m = *reinterpret_cast<const float4*>(k.mean + b * k.C0 + c);
          const float4 r = *reinterpret_cast<const float4*>(k.rstd + b * k.C0 + c);
          v.x = (v.x - m.x) * r.x; v.y = (v.y - m.y) * r.y; v.z = (v.z - m.z) * r.z; v.w = (v.w - m.w) * r.w;
          if (k.scale) {
            const float4 sc = *reinterpret_cast<const float4*>(k.scale + c), sh = *reinterpret_cast<const float4*>(k.shift + c);
            v.x = v.x * sc.x + sh.x; v.y = v.y * sc.y + sh.y; v.z = v.z * sc.z + sh.z; v.w = v.w * sc.w + sh.w;
          }
        }
        if (k.gb) {
          const float* g = k.gb + pix * (2L * k.C0) + (c >> 4) * 32 + (c & 15);   // [gamma x16 | beta x16] chunks
          const float4 ga = *reinterpret_cast<const float4*>(g), be = *reinterpret_cast<const float4*>(g + 16);
          v.x = v.x * (1.f + ga.x) + be.x; v.y = v.y * (1.f + ga.y) + be.y;
          v.z = v.z * (1.f + ga.z) + be.z; v.w = v.w * (1.f + ga.w) + be.w;
        }
      } else {
        v = *reinterpret_cast<const float4*>(k.s1 + b * k.s1b + d * k.s1d + h * k.s1h + w * k.s1w + (c - k.C0));
      }
      if (k.add) {
        const float4 a = *reinterpret_cast<const float4*>(k.add + b * k.ab + d * k.ad + h * k.ah + w * k.aw + c);
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
      }
      v.x = apply_act(v.x, k.act, k.slope); v.y = apply_act(v.y, k.act, k.slope);
      v.z = apply_act(v.z, k.act, k.slope); v.w = apply_act(v.w, k.act, k.slope);
      if (c + 3 >= k.Cl) {                       // ragged tail (plain conversions only): the over-read lanes are pad
        if (c + 1 >= k.Cl) v.y = 0.f;
        if (c + 2 >= k.Cl) v.z = 0.f;
        v.w = 0.f;
      }
    }
    if (k.o32 && c < k.Cl) *reinterpret_cast<float4*>(k.o32 + b * k.ob + d * k.od + h * k.oh + w * k.ow + c) = v;
    if (k.opl) {
      const long o = pix * k.prow + (c >> 5) * 64 + (c & 31);
      uint2 hv, lv;
      split_operand4(v.x * k.amul, v.y * k.amul, v.z * k.amul, v.w * k.amul, hv, lv);
      *reinterpret_cast<uint2*>(k.opl + o) = hv;
      *reinterpret_cast<uint2*>(k.opl + o + 32) = lv;
    }
  }
}

// The common shapes of the input transform -- one source, no pooling / upsampling / SPADE modulation: plain split, pre-activation
// BatchNorm, instance / group normalisation (+ residual, + fp32 copy).  The generic kernel is ISSUE-bound (ncu: 66 % issue
// slots, 2.4 TB/s): a thread there decomposes a 64-bit index, walks a run-time activation switch and splits ONE float4.  Here a
// thread owns 16 consecutive channels of a pixel (4 float4 loads in flight, one 32-bit index decomposition per 64 bytes, hi / lo
// halves written as 32-byte runs) and norm / activation / residual are compile-time.
template <int NORM, bool ADD, int ACT>
__global__ void __launch_bounds__(256) prep_kernel_fast(PrepK k) {
  const unsigned G = (unsigned)(k.Cout + 15) >> 4;                         // 16-channel groups per pixel
  const unsigned total = (unsigned)k.B * k.D * k.H * k.W * G;
  const unsigned uW = k.W, uH = k.H, uD = k.D;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const unsigned pix = idx / G;
    const int c0 = (int)(idx - pix * G) * 16;
    unsigned t = pix / uW; const unsigned w = pix - t * uW;
    unsigned t2 = t / uH; const unsigned h = t - t2 * uH;
    const unsigned b = t2 / uD; const unsigned d = t2 - b * uD;
    const float* sp = k.s0 + b * k.s0b + d * k.s0d + h * k.s0h + w * k.s0w + c0;
    float4 v[4], a[ADD ? 4 : 1];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + 4 * q < k.Cl) v[q] = *reinterpret_cast<const float4*>(sp + 4 * q);
    }
    if constexpr (ADD) {
      const float* ap = k.add + b * k.ab + d * k.ad + h * k.ah + w * k.aw + c0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + 4 * q < k.Cl) a[q] = *reinterpret_cast<const float4*>(ap + 4 * q);
      }
    }
    uint2 hv[4], lv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = c0 + 4 * q;
      float4 x = v[q];
      if (c < k.Cl) {
        if constexpr (NORM == NORM_AFFINE_C) {
          const float4 sc = __ldg(reinterpret_cast<const float4*>(k.scale + c)), sh = __ldg(reinterpret_cast<const float4*>(k.shift + c));
          x.x = x.x * sc.x + sh.x; x.y = x.y * sc.y + sh.y; x.z = x.z * sc.z + sh.z; x.w = x.w * sc.w + sh.w;
        } else if constexpr (NORM == NORM_STATS_BC) {
          const float4 m = __ldg(reinterpret_cast<const float4*>(k.mean + b * k.C0 + c));
          const float4 r = __ldg(reinterpret_cast<const float4*>(k.rstd + b * k.C0 + c));
          x.x = (x.x - m.x) * r.x; x.y = (x.y - m.y) * r.y; x.z = (x.z - m.z) * r.z; x.w = (x.w - m.w) * r.w;
          if (k.scale) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(k.scale + c)), sh = __ldg(reinterpret_cast<const float4*>(k.shift + c));
            x.x = x.x * sc.x + sh.x; x.y = x.y * sc.y + sh.y; x.z = x.z * sc.z + sh.z; x.w = x.w * sc.w + sh.w;
          }
        }
        if constexpr (ADD) { x.x += a[q].x; x.y += a[q].y; x.z += a[q].z; x.w += a[q].w; }
        if constexpr (ACT == ACT_RELU) {
          x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
        } else if constexpr (ACT == ACT_LRELU) {
          x.x = x.x > 0.f ? x.x : x.x * k.slope; x.y = x.y > 0.f ? x.y : x.y * k.slope;
          x.z = x.z > 0.f ? x.z : x.z * k.slope; x.w = x.w > 0.f ? x.w : x.w * k.slope;
        }
        if (c + 3 >= k.Cl) {                       // ragged tail (plain conversions only): the over-read lanes are pad
          if (c + 1 >= k.Cl) x.y = 0.f;
          if (c + 2 >= k.Cl) x.z = 0.f;
          x.w = 0.f;
        }
        if (k.o32) *reinterpret_cast<float4*>(k.o32 + b * k.ob + d * k.od + h * k.oh + w * k.ow + c) = x;
      }
      split_operand4(x.x * k.amul, x.y * k.amul, x.z * k.amul, x.w * k.amul, hv[q], lv[q]);
    }
    if (k.opl) {                                   // 16 channels: hi 32 bytes, lo 32 bytes (64 bytes further)
      __nv_bfloat16* o = k.opl + (long)pix * k.prow + (c0 >> 5) * 64 + (c0 & 31);
      *reinterpret_cast<uint4*>(o) = make_uint4(hv[0].x, hv[0].y, hv[1].x, hv[1].y);
      *reinterpret_cast<uint4*>(o + 8) = make_uint4(hv[2].x, hv[2].y, hv[3].x, hv[3].y);
      *reinterpret_cast<uint4*>(o + 32) = make_uint4(lv[0].x, lv[0].y, lv[1].x, lv[1].y);
      *reinterpret_cast<uint4*>(o + 40) = make_uint4(lv[2].x, lv[2].y, lv[3].x, lv[3].y);
    }
  }
}

template <int NORM, bool ADD>
static void launch_prep_fast(const PrepK& k, unsigned blocks, cudaStream_t st) {
  if (k.act == ACT_RELU) prep_kernel_fast<NORM, ADD, ACT_RELU><<<blocks, 256, 0, st>>>(k);
  else if (k.act == ACT_LRELU) prep_kernel_fast<NORM, ADD, ACT_LRELU><<<blocks, 256, 0, st>>>(k);
  else prep_kernel_fast<NORM, ADD, ACT_NONE><<<blocks, 256, 0, st>>>(k);
}

static bool prep_vec_ok(const PrepK& k) {
  auto a4 = [](long v) { return (v & 3) == 0; };
  auto p16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if (!a4(k.Cout)) return false;
  if (!(a4(k.C0) && a4(k.C1))) {
    // ragged channel count: only a plain single-source conversion into operand planes, reading the (finite or not)
    // pad floats of a row whose stride covers the rounded-up channel count
    const bool plain = !k.s1 && k.norm == NORM_NONE && !k.gb && !k.add && !k.pool2 && !k.o32;
    if (!(plain && k.s0w >= ((k.C0 + 3) & ~3))) return false;
  }
  if (!(a4(k.s0b) && a4(k.s0d) && a4(k.s0h) && a4(k.s0w) && p16(k.s0))) return false;
  if (k.s1 && !(a4(k.s1b) && a4(k.s1d) && a4(k.s1h) && a4(k.s1w) && p16(k.s1))) return false;
  if (k.add && !(a4(k.ab) && a4(k.ad) && a4(k.ah) && a4(k.aw) && p16(k.add))) return false;
  if (k.o32 && !(a4(k.ob) && a4(k.od) && a4(k.oh) && a4(k.ow) && p16(k.o32))) return false;
  if (!(p16(k.scale) && p16(k.shift) && p16(k.mean) && p16(k.rstd) && p16(k.gb) && p16(k.opl))) return false;
  return true;
}

static void launch_prep(const Launcher& L, PrepK& k) {
  L.count();
  if (L.dry) return;
  const long total = (long)k.B * k.D * k.H * k.W * k.Cout;
  // algorithmic bytes: every logical element read once (fp32) and written once (fp32 or hi+lo bf16)
  ProfScope ps(L, PK_PREP, 0.0, (double)k.B * k.D * k.H * k.W * k.Cl * 4.0 * 2.0, "prep");
  // (operand outputs have Cout % 32 == 0; fp32-only outputs need whole 16-channel groups to keep the stores inside the row)
  const bool fast = prep_vec_ok(k) && !k.s1 && !k.pool2 && !k.upshift && !k.gb && total / 4 < (1L << 30) &&
                    (k.act == ACT_NONE || k.act == ACT_RELU || k.act == ACT_LRELU) && (k.opl ? (k.Cout & 31) == 0 : (k.Cout & 15) == 0) &&
                    (k.opl == nullptr || ((uintptr_t)k.opl & 15) == 0);
  if (fast) {
    long blocks = (total / 16 + 255) / 256;
    if (blocks > 148L * 16) blocks = 148L * 16;
    if (blocks < 1) blocks = 1;
    if (k.add) {
      if (k.norm == NORM_STATS_BC) launch_prep_fast<NORM_STATS_BC, true>(k, (unsigned)blocks, L.stream);
      else if (k.norm == NORM_AFFINE_C) launch_prep_fast<NORM_AFFINE_C, true>(k, (unsigned)blocks, L.stream);
      else launch_prep_fast<NORM_NONE, true>(k, (unsigned)blocks, L.stream);
    } else {
      if (k.norm == NORM_STATS_BC) launch_prep_fast<NORM_STATS_BC, false>(k, (unsigned)blocks, L.stream);
      else if (k.norm == NORM_AFFINE_C) launch_prep_fast<NORM_AFFINE_C, false>(k, (unsigned)blocks, L.stream);
      else launch_prep_fast<NORM_NONE, false>(k, (unsigned)blocks, L.stream);
    }
  } else if (prep_vec_ok(k)) {
    long blocks = (total / 4 + 255) / 256;
    if (blocks > 148L * 16) blocks = 148L * 16;
    CS_REQUIRE(total / 4 < (1L << 31), -1, "prep: tensor too large for the 32-bit index math of prep_kernel_v4");
    prep_kernel_v4<<<(unsigned)blocks, 256, 0, L.stream>>>(k);
  } else {
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 32) blocks = 148L * 32;
    CS_REQUIRE(total < (1L << 31), -1, "prep: tensor too large for the 32-bit index math of prep_kernel");
    prep_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(k);
  }
  check_launch("prep");
}

void prep_f32(const Launcher& L, const Prep& p, Act out) {
  PrepK k = make_prepk(p);
  CS_REQUIRE(out.C == k.Cl, -1, "prep_f32: channel mismatch");
  k.B = out.B; k.D = out.D; k.H = out.H; k.W = out.W; k.Cout = k.Cl;
  k.o32 = out.p; k.ob = out.sb; k.od = out.sd; k.oh = out.sh; k.ow = out.sw;
  launch_prep(L, k);
}

void prep_planes(const Launcher& L, const Prep& p, Opd out, const Act* out32) {
  PrepK k = make_prepk(p);
  CS_REQUIRE(out.nblk * 32 >= k.Cl, -1, "prep_planes: padded channels too small");
  k.B = out.B; k.D = out.D; k.H = out.H; k.W = out.W; k.Cout = out.nblk * 32;
  k.opl = out.p; k.prow = out.row(); k.amul = out.amul;
  if (out32) {
    CS_REQUIRE(out32->C == k.Cl, -1, "prep_planes: fp32 copy channel mismatch");
    k.o32 = out32->p; k.ob = out32->sb; k.od = out32->sd; k.oh = out32->sh; k.ow = out32->sw;
  }
  launch_prep(L, k);
}

// ------------------------------------------------------------------------------------------
// instance statistics: scratch[b][c][2] double accumulators, then finalize
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stats_partial_kernel(const float* __restrict__ x, long sb, long sd, long sh, long sw,
                                                           int D, int H, int W, int C, int CT, int chunk,
                                                           double* __restrict__ acc) {
  __shared__ float red[2][256];
  int b = blockIdx.y;
  long S = (long)D * H * W;
  long p0 = (long)blockIdx.x * chunk;
  long p1 = p0 + chunk < S ? p0 + chunk : S;
  int rows = 256 / CT;
  int r = threadIdx.x / CT, cl = threadIdx.x % CT;
  for (int c = cl; c < C; c += CT) {       // uniform trip count across the block (CT divides into C rounds)
    float s1 = 0.f, s2 = 0.f;
    for (long p = p0 + r; p < p1; p += rows) {
      int w = (int)(p % W); long t = p / W; int h = (int)(t % H); int d = (int)(t / H);
      float v = x[b * sb + d * sd + h * sh + w * sw + c];
      s1 += v; s2 += v * v;
    }
    red[0][threadIdx.x] = s1; red[1][threadIdx.x] = s2;
    __syncthreads();
    if (r == 0) {
      double a1 = 0.0, a2 = 0.0;
      for (int i = 0; i < rows; ++i) { a1 += red[0][i * CT + cl]; a2 += red[1][i * CT + cl]; }
      atomicAdd(&acc[((long)b * C + c) * 2 + 0], a1);
      atomicAdd(&acc[((long)b * C + c) * 2 + 1], a2);
    }
    __syncthreads();
  }
}

__global__ void stats_finalize_kernel(const double* __restrict__ acc, float* __restrict__ mean, float* __restrict__ rstd,
                                      int n, double inv_count, float eps) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double m = acc[2 * i] * inv_count;
  double var = acc[2 * i + 1] * inv_count - m * m;
  if (var < 0.0) var = 0.0;
  mean[i] = (float)m;
  rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

// dense [B,S,C] fp32 (C a multiple of 4, C <= 1024): a block takes `rows` consecutive positions, thread = (row group,
// float4 of channels); fp32 partials over <= rows/groups positions, smem reduce over row groups in a fixed order, one fp64
// partial per (sample, block, channel) -- summed in block order by stats_finalize_blocks_kernel.  No atomics: the result is
// bit-reproducible and, because the block size depends on S only, independent of the batch the sample is part of.
__global__ void __launch_bounds__(256) stats_dense_kernel(const float* __restrict__ x, long S, int C, int rows,
                                                          double* __restrict__ part /*[B][gridDim.x][C][2]*/) {
  __shared__ float4 red1[256], red2[256];
  const int b = blockIdx.y;
  const int C4 = C >> 2;
  const int groups = 256 / C4;                   // C4 is a power of two <= 256
  const int g = threadIdx.x / C4, c4 = threadIdx.x % C4;
  const long p0 = (long)blockIdx.x * rows;
  const long p1 = p0 + rows < S ? p0 + rows : S;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  const float4* xp = reinterpret_cast<const float4*>(x + (long)b * S * C) + c4;
  for (long p = p0 + g; p < p1; p += groups) {
    const float4 v = __ldg(xp + p * C4);
    s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
    s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y); s2.z = fmaf(v.z, v.z, s2.z); s2.w = fmaf(v.w, v.w, s2.w);
  }
  red1[threadIdx.x] = s1; red2[threadIdx.x] = s2;
  __syncthreads();
  if (g == 0) {
    double a[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    for (int i = 0; i < groups; ++i) {
      const float4 u = red1[i * C4 + c4], w = red2[i * C4 + c4];
      a[0] += u.x; a[1] += u.y; a[2] += u.z; a[3] += u.w;
      q[0] += w.x; q[1] += w.y; q[2] += w.z; q[3] += w.w;
    }
    double* o = part + (((long)b * gridDim.x + blockIdx.x) * C + c4 * 4) * 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[2 * j] = a[j]; o[2 * j + 1] = q[j]; }
  }
}

// block = 32 (b,c) pairs x 8 slices of the block range; every slice and the final combine run in a fixed order
__global__ void __launch_bounds__(256) stats_finalize_blocks_kernel(const double* __restrict__ part, int nblocks, int C,
                                                                    float* __restrict__ mean, float* __restrict__ rstd, int n /*B*C*/,
                                                                    double inv_count, float eps) {
  __shared__ double r1[8][32], r2[8][32];
  const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  double s1 = 0.0, s2 = 0.0;
  if (i < n) {
    const int b = i / C, c = i % C;
    const double* p = part + ((long)b * nblocks * C + c) * 2;
    for (int t = sl; t < nblocks; t += 8) { s1 += p[(long)t * C * 2]; s2 += p[(long)t * C * 2 + 1]; }
  }
  r1[sl][lane] = s1; r2[sl][lane] = s2;
  __syncthreads();
  if (sl == 0 && i < n) {
    s1 = 0.0; s2 = 0.0;
    for (int q = 0; q < 8; ++q) { s1 += r1[q][lane]; s2 += r2[q][lane]; }
    const double m = s1 * inv_count;
    double var = s2 * inv_count - m * m;
    if (var < 0.0) var = 0.0;
    mean[i] = (float)m;
    rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// mean / rstd from per-block partial sums [B][nblocks][C][2] (doubles) written by a producer kernel (instance_stats' own
// stats_dense_kernel, or the Winograd output transform of the conv that produced the tensor)
void stats_finalize_blocks(const Launcher& L, const double* part, int nblocks, int B, int C, long S, float* mean, float* rstd, float eps) {
  L.count();
  if (L.dry) return;
  const int n = B * C;
  stats_finalize_blocks_kernel<<<(n + 31) / 32, 256, 0, L.stream>>>(part, nblocks, C, mean, rstd, n, 1.0 / (double)S, eps);
  check_launch("stats_finalize");
}

void instance_stats(const Launcher& L, const Act& x, float* mean, float* rstd, float eps, double* scratch) {
  L.count(); L.count();
  if (L.dry) return;
  int C = x.C;
  long S = (long)x.D * x.H * x.W;
  ProfScope ps(L, PK_STATS, 0.0, (double)x.B * S * C * 4.0, "stats");
  const int C4 = C >> 2;
  const bool pow2 = C4 > 0 && (C4 & (C4 - 1)) == 0;
  const int n = x.B * C;
  if ((C & 3) == 0 && pow2 && C4 <= 256 && x.sb == S * C && ((uintptr_t)x.p & 15) == 0) {
    // the tensor is dense per sample with the channel fastest (2-D activations and the [h,w,16,32] volume alike)
    const int groups = 256 / C4;
    long rows = (S + STATS_MAX_BLOCKS - 1) / STATS_MAX_BLOCKS;   // <= STATS_MAX_BLOCKS blocks per sample, a function of S only
    if (rows < groups) rows = groups;
    const long nblocks = (S + rows - 1) / rows;
    CS_REQUIRE(nblocks <= STATS_MAX_BLOCKS && C <= 512, -1, "instance_stats: scratch too small");
    dim3 grid((unsigned)nblocks, x.B);
    stats_dense_kernel<<<grid, 256, 0, L.stream>>>(x.p, S, C, (int)rows, scratch);
    check_launch("stats_dense");
    stats_finalize_blocks_kernel<<<(n + 31) / 32, 256, 0, L.stream>>>(scratch, (int)nblocks, C, mean, rstd, n, 1.0 / (double)S, eps);
    check_launch("stats_finalize");
    return;
  }
  CS_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * x.B * C, L.stream));
  {
    int CT = 1;
    while (CT * 2 <= C && CT * 2 <= 256) CT *= 2;
    // when C is not a multiple of CT the strided loop `c = cl; c < C; c += CT` has a non-uniform trip
    // count, which would break the __syncthreads above; require exact division.
    CS_REQUIRE(C % CT == 0, -1, "instance_stats: C must be a power of two multiple");
    int chunk = 2048;
    int nchunks = (int)((S + chunk - 1) / chunk);
    dim3 grid(nchunks, x.B);
    stats_partial_kernel<<<grid, 256, 0, L.stream>>>(x.p, x.sb, x.sd, x.sh, x.sw, x.D, x.H, x.W, C, CT, chunk, scratch);
    check_launch("stats_partial");
  }
  stats_finalize_kernel<<<(n + 127) / 128, 128, 0, L.stream>>>(scratch, mean, rstd, n, 1.0 / (double)S, eps);
  check_launch("stats_finalize");
}

// ------------------------------------------------------------------------------------------
// adaptive blend: o2 = [out_std(512) | out_mod(512)] per pixel
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adaptive_blend_kernel(const float4* __restrict__ o2, const float* __restrict__ mask,
                                                            const float4* residual, int relu, float4* y,
                                                            __nv_bfloat16* __restrict__ opl, long P, float opl_mul) {
  long total = P * 128;     // 512 channels / 4
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long pix = i >> 7; int q = (int)(i & 127);
    float m = mask[pix];
    float4 s = o2[pix * 256 + q], mo = o2[pix * 256 + 128 + q];
    float4 v;
    v.x = m * mo.x + (1.f - m) * s.x; v.y = m * mo.y + (1.f - m) * s.y;
    v.z = m * mo.z + (1.f - m) * s.z; v.w = m * mo.w + (1.f - m) * s.w;
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (residual) { float4 r = residual[i]; v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
    if (y) y[i] = v;
    if (opl) {                // split-fp16 operand of the next conv: [pix][16 blocks][hi 32 | lo 32]
      const int c = q * 4;
      uint2 hv, lv;
      split_operand4(v.x * opl_mul, v.y * opl_mul, v.z * opl_mul, v.w * opl_mul, hv, lv);
      __nv_bfloat16* o = opl + pix * 1024 + (c >> 5) * 64 + (c & 31);
      *reinterpret_cast<uint2*>(o) = hv;
      *reinterpret_cast<uint2*>(o + 32) = lv;
    }
  }
}

void adaptive_blend(const Launcher& L, const float* o2, const float* mask, const float* residual, int relu, float* y,
                    __nv_bfloat16* opl, long P, float opl_mul) {
  L.count();
  if (L.dry) return;
  long blocks = (P * 128 + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  ProfScope ps(L, PK_OTHER, 0.0, (double)P * (1024 + 512 + 512) * 4.0, "blend");
  adaptive_blend_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(reinterpret_cast<const float4*>(o2), mask,
                                                               reinterpret_cast<const float4*>(residual), relu,
                                                               reinterpret_cast<float4*>(y), opl, P, opl_mul);
  check_launch("adaptive_blend");
}

// ------------------------------------------------------------------------------------------
// calibration (cs_calibrate): max |v| over a split operand, from the hi halves (|lo| <= 2^-11 |hi|), scale divided out
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) operand_absmax_kernel(const __half* __restrict__ p, long nrows /*64-element rows*/, long rstride,
                                                             int nblk, float inv_amul, unsigned* __restrict__ slot) {
  float mx = 0.f;
  const long total = nrows * 32;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i >> 5; const int j = (int)(i & 31);
    const long pix = r / nblk; const int blk = (int)(r % nblk);
    const float v = fabsf(__half2float(p[pix * rstride + blk * 64 + j]));
    if (v > mx && v <= 65504.f) mx = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(slot, __float_as_uint(mx * inv_amul));
}

void operand_absmax(const Launcher& L, const Opd& x, int id) {
  if (L.dry || !L.calib || id < 0) return;
  const long pixels = (long)x.B * x.D * x.H * x.W;
  long blocks = (pixels * x.nblk * 32 + 255) / 256; if (blocks > 148L * 8) blocks = 148L * 8;
  operand_absmax_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(reinterpret_cast<const __half*>(x.p), pixels * x.nblk, x.row(), x.nblk,
                                                               1.0f / x.amul, L.calib + id);
  check_launch("operand_absmax");
}

__global__ void avg2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = (a[i] + b[i]) / 2.f;
}

void avg2(const Launcher& L, const float* a, const float* b, float* y, long n) {
  L.count();
  if (L.dry) return;
  long blocks = (n + 255) / 256; if (blocks > 148L * 8) blocks = 148L * 8;
  avg2_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(a, b, y, n);
  check_launch("avg2");
}

// ------------------------------------------------------------------------------------------
// layout shims: [B,C,S] <-> [B,S,C'] with the optional volume permutation ch=c*16+d <-> ch'=d*32+c
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int vol_perm_fwd(int ch) { return (ch & 15) * 32 + (ch >> 4); }   // c*16+d -> d*32+c

__global__ void __launch_bounds__(256) nchw_to_cl_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                        int C, long S, int vol_perm, int Cstride) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  long s0 = (long)blockIdx.x * 32; int c0 = blockIdx.y * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    int c = c0 + i; long s = s0 + tx;
    tile[i][tx] = (c < C && s < S) ? src[((long)b * C + c) * S + s] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    long s = s0 + i; int c = c0 + tx;
    if (c < C && s < S) {
      int cc = vol_perm ? vol_perm_fwd(c) : c;
      dst[((long)b * S + s) * Cstride + cc] = tile[tx][i];
    }
  }
}

__global__ void __launch_bounds__(256) cl_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                        int C, long S, int vol_perm, int Cstride) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  long s0 = (long)blockIdx.x * 32; int c0 = blockIdx.y * 32;    // c0 indexes the channels-last (internal) channel
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    long s = s0 + i; int c = c0 + tx;
    tile[i][tx] = (c < C && s < S) ? src[((long)b * S + s) * Cstride + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int ci = c0 + i; long s = s0 + tx;
    if (ci < C && s < S) {
      // internal channel ci = d*32+c  ->  reference channel c*16+d
      int cr = vol_perm ? ((ci & 31) * 16 + (ci >> 5)) : ci;
      dst[((long)b * C + cr) * S + s] = tile[tx][i];
    }
  }
}

void nchw_to_cl(const Launcher& L, const float* src, float* dst, int B, int C, long S, int vol_perm) {
  L.count();
  if (L.dry) return;
  CS_REQUIRE(!vol_perm || C == 512, -1, "vol_perm needs 512 channels");
  dim3 grid((unsigned)((S + 31) / 32), (C + 31) / 32, B);
  nchw_to_cl_kernel<<<grid, 256, 0, L.stream>>>(src, dst, C, S, vol_perm, C);
  check_launch("nchw_to_cl");
}

void cl_to_nchw(const Launcher& L, const float* src, float* dst, int B, int C, long S, int vol_perm, int Cstride) {
  L.count();
  if (L.dry) return;
  CS_REQUIRE(!vol_perm || C == 512, -1, "vol_perm needs 512 channels");
  dim3 grid((unsigned)((S + 31) / 32), (C + 31) / 32, B);
  cl_to_nchw_kernel<<<grid, 256, 0, L.stream>>>(src, dst, C, S, vol_perm, Cstride);
  check_launch("cl_to_nchw");
}

// ------------------------------------------------------------------------------------------
// ingest / emit
// ------------------------------------------------------------------------------------------
__global__ void ingest_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    dst[i] = (float)src[i] / 255.f;      // astype(float32) / 255. ; clip is a no-op on u8 input
}

void ingest_u8(const Launcher& L, const uint8_t* src, float* dst, long n) {
  L.count();
  if (L.dry) return;
  long blocks = (n + 255) / 256; if (blocks > 148L * 8) blocks = 148L * 8;
  ingest_u8_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(src, dst, n);
  check_launch("ingest_u8");
}

__global__ void __launch_bounds__(256) emit_image_kernel(const float* __restrict__ y, int Cs, float* __restrict__ img,
                                                        uint8_t* __restrict__ u8, int B, int H, int W) {
  int OH = 2 * H, OW = 2 * W;
  long total = (long)B * OH * OW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int ow = (int)(i % OW); long t = i / OW; int oh = (int)(t % OH); int b = (int)(t / OH);
    const float* q = y + (((long)b * H + (oh >> 1)) * W + (ow >> 1)) * Cs + (oh & 1) * 2 + (ow & 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = 1.f / (1.f + expf(-q[c * 4]));          // PixelShuffle(2): channel c*4 + i*2 + j
      if (img) img[(((long)b * 3 + c) * OH + oh) * OW + ow] = v;
      if (u8) {
        float s = fminf(fmaxf(v, 0.f), 1.f) * 255.f;
        s = fminf(fmaxf(s, 0.f), 255.f);
        u8[i * 3 + c] = (uint8_t)s;                      // truncation, can_swap_e2e.py:320
      }
    }
  }
}

void emit_image(const Launcher& L, const float* y, int Cs, float* img, uint8_t* u8, int B, int H, int W) {
  L.count();
  if (L.dry) return;
  long total = (long)B * 4 * H * W;
  long blocks = (total + 255) / 256; if (blocks > 148L * 16) blocks = 148L * 16;
  ProfScope ps(L, PK_OTHER, 0.0, (double)total * 3 * (4.0 + (img ? 4.0 : 0.0) + (u8 ? 1.0 : 0.0)), "emit_image");
  emit_image_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(y, Cs, img, u8, B, H, W);
  check_launch("emit_image");
}

}  // namespace cs
