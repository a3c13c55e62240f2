// Winograd F(2x2, 3x3) form of the adaptive convs of transfer_model2 (reference adaptive_modulate.py:128-193), the largest
// MMA consumer of the path (28 convs 512 -> 512 per frame, both branches: 24.6 % of the FLOPs).
//
//   Y = A^T [ (G g G^T) .* (B^T d B) ] A          (Lavin & Gray; cross-correlation, 2x2 outputs from a 4x4 input patch)
//
// The 16 element-wise products are 16 independent GEMMs over the input channels:
//   M_k[tile, co] = sum_ci V_k[tile, ci] * U_k[ci, co],   k = 0..15,  tile = 2x2 output block
// i.e. a 1x1x1 "conv" over a tensor whose depth axis is the component k, with depth-dependent weights -- the zrows mode of
// conv_tc_kernel (one depth per M tile, B rows k * Cout .. of the packed weights).  K = 512 channels = 16 pipeline
// iterations: one accumulator, double-buffered TMEM, pair mode.  4 MACs per output instead of 9 (2.25x fewer MMAs); the
// price is the 4x larger fp32 intermediate M (written by the GEMM epilogue, read once by the output transform).
//
//   wino_weights   U = G g G^T of the per-identity combined weights [W | W*s*demod], fp32 [ci][k*1024 + co]
//   wino_in        V = B^T d B of the fp32 activation (zero padding), written as the split-fp16 operand [B,16,H/2,W/2,C]
//   wino_out_blend Y = A^T M A of both branches + bias, mask blend, ReLU / residual (adaptive_modulate.py:186, :337-349)
//
// Numerics (CPU experiment, 512 -> 512 at 64^2, fp32): transforms alone (exact accumulation) 4.8e-7 max / 6e-8 rms, against
// 1.1e-6 / 1.3e-7 of a direct fp32 conv -- the transforms add and subtract fp32 values, the weights are transformed
// from the fp32 master before the fp16 split, and the MMA chain per accumulator shrinks from 288 to 96.
#include "tc_ptx.cuh"
#include <memory>

namespace cs {

namespace {

// G (4x3) rows: [1,0,0], [.5,.5,.5], [.5,-.5,.5], [0,0,1]
__device__ __forceinline__ void g_rows(float a, float b, float c, float* o) {
  o[0] = a; o[1] = 0.5f * ((a + b) + c); o[2] = 0.5f * ((a - b) + c); o[3] = c;
}

// w32 [9][Cin][Cout] (tap = kh*3 + kw) -> U [Cin][16 * Cout], column k * Cout + co, k = i*4 + l, scaled by `mul`
__global__ void __launch_bounds__(256) wino_weights_kernel(const float* __restrict__ w, float* __restrict__ U, int Cin, int Cout,
                                                           float mul) {
  const long total = (long)Cin * Cout;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int co = (int)(idx % Cout); const int ci = (int)(idx / Cout);
    float g[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) g[a][b] = w[((long)(a * 3 + b) * Cin + ci) * Cout + co] * mul;
    float t[4][3];                       // G g
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      float col[4];
      g_rows(g[0][b], g[1][b], g[2][b], col);
#pragma unroll
      for (int i = 0; i < 4; ++i) t[i][b] = col[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {        // (G g) G^T
      float row[4];
      g_rows(t[i][0], t[i][1], t[i][2], row);
#pragma unroll
      for (int l = 0; l < 4; ++l) U[(long)ci * (16L * Cout) + (long)(i * 4 + l) * Cout + co] = row[l];
    }
  }
}

// B^T rows: [1,0,-1,0], [0,1,1,0], [0,-1,1,0], [0,1,0,-1]
__device__ __forceinline__ void bt4(const float4& a, const float4& b, const float4& c, const float4& d, float4* o) {
  o[0] = make_float4(a.x - c.x, a.y - c.y, a.z - c.z, a.w - c.w);
  o[1] = make_float4(b.x + c.x, b.y + c.y, b.z + c.z, b.w + c.w);
  o[2] = make_float4(c.x - b.x, c.y - b.y, c.z - b.z, c.w - b.w);
  o[3] = make_float4(b.x - d.x, b.y - d.y, b.z - d.z, b.w - d.w);
}

// x [B,H,W,C] fp32 dense -> V operand [B,16,H/2,W/2,C/32,64].  One block of C/4 threads per 2x2 output tile (thread = 4
// channels).  MASK: the 3x3 neighbourhoods of the tile's four pixels are exactly the 4x4 patch held in registers, so the
// 512 -> 1 mask conv of AdaptiveSharedWeightConv2d (adaptive_modulate.py:173-180) is computed here as well: per-thread
// partial dot products, shuffle + shared-memory reduction over the block (fixed order: deterministic), sigmoid.
template <bool MASK, int MINB, bool PRE>
__global__ void __launch_bounds__(128, MINB) wino_in_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ V, int B, int H, int W,
                                                      int C, const float* __restrict__ mw /*[9][C]*/, const float* __restrict__ mb,
                                                      float* __restrict__ mask /*[B,H,W]*/, const float* __restrict__ pscale,
                                                      const float* __restrict__ pshift, int pact, float pslope, float amul) {
  __shared__ float red[8][4];
  const int Ht = H >> 1, Wt = W >> 1;
  const long tiles = (long)B * Ht * Wt;
  const long plane = (long)Ht * Wt * (C >> 5) * 64;               // elements per (b, component)
  const int c = threadIdx.x * 4;
  // optional input transform act(x * scale[c] + shift[c]) (pre-activation BatchNorm of a ResBlock2d): the conv's zero padding
  // pads the TRANSFORMED tensor, so out-of-bounds elements stay zero
  float4 psc = make_float4(1.f, 1.f, 1.f, 1.f), psh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pscale) { psc = __ldg(reinterpret_cast<const float4*>(pscale + c)); psh = __ldg(reinterpret_cast<const float4*>(pshift + c)); }
  const float ps = leaky_slope(pact, pslope);
  for (long t0 = blockIdx.x; t0 < tiles; t0 += gridDim.x) {
    const unsigned t = (unsigned)t0, tq = t / (unsigned)Wt;          // 32-bit divisions (tile counts are checked on the host)
    const int tx = (int)(t - tq * (unsigned)Wt);
    const int b = (int)(tq / (unsigned)Ht), ty = (int)(tq - (unsigned)b * (unsigned)Ht);
    float4 d[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ih = 2 * ty - 1 + r;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int iw = 2 * tx - 1 + q;
        const bool in = ih >= 0 && ih < H && iw >= 0 && iw < W;
        d[r][q] = in ? *reinterpret_cast<const float4*>(x + (((long)b * H + ih) * W + iw) * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (PRE && in) {
          d[r][q].x = apply_leaky(fmaf(d[r][q].x, psc.x, psh.x), ps); d[r][q].y = apply_leaky(fmaf(d[r][q].y, psc.y, psh.y), ps);
          d[r][q].z = apply_leaky(fmaf(d[r][q].z, psc.z, psh.z), ps); d[r][q].w = apply_leaky(fmaf(d[r][q].w, psc.w, psh.w), ps);
        }
      }
    }
    if constexpr (MASK) {
      float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int e = 0; e < 3; ++e) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(mw + (a * 3 + e) * C + c));
#pragma unroll
          for (int oy = 0; oy < 2; ++oy)
#pragma unroll
            for (int ox = 0; ox < 2; ++ox) {
              const float4 v = d[oy + a][ox + e];
              acc[oy][ox] = fmaf(v.x, wv.x, fmaf(v.y, wv.y, fmaf(v.z, wv.z, fmaf(v.w, wv.w, acc[oy][ox]))));
            }
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        acc[0][0] += __shfl_xor_sync(0xffffffffu, acc[0][0], o); acc[0][1] += __shfl_xor_sync(0xffffffffu, acc[0][1], o);
        acc[1][0] += __shfl_xor_sync(0xffffffffu, acc[1][0], o); acc[1][1] += __shfl_xor_sync(0xffffffffu, acc[1][1], o);
      }
      __syncthreads();                     // the previous tile's readers are done with `red`
      if ((threadIdx.x & 31) == 0) {
        float* rr = red[threadIdx.x >> 5];
        rr[0] = acc[0][0]; rr[1] = acc[0][1]; rr[2] = acc[1][0]; rr[3] = acc[1][1];
      }
      __syncthreads();
      if (threadIdx.x < 4) {
        float sum = 0.f;
        for (int wq = 0; wq < (int)(blockDim.x >> 5); ++wq) sum += red[wq][threadIdx.x];
        const int oy = threadIdx.x >> 1, ox = threadIdx.x & 1;
        mask[((long)b * H + 2 * ty + oy) * W + 2 * tx + ox] = 1.f / (1.f + expf(-(sum + mb[0])));
      }
    }
    // B^T d : transform along rows, per column, IN PLACE (d is dead after the mask conv): halves the live registers
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 o[4];
      bt4(d[0][q], d[1][q], d[2][q], d[3][q], o);
#pragma unroll
      for (int i = 0; i < 4; ++i) d[i][q] = o[i];
    }
    __nv_bfloat16* base = V + (long)b * 16 * plane + ((long)ty * Wt + tx) * ((C >> 5) * 64) + (c >> 5) * 64 + (c & 31);
#pragma unroll
    for (int i = 0; i < 4; ++i) {        // (B^T d) B, row by row, stored as soon as it is formed
      float4 o[4];
      bt4(d[i][0], d[i][1], d[i][2], d[i][3], o);
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        uint2 hv, lv;
        split_operand4(o[l].x * amul, o[l].y * amul, o[l].z * amul, o[l].w * amul, hv, lv);
        __nv_bfloat16* p = base + (long)(i * 4 + l) * plane;
        *reinterpret_cast<uint2*>(p) = hv;
        *reinterpret_cast<uint2*>(p + 32) = lv;
      }
    }
  }
}

// A^T rows: [1,1,1,0], [0,1,-1,-1]
__device__ __forceinline__ void at4(const float4& a, const float4& b, const float4& c, const float4& d, float4& o0, float4& o1) {
  o0 = make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w);
  o1 = make_float4((b.x - c.x) - d.x, (b.y - c.y) - d.y, (b.z - c.z) - d.z, (b.w - c.w) - d.w);
}

__device__ __forceinline__ void wino_out4(const float* __restrict__ Mt, long plane, float4 (&y)[2][2]) {
  float4 s0[4], s1[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const float4 m0 = *reinterpret_cast<const float4*>(Mt + (long)(0 * 4 + l) * plane);
    const float4 m1 = *reinterpret_cast<const float4*>(Mt + (long)(1 * 4 + l) * plane);
    const float4 m2 = *reinterpret_cast<const float4*>(Mt + (long)(2 * 4 + l) * plane);
    const float4 m3 = *reinterpret_cast<const float4*>(Mt + (long)(3 * 4 + l) * plane);
    at4(m0, m1, m2, m3, s0[l], s1[l]);
  }
  at4(s0[0], s0[1], s0[2], s0[3], y[0][0], y[0][1]);
  at4(s1[0], s1[1], s1[2], s1[3], y[1][0], y[1][1]);
}

// Mt [B,16,H/2,W/2,1024] = [std 512 | mod 512] -> y [B,H,W,512] = mask * (mod + bias) + (1 - mask) * std (+ReLU / +residual)
__global__ void __launch_bounds__(256) wino_out_blend_kernel(const float* __restrict__ Mt, const float* __restrict__ mask,
                                                             const float* __restrict__ bias_mod, const float* residual, int relu,
                                                             float* y, int B, int H, int W) {
  const int Ht = H >> 1, Wt = W >> 1;
  const long total = (long)B * Ht * Wt * 128;
  const long plane = (long)Ht * Wt * 1024;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx & 127) * 4;
    const unsigned t = (unsigned)(idx >> 7), tq = t / (unsigned)Wt;
    const int tx = (int)(t - tq * (unsigned)Wt);
    const int b = (int)(tq / (unsigned)Ht), ty = (int)(tq - (unsigned)b * (unsigned)Ht);
    const float* mp = Mt + (long)b * 16 * plane + ((long)ty * Wt + tx) * 1024 + c;
    float4 ys[2][2], ym[2][2];
    wino_out4(mp, plane, ys);
    wino_out4(mp + 512, plane, ym);
    const float4 bm = __ldg(reinterpret_cast<const float4*>(bias_mod + c));
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const long pix = ((long)b * H + 2 * ty + i) * W + 2 * tx + j;
        const float m = mask[pix];
        const float4 s = ys[i][j];
        const float4 mo = make_float4(ym[i][j].x + bm.x, ym[i][j].y + bm.y, ym[i][j].z + bm.z, ym[i][j].w + bm.w);
        float4 v;
        v.x = m * mo.x + (1.f - m) * s.x; v.y = m * mo.y + (1.f - m) * s.y;
        v.z = m * mo.z + (1.f - m) * s.z; v.w = m * mo.w + (1.f - m) * s.w;
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (residual) {
          const float4 r = *reinterpret_cast<const float4*>(residual + pix * 512 + c);
          v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        *reinterpret_cast<float4*>(y + pix * 512 + c) = v;
      }
  }
}

// generic output transform: Mt [B,16,H/2,W/2,C] -> y [B,H,W,C] = act(Y + bias) (+ residual); y may alias residual.
// grid = (blocks per sample, B).  STATS: also the per-(sample, block, channel) partial sums (sum y, sum y^2) of the OUTPUT for
// the InstanceNorm that follows (reference util.py:286,296): fp32 over a thread's <= ~16 values, fixed-order shared-memory
// combine of the threads that share a channel quad, one fp64 partial per block -- finalised by stats_finalize_blocks.
template <bool STATS>
__global__ void __launch_bounds__(256) wino_out_kernel(const float* __restrict__ Mt, const float* __restrict__ bias, int act, float slope,
                                                       const float* residual, float* y, int H, int W, int C, double* __restrict__ part) {
  __shared__ float4 r1[STATS ? 256 : 1], r2[STATS ? 256 : 1];
  const int Ht = H >> 1, Wt = W >> 1, C4 = C >> 2;
  const int b = blockIdx.y;
  const long per = (long)Ht * Wt * C4;
  const long plane = (long)Ht * Wt * C;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  const float as = leaky_slope(act, slope);
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < per; idx += (long)gridDim.x * blockDim.x) {
    const unsigned ui = (unsigned)idx, t = ui / (unsigned)C4;
    const int c = (int)(ui - t * (unsigned)C4) * 4;
    const int ty = (int)(t / (unsigned)Wt), tx = (int)(t - (unsigned)ty * (unsigned)Wt);
    float4 yy[2][2];
    wino_out4(Mt + (long)b * 16 * plane + ((long)ty * Wt + tx) * C + c, plane, yy);
    float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bz = __ldg(reinterpret_cast<const float4*>(bias + c));
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const long off = (((long)b * H + 2 * ty + i) * W + 2 * tx + j) * C + c;
        float4 v = make_float4(apply_leaky(yy[i][j].x + bz.x, as), apply_leaky(yy[i][j].y + bz.y, as),
                               apply_leaky(yy[i][j].z + bz.z, as), apply_leaky(yy[i][j].w + bz.w, as));
        if (residual) {
          const float4 r = *reinterpret_cast<const float4*>(residual + off);
          v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if constexpr (STATS) {
          s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
          s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y); s2.z = fmaf(v.z, v.z, s2.z); s2.w = fmaf(v.w, v.w, s2.w);
        }
        *reinterpret_cast<float4*>(y + off) = v;
      }
  }
  if constexpr (STATS) {
    // (gridDim.x * 256) % C4 == 0 (host): a thread keeps its channel quad over the loop; threads t, t + C4, ... share it
    r1[threadIdx.x] = s1; r2[threadIdx.x] = s2;
    __syncthreads();
    if ((int)threadIdx.x < C4) {
      double a[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
      for (int i = threadIdx.x; i < 256; i += C4) {
        const float4 u = r1[i], w = r2[i];
        a[0] += u.x; a[1] += u.y; a[2] += u.z; a[3] += u.w;
        q[0] += w.x; q[1] += w.y; q[2] += w.z; q[3] += w.w;
      }
      double* o = part + (((long)b * gridDim.x + blockIdx.x) * C + threadIdx.x * 4) * 2;
#pragma unroll
      for (int j = 0; j < 4; ++j) { o[2 * j] = a[j]; o[2 * j + 1] = q[j]; }
    }
  }
}

}  // namespace

// Winograd form of a static 3x3 conv (weights packed once at load): w.wn, or nothing when the shape does not qualify
void pack_wino_static(cs_ctx* ctx, ConvW& w) {
  if (!(w.KD == 1 && w.KH == 3 && w.KW == 3 && w.w32 && w.Cin % 32 == 0 && w.Cin >= 128 && w.Cin <= 1024 && w.Cout % 256 == 0)) return;
  ConvW* u = new ConvW();
  ctx->wino_convs.emplace_back(u);
  u->w32 = static_cast<float*>(ctx->dmalloc((size_t)w.Cin * 16 * w.Cout * sizeof(float)));
  wino_weights_kernel<<<148 * 8, 256>>>(w.w32, u->w32, w.Cin, w.Cout, 1.0f);
  check_launch("wino_weights");
  u->Cin = w.Cin; u->Cout = 16 * w.Cout; u->KD = u->KH = u->KW = 1;
  u->wmul = w.wmul * 0.25f;
  pack_tc(ctx, *u, nullptr);
  CS_REQUIRE(u->BN > 0 && w.Cout % u->BN == 0 && u->Cout_p == 16 * w.Cout, CS_ERR_WEIGHTS, "pack_wino_static: unexpected N tile");
  u->Cout = w.Cout; u->zrows = w.Cout;
  u->bias = nullptr;                                     // the bias is added by the output transform
  w.wn = u;
}

// y = act(conv3x3(pre(x)) + bias) (+ residual) in Winograd form; x, y, residual dense fp32 [B,1,H,W,C]; V / Mt scratch from `A`
void wino_conv(const Launcher& L, Arena& A, const Act& x, const ConvW& w, const float* pscale, const float* pshift, int pact,
               float pslope, int act, float slope, const float* residual, Act y, const StatsOut* st) {
  CS_REQUIRE(w.wn != nullptr && x.C == w.Cin && y.C == w.Cout && y.H == x.H && y.W == x.W && y.B == x.B && y.sw == y.C &&
                 y.sh == (long)y.W * y.C && y.sb == (long)y.H * y.W * y.C, CS_ERR_INVALID, "wino_conv: unsupported geometry");
  CS_REQUIRE(act_is_leaky(act) && act_is_leaky(pact), CS_ERR_INVALID, "wino_conv: activation must be none / relu / leaky relu");
  const size_t m = A.mark();
  Opd V; V.B = x.B; V.D = 16; V.H = x.H / 2; V.W = x.W / 2; V.nblk = x.C / 32;
  V.amul = w.wn->amul;
  V.p = A.bf16((size_t)x.B * 16 * V.H * V.W * V.nblk * 64);
  Act Mt = make_act(A.f32((size_t)x.B * 16 * V.H * V.W * w.Cout), x.B, 16, V.H, V.W, w.Cout);
  wino_in(L, x, V, nullptr, nullptr, pscale, pshift, pact, pslope);
  ConvGeom g; g.Do = 16; g.Ho = V.H; g.Wo = V.W;
  Epilogue eg;
  eg.alg_flops = 2.0 * (double)x.pixels() * w.Cout * w.Cin * 9.0;   // the 3x3 conv this GEMM stands for
  conv_tc(L, V, *w.wn, g, eg, Mt);
  L.count();
  const int C4 = w.Cout / 4;
  const bool stats = st && st->scratch && C4 <= 256 && 256 % C4 == 0 && w.Cout <= 512;
  const long per = (long)V.H * V.W * C4;
  CS_REQUIRE(per < (1L << 31), CS_ERR_INVALID, "wino_conv: tensor too large for 32-bit index math");
  const long cap = stats ? STATS_MAX_BLOCKS : 512;          // the statistics scratch holds STATS_MAX_BLOCKS partials per sample
  long blocks = (per + 255) / 256; if (blocks > cap) blocks = cap;
  if (!L.dry) {
    ProfScope ps(L, PK_CONV_TC, 0.0, (double)x.pixels() * w.Cout * (4.0 + 1.0 + (residual ? 1.0 : 0.0)) * 4.0, "wino_out");
    dim3 grid((unsigned)blocks, x.B);
    if (stats) wino_out_kernel<true><<<grid, 256, 0, L.stream>>>(Mt.p, w.bias, act, slope, residual, y.p, x.H, x.W, w.Cout, st->scratch);
    else wino_out_kernel<false><<<grid, 256, 0, L.stream>>>(Mt.p, w.bias, act, slope, residual, y.p, x.H, x.W, w.Cout, nullptr);
    check_launch("wino_out");
  }
  if (st && st->mean) {
    if (stats) stats_finalize_blocks(L, st->scratch, (int)blocks, x.B, w.Cout, (long)x.H * x.W, st->mean, st->rstd, st->eps);
    else instance_stats(L, y, st->mean, st->rstd, st->eps, st->scratch);
  }
  A.reset(m);
}

bool wino_ok(const Launcher& L, const ConvW& w, int H, int W) {
  return L.winograd && L.winograd_static && L.conv_impl != 1 && w.wn != nullptr && H % 2 == 0 && W % 2 == 0 && (long)(H / 2) * (W / 2) >= 128;
}

// per identity: U = G g G^T of combined.w32 -> a.wino (a 1x1x1 conv with depth-dependent weights, zrows = 1024)
void pack_wino(cs_ctx* ctx, AdaptiveConvW& a, cudaStream_t stream) {
  ConvW& w = a.wino;
  const int Cin = 512, Cout = 1024;
  if (!w.w32) w.w32 = static_cast<float*>(ctx->dmalloc((size_t)Cin * 16 * Cout * sizeof(float)));
  wino_weights_kernel<<<148 * 8, 256, 0, stream>>>(a.combined.w32, w.w32, Cin, Cout, 1.0f);
  check_launch("wino_weights");
  w.Cin = Cin; w.Cout = 16 * Cout; w.KD = w.KH = w.KW = 1;
  w.zrows = 0;
  w.wmul = a.combined.wmul * 0.25f;                       // |U| <= 2.25 max|g|: a quarter of the direct conv's pre-scale stays inside fp16
  pack_tc(ctx, w, stream);                               // rows k * 1024 + co, K = 512
  CS_REQUIRE(w.BN == 256 && w.Cout_p == 16 * Cout, CS_ERR_WEIGHTS, "pack_wino: unexpected N tile");
  w.Cout = Cout; w.zrows = Cout;
}

// mask_conv != null: also writes mask[B,H,W] = sigmoid(conv3x3(x; mask_conv) + bias) (w32 layout [9][C][1]);
// pscale / pshift / pact: optional per-channel affine + activation applied to x before the transform
void wino_in(const Launcher& L, const Act& x, Opd V, const ConvW* mask_conv, float* mask, const float* pscale, const float* pshift,
             int pact, float pslope) {
  L.count();
  if (L.dry) return;
  CS_REQUIRE(x.D == 1 && x.C % 32 == 0 && x.C <= 1024 && x.H % 2 == 0 && x.W % 2 == 0 && x.sw == x.C && x.sh == (long)x.W * x.C &&
                 x.sb == (long)x.H * x.W * x.C && V.D == 16 && V.H == x.H / 2 && V.W == x.W / 2 && V.nblk == x.C / 32 && V.B == x.B,
             CS_ERR_INVALID, "wino_in: unsupported geometry");
  CS_REQUIRE(act_is_leaky(pact), CS_ERR_INVALID, "wino_in: activation must be none / relu / leaky relu");
  const long tiles = (long)x.B * (x.H / 2) * (x.W / 2);
  CS_REQUIRE(tiles < (1L << 31), CS_ERR_INVALID, "wino_in: tensor too large for 32-bit index math");
  long blocks = tiles; if (blocks > 148L * 16) blocks = 148L * 16;
  ProfScope ps(L, PK_CONV_TC, 0.0, (double)x.pixels() * x.C * 4.0 + (double)x.pixels() * x.C * 4.0 * 4.0, "wino_in");   // part of the conv
  CS_REQUIRE(x.C / 4 <= 128, CS_ERR_INVALID, "wino_in: at most 512 channels");
  // 6 resident blocks of 128 threads per SM (80 registers, a few spilled words): measured 3.95 -> 3.59 ms per step against 4 blocks
  const bool pre = pscale != nullptr || pact != ACT_NONE;     // compile-time in the kernel: the plain form carries none of its code
  if (mask_conv) {
    CS_REQUIRE(mask && mask_conv->w32 && mask_conv->Cin == x.C && mask_conv->Cout == 1 && mask_conv->KD == 1 && mask_conv->KH == 3 &&
                   mask_conv->KW == 3 && mask_conv->bias && !pre, CS_ERR_INVALID, "wino_in: bad mask conv");
    // 5 resident blocks (96 registers, no spills): 6 blocks (80 registers) spill 150 B per thread here -- measured 1.41 -> 1.14 ms per step
    wino_in_kernel<true, 5, false><<<(unsigned)blocks, x.C / 4, 0, L.stream>>>(x.p, V.p, x.B, x.H, x.W, x.C, mask_conv->w32, mask_conv->bias,
                                                                               mask, nullptr, nullptr, ACT_NONE, 0.f, V.amul);
  } else if (pre) {
    wino_in_kernel<false, 6, true><<<(unsigned)blocks, x.C / 4, 0, L.stream>>>(x.p, V.p, x.B, x.H, x.W, x.C, nullptr, nullptr, nullptr, pscale,
                                                                               pshift, pact, pslope, V.amul);
  } else {
    wino_in_kernel<false, 6, false><<<(unsigned)blocks, x.C / 4, 0, L.stream>>>(x.p, V.p, x.B, x.H, x.W, x.C, nullptr, nullptr, nullptr, nullptr,
                                                                                nullptr, ACT_NONE, 0.f, V.amul);
  }
  check_launch("wino_in");
}

void wino_out_blend(const Launcher& L, const float* Mt, const float* mask, const float* bias_mod, const float* residual, int relu,
                    float* y, int B, int H, int W) {
  L.count();
  if (L.dry) return;
  const long total = (long)B * (H / 2) * (W / 2) * 128;
  CS_REQUIRE(total < (1L << 31), CS_ERR_INVALID, "wino_out_blend: tensor too large for 32-bit index math");
  long blocks = (total + 255) / 256; if (blocks > 148L * 16) blocks = 148L * 16;
  ProfScope ps(L, PK_CONV_TC, 0.0, (double)B * H * W * (4.0 * 1024 + 512 + (residual ? 512 : 0)) * 4.0, "wino_out_blend");
  wino_out_blend_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(Mt, mask, bias_mod, residual, relu, y, B, H, W);
  check_launch("wino_out_blend");
}

}  // namespace cs
