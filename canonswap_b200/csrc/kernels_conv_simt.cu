// fp32 SIMT convolutions (sm_100a): the exact-fp32 path for the shapes that do not map onto the
// tcgen05 implicit GEMM (Cin = 3, Cout = 1, Cout = 4) and the debug / cross-check path for all others.
//   conv_simt   implicit GEMM, 64 pixels x 64 couts per CTA, K chunks of 16 channels per filter tap,
//               channels-last input with generic strides, weights [tap][Cin][Cout]
//   conv_cout1  one warp per output pixel, K split over lanes (mask_conv 512->1, occlusion 2272->1;
//               reference adaptive_modulate.py:118-121, dense_motion.py:25,101)
#include "common.cuh"

namespace cs {

struct ConvSimtK {
  const float* x; long xb, xd, xh, xw; int Cin;
  int B, D, H, W;
  int KD, KH, KW, PD, PH, PW, Do, Ho, Wo;
  const float* w; const float* bias; int Cout;
  int act; float slope;
  const float* res; long rb, rd, rh, rw;
  const float* mult;
  float* y; long yb, yd, yh, yw;
  long M;                  // B*Do*Ho*Wo
  int vecA, vecB;          // 128-bit load eligibility
  int xs;                  // input nearest-upsample shift on (H, W); k.H, k.W are the UPSAMPLED extents
};

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) conv_simt_kernel(ConvSimtK k) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const long m0 = (long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // A loader: pixel lm, channel quad lq
  const int lm = t >> 2, lq = (t & 3) * 4;
  long gm = m0 + lm;
  bool mvalid = gm < k.M;
  int ob = 0, od = 0, oh = 0, ow = 0;
  if (mvalid) {
    ow = (int)(gm % k.Wo); long r = gm / k.Wo;
    oh = (int)(r % k.Ho); r /= k.Ho;
    od = (int)(r % k.Do); ob = (int)(r / k.Do);
  }
  // B loader: k row bk, cout quad bq
  const int bk = t >> 4, bq = (t & 15) * 4;

  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int taps = k.KD * k.KH * k.KW;
  for (int tap = 0; tap < taps; ++tap) {
    int kw = tap % k.KW; int r = tap / k.KW; int kh = r % k.KH; int kd = r / k.KH;
    int id = od + kd - k.PD, ih = oh + kh - k.PH, iw = ow + kw - k.PW;
    bool inb = mvalid && id >= 0 && id < k.D && ih >= 0 && ih < k.H && iw >= 0 && iw < k.W;
    const float* xp = k.x + ob * k.xb + id * k.xd + (long)(ih >> k.xs) * k.xh + (long)(iw >> k.xs) * k.xw;
    const float* wp = k.w + (long)tap * k.Cin * k.Cout;
    for (int c0 = 0; c0 < k.Cin; c0 += BK) {
      // ---- load A (transposed into As[k][m]) ----
      float a4[4] = {0.f, 0.f, 0.f, 0.f};
      int c = c0 + lq;
      if (inb) {
        if (k.vecA && c + 3 < k.Cin) {
          float4 v = *reinterpret_cast<const float4*>(xp + c);
          a4[0] = v.x; a4[1] = v.y; a4[2] = v.z; a4[3] = v.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (c + i < k.Cin) a4[i] = xp[c + i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) As[lq + i][lm] = a4[i];
      // ---- load B ----
      float b4[4] = {0.f, 0.f, 0.f, 0.f};
      int ck = c0 + bk, n = n0 + bq;
      if (ck < k.Cin) {
        const float* q = wp + (long)ck * k.Cout + n;
        if (k.vecB && n + 3 < k.Cout) {
          float4 v = *reinterpret_cast<const float4*>(q);
          b4[0] = v.x; b4[1] = v.y; b4[2] = v.z; b4[3] = v.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (n + i < k.Cout) b4[i] = q[i];
        }
      }
      *reinterpret_cast<float4*>(&Bs[bk][bq]) = make_float4(b4[0], b4[1], b4[2], b4[3]);
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long m = m0 + ty * 4 + i;
    if (m >= k.M) continue;
    int pw = (int)(m % k.Wo); long r = m / k.Wo;
    int ph = (int)(r % k.Ho); r /= k.Ho;
    int pd = (int)(r % k.Do); int pb = (int)(r / k.Do);
    float* yp = k.y + pb * k.yb + pd * k.yd + ph * k.yh + pw * k.yw;
    const float* rp = k.res ? k.res + pb * k.rb + pd * k.rd + ph * k.rh + pw * k.rw : nullptr;
    float mu = k.mult ? k.mult[m] : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= k.Cout) continue;
      float v = acc[i][j] + (k.bias ? k.bias[n] : 0.f);
      v = apply_act(v, k.act, k.slope);
      if (rp) v += rp[n];
      if (k.mult) v *= mu;
      yp[n] = v;
    }
  }
}

void conv_simt(const Launcher& L, const Act& x, const ConvW& w, const ConvGeom& g, const Epilogue& e, Act y, int xshift) {
  L.count();
  if (L.dry) return;
  CS_REQUIRE(x.C == w.Cin, -1, "conv_simt: Cin mismatch");
  CS_REQUIRE(y.C == w.Cout, -1, "conv_simt: Cout mismatch");
  CS_REQUIRE(w.w32 != nullptr, -3, "conv_simt: weights not packed");
  ConvSimtK k{};
  k.x = x.p; k.xb = x.sb; k.xd = x.sd; k.xh = x.sh; k.xw = x.sw; k.Cin = x.C;
  k.B = x.B; k.D = x.D; k.H = x.H << xshift; k.W = x.W << xshift; k.xs = xshift;
  k.KD = w.KD; k.KH = w.KH; k.KW = w.KW; k.PD = g.PD; k.PH = g.PH; k.PW = g.PW;
  k.Do = g.Do; k.Ho = g.Ho; k.Wo = g.Wo;
  k.w = w.w32; k.bias = w.bias; k.Cout = w.Cout;
  k.act = e.act; k.slope = e.slope;
  k.res = e.residual; k.rb = e.rs_b; k.rd = e.rs_d; k.rh = e.rs_h; k.rw = e.rs_w;
  k.mult = e.mult;
  k.y = y.p; k.yb = y.sb; k.yd = y.sd; k.yh = y.sh; k.yw = y.sw;
  k.M = (long)x.B * g.Do * g.Ho * g.Wo;
  auto al4 = [](long v) { return (v & 3) == 0; };
  k.vecA = (x.C % 4 == 0) && al4(x.sb) && al4(x.sd) && al4(x.sh) && al4(x.sw) && ((uintptr_t)x.p % 16 == 0);
  k.vecB = (w.Cout % 4 == 0) && ((uintptr_t)w.w32 % 16 == 0);
  dim3 grid((unsigned)((k.M + BM - 1) / BM), (w.Cout + BN - 1) / BN);
  ProfScope ps(L, PK_CONV_SIMT, 2.0 * (double)k.M * w.Cout * w.Cin * w.taps(), 0.0, "simt");
  conv_simt_kernel<<<grid, 256, 0, L.stream>>>(k);
  check_launch("conv_simt");
}

// ------------------------------------------------------------------------------------------
// Cout == 1: one warp per output pixel
// ------------------------------------------------------------------------------------------
struct ConvC1K {
  const float* x; long xb, xd, xh, xw; int Cin;
  int D, H, W, KD, KH, KW, PD, PH, PW, Do, Ho, Wo;
  const float* w; float bias; const float* bias_p; int act;
  float* y; long M;
};

__global__ void __launch_bounds__(256) conv_cout1_kernel(ConvC1K k) {
  const int lane = threadIdx.x & 31;
  long m = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= k.M) return;
  int ow = (int)(m % k.Wo); long r = m / k.Wo;
  int oh = (int)(r % k.Ho); r /= k.Ho;
  int od = (int)(r % k.Do); int ob = (int)(r / k.Do);
  float acc0 = 0.f, acc1 = 0.f;
  const int taps = k.KD * k.KH * k.KW;
  for (int tap = 0; tap < taps; ++tap) {
    int kw = tap % k.KW; int q = tap / k.KW; int kh = q % k.KH; int kd = q / k.KH;
    int id = od + kd - k.PD, ih = oh + kh - k.PH, iw = ow + kw - k.PW;
    if (id < 0 || id >= k.D || ih < 0 || ih >= k.H || iw < 0 || iw >= k.W) continue;   // warp-uniform
    const float* xp = k.x + ob * k.xb + id * k.xd + ih * k.xh + iw * k.xw;
    const float* wp = k.w + (long)tap * k.Cin;
    int c = lane;
    for (; c + 32 < k.Cin; c += 64) {
      acc0 = fmaf(xp[c], wp[c], acc0);
      acc1 = fmaf(xp[c + 32], wp[c + 32], acc1);
    }
    if (c < k.Cin) acc0 = fmaf(xp[c], wp[c], acc0);
  }
  float acc = acc0 + acc1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) k.y[m] = apply_act(acc + (k.bias_p ? k.bias_p[0] : 0.f), k.act, 0.f);
}

void conv_cout1(const Launcher& L, const Act& x, const ConvW& w, const ConvGeom& g, int act, float* y) {
  L.count();
  if (L.dry) return;
  CS_REQUIRE(w.Cout == 1 && x.C == w.Cin, -1, "conv_cout1: shape mismatch");
  ConvC1K k{};
  k.x = x.p; k.xb = x.sb; k.xd = x.sd; k.xh = x.sh; k.xw = x.sw; k.Cin = x.C;
  k.D = x.D; k.H = x.H; k.W = x.W; k.KD = w.KD; k.KH = w.KH; k.KW = w.KW;
  k.PD = g.PD; k.PH = g.PH; k.PW = g.PW; k.Do = g.Do; k.Ho = g.Ho; k.Wo = g.Wo;
  k.w = w.w32; k.bias_p = w.bias; k.act = act; k.y = y;
  k.M = (long)x.B * g.Do * g.Ho * g.Wo;
  ProfScope ps(L, PK_CONV_SIMT, 2.0 * (double)k.M * w.Cin * w.taps(), 0.0, "cout1");
  conv_cout1_kernel<<<(unsigned)((k.M + 7) / 8), 256, 0, L.stream>>>(k);
  check_launch("conv_cout1");
}

}  // namespace cs
