// Paste-back of the swapped crop into the full frame, the step that follows the generator:
//   prepare_paste_back(mask_crop, M_c2o, dsize, if_float=True)    reference src/utils/crop.py:515-521
//   paste_back(img_crop, M_c2o, img_ori, mask_ori)                reference src/utils/crop.py:523-529
//   both through cv2.warpAffine(.., flags=cv2.INTER_LINEAR)       reference src/utils/crop.py:49-63
// The reference runs two full-frame cv2.warpAffine calls and a float blend on the CPU per frame; here it is ONE kernel, one
// thread per destination pixel, no full-frame intermediates: the 10-bit fixed-point source coordinate of OpenCV
// (AB_BITS = 10, INTER_BITS = 5, round_delta = 16, doubles rounded half-to-even), the uint8 bilinear sample with the integer
// weights 32 * p * q and (sum + 2^14) >> 15, the float32 bilinear sample of the mask ((S00*w0 + S01*w1) + S10*w2) + S11*w3,
// and clip(mask*result + (1-mask)*img_ori, 0, 255) truncated to uint8 -- every operation written with explicit
// round-to-nearest intrinsics so that no FMA contraction can change a bit.  Bit-exact against the numpy oracle, which is
// pinned bit for bit against cv2 / the reference functions (oracle/pasteback_oracle.py, tests/test_pasteback.py).
#include "ctx.cuh"

namespace cs {

namespace {

struct PasteK {
  const uint8_t* crop;     // [B,hc,wc,3]
  const float* mask;       // [B,hc,wc]   (the reference stacks it to 3 equal channels)
  const uint8_t* ori;      // [B,H,W,3]
  uint8_t* out;            // [B,H,W,3]
  int B, hc, wc, H, W;
  double iM[CS_PASTE_MAX_BATCH][6];   // inverted 2x3 matrices (host, double, cv::warpAffine's formulas)
};

__device__ __forceinline__ long long cv_round(double v) { return __double2ll_rn(v); }

// one destination pixel: its three output bytes from the frame's bytes o[0..2]
__device__ __forceinline__ void paste_pixel(const PasteK& k, int b, int x, int y, const uint8_t* o, uint8_t* dst) {
  const double* m = k.iM[b];
  // adelta[x] = round(M00 * x * 1024), X0 = round((M01 * y + M02) * 1024) + 16   (no contraction: __dmul_rn / __dadd_rn)
  const long long ad = cv_round(__dmul_rn(__dmul_rn(m[0], (double)x), 1024.0));
  const long long bd = cv_round(__dmul_rn(__dmul_rn(m[3], (double)x), 1024.0));
  const long long X0 = cv_round(__dmul_rn(__dadd_rn(__dmul_rn(m[1], (double)y), m[2]), 1024.0)) + 16;
  const long long Y0 = cv_round(__dmul_rn(__dadd_rn(__dmul_rn(m[4], (double)y), m[5]), 1024.0)) + 16;
  const long long X = (X0 + ad) >> 5, Y = (Y0 + bd) >> 5;
  const long long sx = X >> 5, sy = Y >> 5;
  const int fx = (int)(X & 31), fy = (int)(Y & 31);
  if (sx < -1 || sx >= k.wc || sy < -1 || sy >= k.hc) {      // all four taps outside: result = 0, mask = 0 -> the frame itself
    dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];              // (1 - 0) * img_ori is exact
    return;
  }
  bool ok[4];
  long off[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long yy = sy + (j >> 1), xx = sx + (j & 1);
    ok[j] = yy >= 0 && yy < k.hc && xx >= 0 && xx < k.wc;
    off[j] = ok[j] ? ((long)b * k.hc + yy) * k.wc + xx : 0;
  }
  // float32 mask sample
  const float a = __fmul_rn((float)fx, 1.0f / 32.0f), bb = __fmul_rn((float)fy, 1.0f / 32.0f);
  const float ia = __fsub_rn(1.f, a), ib = __fsub_rn(1.f, bb);
  const float w0 = __fmul_rn(ib, ia), w1 = __fmul_rn(ib, a), w2 = __fmul_rn(bb, ia), w3 = __fmul_rn(bb, a);
  const float s0 = ok[0] ? k.mask[off[0]] : 0.f, s1 = ok[1] ? k.mask[off[1]] : 0.f;
  const float s2 = ok[2] ? k.mask[off[2]] : 0.f, s3 = ok[3] ? k.mask[off[3]] : 0.f;
  float mk = __fmul_rn(s0, w0);
  mk = __fadd_rn(mk, __fmul_rn(s1, w1));
  mk = __fadd_rn(mk, __fmul_rn(s2, w2));
  mk = __fadd_rn(mk, __fmul_rn(s3, w3));
  const float im = __fsub_rn(1.f, mk);
  // uint8 image sample: integer weights 32 * p * q, (sum + 2^14) >> 15
  const int iw0 = 32 * (32 - fy) * (32 - fx), iw1 = 32 * (32 - fy) * fx, iw2 = 32 * fy * (32 - fx), iw3 = 32 * fy * fx;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int acc = 1 << 14;
    if (ok[0]) acc += iw0 * k.crop[off[0] * 3 + c];
    if (ok[1]) acc += iw1 * k.crop[off[1] * 3 + c];
    if (ok[2]) acc += iw2 * k.crop[off[2] * 3 + c];
    if (ok[3]) acc += iw3 * k.crop[off[3] * 3 + c];
    int res = acc >> 15;
    res = res > 255 ? 255 : res;
    float v = __fadd_rn(__fmul_rn(mk, (float)res), __fmul_rn(im, (float)o[c]));
    v = fminf(fmaxf(v, 0.f), 255.f);
    dst[c] = (uint8_t)v;                                  // astype(uint8): truncation
  }
}

// VEC: one thread = 4 consecutive pixels of a row = 12 bytes = three aligned 32-bit words of the frame (W % 4 == 0, 4-byte
// aligned bases); else one thread = one pixel with byte accesses
template <bool VEC>
__global__ void __launch_bounds__(256) paste_back_kernel(const __grid_constant__ PasteK k) {
  if constexpr (VEC) {
    const int W4 = k.W >> 2;
    const long total = (long)k.B * k.H * W4;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
      const int x4 = (int)(idx % W4); long t = idx / W4;
      const int y = (int)(t % k.H); const int b = (int)(t / k.H);
      const long base = (((long)b * k.H + y) * k.W + (long)x4 * 4) * 3;
      union { uint32_t w[3]; uint8_t c[12]; } in, out;
      const uint32_t* ip = reinterpret_cast<const uint32_t*>(k.ori + base);
      in.w[0] = ip[0]; in.w[1] = ip[1]; in.w[2] = ip[2];
#pragma unroll
      for (int p = 0; p < 4; ++p) paste_pixel(k, b, x4 * 4 + p, y, in.c + 3 * p, out.c + 3 * p);
      uint32_t* op = reinterpret_cast<uint32_t*>(k.out + base);
      op[0] = out.w[0]; op[1] = out.w[1]; op[2] = out.w[2];
    }
  } else {
    const long total = (long)k.B * k.H * k.W;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
      const int x = (int)(idx % k.W); long t = idx / k.W;
      const int y = (int)(t % k.H); const int b = (int)(t / k.H);
      uint8_t o[3] = {k.ori[idx * 3], k.ori[idx * 3 + 1], k.ori[idx * 3 + 2]}, d[3];
      paste_pixel(k, b, x, y, o, d);
      k.out[idx * 3] = d[0]; k.out[idx * 3 + 1] = d[1]; k.out[idx * 3 + 2] = d[2];
    }
  }
}

}  // namespace

// M_c2o: [B][6] doubles, row-major 2x3 (the reference's float32 3x3 M_c2o[:2, :] widened to double, as cv2 does)
void paste_back(const Launcher& L, const uint8_t* crop, const float* mask, const double* M_c2o, const uint8_t* ori, uint8_t* out, int B,
                int hc, int wc, int H, int W) {
  L.count();
  if (L.dry) return;
  CS_REQUIRE(B >= 1 && B <= CS_PASTE_MAX_BATCH, CS_ERR_INVALID, "paste_back: batch outside [1, CS_PASTE_MAX_BATCH]");
  PasteK k{};
  k.crop = crop; k.mask = mask; k.ori = ori; k.out = out; k.B = B; k.hc = hc; k.wc = wc; k.H = H; k.W = W;
  for (int b = 0; b < B; ++b) {
    const double* M = M_c2o + 6 * b;                        // cv::warpAffine: invert unless WARP_INVERSE_MAP
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    const double A11 = M[4] * D, A22 = M[0] * D;
    double* q = k.iM[b];
    q[0] = A11; q[1] = M[1] * (-D); q[3] = M[3] * (-D); q[4] = A22;
    q[2] = -q[0] * M[2] - q[1] * M[5];
    q[5] = -q[3] * M[2] - q[4] * M[5];
  }
  const long total = (long)B * H * W;
  const bool vec = (W % 4 == 0) && ((uintptr_t)ori % 4 == 0) && ((uintptr_t)out % 4 == 0);
  long blocks = ((vec ? total / 4 : total) + 255) / 256; if (blocks > 148L * 32) blocks = 148L * 32;
  ProfScope ps(L, PK_OTHER, 0.0, (double)total * 6.0 + (double)B * hc * wc * 7.0, "paste_back");
  if (vec) paste_back_kernel<true><<<(unsigned)blocks, 256, 0, L.stream>>>(k);
  else paste_back_kernel<false><<<(unsigned)blocks, 256, 0, L.stream>>>(k);
  check_launch("paste_back");
}

}  // namespace cs

// ------------------------------------------------------------------------------------------
// SoftErosion (reference src/utils/crop.py:21-47; pipeline_e2e.py:42,275: kernel_size 21, threshold 0.9, iterations 3):
//   for i in range(iterations - 1): x = min(x, conv2d(x, k, padding=r));  x = conv2d(x, k, padding=r)
//   mask = x >= threshold;  x[mask] = 1;  x[~mask] /= x[~mask].max()
// k = (dist.max() - dist) / sum, dist = distance to the kernel centre.  Float work: one block per 32x8 output tile with the
// haloed input tile in shared memory, the kernel weights in constant-like global memory (L1-resident); the maximum of the
// sub-threshold values through an integer atomicMax on the (non-negative) float bit patterns.
// ------------------------------------------------------------------------------------------
namespace cs {

namespace {

constexpr int SE_TW = 32, SE_TH = 8, SE_MAXK = 31;

__global__ void __launch_bounds__(256) soft_erosion_conv_kernel(const float* __restrict__ x, const float* __restrict__ kw, float* __restrict__ y,
                                                                int H, int W, int K, int take_min) {
  extern __shared__ float tile[];                           // (SE_TH + K - 1) x (SE_TW + K - 1)
  const int r = K / 2, tw = SE_TW + K - 1, th = SE_TH + K - 1;
  const int b = blockIdx.z, x0 = blockIdx.x * SE_TW, y0 = blockIdx.y * SE_TH;
  const float* xb = x + (long)b * H * W;
  for (int i = threadIdx.x; i < tw * th; i += blockDim.x) {
    const int ty = i / tw, tx = i % tw;
    const int gy = y0 + ty - r, gx = x0 + tx - r;
    tile[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? xb[(long)gy * W + gx] : 0.f;
  }
  __syncthreads();
  const int lx = threadIdx.x % SE_TW, ly = threadIdx.x / SE_TW;
  const int gx = x0 + lx, gy = y0 + ly;
  if (gx >= W || gy >= H) return;
  float acc = 0.f;
  for (int ky = 0; ky < K; ++ky) {
    const float* trow = tile + (ly + ky) * tw + lx;
    const float* wrow = kw + ky * K;
    for (int kx = 0; kx < K; ++kx) acc = fmaf(trow[kx], __ldg(wrow + kx), acc);
  }
  const float c = tile[(ly + r) * tw + lx + r];
  y[(long)b * H * W + (long)gy * W + gx] = take_min ? fminf(c, acc) : acc;
}

__global__ void __launch_bounds__(256) soft_erosion_max_kernel(const float* __restrict__ x, long n, long per, float thr, int* __restrict__ mx) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = x[i];
    if (!(v >= thr)) atomicMax(mx + i / per, __float_as_int(fmaxf(v, 0.f)));   // conv of a non-negative mask: v >= 0
  }
}

__global__ void __launch_bounds__(256) soft_erosion_apply_kernel(float* __restrict__ x, long n, long per, float thr, const int* __restrict__ mx,
                                                                 uint8_t* __restrict__ hard) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = x[i];
    x[i] = (v >= thr) ? 1.f : v / __int_as_float(mx[i / per]);
    if (hard) hard[i] = v >= thr ? 1 : 0;
  }
}

}  // namespace

// x [B,H,W] -> out [B,H,W]; tmp: 2 * B*H*W floats + B ints of scratch; kw: [K*K] normalised kernel on the device
void soft_erosion(const Launcher& L, const float* x, float* out, uint8_t* hard, float* tmp, const float* kw, int B, int H, int W, int K,
                  float thr, int iterations) {
  for (int i = 0; i < iterations + 2; ++i) L.count();
  if (L.dry) return;
  CS_REQUIRE(K >= 1 && K <= SE_MAXK && (K & 1) && iterations >= 1, CS_ERR_INVALID, "soft_erosion: kernel_size must be odd and <= 31");
  const long n = (long)B * H * W;
  float* a = tmp; float* bbuf = tmp + n;
  int* mx = reinterpret_cast<int*>(tmp + 2 * n);
  const dim3 grid((W + SE_TW - 1) / SE_TW, (H + SE_TH - 1) / SE_TH, B);
  const size_t smem = (size_t)(SE_TH + K - 1) * (SE_TW + K - 1) * sizeof(float);
  ProfScope ps(L, PK_OTHER, 0.0, (double)n * 8.0 * iterations, "soft_erosion");
  const float* src = x;
  for (int i = 0; i < iterations; ++i) {
    float* dst = (i == iterations - 1) ? out : ((i & 1) ? bbuf : a);
    soft_erosion_conv_kernel<<<grid, 256, smem, L.stream>>>(src, kw, dst, H, W, K, i < iterations - 1 ? 1 : 0);
    check_launch("soft_erosion_conv");
    src = dst;
  }
  CS_CUDA(cudaMemsetAsync(mx, 0, sizeof(int) * B, L.stream));
  long blocks = (n + 255) / 256; if (blocks > 148L * 16) blocks = 148L * 16;
  soft_erosion_max_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(out, n, (long)H * W, thr, mx);
  soft_erosion_apply_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(out, n, (long)H * W, thr, mx, hard);
  check_launch("soft_erosion_apply");
}


// ------------------------------------------------------------------------------------------
// Face-parsing post-processing (the step BEFORE the path, reference src/can_swap_pipeline_e2e.py:183-190):
//   upsampled = F.interpolate(logits, size=(H, W), mode='bilinear', align_corners=False)
//   labels    = upsampled.argmax(dim=1);   mask = isin(labels, valid_list)
// fused into one kernel: no [B,19,512,512] upsampled tensor (20 MB per frame), no label tensor unless asked for.  The
// interpolation follows ATen's upsample_bilinear2d arithmetic (source index scale * (dst + 0.5) - 0.5 clamped at 0, weights
// 1 - lambda / lambda, rows combined after columns), with explicit round-to-nearest products and sums so that no FMA
// contraction can move a near-tie; argmax keeps the FIRST maximal class, as torch does.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) parse_mask_kernel(const float* __restrict__ logits, int B, int C, int h, int w, int H, int W,
                                                         unsigned long long valid, float* __restrict__ mask, int* __restrict__ labels) {
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  const long total = (long)B * H * W;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % W); long t = i / W;
    const int oy = (int)(t % H); const int b = (int)(t / H);
    float fy = __fsub_rn(__fmul_rn(sh, (float)oy + 0.5f), 0.5f); if (fy < 0.f) fy = 0.f;
    float fx = __fsub_rn(__fmul_rn(sw, (float)ox + 0.5f), 0.5f); if (fx < 0.f) fx = 0.f;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly1 = __fsub_rn(fy, (float)y0), lx1 = __fsub_rn(fx, (float)x0);
    const float ly0 = __fsub_rn(1.f, ly1), lx0 = __fsub_rn(1.f, lx1);
    const float* base = logits + (long)b * C * h * w;
    float best = 0.f; int arg = 0;
    for (int c = 0; c < C; ++c) {
      const float* pc = base + (long)c * h * w;
      const float v00 = __ldg(pc + (long)y0 * w + x0), v01 = __ldg(pc + (long)y0 * w + x1);
      const float v10 = __ldg(pc + (long)y1 * w + x0), v11 = __ldg(pc + (long)y1 * w + x1);
      const float r0 = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
      const float r1 = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
      const float v = __fadd_rn(__fmul_rn(ly0, r0), __fmul_rn(ly1, r1));
      if (c == 0 || v > best || (v != v && best == best)) { best = v; arg = c; }    // first maximum; a NaN wins, as in torch
    }
    mask[i] = ((valid >> arg) & 1ull) ? 1.f : 0.f;
    if (labels) labels[i] = arg;
  }
}

void parse_mask(const Launcher& L, const float* logits, int B, int C, int h, int w, int H, int W, unsigned long long valid, float* mask,
                int* labels) {
  L.count();
  if (L.dry) return;
  const long total = (long)B * H * W;
  long blocks = (total + 255) / 256; if (blocks > 148L * 16) blocks = 148L * 16;
  ProfScope ps(L, PK_OTHER, 0.0, (double)B * C * h * w * 4.0 + (double)total * 4.0, "parse_mask");
  parse_mask_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(logits, B, C, h, w, H, W, valid, mask, labels);
  check_launch("parse_mask");
}

}  // namespace cs
