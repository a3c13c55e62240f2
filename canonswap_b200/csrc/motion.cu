// Motion extractor M and the keypoint transform that feed the per-frame loop:
//   MotionExtractor.forward -> ConvNeXtV2-tiny   reference src/modules/motion_extractor.py:33-35, convnextv2.py:34-47,110-144
//   LayerNorm (both data formats), GRN            reference src/modules/util.py:356-368,388-396
//   headpose_pred_to_degree, get_rotation_matrix  reference src/utils/camera.py:14-29,32-73
//   can_swapper.transform_keypoint                reference src/can_swap_e2e.py:226-254
//   x_can = scale * kp, x_t = x_s                 reference src/can_swap_pipeline_e2e.py:112-125,231-243
//
// Layout: fp32 channels-last [B,H,W,C] like the generator.  Both LayerNorm flavours of the reference normalise over the
// channel axis of one pixel, so they are one warp-per-pixel kernel here.  The dense contractions (the two Linear layers
// of every block = 94 % of the 11.6 GFLOP, and the 2x2 stride-2 convs as space-to-depth + 1x1) run on the persistent
// tcgen05 kernel of conv_tc.cu as 1x1 convs; everything else is bandwidth-bound:
//   stem_ln      4x4 stride-4 conv 3 -> 96 + LayerNorm, one warp per output pixel
//   dw_ln        depthwise 7x7 conv + bias + LayerNorm + affine -> split-fp16 operand of pwconv1 (one warp per pixel);
//                without the depthwise part and with a space-to-depth output position it is the downsample LayerNorm
//   grn_sumsq / grn_finalize / grn_apply   Global Response Normalization of the GELU output -> operand of pwconv2
//   head         global average pool + LayerNorm(768) + the seven Linear heads (328 outputs), one CTA per frame
//   keypoints    softmax-expectation angles, rotation matrix, x_s = s (kp R + exp) + t_xy, x_can = s kp
#include "ctx.cuh"
#include <cmath>

namespace cs {

namespace {

constexpr float LN_EPS = 1e-6f;
constexpr int M_DIMS[4] = {96, 192, 384, 768};
constexpr int M_DEPTHS[4] = {3, 3, 9, 3};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// stem: Conv2d(3, 96, k=4, s=4) + LayerNorm(channels_first), convnextv2.py:74-77
// img [B,H,W,3] fp32 channels-last, w [48][96] with k = (kh*4 + kw)*3 + ci, out [B,H/4,W/4,96]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_ln_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                      const float* __restrict__ bias, const float* __restrict__ lnw,
                                                      const float* __restrict__ lnb, float* __restrict__ out, int B, int H, int W) {
  const int Ho = H >> 2, Wo = W >> 2;
  const long npix = (long)B * Ho * Wo;
  const int lane = threadIdx.x & 31;
  for (long pix = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pix < npix; pix += (long)gridDim.x * (blockDim.x >> 5)) {
    const int wo = (int)(pix % Wo); long t = pix / Wo;
    const int ho = (int)(t % Ho); const int b = (int)(t / Ho);
    // the 4x4x3 patch: 4 rows of 12 contiguous floats; lane l holds element k = l and k = l + 32 (k < 48)
    const float* p0 = img + (((long)b * H + ho * 4) * W + wo * 4) * 3;
    const float in_a = p0[(long)(lane / 12) * W * 3 + (lane % 12)];
    const int k2 = lane + 32;
    const float in_b = k2 < 48 ? p0[(long)(k2 / 12) * W * 3 + (k2 % 12)] : 0.f;
    float acc[3] = {bias[lane], bias[lane + 32], bias[lane + 64]};
#pragma unroll 8
    for (int k = 0; k < 48; ++k) {
      const float v = k < 32 ? __shfl_sync(0xffffffffu, in_a, k) : __shfl_sync(0xffffffffu, in_b, k - 32);
      const float* wr = w + k * 96;
      acc[0] = fmaf(v, wr[lane], acc[0]); acc[1] = fmaf(v, wr[lane + 32], acc[1]); acc[2] = fmaf(v, wr[lane + 64], acc[2]);
    }
    const float mean = warp_sum(acc[0] + acc[1] + acc[2]) * (1.f / 96.f);
    const float d0 = acc[0] - mean, d1 = acc[1] - mean, d2 = acc[2] - mean;
    const float var = warp_sum(d0 * d0 + d1 * d1 + d2 * d2) * (1.f / 96.f);
    const float r = 1.f / sqrtf(var + LN_EPS);
    float* o = out + pix * 96;
    o[lane] = lnw[lane] * (d0 * r) + lnb[lane];
    o[lane + 32] = lnw[lane + 32] * (d1 * r) + lnb[lane + 32];
    o[lane + 64] = lnw[lane + 64] * (d2 * r) + lnb[lane + 64];
  }
}

// ------------------------------------------------------------------------------------------
// dw_ln: [depthwise 7x7 + bias ->] LayerNorm over C -> affine -> split operand.  One warp per pixel, lane l owns the
// float4 channel groups l, l + 32, ... (C <= 768: at most 6).  S2D: the operand pixel is (h/2, w/2) and the channels land
// at ((h&1)*2 + (w&1))*C (the K order of the 2x2 stride-2 conv packed as a 1x1 conv).
// ------------------------------------------------------------------------------------------
template <bool DW, bool S2D>
__global__ void __launch_bounds__(256) dw_ln_kernel(const float* __restrict__ x, const float* __restrict__ dww /*[49][C]*/,
                                                    const float* __restrict__ dwb, const float* __restrict__ lnw,
                                                    const float* __restrict__ lnb, __nv_bfloat16* __restrict__ opl, int B, int H,
                                                    int W, int C) {
  const long npix = (long)B * H * W;
  const int lane = threadIdx.x & 31;
  const int C4 = C >> 2;
  const float invC = 1.f / (float)C;
  for (long pix = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pix < npix; pix += (long)gridDim.x * (blockDim.x >> 5)) {
    const int w = (int)(pix % W); long t = pix / W;
    const int h = (int)(t % H); const int b = (int)(t / H);
    float4 v[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int g = lane + 32 * j;
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < C4) {
        if constexpr (DW) v[j] = *reinterpret_cast<const float4*>(dwb + g * 4);
        else v[j] = *reinterpret_cast<const float4*>(x + pix * C + g * 4);
      }
    }
    if constexpr (DW) {
      for (int kh = 0; kh < 7; ++kh) {
        const int hh = h + kh - 3;
        if (hh < 0 || hh >= H) continue;
        for (int kw = 0; kw < 7; ++kw) {
          const int ww = w + kw - 3;
          if (ww < 0 || ww >= W) continue;
          const float* xp = x + (((long)b * H + hh) * W + ww) * C;
          const float* wp = dww + (kh * 7 + kw) * C;
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const int g = lane + 32 * j;
            if (g < C4) {
              const float4 a = *reinterpret_cast<const float4*>(xp + g * 4);
              const float4 q = __ldg(reinterpret_cast<const float4*>(wp + g * 4));
              v[j].x = fmaf(a.x, q.x, v[j].x); v[j].y = fmaf(a.y, q.y, v[j].y);
              v[j].z = fmaf(a.z, q.z, v[j].z); v[j].w = fmaf(a.w, q.w, v[j].w);
            }
          }
        }
      }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);      // slots beyond C are zero
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      if (lane + 32 * j < C4) {
        const float dx = v[j].x - mean, dy = v[j].y - mean, dz = v[j].z - mean, dw_ = v[j].w - mean;
        q += (dx * dx + dy * dy) + (dz * dz + dw_ * dw_);
      }
    }
    const float r = 1.f / sqrtf(warp_sum(q) * invC + LN_EPS);
    long opix = pix; int coff = 0; long prow = (long)(C >> 5) * 64;
    if constexpr (S2D) {
      opix = ((long)b * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1);
      coff = ((h & 1) * 2 + (w & 1)) * C;
      prow *= 4;
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int g = lane + 32 * j;
      if (g < C4) {
        const float4 gw = __ldg(reinterpret_cast<const float4*>(lnw + g * 4)), gb = __ldg(reinterpret_cast<const float4*>(lnb + g * 4));
        const float o0 = gw.x * ((v[j].x - mean) * r) + gb.x, o1 = gw.y * ((v[j].y - mean) * r) + gb.y;
        const float o2 = gw.z * ((v[j].z - mean) * r) + gb.z, o3 = gw.w * ((v[j].w - mean) * r) + gb.w;
        uint2 hv, lv;
        split_operand4(o0, o1, o2, o3, hv, lv);
        const int c = coff + g * 4;
        __nv_bfloat16* ep = opl + opix * prow + (c >> 5) * 64 + (c & 31);
        *reinterpret_cast<uint2*>(ep) = hv;
        *reinterpret_cast<uint2*>(ep + 32) = lv;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// GRN, util.py:365-368:  Gx = ||x||_2 over (H,W) per (b,c);  Nx = Gx / (mean_c Gx + 1e-6);  y = gamma (x Nx) + beta + x
// ------------------------------------------------------------------------------------------
// partial sums of squares: block = (slab of pixels) x (all channels of one frame), fp32 per thread, fp64 atomics
__global__ void __launch_bounds__(256) grn_sumsq_kernel(const float* __restrict__ x, double* __restrict__ sumsq, int HW, int C,
                                                        int slab) {
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * slab, p1 = min(HW, p0 + slab);
  const int C4 = C >> 2;
  for (int g = threadIdx.x; g < C4; g += blockDim.x) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* xp = x + ((long)b * HW + p0) * C + g * 4;
    for (int p = p0; p < p1; ++p, xp += C) {
      const float4 v = *reinterpret_cast<const float4*>(xp);
      a.x = fmaf(v.x, v.x, a.x); a.y = fmaf(v.y, v.y, a.y); a.z = fmaf(v.z, v.z, a.z); a.w = fmaf(v.w, v.w, a.w);
    }
    double* d = sumsq + (long)b * C + g * 4;
    atomicAdd(d, (double)a.x); atomicAdd(d + 1, (double)a.y); atomicAdd(d + 2, (double)a.z); atomicAdd(d + 3, (double)a.w);
  }
}

// one block per frame: mult[b,c] = gamma[c] * Nx[b,c] + 1; clears the scratch for the next use
__global__ void __launch_bounds__(256) grn_finalize_kernel(double* __restrict__ sumsq, const float* __restrict__ gamma,
                                                           float* __restrict__ mult, int C) {
  __shared__ float red[8];
  __shared__ float s_mean;
  const int b = blockIdx.x;
  float part = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) part += sqrtf((float)sumsq[(long)b * C + c]);
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
    s_mean = tot / (float)C;
  }
  __syncthreads();
  const float den = s_mean + 1e-6f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float gx = sqrtf((float)sumsq[(long)b * C + c]);
    mult[(long)b * C + c] = gamma[c] * (gx / den) + 1.f;
    sumsq[(long)b * C + c] = 0.0;
  }
}

// y = x * mult[b,c] + beta[c] -> split operand [pixels, C/32, 64]
__global__ void __launch_bounds__(256) grn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mult,
                                                        const float* __restrict__ beta, __nv_bfloat16* __restrict__ opl, long npix,
                                                        int HW, int C) {
  const int C4 = C >> 2;
  const long total = npix * C4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const long pix = i / C4;
    const int b = (int)(pix / HW);
    const float4 v = *reinterpret_cast<const float4*>(x + pix * C + c);
    const float4 m = *reinterpret_cast<const float4*>(mult + (long)b * C + c);
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
    uint2 hv, lv;
    split_operand4(fmaf(v.x, m.x, be.x), fmaf(v.y, m.y, be.y), fmaf(v.z, m.z, be.z), fmaf(v.w, m.w, be.w), hv, lv);
    __nv_bfloat16* ep = opl + pix * ((long)(C >> 5) * 64) + (c >> 5) * 64 + (c & 31);
    *reinterpret_cast<uint2*>(ep) = hv;
    *reinterpret_cast<uint2*>(ep + 32) = lv;
  }
}

// ------------------------------------------------------------------------------------------
// head: x.mean([-2,-1]) -> LayerNorm(768) -> 7 Linear heads (convnextv2.py:110-144), one CTA per frame
// heads [B,328] = kp 63 | scale 1 | pitch 66 | yaw 66 | roll 66 | t 3 | exp 63 (registration order, convnextv2.py:95-103)
// ------------------------------------------------------------------------------------------
constexpr int HEAD_C = 768;
__global__ void __launch_bounds__(256) motion_head_kernel(const float* __restrict__ x, int HW, const float* __restrict__ nw,
                                                          const float* __restrict__ nb, const float* __restrict__ hw /*[328][768]*/,
                                                          const float* __restrict__ hb, float* __restrict__ heads, int n_out) {
  __shared__ float f[HEAD_C];
  __shared__ float red[8];
  __shared__ float s_stat[2];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float loc[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int c = tid + 256 * j;
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += x[((long)b * HW + p) * HEAD_C + c];
    loc[j] = s / (float)HW;
  }
  float part = warp_sum(loc[0] + loc[1] + loc[2]);
  if (lane == 0) red[wid] = part;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; s_stat[0] = t / (float)HEAD_C; }
  __syncthreads();
  const float mean = s_stat[0];
  part = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) part += (loc[j] - mean) * (loc[j] - mean);
  part = warp_sum(part);
  __syncthreads();
  if (lane == 0) red[wid] = part;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; s_stat[1] = 1.f / sqrtf(t / (float)HEAD_C + LN_EPS); }
  __syncthreads();
  const float r = s_stat[1];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int c = tid + 256 * j;
    f[c] = (loc[j] - mean) * r * nw[c] + nb[c];
  }
  __syncthreads();
  for (int o = wid; o < n_out; o += 8) {
    const float* wr = hw + (long)o * HEAD_C;
    float a = 0.f;
    for (int c = lane; c < HEAD_C; c += 32) a = fmaf(f[c], wr[c], a);
    a = warp_sum(a);
    if (lane == 0) heads[(long)b * n_out + o] = a + hb[o];
  }
}

// ------------------------------------------------------------------------------------------
// keypoints: one warp per frame
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float bins_to_degree(const float* __restrict__ logits, int lane) {
  // softmax over 66 bins, sum(p * idx) * 3 - 97.5   (camera.py:19-27)
  const float a = logits[lane], b = logits[lane + 32], c = lane < 2 ? logits[lane + 64] : -INFINITY;
  float m = fmaxf(fmaxf(a, b), c);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float ea = expf(a - m), eb = expf(b - m), ec = lane < 2 ? expf(c - m) : 0.f;
  const float den = warp_sum(ea + eb + ec);
  const float num = warp_sum(ea / den * (float)lane + eb / den * (float)(lane + 32) + ec / den * (float)(lane + 64));
  return num * 3.f - 97.5f;
}

__global__ void __launch_bounds__(32) keypoints_kernel(const float* __restrict__ heads, int n_out, float* __restrict__ x_s,
                                                       float* __restrict__ x_can, float* __restrict__ Rout, float* __restrict__ deg) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const float* h = heads + (long)b * n_out;
  const float* kp = h; const float scale = h[63];
  const float* t = h + 262; const float* ex = h + 265;
  const float PI = 3.14159265358979323846f;
  const float pitch = bins_to_degree(h + 64, lane), yaw = bins_to_degree(h + 130, lane), roll = bins_to_degree(h + 196, lane);
  const float x = pitch / 180.f * PI, y = yaw / 180.f * PI, z = roll / 180.f * PI;
  const float cx = cosf(x), sx = sinf(x), cy = cosf(y), sy = sinf(y), cz = cosf(z), sz = sinf(z);
  // rot = Rz Ry Rx (camera.py:52-72), returned transposed
  const float rzy[9] = {cz * cy, -sz, cz * sy, sz * cy, cz, sz * sy, -sy, 0.f, cy};          // Rz @ Ry
  float rot[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    rot[i * 3 + 0] = rzy[i * 3 + 0];
    rot[i * 3 + 1] = rzy[i * 3 + 1] * cx + rzy[i * 3 + 2] * sx;
    rot[i * 3 + 2] = -rzy[i * 3 + 1] * sx + rzy[i * 3 + 2] * cx;
  }
  // R = rot^T ;  kp @ R : out[j] = sum_i kp[i] * R[i][j] = sum_i kp[i] * rot[j][i]
  if (lane < 21) {
    const float k0 = kp[lane * 3], k1 = kp[lane * 3 + 1], k2 = kp[lane * 3 + 2];
    float o[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float v = (k0 * rot[j * 3] + k1 * rot[j * 3 + 1]) + k2 * rot[j * 3 + 2];
      o[j] = (v + ex[lane * 3 + j]) * scale;
    }
    o[0] += t[0]; o[1] += t[1];
    float* xs = x_s + ((long)b * 21 + lane) * 3;
    xs[0] = o[0]; xs[1] = o[1]; xs[2] = o[2];
    if (x_can) {
      float* xc = x_can + ((long)b * 21 + lane) * 3;
      xc[0] = scale * k0; xc[1] = scale * k1; xc[2] = scale * k2;
    }
  }
  if (Rout && lane < 9) Rout[(long)b * 9 + lane] = rot[(lane % 3) * 3 + lane / 3];
  if (deg && lane == 0) { deg[b * 3] = pitch; deg[b * 3 + 1] = yaw; deg[b * 3 + 2] = roll; }
}

inline unsigned warp_grid(long npix) {
  long blocks = (npix + 7) / 8;
  if (blocks > 148L * 32) blocks = 148L * 32;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

struct NameTable {
  std::map<std::string, const cs_tensor_desc*> m;
  const float* f32(const std::string& key, long numel) const {
    auto it = m.find("motion_extractor.detector." + key);
    if (it == m.end()) throw Error(CS_ERR_WEIGHTS, "missing tensor 'motion_extractor.detector." + key + "'");
    const cs_tensor_desc* d = it->second;
    long n = 1;
    for (int i = 0; i < d->ndim; ++i) n *= d->shape[i];
    if (d->dtype != CS_F32 || n != numel || !d->data)
      throw Error(CS_ERR_WEIGHTS, "tensor 'motion_extractor.detector." + key + "': expected " + std::to_string(numel) + " fp32 elements");
    return static_cast<const float*>(d->data);
  }
};

float* up(cs_ctx* ctx, const float* h, size_t n) {
  float* d = static_cast<float*>(ctx->dmalloc(n * sizeof(float)));
  CS_CUDA(cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice));
  return d;
}

}  // namespace

bool motion_weights_present(const cs_tensor_desc* table, int n) {
  for (int i = 0; i < n; ++i)
    if (table[i].name && std::string(table[i].name) == "motion_extractor.detector.norm.weight") return true;
  return false;
}

// combined_weights['motion_extractor'] (reference can_swap_e2e.py:94), keys as MotionExtractor.state_dict()
void load_motion_weights(cs_ctx* ctx, const cs_tensor_desc* table, int n) {
  NameTable t;
  for (int i = 0; i < n; ++i) if (table[i].name) t.m[table[i].name] = &table[i];
  MotionW& M = ctx->M;
  {  // stem conv [96][3][4][4] -> [k = (kh*4+kw)*3 + ci][96]
    const float* w = t.f32("downsample_layers.0.0.weight", 96L * 48);
    std::vector<float> r(48 * 96);
    for (int co = 0; co < 96; ++co)
      for (int ci = 0; ci < 3; ++ci)
        for (int kk = 0; kk < 16; ++kk) r[(size_t)(kk * 3 + ci) * 96 + co] = w[((long)co * 3 + ci) * 16 + kk];
    M.stem_w = up(ctx, r.data(), r.size());
    M.stem_b = up(ctx, t.f32("downsample_layers.0.0.bias", 96), 96);
    M.stem_ln_w = up(ctx, t.f32("downsample_layers.0.1.weight", 96), 96);
    M.stem_ln_b = up(ctx, t.f32("downsample_layers.0.1.bias", 96), 96);
  }
  for (int i = 0; i < 3; ++i) {  // downsample: LN(C) + Conv2d(C, C2, k=2, s=2) as a 1x1 conv over 4C channels ordered (kh, kw, ci)
    const int C = M_DIMS[i], C2 = M_DIMS[i + 1];
    const std::string p = "downsample_layers." + std::to_string(i + 1);
    M.ds_ln_w[i] = up(ctx, t.f32(p + ".0.weight", C), C);
    M.ds_ln_b[i] = up(ctx, t.f32(p + ".0.bias", C), C);
    const float* w = t.f32(p + ".1.weight", (long)C2 * C * 4);
    std::vector<float> r((size_t)C2 * 4 * C);
    for (int co = 0; co < C2; ++co)
      for (int ci = 0; ci < C; ++ci)
        for (int kk = 0; kk < 4; ++kk) r[((size_t)co * 4 + kk) * C + ci] = w[((long)co * C + ci) * 4 + kk];
    const float* b = t.f32(p + ".1.bias", C2);
    std::vector<float> bias(b, b + C2);
    M.ds[i] = pack_conv_host(ctx, r, &bias, C2, 4 * C, 1, 1, 1);
  }
  int nb = 0;
  for (int i = 0; i < 4; ++i) {
    const int C = M_DIMS[i];
    for (int j = 0; j < M_DEPTHS[i]; ++j, ++nb) {
      const std::string p = "stages." + std::to_string(i) + "." + std::to_string(j);
      MotionBlockW& k = M.blk[nb];
      const float* dw = t.f32(p + ".dwconv.weight", (long)C * 49);           // [C][1][7][7] -> [49][C]
      std::vector<float> r((size_t)49 * C);
      for (int c = 0; c < C; ++c)
        for (int kk = 0; kk < 49; ++kk) r[(size_t)kk * C + c] = dw[(long)c * 49 + kk];
      k.dw_w = up(ctx, r.data(), r.size());
      k.dw_b = up(ctx, t.f32(p + ".dwconv.bias", C), C);
      k.ln_w = up(ctx, t.f32(p + ".norm.weight", C), C);
      k.ln_b = up(ctx, t.f32(p + ".norm.bias", C), C);
      {
        const float* w = t.f32(p + ".pwconv1.weight", 4L * C * C);
        const float* b = t.f32(p + ".pwconv1.bias", 4L * C);
        std::vector<float> wv(w, w + 4L * C * C), bv(b, b + 4L * C);
        k.pw1 = pack_conv_host(ctx, wv, &bv, 4 * C, C, 1, 1, 1);
      }
      k.grn_g = up(ctx, t.f32(p + ".grn.gamma", 4L * C), 4 * C);
      k.grn_b = up(ctx, t.f32(p + ".grn.beta", 4L * C), 4 * C);
      {
        const float* w = t.f32(p + ".pwconv2.weight", 4L * C * C);
        const float* b = t.f32(p + ".pwconv2.bias", C);
        std::vector<float> wv(w, w + 4L * C * C), bv(b, b + C);
        k.pw2 = pack_conv_host(ctx, wv, &bv, C, 4 * C, 1, 1, 1);
      }
    }
  }
  M.norm_w = up(ctx, t.f32("norm.weight", 768), 768);
  M.norm_b = up(ctx, t.f32("norm.bias", 768), 768);
  {
    static const char* names[7] = {"fc_kp", "fc_scale", "fc_pitch", "fc_yaw", "fc_roll", "fc_t", "fc_exp"};
    static const int widths[7] = {63, 1, 66, 66, 66, 3, 63};
    std::vector<float> w((size_t)CS_MOTION_HEADS * 768), b(CS_MOTION_HEADS);
    int o = 0;
    for (int i = 0; i < 7; ++i) {
      const float* hw = t.f32(std::string(names[i]) + ".weight", (long)widths[i] * 768);
      const float* hb = t.f32(std::string(names[i]) + ".bias", widths[i]);
      std::copy(hw, hw + (long)widths[i] * 768, w.begin() + (size_t)o * 768);
      std::copy(hb, hb + widths[i], b.begin() + o);
      o += widths[i];
    }
    M.head_w = up(ctx, w.data(), w.size());
    M.head_b = up(ctx, b.data(), b.size());
  }
  const size_t nsq = (size_t)ctx->max_batch * 3072;
  M.sumsq = static_cast<double*>(ctx->dmalloc(nsq * sizeof(double)));
  CS_CUDA(cudaMemset(M.sumsq, 0, nsq * sizeof(double)));
  M.loaded = true;
}

// 1x1 conv on the tcgen05 kernel from an operand that is already in split form
static void linear_tc(Net& n, const Opd& opd, const ConvW& w, int act, const Act* residual, Act out) {
  ConvGeom g;
  g.Do = 1; g.Ho = out.H; g.Wo = out.W;
  Epilogue e;
  e.act = act;
  if (residual) { e.residual = residual->p; e.rs_b = residual->sb; e.rs_d = residual->sd; e.rs_h = residual->sh; e.rs_w = residual->sw; }
  conv_tc(n.L, opd, w, g, e, out);
}

// MotionExtractor.forward: img_cl [B,H,W,3] fp32 in [0,1] -> heads [B,328]
void run_motion(Net& n, const float* img_cl, int B, float* heads) {
  const MotionW& M = n.ctx->M;
  CS_REQUIRE(M.loaded, CS_ERR_STATE, "motion extractor weights not loaded (no 'motion_extractor.*' tensors in cs_load_weights)");
  n.L.tag = "motion";
  const int H0 = n.ctx->net_h, W0 = n.ctx->net_w;
  size_t m0 = n.A->mark();
  int H = H0 / 4, W = W0 / 4;
  float* xa = n.A->f32((size_t)B * H * W * 96);         // stage buffers shrink by 2x per stage: reuse the first two
  float* xb = n.A->f32((size_t)B * H * W * 96 / 2);
  {
    n.L.count();
    if (!n.L.dry) {
      ProfScope ps(n.L, PK_OTHER, 0.0, (double)B * H0 * W0 * 3 * 4 + (double)B * H * W * 96 * 4, "stem_ln");
      stem_ln_kernel<<<warp_grid((long)B * H * W), 256, 0, n.L.stream>>>(img_cl, M.stem_w, M.stem_b, M.stem_ln_w, M.stem_ln_b, xa, B,
                                                                         H0, W0);
      check_launch("stem_ln");
    }
  }
  float* x = xa;
  int nb = 0;
  for (int s = 0; s < 4; ++s) {
    const int C = M_DIMS[s];
    if (s > 0) {                                        // downsample_layers[s]: LN + 2x2 stride-2 conv (convnextv2.py:79-83)
      const int Cp = M_DIMS[s - 1];
      size_t m = n.A->mark();
      Act geom = make_act(nullptr, B, 1, H / 2, W / 2, C);
      Opd opd = conv_tc_alloc_operand(*n.A, M.ds[s - 1], geom);
      n.L.count();
      if (!n.L.dry) {
        ProfScope ps(n.L, PK_PREP, 0.0, (double)B * H * W * Cp * 8.0, "ln_s2d");
        dw_ln_kernel<false, true><<<warp_grid((long)B * H * W), 256, 0, n.L.stream>>>(x, nullptr, nullptr, M.ds_ln_w[s - 1],
                                                                                      M.ds_ln_b[s - 1], opd.p, B, H, W, Cp);
        check_launch("ln_s2d");
      }
      H /= 2; W /= 2;
      float* xn = (x == xa) ? xb : xa;
      linear_tc(n, opd, M.ds[s - 1], ACT_NONE, nullptr, make_act(xn, B, 1, H, W, C));
      x = xn;
      n.A->reset(m);
    }
    const long npix = (long)B * H * W;
    Act xact = make_act(x, B, 1, H, W, C);
    for (int j = 0; j < M_DEPTHS[s]; ++j, ++nb) {       // Block.forward, convnextv2.py:34-47
      const MotionBlockW& k = M.blk[nb];
      size_t m = n.A->mark();
      Opd o1 = conv_tc_alloc_operand(*n.A, k.pw1, xact);
      n.L.count();
      if (!n.L.dry) {
        ProfScope ps(n.L, PK_PREP, 0.0, (double)npix * C * 8.0, "dw_ln");
        dw_ln_kernel<true, false><<<warp_grid(npix), 256, 0, n.L.stream>>>(x, k.dw_w, k.dw_b, k.ln_w, k.ln_b, o1.p, B, H, W, C);
        check_launch("dw_ln");
      }
      Act hact = make_act(n.A->f32((size_t)npix * 4 * C), B, 1, H, W, 4 * C);
      linear_tc(n, o1, k.pw1, ACT_GELU, nullptr, hact);
      float* mult = n.A->f32((size_t)B * 4 * C);
      Opd o2 = conv_tc_alloc_operand(*n.A, k.pw2, hact);
      n.L.count(); n.L.count(); n.L.count();
      if (!n.L.dry) {
        const int HW = H * W;
        const int slab = HW >= 1024 ? 64 : (HW >= 256 ? 16 : 8);
        {
          ProfScope ps(n.L, PK_STATS, 0.0, (double)npix * 4 * C * 4.0, "grn_sumsq");
          grn_sumsq_kernel<<<dim3((unsigned)((HW + slab - 1) / slab), (unsigned)B), 256, 0, n.L.stream>>>(hact.p, n.grn, HW, 4 * C, slab);
          check_launch("grn_sumsq");
        }
        grn_finalize_kernel<<<(unsigned)B, 256, 0, n.L.stream>>>(n.grn, k.grn_g, mult, 4 * C);
        check_launch("grn_finalize");
        {
          ProfScope ps(n.L, PK_PREP, 0.0, (double)npix * 4 * C * 8.0, "grn_apply");
          long blocks = (npix * C + 255) / 256; if (blocks > 148L * 16) blocks = 148L * 16;
          grn_apply_kernel<<<(unsigned)blocks, 256, 0, n.L.stream>>>(hact.p, mult, k.grn_b, o2.p, npix, HW, 4 * C);
          check_launch("grn_apply");
        }
      }
      linear_tc(n, o2, k.pw2, ACT_NONE, &xact, xact);   // x = input + pwconv2(.), in place
      n.A->reset(m);
    }
  }
  n.L.count();
  if (!n.L.dry) {
    ProfScope ps(n.L, PK_OTHER, 0.0, (double)B * H * W * 768 * 4.0, "motion_head");
    motion_head_kernel<<<(unsigned)B, 256, 0, n.L.stream>>>(x, H * W, M.norm_w, M.norm_b, M.head_w, M.head_b, heads, CS_MOTION_HEADS);
    check_launch("motion_head");
  }
  n.A->reset(m0);
}

// transform_keypoint (+ x_can, R, angles) from the raw heads
void run_keypoints(Net& n, const float* heads, int B, float* x_s, float* x_can, float* R, float* deg) {
  n.L.count();
  if (n.L.dry) return;
  keypoints_kernel<<<(unsigned)B, 32, 0, n.L.stream>>>(heads, CS_MOTION_HEADS, x_s, x_can, R, deg);
  check_launch("keypoints");
}

}  // namespace cs
