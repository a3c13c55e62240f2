// Dense-motion sampling kernels (HBM-bound, sm_100a).
//   dm_input           reference dense_motion.py:29-65,83-84: for every voxel and keypoint k build
//                      [heatmap_k, trilinear sample of the 4-ch compressed feature at the k-th
//                      pure-translation grid] -> hourglass input channel k*5 + {0..4}
//   softmax_flow_warp  reference dense_motion.py:89-94 + warping_network.py:46-47,61: softmax over the
//                      22 mask logits, flow = sum_k mask_k * motion_k, then the 5-D trilinear
//                      grid_sample (zeros padding, align_corners=False) of the 32x16 feature volume
//   grid_sample3d_cl   the bare 5-D grid_sample on the [B,H,W,16,32] volume (unit test entry)
// Sampling semantics follow ATen grid_sampler_3d: ix = ((x+1)*W-1)/2, floor, 8 corners, zeros outside.
#include "common.cuh"

namespace cs {

__device__ __forceinline__ float unnorm(float c, int size) { return ((c + 1.f) * size - 1.f) / 2.f; }

struct Tri {           // trilinear footprint
  int x0, y0, z0;
  float w[8];          // order: tnw, tne, tsw, tse, bnw, bne, bsw, bse  (t = z0, n = y0, w = x0)
};

__device__ __forceinline__ Tri make_tri(float gx, float gy, float gz, int D, int H, int W) {
  Tri t;
  float ix = unnorm(gx, W), iy = unnorm(gy, H), iz = unnorm(gz, D);
  float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  t.x0 = (int)fx; t.y0 = (int)fy; t.z0 = (int)fz;
  float ex = fx + 1.f, ey = fy + 1.f, ez = fz + 1.f;     // east / south / bottom corner coordinates
  float dxw = ex - ix, dxe = ix - fx;
  float dyn = ey - iy, dys = iy - fy;
  float dzt = ez - iz, dzb = iz - fz;
  t.w[0] = dxw * dyn * dzt;  // tnw
  t.w[1] = dxe * dyn * dzt;  // tne
  t.w[2] = dxw * dys * dzt;  // tsw
  t.w[3] = dxe * dys * dzt;  // tse
  t.w[4] = dxw * dyn * dzb;  // bnw
  t.w[5] = dxe * dyn * dzb;  // bne
  t.w[6] = dxw * dys * dzb;  // bsw
  t.w[7] = dxe * dys * dzb;  // bse
  return t;
}

__device__ __forceinline__ void grid_xyz(int d, int h, int w, int D, int H, int W, float& gx, float& gy, float& gz) {
  // make_coordinate_grid, reference util.py:41-50
  gx = 2.f * ((float)w / (float)(W - 1)) - 1.f;
  gy = 2.f * ((float)h / (float)(H - 1)) - 1.f;
  gz = 2.f * ((float)d / (float)(D - 1)) - 1.f;
}

// ------------------------------------------------------------------------------------------
// dm_input: one thread per (voxel, k)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dm_input_kernel(const float4* __restrict__ c4, const float* __restrict__ kpd,
                                                      const float* __restrict__ kps, int K, int B, int D, int H, int W,
                                                      float* __restrict__ out, long ob, long od, long oh, long ow) {
  long total = (long)B * D * H * W * (K + 1);
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    int k = (int)(idx % (K + 1)); long pix = idx / (K + 1);
    int w = (int)(pix % W); long t = pix / W; int h = (int)(t % H); t /= H; int d = (int)(t % D); int b = (int)(t / D);
    float gx, gy, gz;
    grid_xyz(d, h, w, D, H, W, gx, gy, gz);
    float mx = gx, my = gy, mz = gz, heat = 0.f;
    if (k > 0) {
      const float* pd = kpd + ((long)b * K + (k - 1)) * 3;
      const float* ps = kps + ((long)b * K + (k - 1)) * 3;
      float dx = gx - pd[0], dy = gy - pd[1], dz = gz - pd[2];          // identity_grid - kp_driving
      mx = dx + ps[0]; my = dy + ps[1]; mz = dz + ps[2];                 // + kp_source
      float sx = gx - ps[0], sy = gy - ps[1], sz = gz - ps[2];
      float qd = (dx * dx + dy * dy) + dz * dz, qs = (sx * sx + sy * sy) + sz * sz;
      heat = expf(-0.5f * qd / 0.01f) - expf(-0.5f * qs / 0.01f);
    }
    Tri tr = make_tri(mx, my, mz, D, H, W);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int cz = 0; cz < 2; ++cz)
#pragma unroll
      for (int cy = 0; cy < 2; ++cy)
#pragma unroll
        for (int cx = 0; cx < 2; ++cx) {
          int x = tr.x0 + cx, y = tr.y0 + cy, z = tr.z0 + cz;
          if (x >= 0 && x < W && y >= 0 && y < H && z >= 0 && z < D) {
            float wt = tr.w[cz * 4 + cy * 2 + cx];
            float4 v = c4[(((long)b * D + z) * H + y) * W + x];
            acc.x += v.x * wt; acc.y += v.y * wt; acc.z += v.z * wt; acc.w += v.w * wt;
          }
        }
    float* o = out + b * ob + d * od + h * oh + w * ow + k * 5;
    o[0] = heat; o[1] = acc.x; o[2] = acc.y; o[3] = acc.z; o[4] = acc.w;
  }
}

void dm_input(const Launcher& L, const Act& c4, const float* kp_driving, const float* kp_source, int K, Act out) {
  L.count();
  if (L.dry) return;
  CS_REQUIRE(c4.C == 4 && c4.sw == 4, -1, "dm_input: compressed feature must be dense 4-channel");
  long total = c4.pixels() * (K + 1);
  long blocks = (total + 255) / 256; if (blocks > 148L * 32) blocks = 148L * 32;
  ProfScope ps(L, PK_SAMPLE, 0.0, (double)c4.pixels() * (4 + 110) * 4.0, "dm_input");
  dm_input_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(reinterpret_cast<const float4*>(c4.p), kp_driving, kp_source, K,
                                                         c4.B, c4.D, c4.H, c4.W, out.p, out.sb, out.sd, out.sh, out.sw);
  check_launch("dm_input");
}

// The same tensor written directly as the split-fp16 operand [pixel][4 blocks][hi 32 | lo 32] (channels 110..127 zero) of the
// two convs that consume it (hourglass conv0 and, as blocks 1..4 of its 142-channel input, the hourglass' last conv): no fp32
// copy of the 110-channel tensor, no prep passes.
// A block takes 32 consecutive voxels (adjacent in w).  Phase 1: a warp = one keypoint x the 32 voxels -- the motion of a
// keypoint is a pure translation, so the 32 lanes gather 32 ADJACENT float4 of the compressed feature per corner (4..5 lines
// instead of the 32 lines of the one-warp-per-voxel form) and leave their 5 channels in a shared tile [32][128 + 4].
// Phase 2: a warp per voxel splits 4 channels per lane and writes 4 full 128-byte rows.
constexpr int DMI_VOX = 32, DMI_LD = 132;

__global__ void __launch_bounds__(256) dm_input_operand_kernel(const float4* __restrict__ c4, const float* __restrict__ kpd,
                                                               const float* __restrict__ kps, int K, int B, int D, int H, int W,
                                                               __nv_bfloat16* __restrict__ out, long prow, float amul) {
  __shared__ __align__(16) float vals[DMI_VOX * DMI_LD];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long nvox = (long)B * D * H * W;
  for (int i = threadIdx.x; i < DMI_VOX * DMI_LD; i += 256) vals[i] = 0.f;        // channels 110..127 stay zero
  const int items = (K + 1) * DMI_VOX;
  for (long v0 = (long)blockIdx.x * DMI_VOX; v0 < nvox; v0 += (long)gridDim.x * DMI_VOX) {
    __syncthreads();                                       // the previous tile has been written out (and the zero fill)
    // a thread keeps the same voxel (lane) for all its keypoints: decode it once, with 32-bit divisions
    const unsigned upix = (unsigned)v0 + (unsigned)lane;
    const unsigned t1 = upix / (unsigned)W, t2 = t1 / (unsigned)H;
    const int w = (int)(upix - t1 * (unsigned)W), h = (int)(t1 - t2 * (unsigned)H);
    const int b = (int)(t2 / (unsigned)D), d = (int)(t2 - (unsigned)b * (unsigned)D);
    float gx, gy, gz;
    grid_xyz(d, h, w, D, H, W, gx, gy, gz);
    for (int it = threadIdx.x; it < items; it += 256) {
      const int k = it >> 5;
      if (v0 + lane >= nvox) break;
      float mx = gx, my = gy, mz = gz, heat = 0.f;
      if (k > 0) {
        const float* pd = kpd + ((long)b * K + (k - 1)) * 3;
        const float* ps = kps + ((long)b * K + (k - 1)) * 3;
        float dx = gx - pd[0], dy = gy - pd[1], dz = gz - pd[2];          // identity_grid - kp_driving
        mx = dx + ps[0]; my = dy + ps[1]; mz = dz + ps[2];                 // + kp_source
        float sx = gx - ps[0], sy = gy - ps[1], sz = gz - ps[2];
        float qd = (dx * dx + dy * dy) + dz * dz, qs = (sx * sx + sy * sy) + sz * sz;
        heat = expf(-0.5f * qd / 0.01f) - expf(-0.5f * qs / 0.01f);
      }
      Tri tr = make_tri(mx, my, mz, D, H, W);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int cz = 0; cz < 2; ++cz)
#pragma unroll
        for (int cy = 0; cy < 2; ++cy)
#pragma unroll
          for (int cx = 0; cx < 2; ++cx) {
            int x = tr.x0 + cx, y = tr.y0 + cy, z = tr.z0 + cz;
            if (x >= 0 && x < W && y >= 0 && y < H && z >= 0 && z < D) {
              float wt = tr.w[cz * 4 + cy * 2 + cx];
              float4 v = c4[((b * D + z) * H + y) * W + x];                 // < 2^31 voxels (host check)
              acc.x += v.x * wt; acc.y += v.y * wt; acc.z += v.z * wt; acc.w += v.w * wt;
            }
          }
      float* o = vals + (it & 31) * DMI_LD + k * 5;
      o[0] = heat; o[1] = acc.x; o[2] = acc.y; o[3] = acc.z; o[4] = acc.w;
    }
    __syncthreads();
    const int c = lane * 4;
#pragma unroll
    for (int q = 0; q < DMI_VOX / 8; ++q) {
      const int v = wid + 8 * q;
      const long pix = v0 + v;
      if (pix >= nvox) continue;
      const float4 f = *reinterpret_cast<const float4*>(vals + v * DMI_LD + c);
      uint2 hv, lv;
      split_operand4(f.x * amul, f.y * amul, f.z * amul, f.w * amul, hv, lv);
      __nv_bfloat16* o = out + pix * prow + (c >> 5) * 64 + (c & 31);
      *reinterpret_cast<uint2*>(o) = hv;
      *reinterpret_cast<uint2*>(o + 32) = lv;
    }
  }
}

void dm_input_operand(const Launcher& L, const Act& c4, const float* kp_driving, const float* kp_source, int K, Opd out) {
  L.count();
  if (L.dry) return;
  CS_REQUIRE(c4.C == 4 && c4.sw == 4 && (K + 1) * 5 <= 128 && K < 32 && out.nblk == 4 && out.B == c4.B && out.D == c4.D && out.H == c4.H &&
                 out.W == c4.W, -1, "dm_input_operand: bad geometry");
  const long nvox = c4.pixels();
  CS_REQUIRE(nvox < (1L << 31), -1, "dm_input_operand: volume too large for 32-bit index math");
  long blocks = (nvox + DMI_VOX - 1) / DMI_VOX; if (blocks > 148L * 8) blocks = 148L * 8;
  ProfScope ps(L, PK_SAMPLE, 0.0, (double)nvox * (4 + 128) * 4.0, "dm_input");
  dm_input_operand_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(reinterpret_cast<const float4*>(c4.p), kp_driving, kp_source, K, c4.B,
                                                                 c4.D, c4.H, c4.W, out.p, out.row(), out.amul);
  check_launch("dm_input_operand");
}

// ------------------------------------------------------------------------------------------
// sample the [B,H,W,16,32] volume: 8 threads per voxel, 4 channels (one float4) each
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 sample_vol(const float* __restrict__ vol, int b, const Tri& tr, int D, int H, int W, int c4) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int cz = 0; cz < 2; ++cz)
#pragma unroll
    for (int cy = 0; cy < 2; ++cy)
#pragma unroll
      for (int cx = 0; cx < 2; ++cx) {
        int x = tr.x0 + cx, y = tr.y0 + cy, z = tr.z0 + cz;
        if (x >= 0 && x < W && y >= 0 && y < H && z >= 0 && z < D) {
          float wt = tr.w[cz * 4 + cy * 2 + cx];
          float4 v = *reinterpret_cast<const float4*>(vol + (((long)b * H + y) * W + x) * 512 + z * 32 + c4 * 4);
          acc.x += v.x * wt; acc.y += v.y * wt; acc.z += v.z * wt; acc.w += v.w * wt;
        }
      }
  return acc;
}

// 8 lanes per voxel: the lanes SHARE one softmax (lane j takes classes j, j+8, j+16; max / sums by xor-shuffles over the
// 8-lane group) instead of each recomputing all 22 exponentials, the keypoint translations ks_k - kd_k of the block's
// sample sit in shared memory, and every lane then gathers its float4 (4 of the 32 channels) of the 8 corners: a corner is
// one 128-byte line for the group.  Voxel order (b, h, w, d): the 16 depths of a pixel are adjacent in the output volume, so
// a warp (4 voxels) writes 512 contiguous bytes.  A block = 32 voxels = 2 pixels x 16 depths of one sample.
constexpr int SFW_MAXK = 32;

__global__ void __launch_bounds__(256) softmax_flow_warp_kernel(const float* __restrict__ logits, long lb, long ld, long lh, long lw,
                                                               const float* __restrict__ kpd, const float* __restrict__ kps,
                                                               int K, int B, int D, int H, int W,
                                                               const float* __restrict__ vol, float* __restrict__ out,
                                                               float* __restrict__ deformation) {
  __shared__ float s_kd[SFW_MAXK * 3], s_ks[SFW_MAXK * 3];
  const long nvox = (long)B * D * H * W;
  const long per_b = (long)D * H * W;
  const int j = threadIdx.x & 7;                              // lane within the voxel group = channel quad
  for (long v0 = (long)blockIdx.x * 32; v0 < nvox; v0 += (long)gridDim.x * 32) {
    const int b = (int)(v0 / per_b);                          // per_b is a multiple of 32: the block's voxels share b
    __syncthreads();
    if (threadIdx.x < K * 3) {
      s_kd[threadIdx.x] = kpd[(long)b * K * 3 + threadIdx.x];
      s_ks[threadIdx.x] = kps[(long)b * K * 3 + threadIdx.x];
    }
    __syncthreads();
    const long pix = v0 + (threadIdx.x >> 3);
    const unsigned up = (unsigned)pix, u1 = up / (unsigned)D, u2 = u1 / (unsigned)W;       // 32-bit divisions
    const int d = (int)(up - u1 * (unsigned)D), w = (int)(u1 - u2 * (unsigned)W), h = (int)(u2 % (unsigned)H);
    const float* lg = logits + b * lb + d * ld + h * lh + w * lw;
    float gx, gy, gz;
    grid_xyz(d, h, w, D, H, W, gx, gy, gz);
    // this lane's classes
    float l0 = lg[j], l1 = (j + 8 <= K) ? lg[j + 8] : -INFINITY, l2 = (j + 16 <= K) ? lg[j + 16] : -INFINITY;
    float mxl = fmaxf(l0, fmaxf(l1, l2));
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) mxl = fmaxf(mxl, __shfl_xor_sync(0xffffffffu, mxl, o));
    float den = 0.f, fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int k = j + 8 * q;
      if (k <= K) {
        const float e = expf((q == 0 ? l0 : (q == 1 ? l1 : l2)) - mxl);
        den += e;
        float mx = gx, my = gy, mz = gz;
        if (k > 0) {                                          // identity_grid - kp_driving + kp_source (dense_motion.py:38-40)
          mx = (gx - s_kd[(k - 1) * 3 + 0]) + s_ks[(k - 1) * 3 + 0];
          my = (gy - s_kd[(k - 1) * 3 + 1]) + s_ks[(k - 1) * 3 + 1];
          mz = (gz - s_kd[(k - 1) * 3 + 2]) + s_ks[(k - 1) * 3 + 2];
        }
        fx = fmaf(mx, e, fx); fy = fmaf(my, e, fy); fz = fmaf(mz, e, fz);
      }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {                         // xor butterflies: every lane of the group ends with the same sums
      den += __shfl_xor_sync(0xffffffffu, den, o);
      fx += __shfl_xor_sync(0xffffffffu, fx, o);
      fy += __shfl_xor_sync(0xffffffffu, fy, o);
      fz += __shfl_xor_sync(0xffffffffu, fz, o);
    }
    const float inv = 1.f / den;
    fx *= inv; fy *= inv; fz *= inv;
    if (deformation && j == 0) {
      float* df = deformation + ((((long)b * D + d) * H + h) * W + w) * 3;
      df[0] = fx; df[1] = fy; df[2] = fz;
    }
    const Tri tr = make_tri(fx, fy, fz, D, H, W);
    const float4 v = sample_vol(vol, b, tr, D, H, W, j);
    *reinterpret_cast<float4*>(out + (((long)b * H + h) * W + w) * 512 + d * 32 + j * 4) = v;
  }
}

void softmax_flow_warp(const Launcher& L, const Act& logits, const float* kp_driving, const float* kp_source, int K,
                       const float* vol, float* out, float* deformation) {
  L.count();
  if (L.dry) return;
  CS_REQUIRE(logits.D == 16 && logits.C >= K + 1 && K + 1 <= 24 && K <= SFW_MAXK && (logits.H * logits.W) % 2 == 0, -1,
             "softmax_flow_warp: bad logits tensor");
  const long nvox = logits.pixels();
  CS_REQUIRE(nvox < (1L << 31), -1, "softmax_flow_warp: volume too large for 32-bit index math");
  long blocks = nvox / 32; if (blocks > 148L * 32) blocks = 148L * 32;
  // algorithmic bytes / voxel: 32 ch in + 32 ch out + 22 logits (22.5 MB / sample)
  ProfScope ps(L, PK_SAMPLE, 0.0, (double)logits.pixels() * (32 + 32 + (K + 1)) * 4.0, "flow_warp");
  softmax_flow_warp_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(logits.p, logits.sb, logits.sd, logits.sh, logits.sw,
                                                                  kp_driving, kp_source, K, logits.B, logits.D, logits.H,
                                                                  logits.W, vol, out, deformation);
  check_launch("softmax_flow_warp");
}

__global__ void __launch_bounds__(256) grid_sample3d_cl_kernel(const float* __restrict__ vol, const float* __restrict__ grid,
                                                              float* __restrict__ out, int B, int D, int H, int W) {
  long total = (long)B * D * H * W * 8;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    int c4 = (int)(idx & 7); long pix = idx >> 3;
    int d = (int)(pix % D); long t = pix / D; int w = (int)(t % W); t /= W; int h = (int)(t % H); int b = (int)(t / H);
    const float* g = grid + ((((long)b * D + d) * H + h) * W + w) * 3;
    Tri tr = make_tri(g[0], g[1], g[2], D, H, W);
    float4 v = sample_vol(vol, b, tr, D, H, W, c4);
    *reinterpret_cast<float4*>(out + (((long)b * H + h) * W + w) * 512 + d * 32 + c4 * 4) = v;
  }
}

// ------------------------------------------------------------------------------------------
// occlusion map from per-tap projections (reference dense_motion.py:98-102).
// The 7x7 conv over 142*16 channels is split into (1) a 1x1x1 tcgen05 GEMM with depth-dependent weights,
// Y[b,z,h,w,kh*7+kw] = sum_c pred[b,z,h,w,c] * Wocc[c*16+z, kh, kw], and (2) this gather:
// occ[b,h,w] = sigmoid(bias + sum_{z,kh,kw} Y[b,z,h+kh-3,w+kw-3,kh*7+kw]).  One warp = 32 consecutive w of one
// (b,h); per (z,kh) lane l loads the 7 kw-columns of row w0+l-3 (28 contiguous bytes), lanes 0..5 also rows
// w0+29..w0+34, and the diagonal sum is formed with shuffles: every Y element is read once.
// ------------------------------------------------------------------------------------------
// One block = 32 consecutive w of one (b,h); warp j takes the depth slices j, j+16, .. and the partial sums of the warps are
// added in warp order (fixed: deterministic and independent of the batch).
constexpr int OCC_WARPS = 16;

__global__ void __launch_bounds__(32 * OCC_WARPS) occlusion_gather_kernel(const float* __restrict__ Y, const float* __restrict__ bias,
                                                                          float* __restrict__ occ, int B, int D, int H, int W, int ldy) {
  __shared__ float part[OCC_WARPS][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int wtiles = (W + 31) / 32;
  const long tile = blockIdx.x;
  const int wt = (int)(tile % wtiles); const long r = tile / wtiles;
  const int h = (int)(r % H); const int b = (int)(r / H);
  const int w0 = wt * 32;
  const int xa = w0 + lane - 3;                 // row held in v[]
  const int xb = w0 + 32 + lane - 3;            // row held in e[] (lanes 0..5)
  const bool va = xa >= 0 && xa < W, vb = lane < 6 && xb < W;
  float acc = 0.f;
  for (int z = wid; z < D; z += OCC_WARPS) {
    for (int kh = 0; kh < 7; ++kh) {
      const int y = h + kh - 3;
      if (y < 0 || y >= H) continue;            // warp-uniform
      const float* row = Y + ((((long)b * D + z) * H + y) * W) * ldy + kh * 7;
      float v[7], e[7];
#pragma unroll
      for (int kw = 0; kw < 7; ++kw) {
        v[kw] = va ? __ldg(row + (long)xa * ldy + kw) : 0.f;
        e[kw] = vb ? __ldg(row + (long)xb * ldy + kw) : 0.f;
      }
#pragma unroll
      for (int kw = 0; kw < 7; ++kw) {
        // output w0+l needs row index (l + kw) of the 38-row window
        const int src = (lane + kw) & 31;
        const float fv = __shfl_sync(0xffffffffu, v[kw], src);
        const float fe = __shfl_sync(0xffffffffu, e[kw], src);
        acc += (lane + kw < 32) ? fv : fe;
      }
    }
  }
  part[wid][lane] = acc;
  __syncthreads();
  if (wid == 0) {
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < OCC_WARPS; ++j) sum += part[j][lane];
    const int w = w0 + lane;
    if (w < W) occ[((long)b * H + h) * W + w] = 1.f / (1.f + expf(-(sum + (bias ? bias[0] : 0.f))));
  }
}

void occlusion_gather(const Launcher& L, const Act& Y, const float* bias, float* occ) {
  L.count();
  if (L.dry) return;
  CS_REQUIRE(Y.C >= 49 && Y.sh == (long)Y.W * Y.sw && Y.sd == (long)Y.H * Y.sh && Y.sb == (long)Y.D * Y.sd, -1,
             "occlusion_gather: Y must be dense [B,D,H,W,>=49]");
  const long tiles = (long)Y.B * Y.H * ((Y.W + 31) / 32);
  ProfScope ps(L, PK_SAMPLE, 0.0, (double)Y.pixels() * 49 * 4.0, "occ_gather");
  occlusion_gather_kernel<<<(unsigned)tiles, 32 * OCC_WARPS, 0, L.stream>>>(Y.p, bias, occ, Y.B, Y.D, Y.H, Y.W, (int)Y.sw);
  check_launch("occlusion_gather");
}

void grid_sample3d_cl(const Launcher& L, const float* vol, const float* grid, float* out, int B, int D, int H, int W) {
  L.count();
  if (L.dry) return;
  CS_REQUIRE(D == 16, -1, "grid_sample3d_cl: depth must be 16");
  long total = (long)B * D * H * W * 8;
  long blocks = (total + 255) / 256; if (blocks > 148L * 32) blocks = 148L * 32;
  grid_sample3d_cl_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(vol, grid, out, B, D, H, W);
  check_launch("grid_sample3d_cl");
}

}  // namespace cs
