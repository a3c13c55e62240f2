// Internal definitions shared by the canonswap_b200 kernels and host code (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <map>
#include <stdexcept>

namespace cs {

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CS_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      throw cs::Error(-2, std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " +    \
                              __FILE__ + ":" + std::to_string(__LINE__));                  \
  } while (0)

#define CS_REQUIRE(cond, code, msg)                                                        \
  do {                                                                                     \
    if (!(cond)) throw cs::Error((code), std::string(msg) + " [" #cond "]");               \
  } while (0)

// ------------------------------------------------------------------------------------------
// activation tensors: fp32, channels-last, explicit strides (in elements). Channel stride is 1.
// 2-D tensors use D = 1. C is the logical channel count; buffers may carry zero pad channels.
// ------------------------------------------------------------------------------------------
struct Act {
  float* p = nullptr;
  int B = 0, D = 1, H = 0, W = 0, C = 0;
  long sb = 0, sd = 0, sh = 0, sw = 0;
  long pixels() const { return (long)B * D * H * W; }
};

inline Act make_act(float* p, int B, int D, int H, int W, int C, int Cstride = -1) {
  Act a;
  if (Cstride < 0) Cstride = C;
  a.p = p; a.B = B; a.D = D; a.H = H; a.W = W; a.C = C;
  a.sw = Cstride; a.sh = (long)W * a.sw; a.sd = (long)H * a.sh; a.sb = (long)D * a.sd;
  return a;
}

// The 32-channel x 16-depth feature volume lives as [B,H,W,16,32]: a 2-D tensor with 512
// channels ordered ch' = d*32 + c (the reference's view order is ch = c*16 + d).
inline Act vol_as_3d(float* p, int B, int H, int W) {
  Act a; a.p = p; a.B = B; a.D = 16; a.H = H; a.W = W; a.C = 32;
  a.sd = 32; a.sw = 512; a.sh = (long)W * 512; a.sb = (long)H * W * 512;
  return a;
}
inline Act vol_as_2d(float* p, int B, int H, int W) { return make_act(p, B, 1, H, W, 512); }
// channel slice [c0, c0+C) of a channels-last tensor (zero-copy concat)
inline Act slice_c(Act a, int c0, int C) { a.p += c0; a.C = C; return a; }

// split operand of the tcgen05 conv: dense channels-last [B,D,H,W,nblk,64] 16-bit (fp16 pairs) where every
// 32-channel block is the 128-byte row [hi x32 | lo x32], value ~= hi + lo (pad channels are zero)
struct Opd {
  __nv_bfloat16* p = nullptr;
  int B = 0, D = 1, H = 0, W = 0, nblk = 0;
  int pstride = 0;      // 16-bit elements per pixel when the operand is a block range of a wider one (0 = dense: nblk * 64)
  float amul = 1.f;     // power of two the values were multiplied by before the fp16 split (ConvW::amul of the conv the operand
                        // was allocated for): keeps small activations out of fp16's subnormal range; the reader divides it out
  long row() const { return pstride ? pstride : (long)nblk * 64; }
};

// The split itself: hi = fp16(v), lo = fp16(v - hi): v ~= hi + lo to ~2^-23 relative while |v| is in fp16's normal range
// (6.1e-5 .. 65504) and to 3e-8 absolute below it.  fp16 rather than bf16 halves because the measured per-conv error of
// a bf16 split (2^-17 per operand) left only ~25% margin to the 1e-3 end-to-end parity bar over the ~75 stacked convs;
// the price is range: values beyond +-65504 saturate.  Activations of this network are O(1..100); weights are
// pre-scaled per conv by a power of two (ConvW::wmul, undone in the epilogue) so that their remainders stay normal.
// (Mixed formats -- bf16 hi with an fp16 remainder -- are rejected by tcgen05.mma kind::f16: "illegal instruction".)
constexpr uint32_t IDESC_AB_FMT = 0u;          // instruction-descriptor A/B format bits (7-9, 10-12): 0 = f16, 1 = bf16

__host__ __device__ inline void split_operand(float v, __nv_bfloat16& hi_bits, __nv_bfloat16& lo_bits) {
  const float vv = v > 65504.f ? 65504.f : (v < -65504.f ? -65504.f : v);      // NaN passes through
  const __half h = __float2half_rn(vv);
  const __half l = __float2half_rn(vv - __half2float(h));
  hi_bits = *reinterpret_cast<const __nv_bfloat16*>(&h);                        // 16-bit storage slots
  lo_bits = *reinterpret_cast<const __nv_bfloat16*>(&l);
}

#ifdef __CUDACC__
// 4 values -> {hi01, hi23} and {lo01, lo23} packed as two 32-bit words each
__device__ __forceinline__ void split_operand4(float a, float b, float c, float d, uint2& hv, uint2& lv) {
  const float lim = 65504.f;
  a = a > lim ? lim : (a < -lim ? -lim : a); b = b > lim ? lim : (b < -lim ? -lim : b);
  c = c > lim ? lim : (c < -lim ? -lim : c); d = d > lim ? lim : (d < -lim ? -lim : d);
  const __half2 h01 = __floats2half2_rn(a, b), h23 = __floats2half2_rn(c, d);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn(a - f01.x, b - f01.y), l23 = __floats2half2_rn(c - f23.x, d - f23.y);
  hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
  lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
}
__device__ __forceinline__ void split_operand2(float a, float b, uint32_t& hv, uint32_t& lv) {
  const float lim = 65504.f;
  a = a > lim ? lim : (a < -lim ? -lim : a); b = b > lim ? lim : (b < -lim ? -lim : b);
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hv = *reinterpret_cast<const uint32_t*>(&h);
  lv = *reinterpret_cast<const uint32_t*>(&l);
}
#endif

enum ActKind { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_SIGMOID = 3, ACT_GELU = 4 };

__host__ __device__ inline float apply_act(float v, int kind, float slope) {
  switch (kind) {
    case ACT_RELU: return v > 0.f ? v : 0.f;
    case ACT_LRELU: return v > 0.f ? v : v * slope;
    case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));   // nn.GELU() (exact, erf form)
    default: return v;
  }
}

// NONE / RELU / LRELU as one branch-free select (slope 1 / 0 / slope): the hot epilogues hoist the slope once instead of
// running apply_act's switch (with expf / erff inlined at every call site) per element.  Same values as apply_act
// (a negative input under RELU gives -0 instead of +0, which no consumer can tell apart).
__host__ __device__ inline bool act_is_leaky(int kind) { return kind <= ACT_LRELU; }
__host__ __device__ inline float leaky_slope(int kind, float slope) { return kind == ACT_NONE ? 1.f : (kind == ACT_RELU ? 0.f : slope); }
__host__ __device__ inline float apply_leaky(float v, float s) { return v > 0.f ? v : v * s; }

// packed convolution weights (device)
struct ConvW;
struct ConvW {
  int Cin = 0, Cout = 0, KD = 1, KH = 1, KW = 1;
  float* w32 = nullptr;    // [taps][Cin][Cout]   (SIMT path)
  float* bias = nullptr;   // [Cout] or null
  // tcgen05 path: B operand [Cout_p][tap][nblk][hi 32 | lo 32] bf16 (K-major rows), N tile BN
  __nv_bfloat16* wtc = nullptr;
  __nv_bfloat16* w3s = nullptr;    // 3x3x3 32->32 depth-stacked packing (conv3s_tc.cu)
  float w3s_kappa = 0.f;           // truncation pre-compensation folded into w3s
  __nv_bfloat16* w7 = nullptr;     // 7x7x7 depth-stacked packing (conv7_tc.cu), mask conv only
  float w7_kappa = 0.f;            // truncation pre-compensation folded into w7
  int nblk = 0, Cout_p = 0, BN = 0;
  float wmul = 1.f;                // power of two applied to the packed tcgen05 weights (epilogues multiply by 1 / wmul)
  float amul = 1.f;                // power of two applied to this conv's INPUT operand by whoever writes it (Opd::amul), chosen by
                                   // cs_calibrate from the measured max |activation| so that the split halves stay in fp16's normal range
  int id = -1;                     // index of this conv in the ctx's calibration table
  bool amul_fixed = false;         // operand writers that do not take a scale (motion extractor): amul stays 1
  int zrows = 0;                   // > 0: depth-dependent weights (rows d*zrows .. of wtc belong to depth slice d)
  ConvW* wn = nullptr;             // Winograd F(2x2,3x3) form of a 3x3 conv (wino.cu): 16 x Cout rows, K = Cin, zrows = Cout
  // accumulator plan of the tcgen05 kernel, fixed when the weights are packed (pack_tc): the packed weights carry the
  // position-dependent truncation pre-compensation of exactly this MMA issue order (see TcPlan in tc_ptx.cuh)
  int plan_nsets = 0, plan_chunk = 0, plan_nacc = 0, plan_npass = 3;
  bool plan_thin = false;
  float plan_kappa = 0.f;          // pre-compensation per truncation event folded into wtc (0 = none: the epilogue compensates)
  int phase_shift = 0;             // > 0: phase-form conv of an input nearest-upsampled by 2^phase_shift (pack_phase_conv)
  int taps() const { return KD * KH * KW; }
};

// conv epilogue: v = acc + bias; v = act(v); v += residual; v *= mult[pixel]
struct Epilogue {
  int act = ACT_NONE;
  float slope = 0.f;
  const float* residual = nullptr;   // same geometry/strides as the output
  long rs_b = 0, rs_d = 0, rs_h = 0, rs_w = 0;
  const float* mult = nullptr;       // [B, Do*Ho*Wo] per-pixel multiplier
  // tcgen05 path only: additionally emit act2(v * emit_scale[c] + emit_shift[c]) as the split-fp16 operand
  // [pixels, emit_nblk, 64] of the next conv (scale/shift null = identity); the fp32 output pointer may then be null
  __nv_bfloat16* emit = nullptr;
  int emit_nblk = 0;
  float emit_mul = 1.f;              // Opd::amul of the emitted operand
  const float* emit_scale = nullptr;
  const float* emit_shift = nullptr;
  int emit_act = ACT_NONE;
  float emit_slope = 0.f;
  // SPADE mode (tcgen05 path, with `emit`): the conv is a gamma|beta conv whose columns are interleaved in chunks of
  // [gamma x16 | beta x16]; the epilogue emits act(((x - mean) * rstd) * (1 + gamma) + beta) of the tensor x being
  // normalised (dense [B,Hx,Wx,C], read nearest-upsampled by 2^sp_xshift) instead of gamma / beta themselves.
  const float* sp_x = nullptr;
  const float* sp_mean = nullptr;   // [B,C]
  const float* sp_rstd = nullptr;
  int sp_C = 0, sp_xshift = 0, sp_Hx = 0, sp_Wx = 0;
  // phase mode (tcgen05 path): the conv input is the operand nearest-upsampled by 2^phase_shift.  Instead of
  // materialising it, the conv runs on the low-resolution operand once per output phase (a, b) with the taps that fall on
  // the same source pixel pre-summed (ConvW packed by pack_phase_conv: Cout = 4^phase_shift * BN rows): 2.25 (x2) or
  // 4 (x4) times fewer MACs and no upsampled operand.
  int phase_shift = 0;
  // measurement only: algorithmic FLOPs of this launch when they differ from the GEMM's (a Winograd GEMM stands for a 3x3 conv)
  double alg_flops = 0.0;
};

// conv geometry
struct ConvGeom {
  int PD = 0, PH = 0, PW = 0;
  int Do = 1, Ho = 0, Wo = 0;
};

// ------------------------------------------------------------------------------------------
// input transform (the "prep" kernel): gather + normalise + modulate + activate
// ------------------------------------------------------------------------------------------
enum NormKind { NORM_NONE = 0, NORM_AFFINE_C = 1, NORM_STATS_BC = 2 };

struct Prep {
  // sources (concat along C: src0 channels first, then src1)
  Act src0, src1;            // src1.p == nullptr when unused
  int upshift = 0;           // nearest upsample of (h,w) by 2^upshift (src0; src1 must be unused)
  int pool2 = 0;             // 2x2 average pool of (h,w) (src0; src1 must be unused)
  int norm = NORM_NONE;
  const float* scale = nullptr;   // AFFINE_C: per-channel scale / shift. STATS_BC: optional gamma/beta per channel
  const float* shift = nullptr;
  const float* mean = nullptr;    // STATS_BC: [B,C]
  const float* rstd = nullptr;
  const float* gb = nullptr;      // SPADE: [pixels, 2*C] in chunks of [gamma x16 | beta x16], applied after normalisation
  Act add;                        // residual added before activation (add.p == nullptr when unused); geometry of the output
  int act = ACT_NONE;
  float slope = 0.f;
};

// ------------------------------------------------------------------------------------------
// bump arena for per-call scratch
// ------------------------------------------------------------------------------------------
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, high = 0;
  bool measuring = false;    // dry run: no memory, just count
  void* alloc(size_t bytes) {
    size_t a = (off + 255) & ~size_t(255);
    off = a + bytes;
    if (off > high) high = off;
    if (measuring) return reinterpret_cast<void*>(uintptr_t(0x1000) + a);   // fake, never dereferenced
    if (off > cap) throw Error(-5, "workspace arena exhausted");
    return base + a;
  }
  float* f32(size_t n) { return static_cast<float*>(alloc(n * sizeof(float))); }
  __nv_bfloat16* bf16(size_t n) { return static_cast<__nv_bfloat16*>(alloc(n * 2)); }
  size_t mark() const { return off; }
  void reset(size_t m = 0) { off = m; }
};

// optional per-launch CUDA-event timing, aggregated per kernel family (bench.py's roofline leg)
enum ProfKind { PK_CONV_TC = 0, PK_CONV_SIMT = 1, PK_PREP = 2, PK_STATS = 3, PK_SAMPLE = 4, PK_OTHER = 5, PK_N = 6 };
struct Profiler {
  struct Rec { cudaEvent_t a, b; int kind; double flops, bytes; char desc[120]; };
  std::vector<Rec> recs;
  bool on = false;
};

struct Launcher {            // everything a kernel launch helper needs
  cudaStream_t stream = nullptr;
  bool dry = false;          // measuring pass: skip launches
  int64_t* counter = nullptr;
  int conv_impl = 0;         // 0 auto, 1 SIMT, 2 TC
  int npass = 3;             // split-fp16 MMA passes of the tcgen05 conv (3 = hi*hi + lo*hi + hi*lo)
  int max_sets = 0;          // cap on the accumulator sets (0 = automatic)
  bool phase_conv = true;    // convs of a nearest-upsampled seg map in phase form on the low-resolution operand
  bool spade_fused = true;   // SPADE normalise + modulate + activate inside the gamma|beta conv's epilogue
  bool stacked3 = true;      // depth-stacked kernel for the 32 -> 32 3x3x3 volume convs
  bool pair = true;          // tcgen05 pair mode (cta_group::2, 2-CTA clusters) for wide N tiles
  int pair_min_iter = 16;    // ... with at least this many K iterations
  int single_chain = 256;    // convs whose whole MMA chain (hi*hi + corrections) is at most this long use ONE accumulator
  bool winograd = true;      // adaptive convs in Winograd F(2x2,3x3) form (wino.cu)
  bool winograd_static = true;   // ... and the static 3x3 convs with Cout % 256 == 0 (CS_OPT_WINOGRAD = 2: adaptive convs only)
  bool double_buffer = true; // two TMEM accumulator buffers where they fit (epilogue overlaps the next tile's MMAs)
  float acc_comp = 170.f;     // accumulate-truncation compensation per chained MMA, in units of 1e-10 (0 = off): constant epilogue factor,
                              // used by the kernels whose weights are not position-compensated at pack time (ConvW::plan_kappa == 0)
  Profiler* prof = nullptr;
  const char* tag = nullptr;  // stage label attached to profiler records
  unsigned* calib = nullptr;  // calibration pass (cs_calibrate): per-conv max |input activation| as float bits, indexed by ConvW::id
  void count() const { if (counter) ++*counter; }
};

// brackets the launches issued in its scope with two events on the launching stream
struct ProfScope {
  const Launcher& L;
  size_t idx = (size_t)-1;
  ProfScope(const Launcher& l, int kind, double flops, double bytes, const char* desc = nullptr) : L(l) {
    if (!L.prof || !L.prof->on || L.dry) return;
    Profiler::Rec r; r.kind = kind; r.flops = flops; r.bytes = bytes;
    snprintf(r.desc, sizeof(r.desc), "%s%s%s", L.tag ? L.tag : "", L.tag ? " " : "", desc ? desc : "");
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, L.stream);
    idx = L.prof->recs.size();
    L.prof->recs.push_back(r);
  }
  ~ProfScope() { if (idx != (size_t)-1) cudaEventRecord(L.prof->recs[idx].b, L.stream); }
};

inline void check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw Error(-2, std::string("launch ") + what + ": " + cudaGetErrorString(e));
}

// ------------------------------------------------------------------------------------------
// kernel launch helpers (implemented in the .cu files)
// ------------------------------------------------------------------------------------------
// kernels_elem.cu
void prep_f32(const Launcher& L, const Prep& p, Act out);                    // out: fp32 (strides from Act)
void prep_planes(const Launcher& L, const Prep& p, Opd out, const Act* out32);   // split-fp16 planes (+ optional fp32 copy)
void avg2(const Launcher& L, const float* a, const float* b, float* y, long n);   // y = (a + b) / 2
constexpr int STATS_MAX_BLOCKS = 128;      // per-sample partial blocks of instance_stats: scratch = [B][128][C <= 512][2] doubles
void instance_stats(const Launcher& L, const Act& x, float* mean, float* rstd, float eps, double* scratch);
void stats_finalize_blocks(const Launcher& L, const double* part, int nblocks, int B, int C, long S, float* mean, float* rstd, float eps);
// request for the instance statistics of a conv's OUTPUT, computed by the kernel that writes it (no extra pass over the tensor)
struct StatsOut { double* scratch = nullptr; float* mean = nullptr; float* rstd = nullptr; float eps = 1e-5f; };
void adaptive_blend(const Launcher& L, const float* o2 /*[P,1024]*/, const float* mask /*[P]*/,
                    const float* residual /*[P,512] or null*/, int relu, float* y /*[P,512] or null*/,
                    __nv_bfloat16* opl /*next conv operand [P,16,64] or null*/, long P, float opl_mul = 1.f);
// calibration: max |value| of a split operand (hi halves, scale divided out) into slot `id` of the table
void operand_absmax(const Launcher& L, const Opd& x, int id);
void nchw_to_cl(const Launcher& L, const float* src, float* dst, int B, int C, long S, int vol_perm);
void cl_to_nchw(const Launcher& L, const float* src, float* dst, int B, int C, long S, int vol_perm, int Cstride);
void ingest_u8(const Launcher& L, const uint8_t* src, float* dst, long n);
void emit_image(const Launcher& L, const float* y /*[B,H,W,Cs] conv_img out, 12 valid*/, int Cs,
                float* img /*[B,3,2H,2W] or null*/, uint8_t* u8 /*[B,2H,2W,3] or null*/, int B, int H, int W);
// kernels_conv_simt.cu
// xshift: the conv input is x nearest-upsampled by 2^xshift along (H, W) (read through index math)
void conv_simt(const Launcher& L, const Act& x, const ConvW& w, const ConvGeom& g, const Epilogue& e, Act y, int xshift = 0);
void conv_cout1(const Launcher& L, const Act& x, const ConvW& w, const ConvGeom& g, int act, float* y /*[B*Do*Ho*Wo]*/);
// kernels_motion.cu
void dm_input(const Launcher& L, const Act& c4, const float* kp_driving, const float* kp_source, int K,
              Act out /*[B,D,H,W,(K+1)*5 (+pad)]*/);
// the same tensor written directly as the split operand of the convs that read it (4 blocks, pad channels zero)
void dm_input_operand(const Launcher& L, const Act& c4, const float* kp_driving, const float* kp_source, int K, Opd out);
void softmax_flow_warp(const Launcher& L, const Act& logits /*[B,D,H,W,K+1]*/, const float* kp_driving,
                       const float* kp_source, int K, const float* vol /*[B,H,W,16,32]*/,
                       float* out /*[B,H,W,16,32]*/, float* deformation /*[B,D,H,W,3] or null*/);
void grid_sample3d_cl(const Launcher& L, const float* vol, const float* grid, float* out, int B, int D, int H, int W);
// conv_tc.cu
// "same" convolution (stride 1, pad = k/2) of a dense channels-last tensor with the geometry of `out`
bool conv_tc_supported(const ConvW& w, const Act& out);
Opd conv_tc_alloc_operand(Arena& A, const ConvW& w, const Act& out);   // split-fp16 operand with the geometry of `out`
void conv_tc(const Launcher& L, const Opd& x, const ConvW& w, const ConvGeom& g, const Epilogue& e, Act y);
// conv3s_tc.cu : the 32 -> 32 3x3x3 volume convs (depth-stacked, weights resident in shared memory)
bool conv3s_supported(const ConvW& w, int H, int W);
void conv3s_tc(const Launcher& L, const Opd& x, const ConvW& w, const Epilogue& e, Act y, float* stats_part = nullptr);
size_t conv3s_stats_floats(int B, int H, int W);            // per-tile partial sums written when stats_part != null
void conv3s_stats(const Launcher& L, const float* stats_part, int B, int H, int W, float* mean, float* rstd, float eps);
// conv7_tc.cu : the 7x7x7 mask conv (depth-stacked, kh-split); scratch holds the 7 partial logit tensors
bool conv7_supported(const ConvW& w, const Act& out);
size_t conv7_scratch_floats(const Act& out);
void conv7_tc(const Launcher& L, const Opd& x, const ConvW& w, Act out, float* scratch);
// occlusion map = sigmoid(bias + sum_{z,kh,kw} Y[b,z,h+kh-3,w+kw-3,kh*7+kw]) from the per-tap projections Y [B,16,H,W,64]
void occlusion_gather(const Launcher& L, const Act& Y, const float* bias, float* occ);

}  // namespace cs
