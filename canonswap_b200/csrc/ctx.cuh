// Context, packed-weight inventory and per-call helpers of the canonswap_b200 library.
#pragma once
#include "common.cuh"
#include <memory>
#include "../../include/canonswap_b200.h"

namespace cs {

constexpr int NUM_KP = 21;
constexpr int HG_IN = (NUM_KP + 1) * 5;   // 110
constexpr int HG_OUT = 32 + HG_IN;        // 142

struct Affine {              // per-channel scale / shift (folded eval-mode BN, or GN gamma/beta)
  float* scale = nullptr;
  float* shift = nullptr;
};

struct ResBlock3dW {         // reference util.py:80-102
  Affine bn1;                // pre-activation BN on the block input
  ConvW conv1;               // norm2 folded in (post-conv), ReLU epilogue
  ConvW conv2;
};

struct GnResBlockW {         // reference util.py:515-544
  ConvW conv1, conv2;
  Affine gn1, gn2;
};

struct ResBlock2dW {         // reference util.py:105-128 at 512 channels (volume channel order)
  Affine bn1;
  ConvW conv1;               // norm2 folded, LeakyReLU(0.01) epilogue
  ConvW conv2;
};

struct AdaptiveConvW {       // reference adaptive_modulate.py:73-193
  float* w_base = nullptr;   // [9][512][512] fp32 master (internal channel order both sides)
  float* bias_param = nullptr;   // [512]
  float* fc0_w = nullptr; float* fc0_b = nullptr;   // style MLP (rows of fc2 permuted to internal order)
  float* fc2_w = nullptr; float* fc2_b = nullptr;
  ConvW mask_conv;           // 512 -> 1, sigmoid
  ConvW combined;            // per identity: Cout = 1024 = [W | W * s * demod], bias = [0 | bias_param]
  ConvW wino;                // per identity: Winograd F(2x2,3x3) transform of `combined` (wino.cu), 16 x 1024 rows, K = 512
  float* style = nullptr;    // [512] per identity
  float* demod = nullptr;    // [512] per identity
};

struct SpadeNormW {          // reference util.py:282-302
  ConvW shared;              // label_nc(256) -> 128, ReLU
  ConvW shared_ph;           // the same conv applied to a nearest-upsampled seg, in phase form (up blocks only)
  int phase_shift = 0;
  ConvW gamma_beta;          // 128 -> 2*C : gamma | beta stacked along Cout
  int C = 0;
};

struct SpadeBlockW {         // reference util.py:305-344
  int fin = 0, fout = 0, fmid = 0;
  bool learned_shortcut = false;
  SpadeNormW norm_0, norm_1, norm_s;
  ConvW conv_0, conv_1, conv_s;    // spectral-norm sigma folded
};

struct MotionBlockW {        // ConvNeXtV2 Block, reference convnextv2.py:23-47
  float* dw_w = nullptr;     // depthwise 7x7 [49][C]
  float* dw_b = nullptr;
  float* ln_w = nullptr; float* ln_b = nullptr;
  ConvW pw1, pw2;            // Linear C -> 4C (+GELU), Linear 4C -> C, as 1x1 convs
  float* grn_g = nullptr; float* grn_b = nullptr;
};

struct MotionW {             // MotionExtractor.detector, reference convnextv2.py:62-103
  bool loaded = false;
  float* stem_w = nullptr; float* stem_b = nullptr; float* stem_ln_w = nullptr; float* stem_ln_b = nullptr;
  float* ds_ln_w[3] = {}; float* ds_ln_b[3] = {};
  ConvW ds[3];               // 2x2 stride-2 convs as 1x1 convs over space-to-depth channels (kh, kw, ci)
  MotionBlockW blk[18];
  float* norm_w = nullptr; float* norm_b = nullptr;
  float* head_w = nullptr; float* head_b = nullptr;   // [328][768]: kp | scale | pitch | yaw | roll | t | exp
  double* sumsq = nullptr;   // GRN scratch [max_batch][3072], kept zero between uses
};

struct Weights {
  // F
  ConvW f_first, f_down[2], f_second;
  ResBlock3dW f_res[6];
  // W
  ConvW dm_compress;
  ConvW hg_enc[5], hg_dec[5], hg_final, dm_mask, dm_occlusion;
  ConvW hg_dec_ph[5];        // decoder convs in phase form on the low-resolution operand
  ConvW dm_occ_y;            // occlusion conv as per-tap projections (1x1x1, depth-dependent weights; tcgen05 only)
  ConvW w_third, w_fourth;
  // swap
  AdaptiveConvW ad[14];
  ResBlock3dW t_res[6];
  // refine
  GnResBlockW r_gn1[3], r_gn3[3];
  ResBlock2dW r_res2[3];
  // G
  ConvW g_fc, g_img;
  SpadeBlockW g_blocks[8];
};

}  // namespace cs

struct cs_ctx {
  int device = 0, max_batch = 1, net_h = 0, net_w = 0, h = 0, w = 0;
  std::string err;
  bool weights_loaded = false, identity_set = false;
  int conv_impl = 0, use_graph = 0, tc_passes = 3, tc_sets = 0, tc_comp = 170, tc_poscomp = 330, tc_pair = 1, tc_stacked3 = 1, tc_dbuf = 1, winograd = 1, tc_chain_max = 0, tc_single_chain = 256, tc_bn_max = 0, tc_bn_min = 128;
  int64_t launches = 0;
  std::vector<void*> owned;        // device allocations owned by the ctx
  size_t owned_bytes = 0;
  cs::Arena arena;
  cs::Weights W;
  cs::MotionW M;
  float* se_scratch = nullptr; size_t se_cap = 0;   // SoftErosion scratch (grown on demand, outside the hot path's arena)
  std::vector<std::unique_ptr<cs::ConvW>> wino_convs;   // Winograd forms of static convs (ConvW::wn)
  double* stats_scratch = nullptr; // [max_batch][STATS_MAX_BLOCKS][512][2] double: per-block partial sums of instance_stats
  double* stats_lane[4] = {};      // per-lane statistics scratch (CS_OPT_LANES); [0] == stats_scratch
  int lanes = 2;                   // CS_OPT_LANES: a graph-captured cs_frame runs as this many concurrent sub-batches
  cudaStream_t lane_stream[4] = {}; // [0] unused (the capture stream itself)
  cudaEvent_t ev_lane[4] = {};
  cudaEvent_t ev_fork = nullptr;
  cs::Profiler prof;
  // activation-scale calibration (cs_calibrate): per-conv max |input activation| as float bits, indexed by ConvW::id
  unsigned* calib_tab = nullptr; int n_conv_ids = 0; bool calib_on = false; int test_amul_log2 = 0;
  // CUDA-graph replay of cs_frame (CS_OPT_USE_GRAPH): the whole loop body is captured once per (B, flags, outputs)
  // on fixed staging buffers; a call then is copy-in -> graph launch -> copy-out on the caller's stream.
  struct FrameGraph { cudaGraphExec_t exec = nullptr; int B = 0, flags = 0; bool f32 = false, u8 = false; int seen = 0; int64_t launches = 0; };
  std::vector<FrameGraph> graphs;
  cudaStream_t cap_stream = nullptr;   // capture happens on a private stream (the legacy default stream cannot be captured)
  void* g_frames = nullptr; float* g_kpt = nullptr; float* g_kpc = nullptr; float* g_out32 = nullptr; uint8_t* g_outu8 = nullptr;
  void drop_graphs() {
    for (auto& g : graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    graphs.clear();
  }
  void* dmalloc(size_t bytes) {
    void* p = nullptr;
    CS_CUDA(cudaMalloc(&p, bytes ? bytes : 256));
    owned.push_back(p); owned_bytes += bytes;
    return p;
  }
};

namespace cs {

struct Net {                 // per-call view
  cs_ctx* ctx;
  Launcher L;
  Arena* A;
  double* stats = nullptr;   // instance-statistics scratch of this lane
  double* grn = nullptr;     // GRN sum-of-squares scratch of this lane (motion.cu)
  const Weights& W() const { return ctx->W; }
};

// weights.cu
void load_weights(cs_ctx* ctx, const cs_tensor_desc* table, int n);
void set_identity(cs_ctx* ctx, const float* id_dev, cudaStream_t stream);
void calibrate_begin(cs_ctx* ctx);                     // weights.cu
int calibrate_end(cs_ctx* ctx, float* maxima, int cap);
void reset_activation_scales(cs_ctx* ctx);
ConvW pack_conv_host(cs_ctx* ctx, const std::vector<float>& w_pt /*[Cout][Cin][taps]*/, const std::vector<float>* bias,
                     int Cout, int Cin, int KD, int KH, int KW, int phase_shift = 0);
ConvW pack_phase_conv_host(cs_ctx* ctx, const std::vector<float>& w_pt, const std::vector<float>* bias, int Cout, int Cin, int KD, int shift);
float weight_prescale(const float* w, size_t n);             // weights.cu
void pack_tc(cs_ctx* ctx, ConvW& w, cudaStream_t stream);
void pack_conv3s(cs_ctx* ctx, ConvW& w);                     // conv3s_tc.cu
void pack_conv7(cs_ctx* ctx, ConvW& w);                      // conv7_tc.cu   // derive the split-fp16 B operand from w32 (conv_tc.cu)

// wino.cu : Winograd F(2x2,3x3) form of the adaptive convs
void pack_wino(cs_ctx* ctx, AdaptiveConvW& a, cudaStream_t stream);
void wino_in(const Launcher& L, const Act& x, Opd V, const ConvW* mask_conv, float* mask, const float* pscale = nullptr,
             const float* pshift = nullptr, int pact = ACT_NONE, float pslope = 0.f);
void pack_wino_static(cs_ctx* ctx, ConvW& w);
bool wino_ok(const Launcher& L, const ConvW& w, int H, int W);
void wino_conv(const Launcher& L, Arena& A, const Act& x, const ConvW& w, const float* pscale, const float* pshift, int pact,
               float pslope, int act, float slope, const float* residual, Act y, const StatsOut* st = nullptr);
void wino_out_blend(const Launcher& L, const float* Mt, const float* mask, const float* bias_mod, const float* residual, int relu,
                    float* y, int B, int H, int W);

// pasteback.cu
void soft_erosion(const Launcher& L, const float* x, float* out, uint8_t* hard, float* tmp, const float* kw, int B, int H, int W, int K,
                  float thr, int iterations);
void paste_back(const Launcher& L, const uint8_t* crop, const float* mask, const double* M_c2o, const uint8_t* ori, uint8_t* out, int B,
                int hc, int wc, int H, int W);

void parse_mask(const Launcher& L, const float* logits, int B, int C, int h, int w, int H, int W, unsigned long long valid, float* mask,
                int* labels);

// net.cu : stages on the internal (channels-last) layout
void run_F(Net& n, const float* img_cl, int B, float* vol_out);
void run_warp(Net& n, const float* vol_in, const float* kp_source, const float* kp_driving, int B,
              float* vol_out, float* occ, float* deformation);
void run_warp_out(Net& n, const float* vol_in, const float* occ, int B, float* out256);
void run_swap(Net& n, const float* vol_in, int B, float* vol_out, float* masks);
void run_refine(Net& n, const float* vol_in, int B, float* vol_out);
void run_spade(Net& n, const float* feat256, int B, float* img_nchw, uint8_t* img_u8);
// motion.cu : motion extractor M + keypoint transform
bool motion_weights_present(const cs_tensor_desc* table, int n);
void load_motion_weights(cs_ctx* ctx, const cs_tensor_desc* table, int n);
void run_motion(Net& n, const float* img_cl, int B, float* heads /*[B,CS_MOTION_HEADS]*/);
void run_keypoints(Net& n, const float* heads, int B, float* x_s, float* x_can, float* R, float* deg);

}  // namespace cs
