// extern "C" entry points of libcanonswap_b200.so (declared in include/canonswap_b200.h).
// Every call validates its arguments, enqueues kernels on the caller's stream and returns a
// cs_status; C++ exceptions never cross the ABI. There is no CPU or ATen/cuDNN fallback: a missing
// device, unsupported shape or CUDA error is reported, not worked around.
#include "ctx.cuh"
#include <cstring>
#include <cmath>

using namespace cs;

namespace {

thread_local std::string g_create_error;

int fail(cs_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg; else g_create_error = msg;
  return code;
}

#define CS_API_BEGIN(ctx)                                                              \
  if (!(ctx)) return fail(nullptr, CS_ERR_INVALID, "null context");                    \
  try {                                                                                \
    CS_CUDA(cudaSetDevice((ctx)->device));
#define CS_API_END(ctx)                                                                \
    return CS_OK;                                                                      \
  } catch (const cs::Error& e) {                                                       \
    (ctx)->arena.reset(0);                                                             \
    return fail((ctx), e.code, e.what());                                              \
  } catch (const std::exception& e) {                                                  \
    (ctx)->arena.reset(0);                                                             \
    return fail((ctx), CS_ERR_INVALID, e.what());                                      \
  }

Net make_net(cs_ctx* ctx, void* stream, bool dry) {
  Net n;
  n.ctx = ctx;
  n.A = &ctx->arena;
  n.stats = ctx->stats_scratch;
  n.grn = ctx->M.sumsq;
  n.L.stream = static_cast<cudaStream_t>(stream);
  n.L.dry = dry;
  n.L.counter = dry ? nullptr : &ctx->launches;
  n.L.conv_impl = ctx->conv_impl;
  n.L.npass = ctx->tc_passes;
  n.L.max_sets = ctx->tc_sets;
  n.L.acc_comp = (float)ctx->tc_comp;
  n.L.pair = ctx->tc_pair != 0;
  n.L.pair_min_iter = ctx->tc_pair > 1 ? ctx->tc_pair : 16;
  n.L.stacked3 = ctx->tc_stacked3 != 0;
  n.L.double_buffer = ctx->tc_dbuf != 0;
  n.L.winograd = ctx->winograd != 0;
  n.L.winograd_static = ctx->winograd == 1;
  n.L.single_chain = ctx->tc_single_chain;
  n.L.prof = dry ? nullptr : &ctx->prof;
  n.L.calib = (!dry && ctx->calib_on) ? ctx->calib_tab : nullptr;
  return n;
}

void check_batch(cs_ctx* ctx, int B) {
  CS_REQUIRE(ctx->weights_loaded, CS_ERR_STATE, "weights not loaded (call cs_load_weights first)");
  CS_REQUIRE(B >= 1 && B <= ctx->max_batch, CS_ERR_INVALID, "batch size outside [1, max_batch]");
}

long vol_elems(const cs_ctx* ctx, int B) { return (long)B * ctx->h * ctx->w * 512; }

// ---- stage bodies shared by the real calls and the workspace-measuring dry run ------------------
void body_appearance(Net& n, const float* img, float* f3d, int B) {
  cs_ctx* c = n.ctx;
  float* img_cl = n.A->f32((size_t)B * c->net_h * c->net_w * 3);
  float* vol = n.A->f32(vol_elems(c, B));
  nchw_to_cl(n.L, img, img_cl, B, 3, (long)c->net_h * c->net_w, 0);
  run_F(n, img_cl, B, vol);
  cl_to_nchw(n.L, vol, f3d, B, 512, (long)c->h * c->w, 1, 512);
}

void body_warp(Net& n, const float* f3d, const float* kp_source, const float* kp_driving, float* out3d, float* occ,
               float* deformation, int B) {
  cs_ctx* c = n.ctx;
  float* vin = n.A->f32(vol_elems(c, B));
  float* vout = n.A->f32(vol_elems(c, B));
  nchw_to_cl(n.L, f3d, vin, B, 512, (long)c->h * c->w, 1);
  run_warp(n, vin, kp_source, kp_driving, B, vout, occ, deformation);
  cl_to_nchw(n.L, vout, out3d, B, 512, (long)c->h * c->w, 1, 512);
}

void body_warp_out(Net& n, const float* f3d, const float* occ, float* out, int B) {
  cs_ctx* c = n.ctx;
  float* vin = n.A->f32(vol_elems(c, B));
  float* o = n.A->f32((size_t)B * c->h * c->w * 256);
  nchw_to_cl(n.L, f3d, vin, B, 512, (long)c->h * c->w, 1);
  run_warp_out(n, vin, occ, B, o);
  cl_to_nchw(n.L, o, out, B, 256, (long)c->h * c->w, 0, 256);
}

void body_warp_forward(Net& n, const float* f3d, const float* kp_driving, const float* kp_source, float* out, float* occ,
                       float* deformation, int B) {
  cs_ctx* c = n.ctx;
  float* vin = n.A->f32(vol_elems(c, B));
  float* vout = n.A->f32(vol_elems(c, B));
  float* o = n.A->f32((size_t)B * c->h * c->w * 256);
  float* occ_buf = occ ? occ : n.A->f32((size_t)B * c->h * c->w);
  nchw_to_cl(n.L, f3d, vin, B, 512, (long)c->h * c->w, 1);
  run_warp(n, vin, kp_source, kp_driving, B, vout, occ_buf, deformation);
  run_warp_out(n, vout, occ_buf, B, o);
  cl_to_nchw(n.L, o, out, B, 256, (long)c->h * c->w, 0, 256);
}

void body_swap(Net& n, const float* f3d, float* out3d, float* masks, int B) {
  cs_ctx* c = n.ctx;
  float* v = n.A->f32(vol_elems(c, B));
  nchw_to_cl(n.L, f3d, v, B, 512, (long)c->h * c->w, 1);
  run_swap(n, v, B, v, masks);
  cl_to_nchw(n.L, v, out3d, B, 512, (long)c->h * c->w, 1, 512);
}

void body_refine(Net& n, const float* f3d, float* out3d, int B) {
  cs_ctx* c = n.ctx;
  float* v = n.A->f32(vol_elems(c, B));
  nchw_to_cl(n.L, f3d, v, B, 512, (long)c->h * c->w, 1);
  run_refine(n, v, B, v);
  cl_to_nchw(n.L, v, out3d, B, 512, (long)c->h * c->w, 1, 512);
}

void body_spade(Net& n, const float* feat, float* img, uint8_t* img_u8, int B) {
  cs_ctx* c = n.ctx;
  float* f = n.A->f32((size_t)B * c->h * c->w * 256);
  nchw_to_cl(n.L, feat, f, B, 256, (long)c->h * c->w, 0);
  run_spade(n, f, B, img, img_u8);
}

// The per-frame loop body, reference can_swap_pipeline_e2e.py:242-267, entirely in the internal layout.
void body_frame(Net& n, const void* frames, const float* kp_t, const float* kp_can, float* out_f32, uint8_t* out_u8, int B,
                int flags) {
  cs_ctx* c = n.ctx;
  const long npix = (long)B * c->net_h * c->net_w;
  float* img_cl = n.A->f32((size_t)npix * 3);
  float* va = n.A->f32(vol_elems(c, B));
  float* vb = n.A->f32(vol_elems(c, B));
  float* occ = n.A->f32((size_t)B * c->h * c->w);
  float* o256 = n.A->f32((size_t)B * c->h * c->w * 256);
  if ((flags & CS_FRAME_V2I) && (flags & CS_FRAME_V2I_FEATURE)) {  // v2i with the appearance volume resident (one per source)
    nchw_to_cl(n.L, static_cast<const float*>(frames), va, 1, 512, (long)c->h * c->w, 1);
    if (!n.L.dry)
      for (int b = 1; b < B; ++b)
        CS_CUDA(cudaMemcpyAsync(va + vol_elems(c, b), va, (size_t)vol_elems(c, 1) * sizeof(float), cudaMemcpyDeviceToDevice, n.L.stream));
    run_warp(n, va, /*kp_source=*/kp_t, /*kp_driving=*/kp_can, B, vb, occ, nullptr);
    run_warp_out(n, vb, occ, B, o256);
    run_spade(n, o256, B, out_f32, out_u8);
    return;
  }
  if (flags & CS_FRAME_IN_U8_HWC) ingest_u8(n.L, static_cast<const uint8_t*>(frames), img_cl, npix * 3);   // prepare_videos
  else nchw_to_cl(n.L, static_cast<const float*>(frames), img_cl, B, 3, (long)c->net_h * c->net_w, 0);
  if (flags & CS_FRAME_MOTION) {                                  // x_t / x_can of the frames from M (:112-125, :231-243)
    float* heads = n.A->f32((size_t)B * CS_MOTION_HEADS);
    float* kt = n.A->f32((size_t)B * NUM_KP * 3);
    float* kc = n.A->f32((size_t)B * NUM_KP * 3);
    run_motion(n, img_cl, B, heads);
    run_keypoints(n, heads, B, kt, kc, nullptr, nullptr);
    kp_t = kt; kp_can = kc;
  }
  run_F(n, img_cl, B, va);                                        // :242 f_s = extract_feature_3d(I_s)
  if (flags & CS_FRAME_V2I) {                                     // can_swap_pipeline_v2i.py:308-309
    run_warp(n, va, /*kp_source=*/kp_t, /*kp_driving=*/kp_can, B, vb, occ, nullptr);
    run_warp_out(n, vb, occ, B, o256);
    run_spade(n, o256, B, out_f32, out_u8);
    return;
  }
  run_warp(n, va, /*kp_source=*/kp_t, /*kp_driving=*/kp_can, B, vb, occ, nullptr);   // :244 warp(f_s, x_t, x_can)
  if (flags & CS_FRAME_DEBUG_DECODES) {                           // :248 rec_can = conv_decode(f_can, occ)
    run_warp_out(n, vb, occ, B, o256);
    run_spade(n, o256, B, nullptr, nullptr);
  }
  run_swap(n, vb, B, vb, nullptr);                                // :253 swap_module(f_can, source_id)
  if (flags & CS_FRAME_DEBUG_DECODES) {                           // :257 swap_can = conv_decode(f_swap, occ)
    run_warp_out(n, vb, occ, B, o256);
    run_spade(n, o256, B, nullptr, nullptr);
  }
  run_refine(n, vb, B, vb);                                       // :262 refine_module(f_swap)
  run_warp(n, vb, /*kp_source=*/kp_can, /*kp_driving=*/kp_t, B, va, occ, nullptr);   // :263 warp_decode(f_swap, x_can, x_t)
  run_warp_out(n, va, occ, B, o256);
  run_spade(n, o256, B, out_f32, out_u8);                         // :267 parse_output fused into the emit kernel
}

void body_motion(Net& n, const float* img, float* heads, int B) {
  cs_ctx* c = n.ctx;
  float* img_cl = n.A->f32((size_t)B * c->net_h * c->net_w * 3);
  nchw_to_cl(n.L, img, img_cl, B, 3, (long)c->net_h * c->net_w, 0);
  run_motion(n, img_cl, B, heads);
}

// size the arena: dry-run every entry point at max_batch and keep the high-water mark
void size_workspace(cs_ctx* ctx) {
  Arena& A = ctx->arena;
  A.measuring = true; A.base = nullptr; A.cap = 0; A.off = 0; A.high = 0;
  bool id = ctx->identity_set;
  ctx->identity_set = true;
  const int B = ctx->max_batch;
  float* fake = reinterpret_cast<float*>(uintptr_t(0x1000));
  {
    Net n = make_net(ctx, nullptr, true);
    // both conv implementations must fit (CS_OPT_CONV_IMPL may be flipped after loading)
    for (int impl = 0; impl < 3; ++impl) {                 // SIMT, tcgen05 direct, tcgen05 + Winograd adaptive convs
      n.L.conv_impl = impl == 0 ? 1 : 0;
      n.L.winograd = impl == 2;
      A.reset(0); body_appearance(n, fake, fake, B);
      A.reset(0); body_warp(n, fake, fake, fake, fake, fake, fake, B);
      A.reset(0); body_warp_out(n, fake, fake, fake, B);
      A.reset(0); body_warp_forward(n, fake, fake, fake, fake, nullptr, fake, B);
      A.reset(0); body_swap(n, fake, fake, fake, B);
      A.reset(0); body_refine(n, fake, fake, B);
      A.reset(0); body_spade(n, fake, fake, reinterpret_cast<uint8_t*>(fake), B);
      A.reset(0); body_frame(n, fake, fake, fake, fake, reinterpret_cast<uint8_t*>(fake), B,
                             CS_FRAME_IN_U8_HWC | CS_FRAME_DEBUG_DECODES | (ctx->M.loaded ? CS_FRAME_MOTION : 0));
      if (ctx->M.loaded) { A.reset(0); body_motion(n, fake, fake, B); }
    }
  }
  // multi-lane replay (CS_OPT_LANES = 2 | 4): each 1/L of the arena must hold cs_frame at ceil(max_batch / L)
  size_t full_high = A.high;
  for (int lanes = 2; lanes <= 4; lanes *= 2) {
    Net n = make_net(ctx, nullptr, true);
    A.high = 0;
    for (int impl = 0; impl < 3; ++impl) {
      n.L.conv_impl = impl == 0 ? 1 : 0;
      n.L.winograd = impl == 2;
      A.reset(0); body_frame(n, fake, fake, fake, fake, reinterpret_cast<uint8_t*>(fake), (B + lanes - 1) / lanes,
                             CS_FRAME_IN_U8_HWC | CS_FRAME_DEBUG_DECODES | (ctx->M.loaded ? CS_FRAME_MOTION : 0));
    }
    const size_t lane_need = (A.high + 4096) * lanes;
    if (lane_need > full_high) full_high = lane_need;
  }
  A.high = full_high;
  ctx->identity_set = id;
  size_t need = A.high + (1 << 20);
  A.measuring = false; A.off = 0; A.high = 0;
  A.base = static_cast<char*>(ctx->dmalloc(need));
  A.cap = need;
}

}  // namespace

extern "C" {

int cs_create(cs_ctx** out, int device, int max_batch, int net_h, int net_w) {
  if (!out) return fail(nullptr, CS_ERR_INVALID, "cs_create: null out pointer");
  *out = nullptr;
  if (max_batch < 1 || max_batch > 1024) return fail(nullptr, CS_ERR_INVALID, "cs_create: max_batch outside [1, 1024]");
  if (net_h < 128 || net_w < 128 || net_h % 128 || net_w % 128 || net_h > 2048 || net_w > 2048)
    return fail(nullptr, CS_ERR_INVALID, "cs_create: net_h / net_w must be multiples of 128 in [128, 2048]");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, CS_ERR_CUDA, std::string("cs_create: no CUDA device (") + cudaGetErrorString(e) +
                                          "); this library has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(nullptr, CS_ERR_INVALID, "cs_create: bad device index");
  cs_ctx* ctx = nullptr;
  try {
    CS_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
      throw cs::Error(CS_ERR_CUDA, "cs_create: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                       ", this library is built for sm_100a (B200) only");
    ctx = new cs_ctx();
    ctx->device = device; ctx->max_batch = max_batch; ctx->net_h = net_h; ctx->net_w = net_w;
    ctx->h = net_h / 4; ctx->w = net_w / 4;
    const size_t sb = sizeof(double) * 2 * (size_t)max_batch * 512 * STATS_MAX_BLOCKS;   // per-block partials (deterministic reduction)
    ctx->stats_scratch = static_cast<double*>(ctx->dmalloc(sb));
    ctx->stats_lane[0] = ctx->stats_scratch;
    for (int l = 1; l < 4; ++l) ctx->stats_lane[l] = static_cast<double*>(ctx->dmalloc(sb / 2 + 4096));   // a lane runs <= ceil(B / 2) frames
  } catch (const std::exception& ex) {
    if (ctx) cs_destroy(ctx);
    return fail(nullptr, CS_ERR_CUDA, ex.what());
  }
  *out = ctx;
  return CS_OK;
}

void cs_destroy(cs_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  ctx->drop_graphs();
  if (ctx->cap_stream) cudaStreamDestroy(ctx->cap_stream);
  for (int l = 1; l < 4; ++l) {
    if (ctx->lane_stream[l]) cudaStreamDestroy(ctx->lane_stream[l]);
    if (ctx->ev_lane[l]) cudaEventDestroy(ctx->ev_lane[l]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  for (void* p : ctx->owned) cudaFree(p);
  delete ctx;
}

const char* cs_last_error(const cs_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int cs_set_option(cs_ctx* ctx, int option, int value) {
  if (!ctx) return fail(nullptr, CS_ERR_INVALID, "null context");
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  ctx->drop_graphs();                                       // captured graphs bake the kernel selection in
  // options that shape the packed weights (accumulator plan, truncation pre-compensation) are fixed by cs_load_weights
  const bool pack_time = option == CS_OPT_TC_PASSES || option == CS_OPT_TC_SETS || option == CS_OPT_TC_SINGLE_CHAIN ||
                         option == CS_OPT_TC_DOUBLE_BUFFER || option == CS_OPT_TC_POSCOMP || option == CS_OPT_TC_BN_MAX ||
                         option == CS_OPT_TC_CHAIN_MAX;
  if (pack_time && ctx->weights_loaded)
    return fail(ctx, CS_ERR_STATE, "this option shapes the packed weights: set it before cs_load_weights");
  switch (option) {
    case CS_OPT_TC_POSCOMP:
      if (value < 0 || value > 2000) return fail(ctx, CS_ERR_INVALID, "CS_OPT_TC_POSCOMP: value must be in [0, 2000]");
      ctx->tc_poscomp = value; return CS_OK;
    case CS_OPT_CONV_IMPL:
      if (value < 0 || value > 1) return fail(ctx, CS_ERR_INVALID, "CS_OPT_CONV_IMPL: value must be 0 or 1");
      ctx->conv_impl = value; return CS_OK;
    case CS_OPT_TC_PASSES:
      if (value < 1 || value > 3) return fail(ctx, CS_ERR_INVALID, "CS_OPT_TC_PASSES: value must be 1, 2 or 3");
      ctx->tc_passes = value; return CS_OK;
    case CS_OPT_TC_SETS:
      if (value < 0 || value > 16) return fail(ctx, CS_ERR_INVALID, "CS_OPT_TC_SETS: value must be in [0, 16]");
      ctx->tc_sets = value; return CS_OK;
    case CS_OPT_TC_COMP:
      if (value < 0 || value > 1000) return fail(ctx, CS_ERR_INVALID, "CS_OPT_TC_COMP: value must be in [0, 1000]");
      ctx->tc_comp = value; return CS_OK;
    case CS_OPT_TC_BN_MAX:
      if (value != 0 && (value < 16 || value > 256 || value % 16)) return fail(ctx, CS_ERR_INVALID, "CS_OPT_TC_BN_MAX: 0 or a multiple of 16 in [16, 256]");
      ctx->tc_bn_max = value; return CS_OK;
    case CS_OPT_TC_SINGLE_CHAIN:
      if (value < 0 || value > 4096) return fail(ctx, CS_ERR_INVALID, "CS_OPT_TC_SINGLE_CHAIN: value must be in [0, 4096]");
      ctx->tc_single_chain = value; return CS_OK;
    case CS_OPT_TC_CHAIN_MAX:
      if (value < 0 || value > 100000) return fail(ctx, CS_ERR_INVALID, "CS_OPT_TC_CHAIN_MAX: value must be in [0, 100000]");
      ctx->tc_chain_max = value; return CS_OK;
    case CS_OPT_TEST_AMUL:
      if (value < -14 || value > 14) return fail(ctx, CS_ERR_INVALID, "CS_OPT_TEST_AMUL: log2 of the scale, in [-14, 14]");
      ctx->test_amul_log2 = value; return CS_OK;
    case CS_OPT_WINOGRAD:
      if (value < 0 || value > 2) return fail(ctx, CS_ERR_INVALID, "CS_OPT_WINOGRAD: 0 off, 1 on (default), 2 adaptive convs only");
      ctx->winograd = value; return CS_OK;
    case CS_OPT_TC_DOUBLE_BUFFER:
      ctx->tc_dbuf = value ? 1 : 0; return CS_OK;
    case CS_OPT_TC_STACKED3:
      ctx->tc_stacked3 = value ? 1 : 0; return CS_OK;
    case CS_OPT_TC_PAIR:
      if (value < 0 || value > 100000) return fail(ctx, CS_ERR_INVALID, "CS_OPT_TC_PAIR: value must be in [0, 100000]");
      ctx->tc_pair = value; return CS_OK;
    case CS_OPT_USE_GRAPH:
      ctx->use_graph = value ? 1 : 0; return CS_OK;
    case CS_OPT_LANES:
      if (value != 1 && value != 2 && value != 4) return fail(ctx, CS_ERR_INVALID, "CS_OPT_LANES: value must be 1, 2 or 4");
      ctx->lanes = value; return CS_OK;
    default: return fail(ctx, CS_ERR_INVALID, "unknown option");
  }
}

int64_t cs_launch_count(const cs_ctx* ctx) { return ctx ? ctx->launches : 0; }
size_t cs_workspace_bytes(const cs_ctx* ctx) { return ctx ? ctx->owned_bytes : 0; }

int cs_load_weights(cs_ctx* ctx, const cs_tensor_desc* table, int n) {
  CS_API_BEGIN(ctx)
  // a failure anywhere below (malformed optional tensors, arena allocation) must leave the ctx UNLOADED and free what this call
  // allocated, so that the caller can retry
  const size_t owned0 = ctx->owned.size();
  const size_t bytes0 = ctx->owned_bytes;
  try {
    load_weights(ctx, table, n);
    ctx->weights_loaded = false;                             // not usable before the workspace exists
    if (motion_weights_present(table, n)) load_motion_weights(ctx, table, n);   // optional: combined_weights['motion_extractor']
    size_workspace(ctx);
    ctx->weights_loaded = true;
  } catch (...) {
    cudaDeviceSynchronize();
    while (ctx->owned.size() > owned0) { cudaFree(ctx->owned.back()); ctx->owned.pop_back(); }
    ctx->owned_bytes = bytes0;
    ctx->W = cs::Weights();
    ctx->M = cs::MotionW();
    ctx->wino_convs.clear();
    ctx->arena = cs::Arena();
    ctx->weights_loaded = false; ctx->identity_set = false; ctx->calib_tab = nullptr;
    throw;
  }
  CS_API_END(ctx)
}

int cs_set_identity(cs_ctx* ctx, const float* id_dev, void* stream) {
  CS_API_BEGIN(ctx)
  set_identity(ctx, id_dev, static_cast<cudaStream_t>(stream));
  CS_API_END(ctx)
}

int cs_appearance(cs_ctx* ctx, const float* img, float* f3d, int B, void* stream) {
  CS_API_BEGIN(ctx)
  check_batch(ctx, B);
  CS_REQUIRE(img && f3d, CS_ERR_INVALID, "cs_appearance: null tensor");
  Net n = make_net(ctx, stream, false);
  ctx->arena.reset(0);
  body_appearance(n, img, f3d, B);
  CS_API_END(ctx)
}

int cs_warp(cs_ctx* ctx, const float* f3d, const float* kp_source, const float* kp_driving, float* out3d, float* occ,
            float* deformation, int B, void* stream) {
  CS_API_BEGIN(ctx)
  check_batch(ctx, B);
  CS_REQUIRE(f3d && kp_source && kp_driving && out3d && occ, CS_ERR_INVALID, "cs_warp: null tensor");
  Net n = make_net(ctx, stream, false);
  ctx->arena.reset(0);
  body_warp(n, f3d, kp_source, kp_driving, out3d, occ, deformation, B);
  CS_API_END(ctx)
}

int cs_warp_out(cs_ctx* ctx, const float* f3d, const float* occ, float* out, int B, void* stream) {
  CS_API_BEGIN(ctx)
  check_batch(ctx, B);
  CS_REQUIRE(f3d && out, CS_ERR_INVALID, "cs_warp_out: null tensor");
  Net n = make_net(ctx, stream, false);
  ctx->arena.reset(0);
  body_warp_out(n, f3d, occ, out, B);
  CS_API_END(ctx)
}

int cs_warp_forward(cs_ctx* ctx, const float* f3d, const float* kp_driving, const float* kp_source, float* out, float* occ,
                    float* deformation, int B, void* stream) {
  CS_API_BEGIN(ctx)
  check_batch(ctx, B);
  CS_REQUIRE(f3d && kp_source && kp_driving && out, CS_ERR_INVALID, "cs_warp_forward: null tensor");
  Net n = make_net(ctx, stream, false);
  ctx->arena.reset(0);
  body_warp_forward(n, f3d, kp_driving, kp_source, out, occ, deformation, B);
  CS_API_END(ctx)
}

int cs_swap(cs_ctx* ctx, const float* f3d, float* out3d, float* masks, int B, void* stream) {
  CS_API_BEGIN(ctx)
  check_batch(ctx, B);
  CS_REQUIRE(f3d && out3d, CS_ERR_INVALID, "cs_swap: null tensor");
  CS_REQUIRE(ctx->identity_set, CS_ERR_STATE, "cs_swap before cs_set_identity");
  Net n = make_net(ctx, stream, false);
  ctx->arena.reset(0);
  body_swap(n, f3d, out3d, masks, B);
  CS_API_END(ctx)
}

int cs_refine(cs_ctx* ctx, const float* f3d, float* out3d, int B, void* stream) {
  CS_API_BEGIN(ctx)
  check_batch(ctx, B);
  CS_REQUIRE(f3d && out3d, CS_ERR_INVALID, "cs_refine: null tensor");
  Net n = make_net(ctx, stream, false);
  ctx->arena.reset(0);
  body_refine(n, f3d, out3d, B);
  CS_API_END(ctx)
}

int cs_spade(cs_ctx* ctx, const float* feat, float* img, uint8_t* img_u8, int B, void* stream) {
  CS_API_BEGIN(ctx)
  check_batch(ctx, B);
  CS_REQUIRE(feat && (img || img_u8), CS_ERR_INVALID, "cs_spade: null tensor");
  Net n = make_net(ctx, stream, false);
  ctx->arena.reset(0);
  body_spade(n, feat, img, img_u8, B);
  CS_API_END(ctx)
}

int cs_frame(cs_ctx* ctx, const void* frames, const float* kp_t, const float* kp_can, float* out_f32, uint8_t* out_u8, int B,
             int flags, void* stream) {
  CS_API_BEGIN(ctx)
  check_batch(ctx, B);
  CS_REQUIRE(frames && (out_f32 || out_u8), CS_ERR_INVALID, "cs_frame: null tensor");
  CS_REQUIRE((flags & CS_FRAME_MOTION) || (kp_t && kp_can), CS_ERR_INVALID, "cs_frame: null keypoints without CS_FRAME_MOTION");
  CS_REQUIRE(!(flags & CS_FRAME_MOTION) || ctx->M.loaded, CS_ERR_STATE, "cs_frame: CS_FRAME_MOTION without motion extractor weights");
  CS_REQUIRE(ctx->identity_set || (flags & CS_FRAME_V2I), CS_ERR_STATE, "cs_frame before cs_set_identity");
  CS_REQUIRE(!(flags & CS_FRAME_V2I_FEATURE) || ((flags & CS_FRAME_V2I) && !(flags & (CS_FRAME_IN_U8_HWC | CS_FRAME_MOTION))),
             CS_ERR_INVALID, "cs_frame: CS_FRAME_V2I_FEATURE needs CS_FRAME_V2I and fp32 input without CS_FRAME_MOTION");
  Net n = make_net(ctx, stream, false);
  ctx->arena.reset(0);
  if (!ctx->use_graph || ctx->prof.on || ctx->calib_on) {
    body_frame(n, frames, kp_t, kp_can, out_f32, out_u8, B, flags);
  } else {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t npix = (size_t)B * ctx->net_h * ctx->net_w;
    const size_t vol_bytes = (size_t)ctx->h * ctx->w * 512 * sizeof(float);
    const size_t in_bytes = (flags & CS_FRAME_V2I_FEATURE) ? vol_bytes : (flags & CS_FRAME_IN_U8_HWC) ? npix * 3 : npix * 3 * sizeof(float);
    const size_t kp_bytes = (size_t)B * NUM_KP * 3 * sizeof(float);
    const size_t o32_bytes = npix * 4 * 3 * sizeof(float), ou8_bytes = npix * 4 * 3;
    if (!ctx->g_frames) {                                   // staging buffers sized for max_batch
      const size_t mp = (size_t)ctx->max_batch * ctx->net_h * ctx->net_w;
      ctx->g_frames = ctx->dmalloc(mp * 3 * sizeof(float) > vol_bytes ? mp * 3 * sizeof(float) : vol_bytes);
      ctx->g_kpt = static_cast<float*>(ctx->dmalloc((size_t)ctx->max_batch * NUM_KP * 3 * sizeof(float)));
      ctx->g_kpc = static_cast<float*>(ctx->dmalloc((size_t)ctx->max_batch * NUM_KP * 3 * sizeof(float)));
      ctx->g_out32 = static_cast<float*>(ctx->dmalloc(mp * 4 * 3 * sizeof(float)));
      ctx->g_outu8 = static_cast<uint8_t*>(ctx->dmalloc(mp * 4 * 3));
    }
    cs_ctx::FrameGraph* fg = nullptr;
    for (auto& g : ctx->graphs)
      if (g.B == B && g.flags == flags && g.f32 == (out_f32 != nullptr) && g.u8 == (out_u8 != nullptr)) fg = &g;
    if (!fg) {
      ctx->graphs.emplace_back();
      fg = &ctx->graphs.back();
      fg->B = B; fg->flags = flags; fg->f32 = out_f32 != nullptr; fg->u8 = out_u8 != nullptr;
    }
    if (fg->seen == 0) {
      // first call of this shape runs eagerly (lazy one-time initialisation must not happen under capture)
      fg->seen = 1;
      body_frame(n, frames, kp_t, kp_can, out_f32, out_u8, B, flags);
    } else {
      if (!fg->exec) {
        const int64_t l0 = ctx->launches;
        cudaGraph_t graph = nullptr;
        if (!ctx->cap_stream) CS_CUDA(cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking));
        cudaStream_t cst = ctx->cap_stream;
        n.L.stream = cst;
        CS_CUDA(cudaStreamBeginCapture(cst, cudaStreamCaptureModeThreadLocal));
        try {
          const int lanes = ctx->lanes <= B ? ctx->lanes : (B >= 2 ? 2 : 1);
          if (lanes >= 2) {
            // concurrent sub-batches on forked capture streams: frames are independent, so the tail wave of one lane's kernel
            // is filled by CTAs of another lane's, and one lane's bandwidth-bound kernels overlap another's MMA-bound ones
            // (each lane owns 1/L of the arena and its own statistics scratch)
            if (!ctx->ev_fork) CS_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
            CS_CUDA(cudaEventRecord(ctx->ev_fork, cst));
            const size_t part = (ctx->arena.cap / lanes) & ~size_t(255);
            Arena arenas[4];
            int b_lo = 0;
            for (int l = 0; l < lanes; ++l) {
              const int Bl = (B - b_lo + (lanes - l) - 1) / (lanes - l);          // even split of what is left
              cudaStream_t sl = cst;
              if (l > 0) {
                if (!ctx->lane_stream[l]) CS_CUDA(cudaStreamCreateWithFlags(&ctx->lane_stream[l], cudaStreamNonBlocking));
                if (!ctx->ev_lane[l]) CS_CUDA(cudaEventCreateWithFlags(&ctx->ev_lane[l], cudaEventDisableTiming));
                sl = ctx->lane_stream[l];
                CS_CUDA(cudaStreamWaitEvent(sl, ctx->ev_fork, 0));
              }
              arenas[l] = ctx->arena;
              arenas[l].base = ctx->arena.base + part * l; arenas[l].cap = part; arenas[l].off = 0; arenas[l].high = 0;
              const size_t px0 = (size_t)b_lo * ctx->net_h * ctx->net_w;
              const size_t in_off = (flags & CS_FRAME_V2I_FEATURE) ? 0 : (flags & CS_FRAME_IN_U8_HWC) ? px0 * 3 : px0 * 3 * sizeof(float);
              Net nl = n; nl.A = &arenas[l]; nl.L.stream = sl; nl.stats = ctx->stats_lane[l];
              if (ctx->M.sumsq) nl.grn = ctx->M.sumsq + (size_t)b_lo * 3072;
              body_frame(nl, static_cast<char*>(ctx->g_frames) + in_off, ctx->g_kpt + (size_t)b_lo * NUM_KP * 3,
                         ctx->g_kpc + (size_t)b_lo * NUM_KP * 3, out_f32 ? ctx->g_out32 + px0 * 4 * 3 : nullptr,
                         out_u8 ? ctx->g_outu8 + px0 * 4 * 3 : nullptr, Bl, flags);
              if (l > 0) {
                CS_CUDA(cudaEventRecord(ctx->ev_lane[l], sl));
                CS_CUDA(cudaStreamWaitEvent(cst, ctx->ev_lane[l], 0));
              }
              b_lo += Bl;
            }
          } else {
            body_frame(n, ctx->g_frames, ctx->g_kpt, ctx->g_kpc, out_f32 ? ctx->g_out32 : nullptr, out_u8 ? ctx->g_outu8 : nullptr, B,
                       flags);
          }
        } catch (...) {
          cudaStreamEndCapture(cst, &graph);
          if (graph) cudaGraphDestroy(graph);
          throw;
        }
        CS_CUDA(cudaStreamEndCapture(cst, &graph));
        cudaError_t ie = cudaGraphInstantiate(&fg->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) { fg->exec = nullptr; throw cs::Error(CS_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); }
        fg->launches = ctx->launches - l0;
        ctx->launches = l0;
      }
      CS_CUDA(cudaMemcpyAsync(ctx->g_frames, frames, in_bytes, cudaMemcpyDeviceToDevice, st));
      if (!(flags & CS_FRAME_MOTION)) {
        CS_CUDA(cudaMemcpyAsync(ctx->g_kpt, kp_t, kp_bytes, cudaMemcpyDeviceToDevice, st));
        CS_CUDA(cudaMemcpyAsync(ctx->g_kpc, kp_can, kp_bytes, cudaMemcpyDeviceToDevice, st));
      }
      CS_CUDA(cudaGraphLaunch(fg->exec, st));
      if (out_f32) CS_CUDA(cudaMemcpyAsync(out_f32, ctx->g_out32, o32_bytes, cudaMemcpyDeviceToDevice, st));
      if (out_u8) CS_CUDA(cudaMemcpyAsync(out_u8, ctx->g_outu8, ou8_bytes, cudaMemcpyDeviceToDevice, st));
      ctx->launches += fg->launches;
    }
  }
  CS_API_END(ctx)
}

int cs_motion(cs_ctx* ctx, const float* img, float* heads, int B, void* stream) {
  CS_API_BEGIN(ctx)
  check_batch(ctx, B);
  CS_REQUIRE(img && heads, CS_ERR_INVALID, "cs_motion: null tensor");
  CS_REQUIRE(ctx->M.loaded, CS_ERR_STATE, "cs_motion: motion extractor weights not loaded");
  Net n = make_net(ctx, stream, false);
  ctx->arena.reset(0);
  body_motion(n, img, heads, B);
  CS_API_END(ctx)
}

int cs_keypoints(cs_ctx* ctx, const float* heads, float* x_s, float* x_can, float* R, float* deg, int B, void* stream) {
  CS_API_BEGIN(ctx)
  CS_REQUIRE(B >= 1 && B <= 65535, CS_ERR_INVALID, "cs_keypoints: bad batch");
  CS_REQUIRE(heads && x_s, CS_ERR_INVALID, "cs_keypoints: null tensor");
  Net n = make_net(ctx, stream, false);
  run_keypoints(n, heads, B, x_s, x_can, R, deg);
  CS_API_END(ctx)
}

int cs_paste_back(cs_ctx* ctx, const uint8_t* img_crop, const float* mask_crop, const double* M_c2o, const uint8_t* img_ori,
                  uint8_t* out, int B, int hc, int wc, int H, int W, void* stream) {
  CS_API_BEGIN(ctx)
  CS_REQUIRE(img_crop && mask_crop && M_c2o && img_ori && out, CS_ERR_INVALID, "cs_paste_back: null argument");
  CS_REQUIRE(B >= 1 && B <= CS_PASTE_MAX_BATCH, CS_ERR_INVALID, "cs_paste_back: batch outside [1, CS_PASTE_MAX_BATCH]");
  CS_REQUIRE(hc >= 1 && wc >= 1 && H >= 1 && W >= 1 && hc <= 16384 && wc <= 16384 && H <= 16384 && W <= 16384, CS_ERR_INVALID,
             "cs_paste_back: image size outside [1, 16384]");
  Net n = make_net(ctx, stream, false);
  paste_back(n.L, img_crop, mask_crop, M_c2o, img_ori, out, B, hc, wc, H, W);
  CS_API_END(ctx)
}

int cs_soft_erosion(cs_ctx* ctx, const float* mask, const float* kernel, float* out, uint8_t* hard, int B, int H, int W, int kernel_size,
                    float threshold, int iterations, void* stream) {
  CS_API_BEGIN(ctx)
  CS_REQUIRE(mask && kernel && out && B >= 1 && H >= 1 && W >= 1 && B <= 65535, CS_ERR_INVALID, "cs_soft_erosion: bad argument");
  const size_t need = ((size_t)2 * B * H * W + B + 64) * sizeof(float);
  if (need > ctx->se_cap) {                                  // not on the per-frame path of the generator: grown on demand
    CS_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    if (ctx->se_scratch) {                                    // the smaller buffer is not kept
      for (size_t i = 0; i < ctx->owned.size(); ++i)
        if (ctx->owned[i] == ctx->se_scratch) { ctx->owned.erase(ctx->owned.begin() + (long)i); break; }
      cudaFree(ctx->se_scratch);
      ctx->owned_bytes -= ctx->se_cap;
      ctx->se_scratch = nullptr; ctx->se_cap = 0;
    }
    ctx->se_scratch = static_cast<float*>(ctx->dmalloc(need));
    ctx->se_cap = need;
  }
  Net n = make_net(ctx, stream, false);
  soft_erosion(n.L, mask, out, hard, ctx->se_scratch, kernel, B, H, W, kernel_size, threshold, iterations);
  CS_API_END(ctx)
}

int cs_calibrate(cs_ctx* ctx, int phase, float* maxima, int cap) {
  CS_API_BEGIN(ctx)
  CS_CUDA(cudaDeviceSynchronize());
  ctx->drop_graphs();                                        // captured graphs bake the operand scales in
  if (phase == 1) {
    calibrate_begin(ctx);
  } else if (phase == 0) {
    calibrate_end(ctx, maxima, cap);
  } else if (phase == 2) {                                   // back to the unscaled operands
    ctx->calib_on = false;
    reset_activation_scales(ctx);
  } else {
    throw cs::Error(CS_ERR_INVALID, "cs_calibrate: phase must be 1 (begin), 0 (end) or 2 (reset)");
  }
  CS_API_END(ctx)
}

int cs_parse_mask(cs_ctx* ctx, const float* logits, int B, int C, int h, int w, int H, int W, uint64_t valid_classes, float* mask,
                  int32_t* labels, void* stream) {
  CS_API_BEGIN(ctx)
  CS_REQUIRE(logits && mask && B >= 1 && B <= 65535 && C >= 1 && C <= 64 && h >= 1 && w >= 1 && H >= 1 && W >= 1 && h <= 16384 &&
                 w <= 16384 && H <= 16384 && W <= 16384, CS_ERR_INVALID, "cs_parse_mask: bad argument (1 <= C <= 64 classes)");
  Net n = make_net(ctx, stream, false);
  parse_mask(n.L, logits, B, C, h, w, H, W, (unsigned long long)valid_classes, mask, labels);
  CS_API_END(ctx)
}

// ---- per-kernel-family timing (bench.py roofline leg) ---------------------------------------------
int cs_profile(cs_ctx* ctx, int enable) {
  CS_API_BEGIN(ctx)
  CS_CUDA(cudaDeviceSynchronize());
  for (auto& r : ctx->prof.recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  ctx->prof.recs.clear();
  ctx->prof.on = enable != 0;
  CS_API_END(ctx)
}

int cs_profile_read(cs_ctx* ctx, double* out) {
  CS_API_BEGIN(ctx)
  CS_REQUIRE(out != nullptr, CS_ERR_INVALID, "cs_profile_read: null output");
  CS_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < PK_N * 4; ++i) out[i] = 0.0;
  for (auto& r : ctx->prof.recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      out[r.kind * 4 + 0] += ms; out[r.kind * 4 + 1] += r.flops; out[r.kind * 4 + 2] += r.bytes; out[r.kind * 4 + 3] += 1.0;
    }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  ctx->prof.recs.clear();
  CS_API_END(ctx)
}

int cs_profile_dump(cs_ctx* ctx, char* buf, int cap) {
  CS_API_BEGIN(ctx)
  CS_REQUIRE(buf != nullptr && cap > 0, CS_ERR_INVALID, "cs_profile_dump: bad buffer");
  CS_CUDA(cudaDeviceSynchronize());
  int off = 0;
  buf[0] = 0;
  int idx = 0;
  for (auto& r : ctx->prof.recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) ms = -1.f;
    int nw = snprintf(buf + off, (size_t)(cap - off), "%d,%d,%.6f,%.6e,%.6e,%s\n", idx++, r.kind, ms, r.flops, r.bytes, r.desc);
    if (nw < 0 || nw >= cap - off) { buf[off] = 0; break; }
    off += nw;
  }
  CS_API_END(ctx)
}

// ---- kernel-level entry points -------------------------------------------------------------------
int cs_test_conv(cs_ctx* ctx, const float* x, const float* w, const float* bias, float* y, int B, int D, int H, int W, int Cin,
                 int Cout, int KD, int KH, int KW, int PD, int PH, int PW, int act, float slope, int impl, void* stream) {
  CS_API_BEGIN(ctx)
  CS_REQUIRE(x && w && y && B > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, CS_ERR_INVALID, "cs_test_conv: bad argument");
  CS_REQUIRE(KD > 0 && KH > 0 && KW > 0 && impl >= 0 && impl <= 6, CS_ERR_INVALID, "cs_test_conv: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CS_CUDA(cudaStreamSynchronize(st));
  const size_t nw = (size_t)Cout * Cin * KD * KH * KW;
  std::vector<float> hw(nw), hb;
  CS_CUDA(cudaMemcpy(hw.data(), w, nw * sizeof(float), cudaMemcpyDeviceToHost));
  if (bias) { hb.resize(Cout); CS_CUDA(cudaMemcpy(hb.data(), bias, Cout * sizeof(float), cudaMemcpyDeviceToHost)); }
  const size_t owned0 = ctx->owned.size();
  const size_t bytes0 = ctx->owned_bytes;
  auto cleanup = [&]() {
    cudaStreamSynchronize(st);
    while (ctx->owned.size() > owned0) { cudaFree(ctx->owned.back()); ctx->owned.pop_back(); }
    ctx->owned_bytes = bytes0;
  };
  try {
    ConvW cw = pack_conv_host(ctx, hw, bias ? &hb : nullptr, Cout, Cin, KD, KH, KW);
    cw.amul = std::ldexp(1.f, ctx->test_amul_log2);          // CS_OPT_TEST_AMUL: exercise the activation pre-scale
    CS_CUDA(cudaDeviceSynchronize());          // the packing kernels ran on the null stream
    const int Do = D + 2 * PD - KD + 1, Ho = H + 2 * PH - KH + 1, Wo = W + 2 * PW - KW + 1;
    CS_REQUIRE(Do > 0 && Ho > 0 && Wo > 0, CS_ERR_INVALID, "cs_test_conv: empty output");
    Act xa = make_act(const_cast<float*>(x), B, D, H, W, Cin);
    Act ya = make_act(y, B, Do, Ho, Wo, Cout);
    ConvGeom g; g.PD = PD; g.PH = PH; g.PW = PW; g.Do = Do; g.Ho = Ho; g.Wo = Wo;
    Epilogue e; e.act = act; e.slope = slope;
    Launcher L; L.stream = st; L.counter = &ctx->launches; L.npass = ctx->tc_passes; L.prof = &ctx->prof; L.max_sets = ctx->tc_sets; L.acc_comp = (float)ctx->tc_comp; L.pair = ctx->tc_pair != 0; L.pair_min_iter = ctx->tc_pair > 1 ? ctx->tc_pair : 16; L.double_buffer = ctx->tc_dbuf != 0; L.single_chain = ctx->tc_single_chain;
    const bool same = (Ho == H && Wo == W && PH == KH / 2 && PW == KW / 2) &&
                      ((Do == D && PD == KD / 2) || (Do == 1 && KD == D && PD == 0));
    bool tc = same && conv_tc_supported(cw, ya);
    if (impl == 2) CS_REQUIRE(tc, CS_ERR_INVALID, "cs_test_conv: shape not supported by the tcgen05 conv");
    if (impl == 1) tc = false;
    if (impl == 6) {
      // phase form: x is the LOW-resolution input of a conv applied to nearest-upsample(x, (1,2,2)); y [B,D,2H,2W,Cout]
      CS_REQUIRE(KH == 3 && KW == 3 && (KD == 1 || KD == 3) && PH == 1 && PW == 1 && PD == KD / 2 && Cout % 16 == 0 && (Cout <= 256 || Cout % 256 == 0),
                 CS_ERR_INVALID, "cs_test_conv: shape not supported by the phase-form conv");
      ConvW pw = pack_phase_conv_host(ctx, hw, bias ? &hb : nullptr, Cout, Cin, KD, 1);
      CS_CUDA(cudaDeviceSynchronize());
      Act yup = make_act(y, B, D, 2 * H, 2 * W, Cout);
      Arena tmp; tmp.measuring = true;
      conv_tc_alloc_operand(tmp, pw, xa);
      Arena real; real.cap = tmp.high + 4096; real.base = static_cast<char*>(ctx->dmalloc(real.cap));
      Opd opd = conv_tc_alloc_operand(real, pw, xa);
      Prep p; p.src0 = xa;
      prep_planes(L, p, opd, nullptr);
      ConvGeom gp; gp.PD = PD; gp.PH = 1; gp.PW = 1; gp.Do = D; gp.Ho = 2 * H; gp.Wo = 2 * W;
      Epilogue ep; ep.act = act; ep.slope = slope; ep.phase_shift = 1;
      conv_tc(L, opd, pw, gp, ep, yup);
    } else if (impl == 5) {
      // Winograd F(2x2,3x3) form (wino.cu): input transform -> 16 GEMMs on the tcgen05 kernel -> output transform
      pack_wino_static(ctx, cw);
      if (cw.wn) cw.wn->amul = cw.amul;                     // the pre-scale applies to the transformed operand V
      CS_CUDA(cudaDeviceSynchronize());
      L.winograd = true;
      CS_REQUIRE(same && D == 1 && wino_ok(L, cw, H, W), CS_ERR_INVALID, "cs_test_conv: shape not supported by the Winograd conv");
      Arena tmp; tmp.measuring = true;
      { Launcher dry = L; dry.dry = true; dry.counter = nullptr; dry.prof = nullptr;
        wino_conv(dry, tmp, xa, cw, nullptr, nullptr, ACT_NONE, 0.f, act, slope, nullptr, ya); }
      Arena real; real.cap = tmp.high + 4096; real.base = static_cast<char*>(ctx->dmalloc(real.cap));
      wino_conv(L, real, xa, cw, nullptr, nullptr, ACT_NONE, 0.f, act, slope, nullptr, ya);
    } else if (impl == 4) {
      // the depth-stacked 32 -> 32 3x3x3 kernel
      pack_conv3s(ctx, cw);
      CS_CUDA(cudaDeviceSynchronize());
      CS_REQUIRE(same && D == 16 && conv3s_supported(cw, H, W), CS_ERR_INVALID, "cs_test_conv: shape not supported by the stacked 3x3x3 kernel");
      Arena tmp; tmp.measuring = true;
      conv_tc_alloc_operand(tmp, cw, xa);
      Arena real; real.cap = tmp.high + 4096; real.base = static_cast<char*>(ctx->dmalloc(real.cap));
      Opd opd = conv_tc_alloc_operand(real, cw, xa);
      Prep p; p.src0 = xa;
      prep_planes(L, p, opd, nullptr);
      conv3s_tc(L, opd, cw, e, ya);
    } else if (impl == 3) {
      // the depth-stacked 7x7x7 kernel: output rows padded to a multiple of 4 channels, then compacted
      pack_conv7(ctx, cw);
      CS_CUDA(cudaDeviceSynchronize());
      const int ldo = (Cout + 3) / 4 * 4;
      float* tmp_out = static_cast<float*>(ctx->dmalloc((size_t)B * D * H * W * ldo * sizeof(float)));
      Act oa = make_act(tmp_out, B, D, H, W, Cout, ldo);
      CS_REQUIRE(same && conv7_supported(cw, oa), CS_ERR_INVALID, "cs_test_conv: shape not supported by the 7x7x7 kernel");
      Arena tmp; tmp.measuring = true;
      conv_tc_alloc_operand(tmp, cw, xa);
      Arena real; real.cap = tmp.high + 4096; real.base = static_cast<char*>(ctx->dmalloc(real.cap));
      Opd opd = conv_tc_alloc_operand(real, cw, xa);
      float* parts = static_cast<float*>(ctx->dmalloc(conv7_scratch_floats(oa) * sizeof(float)));
      Prep p; p.src0 = xa;
      prep_planes(L, p, opd, nullptr);
      conv7_tc(L, opd, cw, oa, parts);
      CS_CUDA(cudaMemcpy2DAsync(y, (size_t)Cout * 4, tmp_out, (size_t)ldo * 4, (size_t)Cout * 4, (size_t)B * D * H * W,
                                cudaMemcpyDeviceToDevice, st));
    } else if (tc) {
      // private scratch for the operand planes (the ctx arena may not exist before cs_load_weights)
      Arena tmp; tmp.measuring = true;
      conv_tc_alloc_operand(tmp, cw, xa);
      Arena real; real.cap = tmp.high + 4096; real.base = static_cast<char*>(ctx->dmalloc(real.cap));
      Opd opd = conv_tc_alloc_operand(real, cw, xa);
      Prep p; p.src0 = xa;
      prep_planes(L, p, opd, nullptr);
      conv_tc(L, opd, cw, g, e, ya);
    } else if (Cout == 1) {
      conv_cout1(L, xa, cw, g, act, y);
    } else {
      conv_simt(L, xa, cw, g, e, ya, 0);
    }
  } catch (...) { cleanup(); throw; }
  cleanup();
  CS_API_END(ctx)
}

int cs_test_grid_sample3d(cs_ctx* ctx, const float* inp, const float* grid, float* out, int B, int C, int D, int H, int W,
                          void* stream) {
  CS_API_BEGIN(ctx)
  CS_REQUIRE(inp && grid && out && B > 0, CS_ERR_INVALID, "cs_test_grid_sample3d: bad argument");
  CS_REQUIRE(C == 32 && D == 16, CS_ERR_INVALID, "cs_test_grid_sample3d: only the 32x16 feature volume is supported");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Launcher L; L.stream = st; L.counter = &ctx->launches;
  const size_t n = (size_t)B * 512 * H * W;
  float *vin = nullptr, *vout = nullptr;
  CS_CUDA(cudaMalloc(&vin, n * sizeof(float)));
  if (cudaMalloc(&vout, n * sizeof(float)) != cudaSuccess) { cudaFree(vin); throw cs::Error(CS_ERR_NOMEM, "cudaMalloc failed"); }
  try {
    nchw_to_cl(L, inp, vin, B, 512, (long)H * W, 1);
    grid_sample3d_cl(L, vin, grid, vout, B, D, H, W);
    cl_to_nchw(L, vout, out, B, 512, (long)H * W, 1, 512);
  } catch (...) { cudaStreamSynchronize(st); cudaFree(vin); cudaFree(vout); throw; }
  cudaStreamSynchronize(st);
  cudaFree(vin); cudaFree(vout);
  CS_API_END(ctx)
}

int cs_test_instance_stats(cs_ctx* ctx, const float* x, float* mean, float* rstd, int B, int C, int S, float eps, void* stream) {
  CS_API_BEGIN(ctx)
  CS_REQUIRE(x && mean && rstd && B > 0 && C > 0 && S > 0, CS_ERR_INVALID, "cs_test_instance_stats: bad argument");
  CS_REQUIRE(B <= ctx->max_batch && C <= 512, CS_ERR_INVALID, "cs_test_instance_stats: B*C exceeds the ctx scratch");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Launcher L; L.stream = st; L.counter = &ctx->launches;
  float* cl = nullptr;
  CS_CUDA(cudaMalloc(&cl, (size_t)B * C * S * sizeof(float)));
  try {
    nchw_to_cl(L, x, cl, B, C, S, 0);
    Act a = make_act(cl, B, 1, 1, S, C);
    instance_stats(L, a, mean, rstd, eps, ctx->stats_scratch);
  } catch (...) { cudaStreamSynchronize(st); cudaFree(cl); throw; }
  cudaStreamSynchronize(st);
  cudaFree(cl);
  CS_API_END(ctx)
}

}  // extern "C"
