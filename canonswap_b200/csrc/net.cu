// The five hot-path networks expressed on the library's kernels (internal channels-last layout).
// Every function cites the reference forward it implements; the arithmetic order inside each
// fused step follows the reference (conv -> +bias -> BN -> act, IN -> (1+gamma) -> +beta -> act ...).
//
// Layout recap (DESIGN.md "Data layout"):
//   2-D activations  [B,H,W,C] fp32
//   feature volume   [B,H,W,16,32] fp32  == 2-D tensor with 512 channels ordered d*32+c
//                                        == 3-D tensor (D=16, C=32) with strides sd=32, sw=512
//   hourglass        [B,16,H,W,C] fp32 (NDHWC), skip concats are channel slices of one buffer
#include "ctx.cuh"

namespace cs {

namespace {

Act new_act(Net& n, int B, int D, int H, int W, int C, int Cstride = -1) {
  if (Cstride < 0) Cstride = C;
  float* p = n.A->f32((size_t)B * D * H * W * Cstride);
  return make_act(p, B, D, H, W, C, Cstride);
}

struct ConvOpts {
  int act = ACT_NONE;
  float slope = 0.f;
  const Act* residual = nullptr;
  const float* mult = nullptr;
  int xshift = 0;              // input is x nearest-upsampled by 2^xshift (no prep only)
  // tcgen05 path: also emit the next conv's operand from the epilogue (see Epilogue in common.cuh)
  const Opd* emit = nullptr;
  const Affine* emit_affine = nullptr;
  int emit_act = ACT_NONE;
  float emit_slope = 0.f;
  // SPADE-fused epilogue of a gamma|beta conv (see Epilogue::sp_x): the tensor being normalised and its statistics
  const Act* sp_x = nullptr;
  int sp_xshift = 0;
  const float* sp_mean = nullptr;
  const float* sp_rstd = nullptr;
  int phase_shift = 0;         // tcgen05 phase-mode conv of a nearest-upsampled operand (see Epilogue::phase_shift)
};

Prep prep_of(const Act& src) {
  Prep p;
  p.src0 = src;
  return p;
}

}  // namespace

// One convolution layer: optional input transform (prep) -> conv -> fused epilogue.
// With prep the conv input has the geometry of `out` and w.Cin channels; without, it is `x`.
void conv_layer(Net& n, const Prep* prep, const Act& x, const ConvW& w, const ConvOpts& o, Act out) {
  CS_REQUIRE(out.C == w.Cout, CS_ERR_INVALID, "conv_layer: Cout mismatch");
  ConvGeom g;
  g.PD = w.KD / 2; g.PH = w.KH / 2; g.PW = w.KW / 2;
  g.Do = out.D; g.Ho = out.H; g.Wo = out.W;
  Epilogue e;
  e.act = o.act; e.slope = o.slope; e.mult = o.mult;
  if (o.residual) {
    e.residual = o.residual->p;
    e.rs_b = o.residual->sb; e.rs_d = o.residual->sd; e.rs_h = o.residual->sh; e.rs_w = o.residual->sw;
  }
  size_t m = n.A->mark();
  // ---- tcgen05 path: split-fp16 operand planes produced by the prep kernel ----
  if (n.L.conv_impl != 1 && !o.xshift && conv_tc_supported(w, out)) {
    Opd opd = conv_tc_alloc_operand(*n.A, w, out);
    Prep ident;
    if (!prep) { CS_REQUIRE(x.C == w.Cin, CS_ERR_INVALID, "conv_layer: Cin mismatch"); ident = prep_of(x); }
    prep_planes(n.L, prep ? *prep : ident, opd, nullptr);
    if (n.L.stacked3 && !o.mult && out.D == 16 && conv3s_supported(w, out.H, out.W)) conv3s_tc(n.L, opd, w, e, out);
    else conv_tc(n.L, opd, w, g, e, out);
    n.A->reset(m);
    return;
  }
  // ---- fp32 SIMT path ----
  if (prep) {
    Act in = new_act(n, out.B, out.D, out.H, out.W, w.Cin);
    prep_f32(n.L, *prep, in);
    conv_simt(n.L, in, w, g, e, out, 0);
  } else {
    CS_REQUIRE(x.C == w.Cin, CS_ERR_INVALID, "conv_layer: Cin mismatch");
    conv_simt(n.L, x, w, g, e, out, o.xshift);
  }
  n.A->reset(m);
}

// tcgen05 conv on an operand that is already in split-fp16 form (several convs can share one operand)
static void conv_from_operand(Net& n, const Opd& opd, const ConvW& w, const ConvOpts& o, Act out) {
  ConvGeom g;
  g.PD = (out.D == opd.D) ? w.KD / 2 : 0; g.PH = w.KH / 2; g.PW = w.KW / 2;
  g.Do = out.D; g.Ho = out.H; g.Wo = out.W;
  Epilogue e;
  e.act = o.act; e.slope = o.slope; e.mult = o.mult; e.phase_shift = o.phase_shift;
  if (o.residual) {
    e.residual = o.residual->p;
    e.rs_b = o.residual->sb; e.rs_d = o.residual->sd; e.rs_h = o.residual->sh; e.rs_w = o.residual->sw;
  }
  if (o.emit) {
    e.emit = o.emit->p; e.emit_nblk = o.emit->nblk; e.emit_mul = o.emit->amul;
    if (o.emit_affine) { e.emit_scale = o.emit_affine->scale; e.emit_shift = o.emit_affine->shift; }
    e.emit_act = o.emit_act; e.emit_slope = o.emit_slope;
  }
  if (o.sp_x) {
    e.sp_x = o.sp_x->p; e.sp_mean = o.sp_mean; e.sp_rstd = o.sp_rstd; e.sp_C = o.sp_x->C; e.sp_xshift = o.sp_xshift;
    e.sp_Hx = o.sp_x->H; e.sp_Wx = o.sp_x->W;
    e.emit_act = o.emit_act; e.emit_slope = o.emit_slope;      // the activation after the modulation, operand or fp32 output alike
  }
  if (n.L.stacked3 && !o.mult && out.D == 16 && opd.D == 16 && conv3s_supported(w, opd.H, opd.W)) {
    conv3s_tc(n.L, opd, w, e, out);                      // 32 -> 32 3x3x3 volume conv: depth-stacked kernel
    return;
  }
  conv_tc(n.L, opd, w, g, e, out);
}

static bool use_tc(const Net& n, const ConvW& w, const Act& out) { return n.L.conv_impl != 1 && conv_tc_supported(w, out); }

// ------------------------------------------------------------------------------------------
// blocks
// ------------------------------------------------------------------------------------------
// ResBlock3d, reference util.py:94-102, in place on the volume `vol` ([B,h,w,16,32]).
static void resblock3d(Net& n, const ResBlock3dW& w, float* vol, int B, int h, int wd) {
  size_t m = n.A->mark();
  Act x = vol_as_3d(vol, B, h, wd);
  Act t = vol_as_3d(n.A->f32((size_t)B * h * wd * 512), B, h, wd);
  Prep p = prep_of(x);
  p.norm = NORM_AFFINE_C; p.scale = w.bn1.scale; p.shift = w.bn1.shift; p.act = ACT_RELU;
  ConvOpts o1; o1.act = ACT_RELU;                       // norm2 folded into conv1
  conv_layer(n, &p, x, w.conv1, o1, t);
  ConvOpts o2; o2.residual = &x;
  conv_layer(n, nullptr, t, w.conv2, o2, x);            // out = conv2(.) + x, written over x
  n.A->reset(m);
}

// A run of pre-activation residual blocks (ResBlock3d util.py:94-102 on the 3-D view, ResBlock2d util.py:120-128 on the
// 2-D view) on the tcgen05 path: only the first block's input goes through a prep kernel; every conv epilogue emits
// the next conv's split-fp16 operand (conv1: act(.) as is, conv2: act(bn1_next(x + conv2(.)))), so the intermediate
// `t` never exists in fp32 and the volume is read once per block (as the residual).
struct PreActBlock { const Affine* bn1; const ConvW* conv1; const ConvW* conv2; };

static void preact_chain_tc(Net& n, const PreActBlock* blk, int nb, Act x, int act, float slope, bool emit_from_conv2) {
  size_t m = n.A->mark();
  Opd a = conv_tc_alloc_operand(*n.A, *blk[0].conv1, x);
  Opd t = conv_tc_alloc_operand(*n.A, *blk[0].conv2, x);
  Opd b = conv_tc_alloc_operand(*n.A, *blk[0].conv1, x);
  Prep p = prep_of(x);
  p.norm = NORM_AFFINE_C; p.scale = blk[0].bn1->scale; p.shift = blk[0].bn1->shift; p.act = act; p.slope = slope;
  prep_planes(n.L, p, a, nullptr);
  Act none = x; none.p = nullptr;                        // geometry only: conv1 writes no fp32 output
  for (int i = 0; i < nb; ++i) {
    ConvOpts o1; o1.act = act; o1.slope = slope;         // norm2 folded into conv1
    o1.emit = &t;
    conv_from_operand(n, a, *blk[i].conv1, o1, none);
    ConvOpts o2; o2.residual = &x;
    const bool fuse = emit_from_conv2 && i + 1 < nb;
    if (fuse) { o2.emit = &b; o2.emit_affine = blk[i + 1].bn1; o2.emit_act = act; o2.emit_slope = slope; }
    conv_from_operand(n, t, *blk[i].conv2, o2, x);       // x = conv2(.) + x, in place
    if (fuse) {
      Opd tmp = a; a = b; b = tmp;
    } else if (i + 1 < nb) {
      // wide tiles: residual + fp32 + operand in one (non-overlapped) epilogue costs more than a bandwidth-bound prep
      Prep pn = prep_of(x);
      pn.norm = NORM_AFFINE_C; pn.scale = blk[i + 1].bn1->scale; pn.shift = blk[i + 1].bn1->shift; pn.act = act; pn.slope = slope;
      prep_planes(n.L, pn, a, nullptr);
    }
  }
  n.A->reset(m);
}

static void resblock3d_run(Net& n, const ResBlock3dW* w, int nb, float* vol, int B, int h, int wd) {
  Act x = vol_as_3d(vol, B, h, wd);
  if (use_tc(n, w[0].conv1, x)) {
    PreActBlock blk[8];
    for (int i = 0; i < nb; ++i) blk[i] = PreActBlock{&w[i].bn1, &w[i].conv1, &w[i].conv2};
    preact_chain_tc(n, blk, nb, x, ACT_RELU, 0.f, true);
  } else {
    for (int i = 0; i < nb; ++i) resblock3d(n, w[i], vol, B, h, wd);
  }
}

// ResBlock2d, reference util.py:120-128, in place on the volume seen as [B,h,w,512].
static void resblock2d(Net& n, const ResBlock2dW& w, float* vol, int B, int h, int wd) {
  size_t m = n.A->mark();
  Act x = vol_as_2d(vol, B, h, wd);
  Act t = new_act(n, B, 1, h, wd, 512);
  Prep p = prep_of(x);
  p.norm = NORM_AFFINE_C; p.scale = w.bn1.scale; p.shift = w.bn1.shift; p.act = ACT_LRELU; p.slope = 0.01f;
  ConvOpts o1; o1.act = ACT_LRELU; o1.slope = 0.01f;
  conv_layer(n, &p, x, w.conv1, o1, t);
  ConvOpts o2; o2.residual = &x;
  conv_layer(n, nullptr, t, w.conv2, o2, x);
  n.A->reset(m);
}

// ResBlock3D_stage3_leak, reference util.py:528-544 (GroupNorm(32,32) == per-(b,c) instance norm), in place.
static void gn_resblock3d(Net& n, const GnResBlockW& w, float* vol, int B, int h, int wd) {
  size_t m = n.A->mark();
  Act x = vol_as_3d(vol, B, h, wd);
  Act t1 = vol_as_3d(n.A->f32((size_t)B * h * wd * 512), B, h, wd);
  Act t2 = vol_as_3d(n.A->f32((size_t)B * h * wd * 512), B, h, wd);
  float* mean = n.A->f32((size_t)B * 32);
  float* rstd = n.A->f32((size_t)B * 32);
  conv_layer(n, nullptr, x, w.conv1, ConvOpts(), t1);
  instance_stats(n.L, t1, mean, rstd, 1e-5f, n.stats);
  Prep p = prep_of(t1);
  p.norm = NORM_STATS_BC; p.mean = mean; p.rstd = rstd; p.scale = w.gn1.scale; p.shift = w.gn1.shift;
  p.act = ACT_LRELU; p.slope = 0.01f;
  conv_layer(n, &p, t1, w.conv2, ConvOpts(), t2);
  instance_stats(n.L, t2, mean, rstd, 1e-5f, n.stats);
  Prep q = prep_of(t2);
  q.norm = NORM_STATS_BC; q.mean = mean; q.rstd = rstd; q.scale = w.gn2.scale; q.shift = w.gn2.shift;
  q.add = x; q.act = ACT_LRELU; q.slope = 0.01f;
  prep_f32(n.L, q, x);                                   // lrelu(gn2(.) + x) -> x (element-wise, safe in place)
  n.A->reset(m);
}

// A run of ResBlock3D_stage3_leak blocks (util.py:528-544) on the depth-stacked tcgen05 kernel: the conv epilogues write the
// per-tile partial sums of their output, so the GroupNorm statistics need no extra pass over the tensor (and are
// deterministic / independent of the batch size); the block's last element-wise pass lrelu(gn2(.) + x) also writes the next
// block's conv1 operand.
static void gn_resblock3d_chain_tc(Net& n, const GnResBlockW* w, int nb, float* vol, int B, int h, int wd) {
  size_t m = n.A->mark();
  Act x = vol_as_3d(vol, B, h, wd);
  Act t1 = vol_as_3d(n.A->f32((size_t)B * h * wd * 512), B, h, wd);
  Act t2 = vol_as_3d(n.A->f32((size_t)B * h * wd * 512), B, h, wd);
  float* mean = n.A->f32((size_t)B * 32);
  float* rstd = n.A->f32((size_t)B * 32);
  float* part = n.A->f32(conv3s_stats_floats(B, h, wd));
  Opd a = conv_tc_alloc_operand(*n.A, w[0].conv1, x);
  Opd b = conv_tc_alloc_operand(*n.A, w[0].conv2, x);
  prep_planes(n.L, prep_of(x), a, nullptr);
  for (int i = 0; i < nb; ++i) {
    conv3s_tc(n.L, a, w[i].conv1, Epilogue(), t1, part);
    conv3s_stats(n.L, part, B, h, wd, mean, rstd, 1e-5f);
    Prep p = prep_of(t1);
    p.norm = NORM_STATS_BC; p.mean = mean; p.rstd = rstd; p.scale = w[i].gn1.scale; p.shift = w[i].gn1.shift;
    p.act = ACT_LRELU; p.slope = 0.01f;
    prep_planes(n.L, p, b, nullptr);
    conv3s_tc(n.L, b, w[i].conv2, Epilogue(), t2, part);
    conv3s_stats(n.L, part, B, h, wd, mean, rstd, 1e-5f);
    Prep q = prep_of(t2);
    q.norm = NORM_STATS_BC; q.mean = mean; q.rstd = rstd; q.scale = w[i].gn2.scale; q.shift = w[i].gn2.shift;
    q.add = x; q.act = ACT_LRELU; q.slope = 0.01f;
    if (i + 1 < nb) prep_planes(n.L, q, a, &x);          // x = lrelu(gn2(.) + x) (element-wise, safe in place) + next operand
    else prep_f32(n.L, q, x);
  }
  n.A->reset(m);
}

static void gn_resblock3d_run(Net& n, const GnResBlockW* w, int nb, float* vol, int B, int h, int wd) {
  if (n.L.conv_impl != 1 && n.L.stacked3 && conv3s_supported(w[0].conv1, h, wd)) gn_resblock3d_chain_tc(n, w, nb, vol, B, h, wd);
  else for (int i = 0; i < nb; ++i) gn_resblock3d(n, w[i], vol, B, h, wd);
}

// ------------------------------------------------------------------------------------------
// F : AppearanceFeatureExtractor.forward, reference appearance_feature_extractor.py:38-48
// ------------------------------------------------------------------------------------------
void run_F(Net& n, const float* img_cl, int B, float* vol_out) {
  n.L.tag = "F";
  const Weights& W = n.W();
  const int H = n.ctx->net_h, Wd = n.ctx->net_w, h = n.ctx->h, w = n.ctx->w;
  size_t m = n.A->mark();
  Act img = make_act(const_cast<float*>(img_cl), B, 1, H, Wd, 3);
  ConvOpts relu; relu.act = ACT_RELU;
  Act a0 = new_act(n, B, 1, H, Wd, 64);
  conv_layer(n, nullptr, img, W.f_first, relu, a0);                       // SameBlock2d util.py:207-211
  Act a1 = new_act(n, B, 1, H, Wd, 128);
  conv_layer(n, nullptr, a0, W.f_down[0], relu, a1);                      // DownBlock2d util.py:161-166 (pool below)
  Act a2 = new_act(n, B, 1, H / 2, Wd / 2, 256);
  Prep p1 = prep_of(a1); p1.pool2 = 1;
  conv_layer(n, &p1, a1, W.f_down[1], relu, a2);
  Prep p2 = prep_of(a2); p2.pool2 = 1;
  Act vol2d = vol_as_2d(vol_out, B, h, w);
  conv_layer(n, &p2, a2, W.f_second, ConvOpts(), vol2d);                  // 1x1, Cout pre-permuted to d*32+c
  n.A->reset(m);
  resblock3d_run(n, W.f_res, 6, vol_out, B, h, w);
}

// ------------------------------------------------------------------------------------------
// DenseMotionNetwork.forward, reference dense_motion.py:67-104.
// Produces mask logits [B,16,h,w,24 (22 valid)], and the occlusion map [B,h,w].
// ------------------------------------------------------------------------------------------
static void dense_motion(Net& n, const float* vol_in, const float* kp_driving, const float* kp_source, int B,
                         Act* logits_out, float* occ) {
  const Weights& W = n.W();
  const int h = n.ctx->h, w = n.ctx->w, D = 16;
  ConvOpts relu; relu.act = ACT_RELU;
  // compress 1x1x1 32->4 + BN + ReLU (:70-72)
  Act vol3 = vol_as_3d(const_cast<float*>(vol_in), B, h, w);
  Act c4 = new_act(n, B, D, h, w, 4);
  conv_layer(n, nullptr, vol3, W.dm_compress, relu, c4);
  // concat buffers of the hourglass decoder (reference util.py:259-260: cat([out, skip]))
  const int hs[6] = {h, h / 2, h / 4, h / 8, h / 16, h / 32};
  const int ws[6] = {w, w / 2, w / 4, w / 8, w / 16, w / 32};
  CS_REQUIRE(hs[5] >= 1 && ws[5] >= 1, CS_ERR_INVALID, "feature resolution too small for the 5-level hourglass");
  // tcgen05 path: the 142-channel input of the hourglass' last conv [up4 32 | x 110] exists ONLY as its split operand
  // (5 blocks): block 0 is emitted by the last decoder conv's epilogue, blocks 1..4 are written by dm_input_operand and are,
  // as a block range, also the operand of the first encoder conv.  No fp32 copy of either tensor, no prep pass over them.
  const bool opd_path = n.L.conv_impl != 1 && conv_tc_supported(W.hg_enc[0], make_act(nullptr, B, D, h, w, 64)) &&
                        conv_tc_supported(W.hg_dec[4], make_act(nullptr, B, D, h, w, 32)) && conv_tc_supported(W.hg_final, make_act(nullptr, B, D, h, w, HG_OUT));
  Opd cat5_op, x_op;
  Act cat5, x;
  if (opd_path) {
    cat5_op = conv_tc_alloc_operand(*n.A, W.hg_final, make_act(nullptr, B, D, h, w, HG_OUT));     // 5 blocks
    x_op = cat5_op; x_op.p = cat5_op.p + 64; x_op.nblk = 4; x_op.pstride = cat5_op.nblk * 64;
  } else {
    cat5 = new_act(n, B, D, hs[0], ws[0], HG_OUT, 144);  // [up4 32 | x 110] (+2 pad floats: 16-byte rows)
    x = slice_c(cat5, 32, HG_IN);
  }
  Act cat4 = new_act(n, B, D, hs[1], ws[1], 128);        // [up3 64 | p0 64]
  Act cat3 = new_act(n, B, D, hs[2], ws[2], 256);        // [up2 128 | p1 128]
  Act cat2 = new_act(n, B, D, hs[3], ws[3], 512);        // [up1 256 | p2 256]
  Act cat1 = new_act(n, B, D, hs[4], ws[4], 1024);       // [up0 512 | p3 512]
  Act p4 = new_act(n, B, D, hs[5], ws[5], 1024);
  // hourglass input: [heat_k, deformed_k(4)] per keypoint (:29-65, :83-84)
  if (opd_path) dm_input_operand(n.L, c4, kp_driving, kp_source, NUM_KP, x_op);
  else dm_input(n.L, c4, kp_driving, kp_source, NUM_KP, x);
  // encoder: conv-BN-ReLU at full res, then avg-pool (1,2,2) into the skip slice (util.py:185-190)
  Act skips[5] = {slice_c(cat4, 64, 64), slice_c(cat3, 128, 128), slice_c(cat2, 256, 256), slice_c(cat1, 512, 512), p4};
  Act cur = x;
  for (int i = 0; i < 5; ++i) {
    size_t m = n.A->mark();
    Act e = new_act(n, B, D, hs[i], ws[i], W.hg_enc[i].Cout);
    if (i == 0 && opd_path) conv_from_operand(n, x_op, W.hg_enc[0], relu, e);
    else conv_layer(n, nullptr, cur, W.hg_enc[i], relu, e);
    Prep pp = prep_of(e); pp.pool2 = 1;
    prep_f32(n.L, pp, skips[i]);
    n.A->reset(m);
    cur = skips[i];
  }
  // decoder: nearest (1,2,2) upsample -> conv-BN-ReLU -> written next to its skip (util.py:142-147,255-262)
  Act cats[5] = {cat1, cat2, cat3, cat4, cat5};
  Act dcur = p4;
  for (int i = 0; i < 5; ++i) {
    const bool phase = n.L.conv_impl != 1 && n.L.phase_conv && W.hg_dec_ph[i].wtc != nullptr;
    if (phase || (i == 4 && opd_path)) {
      // phase form: the conv of the (1,2,2)-upsampled tensor runs on the low-resolution operand, one N tile per output phase;
      // the last decoder conv writes operand block 0 of cat5 only (no fp32 output)
      size_t m = n.A->mark();
      const ConvW& wc = phase ? W.hg_dec_ph[i] : W.hg_dec[i];
      const bool last = i == 4 && opd_path;
      Act dst = last ? make_act(nullptr, B, D, h, w, W.hg_dec[4].Cout) : slice_c(cats[i], 0, W.hg_dec[i].Cout);
      Act ingeom = phase ? make_act(nullptr, dcur.B, dcur.D, dcur.H, dcur.W, wc.Cin) : make_act(nullptr, B, D, dst.H, dst.W, wc.Cin);
      Opd in_op = conv_tc_alloc_operand(*n.A, wc, ingeom);
      Prep up = prep_of(dcur); up.upshift = phase ? 0 : 1;
      prep_planes(n.L, up, in_op, nullptr);
      ConvOpts o = relu;
      if (phase) o.phase_shift = 1;
      if (last) o.emit = &cat5_op;
      conv_from_operand(n, in_op, wc, o, dst);
      n.A->reset(m);
      dcur = cats[i];
      continue;
    }
    Act dst = slice_c(cats[i], 0, W.hg_dec[i].Cout);
    if (n.L.conv_impl == 1 || !conv_tc_supported(W.hg_dec[i], dst)) {
      ConvOpts o = relu; o.xshift = 1;
      conv_layer(n, nullptr, dcur, W.hg_dec[i], o, dst);
    } else {
      Prep up = prep_of(dcur); up.upshift = 1;
      conv_layer(n, &up, dcur, W.hg_dec[i], relu, dst);
    }
    dcur = cats[i];
  }
  Act pred = new_act(n, B, D, h, w, HG_OUT, 144);
  if (opd_path) conv_from_operand(n, cat5_op, W.hg_final, relu, pred);
  else conv_layer(n, nullptr, cat5, W.hg_final, relu, pred);
  // mask logits 7x7x7 (:88); softmax is fused into the flow/warp kernel
  // occlusion 7x7 over (c*16+d) channels == conv3d kernel (16,7,7), pad (0,3,3), then sigmoid (:98-102)
  Act logits = new_act(n, B, D, h, w, NUM_KP + 1, 24);
  *logits_out = logits;
  Act occ_act = make_act(occ, B, 1, h, w, 1);
  if (use_tc(n, W.dm_mask, logits) && use_tc(n, W.dm_occlusion, occ_act)) {
    size_t m = n.A->mark();
    Opd opd = conv_tc_alloc_operand(*n.A, W.dm_mask, pred);             // both convs read the same operand
    prep_planes(n.L, prep_of(pred), opd, nullptr);
    if (conv7_supported(W.dm_mask, logits)) {
      float* parts = n.A->f32(conv7_scratch_floats(logits));
      conv7_tc(n.L, opd, W.dm_mask, logits, parts);
    } else {
      conv_from_operand(n, opd, W.dm_mask, ConvOpts(), logits);
    }
    if (opd.H * opd.W >= 128) {                                         // one depth slice per tile
      Act Y = new_act(n, B, D, h, w, 64);
      conv_from_operand(n, opd, W.dm_occ_y, ConvOpts(), Y);
      occlusion_gather(n.L, Y, W.dm_occlusion.bias, occ);
    } else {
      ConvOpts sg; sg.act = ACT_SIGMOID;
      conv_from_operand(n, opd, W.dm_occlusion, sg, occ_act);
    }
    n.A->reset(m);
  } else {
    conv_layer(n, nullptr, pred, W.dm_mask, ConvOpts(), logits);
    ConvGeom g; g.PD = 0; g.PH = 3; g.PW = 3; g.Do = 1; g.Ho = h; g.Wo = w;
    conv_cout1(n.L, pred, W.dm_occlusion, g, ACT_SIGMOID, occ);
  }
}

// WarpingNetwork.warp, reference warping_network.py:49-62
void run_warp(Net& n, const float* vol_in, const float* kp_source, const float* kp_driving, int B, float* vol_out,
              float* occ, float* deformation) {
  n.L.tag = "warp";
  size_t m = n.A->mark();
  Act logits;
  float* occ_buf = occ ? occ : n.A->f32((size_t)B * n.ctx->h * n.ctx->w);
  dense_motion(n, vol_in, kp_driving, kp_source, B, &logits, occ_buf);
  softmax_flow_warp(n.L, logits, kp_driving, kp_source, NUM_KP, vol_in, vol_out, deformation);
  n.A->reset(m);
}

// WarpingNetwork.warp_out, reference warping_network.py:64-71
void run_warp_out(Net& n, const float* vol_in, const float* occ, int B, float* out256) {
  n.L.tag = "warp_out";
  const Weights& W = n.W();
  const int h = n.ctx->h, w = n.ctx->w;
  size_t m = n.A->mark();
  Act x = vol_as_2d(const_cast<float*>(vol_in), B, h, w);
  Act t = new_act(n, B, 1, h, w, 256);
  ConvOpts o1; o1.act = ACT_LRELU; o1.slope = 0.01f;                      // SameBlock2d(lrelu=True), BN folded
  if (wino_ok(n.L, W.w_third, h, w)) wino_conv(n.L, *n.A, x, W.w_third, nullptr, nullptr, ACT_NONE, 0.f, ACT_LRELU, 0.01f, nullptr, t);
  else conv_layer(n, nullptr, x, W.w_third, o1, t);
  ConvOpts o2; o2.mult = occ;                                             // out * occlusion_map
  conv_layer(n, nullptr, t, W.w_fourth, o2, make_act(out256, B, 1, h, w, 256));
  n.A->reset(m);
}

// ------------------------------------------------------------------------------------------
// swap : transfer_model2.forward, reference adaptive_modulate.py:522-554
// ------------------------------------------------------------------------------------------
// AdaptiveSharedWeightConv2d.forward (:128-193) as ONE conv with Cout = 1024 = [W | W*s*demod],
// a 512->1 mask conv and the blend  mask*out_mod + (1-mask)*out_std (+ReLU / +residual).
static void adaptive_conv(Net& n, const AdaptiveConvW& a, const Act& x, const float* residual, int relu, float* y,
                          float* mask) {
  size_t m = n.A->mark();
  long P = x.pixels();
  Act o2 = new_act(n, x.B, 1, x.H, x.W, 1024);
  conv_layer(n, nullptr, x, a.combined, ConvOpts(), o2);
  ConvGeom g; g.PD = 0; g.PH = 1; g.PW = 1; g.Do = 1; g.Ho = x.H; g.Wo = x.W;
  conv_cout1(n.L, x, a.mask_conv, g, ACT_SIGMOID, mask);
  adaptive_blend(n.L, o2.p, mask, residual, relu, y, nullptr, P);
  n.A->reset(m);
}

// tcgen05 variant: input already in operand form (shared by the combined conv and the mask conv); the blend
// writes the fp32 result (if wanted) and / or the next adaptive conv's operand.
static void adaptive_conv_tc(Net& n, const AdaptiveConvW& a, const Act& geom, const Opd& in, const float* residual, int relu,
                             float* y, const Opd* out, float* mask) {
  size_t m = n.A->mark();
  Act o2 = new_act(n, geom.B, 1, geom.H, geom.W, 1024);
  conv_from_operand(n, in, a.combined, ConvOpts(), o2);
  ConvOpts sg; sg.act = ACT_SIGMOID;
  conv_from_operand(n, in, a.mask_conv, sg, make_act(mask, geom.B, 1, geom.H, geom.W, 1));
  adaptive_blend(n.L, o2.p, mask, residual, relu, y, out ? out->p : nullptr, geom.pixels(), out ? out->amul : 1.f);
  n.A->reset(m);
}

void run_swap(Net& n, const float* vol_in, int B, float* vol_out, float* masks) {
  n.L.tag = "swap";
  CS_REQUIRE(n.ctx->identity_set, CS_ERR_STATE, "cs_swap / cs_frame before cs_set_identity");
  const Weights& W = n.W();
  const int h = n.ctx->h, w = n.ctx->w;
  const long P = (long)B * h * w;
  size_t m = n.A->mark();
  if (vol_out != vol_in && !n.L.dry)
    CS_CUDA(cudaMemcpyAsync(vol_out, vol_in, (size_t)P * 512 * sizeof(float), cudaMemcpyDeviceToDevice, n.L.stream));
  float* m1 = n.A->f32((size_t)P);
  float* m2 = n.A->f32((size_t)P);
  Act x = vol_as_2d(vol_out, B, h, w);
  Act mk = make_act(m1, B, 1, h, w, 1);
  if (n.L.winograd && n.L.conv_impl != 1 && (W.ad[0].wino.wtc || n.L.dry) && h % 2 == 0 && w % 2 == 0 && (long)(h / 2) * (w / 2) >= 128) {
    // Winograd F(2x2,3x3) form (wino.cu): input transform -> 16 GEMMs over the channels (depth-dependent weights) ->
    // output transform fused with the mask blend; the 512 -> 1 mask conv is computed by the input transform
    float* y1 = n.A->f32((size_t)P * 512);
    Opd V; V.B = B; V.D = 16; V.H = h / 2; V.W = w / 2; V.nblk = 16;
    V.p = n.A->bf16((size_t)B * 16 * V.H * V.W * 16 * 64);
    Act Mt = make_act(n.A->f32((size_t)B * 16 * V.H * V.W * 1024), B, 16, V.H, V.W, 1024);
    ConvGeom gg; gg.Do = 16; gg.Ho = V.H; gg.Wo = V.W;
    auto wino_adaptive = [&](const AdaptiveConvW& a, float* in, const float* residual, int relu, float* out, float* mask) {
      Act xin = vol_as_2d(in, B, h, w);
      V.amul = a.wino.amul;
      wino_in(n.L, xin, V, &a.mask_conv, mask);              // + the 512 -> 1 mask conv on the same patches
      Epilogue eg;
      eg.alg_flops = 2.0 * (double)P * 1024 * 512 * 9.0;       // the two 3x3 branches this GEMM stands for
      conv_tc(n.L, V, a.wino, gg, eg, Mt);
      wino_out_blend(n.L, Mt.p, mask, a.bias_param, residual, relu, out, B, h, w);
    };
    for (int i = 0; i < 7; ++i) {                                             // ResnetBlock_Adaptive2D :337-349
      wino_adaptive(W.ad[2 * i], vol_out, nullptr, 1, y1, m1);                // y = relu(conv1(x))
      wino_adaptive(W.ad[2 * i + 1], y1, vol_out, 0, vol_out, m2);            // x + conv2(y), in place
      if (masks) avg2(n.L, m1, m2, masks + (long)i * P, P);
    }
  } else if (use_tc(n, W.ad[0].combined, make_act(nullptr, B, 1, h, w, 1024)) && use_tc(n, W.ad[0].mask_conv, mk)) {
    Opd oa = conv_tc_alloc_operand(*n.A, W.ad[0].combined, x);
    Opd ob = conv_tc_alloc_operand(*n.A, W.ad[0].combined, x);
    prep_planes(n.L, prep_of(x), oa, nullptr);
    for (int i = 0; i < 7; ++i) {                                           // ResnetBlock_Adaptive2D :337-349
      adaptive_conv_tc(n, W.ad[2 * i], x, oa, nullptr, 1, nullptr, &ob, m1);              // y = relu(conv1(x)): operand only
      adaptive_conv_tc(n, W.ad[2 * i + 1], x, ob, vol_out, 0, vol_out, i < 6 ? &oa : nullptr, m2);   // x + conv2(y), in place
      if (masks) avg2(n.L, m1, m2, masks + (long)i * P, P);
    }
  } else {
    float* y1 = n.A->f32((size_t)P * 512);
    Act t = vol_as_2d(y1, B, h, w);
    for (int i = 0; i < 7; ++i) {
      adaptive_conv(n, W.ad[2 * i], x, nullptr, 1, y1, m1);
      adaptive_conv(n, W.ad[2 * i + 1], t, vol_out, 0, vol_out, m2);        // x + y, in place
      if (masks) avg2(n.L, m1, m2, masks + (long)i * P, P);
    }
  }
  n.A->reset(m);
  resblock3d_run(n, W.t_res, 6, vol_out, B, h, w);
}

// ------------------------------------------------------------------------------------------
// refine : G3d.forward, reference adaptive_modulate.py:721-733
// ------------------------------------------------------------------------------------------
void run_refine(Net& n, const float* vol_in, int B, float* vol_out) {
  n.L.tag = "refine";
  const Weights& W = n.W();
  const int h = n.ctx->h, w = n.ctx->w;
  if (vol_out != vol_in && !n.L.dry)
    CS_CUDA(cudaMemcpyAsync(vol_out, vol_in, (size_t)B * h * w * 512 * sizeof(float), cudaMemcpyDeviceToDevice, n.L.stream));
  gn_resblock3d_run(n, W.r_gn1, 3, vol_out, B, h, w);
  {
    Act x2 = vol_as_2d(vol_out, B, h, w);
    if (wino_ok(n.L, W.r_res2[0].conv1, h, w) && wino_ok(n.L, W.r_res2[0].conv2, h, w)) {
      // ResBlock2d (util.py:120-128) in Winograd form: bn1 + lrelu in the input transform, norm2 folded into conv1
      size_t m = n.A->mark();
      Act t = new_act(n, B, 1, h, w, 512);
      for (int i = 0; i < 3; ++i) {
        const ResBlock2dW& r = W.r_res2[i];
        wino_conv(n.L, *n.A, x2, r.conv1, r.bn1.scale, r.bn1.shift, ACT_LRELU, 0.01f, ACT_LRELU, 0.01f, nullptr, t);
        wino_conv(n.L, *n.A, t, r.conv2, nullptr, nullptr, ACT_NONE, 0.f, ACT_NONE, 0.f, x2.p, x2);   // x = conv2(.) + x, in place
      }
      n.A->reset(m);
    } else if (use_tc(n, W.r_res2[0].conv1, x2)) {
      PreActBlock blk[3];
      for (int i = 0; i < 3; ++i) blk[i] = PreActBlock{&W.r_res2[i].bn1, &W.r_res2[i].conv1, &W.r_res2[i].conv2};
      preact_chain_tc(n, blk, 3, x2, ACT_LRELU, 0.01f, false);
    } else {
      for (int i = 0; i < 3; ++i) resblock2d(n, W.r_res2[i], vol_out, B, h, w);
    }
  }
  gn_resblock3d_run(n, W.r_gn3, 3, vol_out, B, h, w);
}

// ------------------------------------------------------------------------------------------
// G : SPADEDecoder.forward, reference spade_generator.py:41-59
// ------------------------------------------------------------------------------------------
// SPADE (util.py:295-302): gamma|beta = conv(relu(conv(nearest(seg)))) as one Cout = 2C tensor [P, 2C]
static float* spade_gamma_beta(Net& n, const SpadeNormW& s, const Act& seg, int segshift, int B, int H, int W) {
  Act gb = new_act(n, B, 1, H, W, 2 * s.C);
  size_t m = n.A->mark();
  Act actv = new_act(n, B, 1, H, W, 128);
  ConvOpts relu; relu.act = ACT_RELU;
  if (segshift == 0) {
    conv_layer(n, nullptr, seg, s.shared, relu, actv);
  } else if (n.L.conv_impl == 1 || !conv_tc_supported(s.shared, actv)) {
    relu.xshift = segshift;
    conv_layer(n, nullptr, seg, s.shared, relu, actv);
  } else {
    Prep up = prep_of(seg); up.upshift = segshift;
    conv_layer(n, &up, seg, s.shared, relu, actv);
  }
  conv_layer(n, nullptr, actv, s.gamma_beta, ConvOpts(), gb);
  n.A->reset(m);
  return gb.p;
}

// tcgen05 form of one SPADE normalisation + activation (util.py:295-302) feeding `consumer`: two kernels, no fp32
// intermediates.  mlp_shared runs on the (shared) seg operand and emits relu(.) as the operand of the gamma|beta conv,
// whose SPADE epilogue reads x and its instance statistics and emits act(x_hat * (1 + gamma) + beta) directly as the
// consumer conv's operand.
// out32 != null: the modulated activation is written as fp32 [B,H,W,C] instead (input of a Winograd conv)
static Opd spade_norm_tc(Net& n, const SpadeNormW& s, const Opd& seg_op, int seg_phase, const Act& x, int xup, const float* mean,
                         const float* rstd, int act, float slope, const ConvW& consumer, int B, int H, int W,
                         const Act* out32 = nullptr) {
  Act geom128 = make_act(nullptr, B, 1, H, W, 128);
  Act geom2c = make_act(nullptr, B, 1, H, W, 2 * s.C);
  Opd mod;
  if (!out32) mod = conv_tc_alloc_operand(*n.A, consumer, geom2c);    // survives this call (caller resets the arena)
  size_t m = n.A->mark();
  Opd actv = conv_tc_alloc_operand(*n.A, s.gamma_beta, geom128);
  ConvOpts o1; o1.act = ACT_RELU; o1.emit = &actv;
  if (seg_phase > 0) {                                   // seg_op is the LOW-resolution seg: phase-form conv, upsampled output
    o1.phase_shift = seg_phase;
    conv_from_operand(n, seg_op, s.shared_ph, o1, geom128);
  } else {
    conv_from_operand(n, seg_op, s.shared, o1, geom128);
  }
  ConvOpts o2; o2.emit = out32 ? nullptr : &mod; o2.emit_act = act; o2.emit_slope = slope;
  o2.sp_x = &x; o2.sp_xshift = xup; o2.sp_mean = mean; o2.sp_rstd = rstd;
  conv_from_operand(n, actv, s.gamma_beta, o2, out32 ? *out32 : geom2c);
  n.A->reset(m);
  return mod;
}

static bool spade_block_tc_ok(const Net& n, const SpadeBlockW& b) {
  if (n.L.conv_impl == 1 || !n.L.spade_fused) return false;
  const ConvW* ws[] = {&b.norm_0.shared, &b.norm_0.gamma_beta, &b.norm_1.shared, &b.norm_1.gamma_beta, &b.conv_0, &b.conv_1};
  for (const ConvW* w : ws) if (!w->wtc) return false;
  if (b.learned_shortcut && !(b.norm_s.shared.wtc && b.norm_s.gamma_beta.wtc && b.conv_s.wtc)) return false;
  return b.fin % 32 == 0 && b.fmid % 32 == 0;
}

// SPADEResnetBlock (util.py:329-344) on the tcgen05 path. seg_op: split operand of the (upsampled) seg map at this
// block's resolution, shared by the block's mlp_shared convs.
// xstat: statistics of x when its producer already computed them (else they are computed here); ostat: request for the
// statistics of the block's output (filled by the Winograd output transform of conv_1 when it runs in that form)
struct StatPair { float* mean = nullptr; float* rstd = nullptr; bool valid = false; };

static Act spade_block_tc(Net& n, const SpadeBlockW& b, const Act& x, int xup, const Opd& seg_op, int seg_phase, Act out,
                          const StatPair* xstat, StatPair* ostat) {
  const int B = x.B, H = x.H << xup, W = x.W << xup;
  size_t m = n.A->mark();
  float* mean = n.A->f32((size_t)B * b.fin);
  float* rstd = n.A->f32((size_t)B * b.fin);
  if (xstat && xstat->valid) { mean = xstat->mean; rstd = xstat->rstd; }
  else instance_stats(n.L, x, mean, rstd, 1e-5f, n.stats);
  if (ostat) ostat->valid = false;
  Act xs;
  if (b.learned_shortcut) {
    xs = new_act(n, B, 1, H, W, b.fout);
    size_t m2 = n.A->mark();
    Opd ms = spade_norm_tc(n, b.norm_s, seg_op, seg_phase, x, xup, mean, rstd, ACT_NONE, 0.f, b.conv_s, B, H, W);
    conv_from_operand(n, ms, b.conv_s, ConvOpts(), xs);
    n.A->reset(m2);
  } else {
    CS_REQUIRE(xup == 0, CS_ERR_INVALID, "identity shortcut needs an un-upsampled input");
    xs = x;
  }
  Act dx = new_act(n, B, 1, H, W, b.fmid);
  float* mean1 = n.A->f32((size_t)B * b.fmid);
  float* rstd1 = n.A->f32((size_t)B * b.fmid);
  bool have1 = false;
  {
    size_t m2 = n.A->mark();
    if (wino_ok(n.L, b.conv_0, H, W)) {                    // Winograd F(2x2,3x3): the SPADE epilogue writes fp32
      Act mod = new_act(n, B, 1, H, W, b.fin);
      spade_norm_tc(n, b.norm_0, seg_op, seg_phase, x, xup, mean, rstd, ACT_LRELU, 0.2f, b.conv_0, B, H, W, &mod);
      StatsOut so; so.scratch = n.stats; so.mean = mean1; so.rstd = rstd1;       // statistics of dx from the output transform
      wino_conv(n.L, *n.A, mod, b.conv_0, nullptr, nullptr, ACT_NONE, 0.f, ACT_NONE, 0.f, nullptr, dx, &so);
      have1 = true;
    } else {
      Opd m0 = spade_norm_tc(n, b.norm_0, seg_op, seg_phase, x, xup, mean, rstd, ACT_LRELU, 0.2f, b.conv_0, B, H, W);
      conv_from_operand(n, m0, b.conv_0, ConvOpts(), dx);
    }
    n.A->reset(m2);
  }
  {
    if (!have1) instance_stats(n.L, dx, mean1, rstd1, 1e-5f, n.stats);
    const bool xs_dense = xs.sw == xs.C && xs.sh == (long)xs.W * xs.C && xs.sb == (long)xs.H * xs.W * xs.C && xs.H == H && xs.W == W;
    if (wino_ok(n.L, b.conv_1, H, W) && xs_dense) {
      Act mod = new_act(n, B, 1, H, W, b.fmid);
      spade_norm_tc(n, b.norm_1, seg_op, seg_phase, dx, 0, mean1, rstd1, ACT_LRELU, 0.2f, b.conv_1, B, H, W, &mod);
      StatsOut so; so.scratch = n.stats;
      if (ostat && ostat->mean) { so.mean = ostat->mean; so.rstd = ostat->rstd; ostat->valid = true; }
      wino_conv(n.L, *n.A, mod, b.conv_1, nullptr, nullptr, ACT_NONE, 0.f, ACT_NONE, 0.f, xs.p, out, ostat && ostat->mean ? &so : nullptr);
    } else {
      Opd m1 = spade_norm_tc(n, b.norm_1, seg_op, seg_phase, dx, 0, mean1, rstd1, ACT_LRELU, 0.2f, b.conv_1, B, H, W);
      ConvOpts o; o.residual = &xs;
      conv_from_operand(n, m1, b.conv_1, o, out);
    }
  }
  n.A->reset(m);
  return out;
}

// SPADEResnetBlock (util.py:329-344). x is read nearest-upsampled by 2^xup; returns [B,H,W,fout].
static Act spade_block(Net& n, const SpadeBlockW& b, const Act& x, int xup, const Act& seg, int segshift, Act out,
                       const StatPair* xstat = nullptr, StatPair* ostat = nullptr) {
  const int B = x.B, H = x.H << xup, W = x.W << xup;
  if (ostat) ostat->valid = false;
  if (spade_block_tc_ok(n, b)) {
    size_t m0 = n.A->mark();
    // the seg map's operand at its own resolution; up blocks convolve it in phase form instead of upsampling it
    const bool phase = segshift > 0 && n.L.phase_conv && b.norm_0.shared_ph.wtc && b.norm_1.shared_ph.wtc &&
                       (!b.learned_shortcut || b.norm_s.shared_ph.wtc) && b.norm_0.phase_shift == segshift;
    const int sh = phase ? 0 : segshift;
    Opd seg_op = conv_tc_alloc_operand(*n.A, b.norm_0.shared, make_act(nullptr, B, 1, seg.H << sh, seg.W << sh, 128));
    Prep up = prep_of(seg); up.upshift = sh;
    prep_planes(n.L, up, seg_op, nullptr);
    spade_block_tc(n, b, x, xup, seg_op, phase ? segshift : 0, out, xstat, ostat);
    n.A->reset(m0);
    return out;
  }
  size_t m = n.A->mark();
  // InstanceNorm statistics of x (nearest upsampling leaves mean / biased variance unchanged)
  float* mean = n.A->f32((size_t)B * b.fin);
  float* rstd = n.A->f32((size_t)B * b.fin);
  instance_stats(n.L, x, mean, rstd, 1e-5f, n.stats);
  // shortcut
  Act xs;
  if (b.learned_shortcut) {
    xs = new_act(n, B, 1, H, W, b.fout);
    size_t m2 = n.A->mark();
    float* gbs = spade_gamma_beta(n, b.norm_s, seg, segshift, B, H, W);
    Prep ps = prep_of(x); ps.upshift = xup; ps.norm = NORM_STATS_BC; ps.mean = mean; ps.rstd = rstd; ps.gb = gbs;
    conv_layer(n, &ps, x, b.conv_s, ConvOpts(), xs);
    n.A->reset(m2);
  } else {
    CS_REQUIRE(xup == 0, CS_ERR_INVALID, "identity shortcut needs an un-upsampled input");
    xs = x;
  }
  // dx = conv_0(lrelu(norm_0(x, seg), 0.2))
  Act dx = new_act(n, B, 1, H, W, b.fmid);
  {
    size_t m2 = n.A->mark();
    float* gb0 = spade_gamma_beta(n, b.norm_0, seg, segshift, B, H, W);
    Prep p0 = prep_of(x); p0.upshift = xup; p0.norm = NORM_STATS_BC; p0.mean = mean; p0.rstd = rstd; p0.gb = gb0;
    p0.act = ACT_LRELU; p0.slope = 0.2f;
    conv_layer(n, &p0, x, b.conv_0, ConvOpts(), dx);
    n.A->reset(m2);
  }
  // out = x_s + conv_1(lrelu(norm_1(dx, seg), 0.2))
  {
    float* mean1 = n.A->f32((size_t)B * b.fmid);
    float* rstd1 = n.A->f32((size_t)B * b.fmid);
    instance_stats(n.L, dx, mean1, rstd1, 1e-5f, n.stats);
    float* gb1 = spade_gamma_beta(n, b.norm_1, seg, segshift, B, H, W);
    Prep p1 = prep_of(dx); p1.norm = NORM_STATS_BC; p1.mean = mean1; p1.rstd = rstd1; p1.gb = gb1;
    p1.act = ACT_LRELU; p1.slope = 0.2f;
    ConvOpts o; o.residual = &xs;
    conv_layer(n, &p1, dx, b.conv_1, o, out);
  }
  n.A->reset(m);
  return out;
}

void run_spade(Net& n, const float* feat256, int B, float* img_nchw, uint8_t* img_u8) {
  n.L.tag = "spade";
  const Weights& W = n.W();
  const int h = n.ctx->h, w = n.ctx->w;
  size_t m = n.A->mark();
  Act seg = make_act(const_cast<float*>(feat256), B, 1, h, w, 256);
  Act xa = new_act(n, B, 1, h, w, 512);
  Act xb = new_act(n, B, 1, h, w, 512);
  // the InstanceNorm statistics of every block input travel with the tensor: they are produced by the Winograd output
  // transform of the conv that wrote it (no separate pass; nearest upsampling leaves mean / biased variance unchanged)
  StatPair sa, sb;
  sa.mean = n.A->f32((size_t)B * 512); sa.rstd = n.A->f32((size_t)B * 512);
  sb.mean = n.A->f32((size_t)B * 512); sb.rstd = n.A->f32((size_t)B * 512);
  if (wino_ok(n.L, W.g_fc, h, w)) {
    StatsOut so; so.scratch = n.stats; so.mean = sa.mean; so.rstd = sa.rstd;
    wino_conv(n.L, *n.A, seg, W.g_fc, nullptr, nullptr, ACT_NONE, 0.f, ACT_NONE, 0.f, nullptr, xa, &so);
    sa.valid = true;
  } else {
    conv_layer(n, nullptr, seg, W.g_fc, ConvOpts(), xa);
  }
  for (int i = 0; i < 6; ++i) {                                           // G_middle_0..5
    spade_block(n, W.g_blocks[i], xa, 0, seg, 0, xb, &sa, &sb);
    Act t = xa; xa = xb; xb = t;
    StatPair ts = sa; sa = sb; sb = ts;
  }
  Act u0 = new_act(n, B, 1, 2 * h, 2 * w, 256);
  spade_block(n, W.g_blocks[6], xa, 1, seg, 1, u0, &sa, &sb);             // up -> up_0
  Act u1 = new_act(n, B, 1, 4 * h, 4 * w, 64);
  spade_block(n, W.g_blocks[7], u0, 1, seg, 2, u1, &sb, nullptr);         // up -> up_1
  Act y = new_act(n, B, 1, 4 * h, 4 * w, 12);
  Prep pl = prep_of(u1); pl.act = ACT_LRELU; pl.slope = 0.2f;             // conv_img(leaky_relu(x, 0.2))
  conv_layer(n, &pl, u1, W.g_img, ConvOpts(), y);
  emit_image(n.L, y.p, 12, img_nchw, img_u8, B, 4 * h, 4 * w);            // PixelShuffle(2) + sigmoid (+ u8)
  n.A->reset(m);
}

}  // namespace cs
