// Weight ingestion for the five hot-path networks (replaces can_swapper.load_cpk, reference
// src/can_swap_e2e.py:87-100) and the per-identity derivation of the modulated weight sets of
// transfer_model2 (reference src/modules/adaptive_modulate.py:148-170).
//
// load_weights() takes the raw reference state_dict entries (host pointers), and on the host
//   - folds eval-mode BatchNorm that FOLLOWS a conv into the conv's weight / bias (eps 1e-5),
//   - turns BatchNorm / GroupNorm that PRECEDES a conv into per-channel scale / shift tables,
//   - folds the spectral-norm sigma = u . (W v) of the SPADE convs (reference util.py:318-322),
//   - stacks mlp_gamma | mlp_beta of every SPADE into one Cout = 2C conv,
//   - permutes every 512-channel axis that is a view of the 32x16 volume from the reference order
//     (c*16 + d) to the internal channels-last order (d*32 + c),
//   - re-lays every conv as [tap][Cin][Cout] fp32 (SIMT path) and derives the split-fp16 tcgen05
//     operand [tap][Cout_p][Cin_p] from it on the device.
#include "tc_ptx.cuh"
#include <cmath>
#include <cstring>

namespace cs {

namespace {

constexpr double BN_EPS = 1e-5;

struct Table {
  std::map<std::string, const cs_tensor_desc*> m;
  std::string net;   // current prefix, e.g. "warping_module."
  const cs_tensor_desc* find(const std::string& key) const {
    auto it = m.find(net + key);
    if (it == m.end()) throw Error(CS_ERR_WEIGHTS, "missing tensor '" + net + key + "'");
    return it->second;
  }
  bool has(const std::string& key) const { return m.count(net + key) != 0; }
  // fp32 tensor with an exact element count
  const float* f32(const std::string& key, long numel) const {
    const cs_tensor_desc* d = find(key);
    if (d->dtype != CS_F32) throw Error(CS_ERR_WEIGHTS, "tensor '" + net + key + "' is not fp32");
    long n = 1;
    for (int i = 0; i < d->ndim; ++i) n *= d->shape[i];
    if (n != numel)
      throw Error(CS_ERR_WEIGHTS, "tensor '" + net + key + "' has " + std::to_string(n) + " elements, expected " +
                                      std::to_string(numel));
    if (!d->data) throw Error(CS_ERR_WEIGHTS, "tensor '" + net + key + "' has a null data pointer");
    return static_cast<const float*>(d->data);
  }
};

// host-side conv in PyTorch layout [Cout][Cin][taps]
struct HostConv {
  int Cout = 0, Cin = 0, KD = 1, KH = 1, KW = 1;
  std::vector<float> w;
  std::vector<float> b;     // empty when the conv has no bias
  int taps() const { return KD * KH * KW; }
};

HostConv read_conv(const Table& t, const std::string& p, int Cout, int Cin, int KD, int KH, int KW, bool bias = true,
                   const char* wname = ".weight") {
  HostConv c;
  c.Cout = Cout; c.Cin = Cin; c.KD = KD; c.KH = KH; c.KW = KW;
  long n = (long)Cout * Cin * c.taps();
  const float* w = t.f32(p + wname, n);
  c.w.assign(w, w + n);
  if (bias) {
    const float* b = t.f32(p + ".bias", Cout);
    c.b.assign(b, b + Cout);
  }
  return c;
}

struct HostAffine { std::vector<float> scale, shift; };

HostAffine read_bn(const Table& t, const std::string& p, int C) {
  const float* g = t.f32(p + ".weight", C);
  const float* be = t.f32(p + ".bias", C);
  const float* mu = t.f32(p + ".running_mean", C);
  const float* var = t.f32(p + ".running_var", C);
  HostAffine a;
  a.scale.resize(C); a.shift.resize(C);
  for (int c = 0; c < C; ++c) {
    double s = (double)g[c] / std::sqrt((double)var[c] + BN_EPS);
    a.scale[c] = (float)s;
    a.shift[c] = (float)((double)be[c] - (double)mu[c] * s);
  }
  return a;
}

// y = BN(conv(x)) -> conv'
void fold_bn_post(HostConv& c, const HostAffine& a) {
  long per = (long)c.Cin * c.taps();
  if (c.b.empty()) c.b.assign(c.Cout, 0.f);
  for (int o = 0; o < c.Cout; ++o) {
    float s = a.scale[o];
    float* w = c.w.data() + (long)o * per;
    for (long i = 0; i < per; ++i) w[i] *= s;
    c.b[o] = c.b[o] * s + a.shift[o];
  }
}

inline int vol_ref(int internal) { return (internal & 31) * 16 + (internal >> 5); }   // d*32+c -> c*16+d

void permute_out_vol(HostConv& c) {
  if (c.Cout != 512) throw Error(CS_ERR_WEIGHTS, "permute_out_vol: Cout != 512");
  long per = (long)c.Cin * c.taps();
  std::vector<float> w(c.w.size());
  for (int o = 0; o < 512; ++o) std::memcpy(w.data() + (long)o * per, c.w.data() + (long)vol_ref(o) * per, per * sizeof(float));
  c.w.swap(w);
  if (!c.b.empty()) {
    std::vector<float> b(512);
    for (int o = 0; o < 512; ++o) b[o] = c.b[vol_ref(o)];
    c.b.swap(b);
  }
}

void permute_in_vol(HostConv& c) {
  if (c.Cin != 512) throw Error(CS_ERR_WEIGHTS, "permute_in_vol: Cin != 512");
  int taps = c.taps();
  std::vector<float> w(c.w.size());
  for (int o = 0; o < c.Cout; ++o)
    for (int i = 0; i < 512; ++i)
      std::memcpy(w.data() + ((long)o * 512 + i) * taps, c.w.data() + ((long)o * 512 + vol_ref(i)) * taps, taps * sizeof(float));
  c.w.swap(w);
}

std::vector<float> permute_vec_vol(const float* v) {
  std::vector<float> r(512);
  for (int i = 0; i < 512; ++i) r[i] = v[vol_ref(i)];
  return r;
}

float* upload(cs_ctx* ctx, const float* h, size_t n) {
  float* d = static_cast<float*>(ctx->dmalloc(n * sizeof(float)));
  CS_CUDA(cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice));
  return d;
}
float* upload(cs_ctx* ctx, const std::vector<float>& h) { return upload(ctx, h.data(), h.size()); }

Affine upload_affine(cs_ctx* ctx, const HostAffine& a) {
  Affine r;
  r.scale = upload(ctx, a.scale);
  r.shift = upload(ctx, a.shift);
  return r;
}

ConvW pack(cs_ctx* ctx, const HostConv& c, int phase_shift = 0) {
  return pack_conv_host(ctx, c.w, c.b.empty() ? nullptr : &c.b, c.Cout, c.Cin, c.KD, c.KH, c.KW, phase_shift);
}

ResBlock3dW read_resblock3d(cs_ctx* ctx, const Table& t, const std::string& p) {
  ResBlock3dW r;                                                 // reference util.py:85-102
  r.bn1 = upload_affine(ctx, read_bn(t, p + ".norm1", 32));
  HostConv c1 = read_conv(t, p + ".conv1", 32, 32, 3, 3, 3);
  fold_bn_post(c1, read_bn(t, p + ".norm2", 32));
  r.conv1 = pack(ctx, c1);
  r.conv2 = pack(ctx, read_conv(t, p + ".conv2", 32, 32, 3, 3, 3));
  pack_conv3s(ctx, r.conv1);
  pack_conv3s(ctx, r.conv2);
  return r;
}

GnResBlockW read_gn_resblock(cs_ctx* ctx, const Table& t, const std::string& p) {
  GnResBlockW r;                                                 // reference util.py:515-544
  r.conv1 = pack(ctx, read_conv(t, p + ".conv1", 32, 32, 3, 3, 3));
  r.conv2 = pack(ctx, read_conv(t, p + ".conv2", 32, 32, 3, 3, 3));
  pack_conv3s(ctx, r.conv1);
  pack_conv3s(ctx, r.conv2);
  r.gn1.scale = upload(ctx, t.f32(p + ".gn1.weight", 32), 32);
  r.gn1.shift = upload(ctx, t.f32(p + ".gn1.bias", 32), 32);
  r.gn2.scale = upload(ctx, t.f32(p + ".gn2.weight", 32), 32);
  r.gn2.shift = upload(ctx, t.f32(p + ".gn2.bias", 32), 32);
  return r;
}

// spectral-norm conv in eval mode: W / (u . (W_mat v))
HostConv read_sn_conv(const Table& t, const std::string& p, int Cout, int Cin, int k, bool bias) {
  HostConv c = read_conv(t, p, Cout, Cin, 1, k, k, bias, ".weight_orig");
  long per = (long)Cin * k * k;
  const float* u = t.f32(p + ".weight_u", Cout);
  const float* v = t.f32(p + ".weight_v", per);
  double sigma = 0.0;
  for (int o = 0; o < Cout; ++o) {
    double acc = 0.0;
    const float* w = c.w.data() + (long)o * per;
    for (long i = 0; i < per; ++i) acc += (double)w[i] * (double)v[i];
    sigma += (double)u[o] * acc;
  }
  if (!(std::fabs(sigma) > 0.0)) throw Error(CS_ERR_WEIGHTS, "spectral norm sigma is zero for '" + t.net + p + "'");
  float sig32 = (float)sigma;          // the reference divides in fp32: w / sigma
  for (auto& x : c.w) x = x / sig32;
  return c;
}

// 3x3 conv applied to an input nearest-upsampled by f = 2^shift, rewritten on the low-resolution input: output phase
// (a, b) sees the source rows {y-1, y} (a = 0), {y} (0 < a < f-1) or {y, y+1} (a = f-1), so the taps that hit the same
// source pixel are summed.  Result: a 3x3 conv with f*f*Cout output rows (phase-major), zero where a phase has no tap.
// (KD = 3: the hourglass decoder convs, whose input is upsampled in (h, w) only -- the same in-plane sums at every depth tap)
HostConv phase_conv(const HostConv& c, int shift) {
  const int f = 1 << shift;
  const int KD = c.KD, T = KD * 9;
  HostConv r;
  r.Cout = c.Cout * f * f; r.Cin = c.Cin; r.KD = KD; r.KH = 3; r.KW = 3;
  r.w.assign((size_t)r.Cout * c.Cin * T, 0.f);
  auto src_tap = [&](int a, int d) { return a == 0 ? (d == 0 ? 0 : 1) : (a == f - 1 ? (d == 2 ? 2 : 1) : 1); };
  for (int a = 0; a < f; ++a)
    for (int b = 0; b < f; ++b)
      for (int co = 0; co < c.Cout; ++co)
        for (int ci = 0; ci < c.Cin; ++ci) {
          const float* w = c.w.data() + ((long)co * c.Cin + ci) * T;
          float* o = r.w.data() + (((long)(a * f + b) * c.Cout + co) * c.Cin + ci) * T;
          for (int kd = 0; kd < KD; ++kd)
            for (int dy = 0; dy < 3; ++dy)
              for (int dx = 0; dx < 3; ++dx) o[kd * 9 + src_tap(a, dy) * 3 + src_tap(b, dx)] += w[kd * 9 + dy * 3 + dx];
        }
  if (!c.b.empty()) {
    r.b.resize(r.Cout);
    for (int p = 0; p < f * f; ++p) std::memcpy(r.b.data() + (long)p * c.Cout, c.b.data(), c.Cout * sizeof(float));
  }
  return r;
}

SpadeNormW read_spade(cs_ctx* ctx, const Table& t, const std::string& p, int C, int phase_shift = 0) {
  SpadeNormW s;                                                  // reference util.py:282-302
  s.C = C;
  HostConv sh = read_conv(t, p + ".mlp_shared.0", 128, 256, 1, 3, 3);
  s.shared = pack(ctx, sh);
  if (phase_shift > 0) {
    s.phase_shift = phase_shift;
    s.shared_ph = pack(ctx, phase_conv(sh, phase_shift), phase_shift);     // one N tile (BN = 128) per output phase
  }
  HostConv g = read_conv(t, p + ".mlp_gamma", C, 128, 1, 3, 3);
  HostConv b = read_conv(t, p + ".mlp_beta", C, 128, 1, 3, 3);
  // gamma | beta as ONE conv whose output channels are interleaved in chunks of [gamma x16 | beta x16]: a 32-column
  // chunk of the tcgen05 epilogue then holds both modulation terms of 16 channels (SPADE-fused epilogue, conv_tc.cu);
  // channel c: gamma at (c/16)*32 + c%16, beta 16 further.
  HostConv gb;
  gb.Cout = 2 * C; gb.Cin = 128; gb.KD = 1; gb.KH = 3; gb.KW = 3;
  const long per = 128L * 9;
  gb.w.resize((size_t)2 * C * per);
  gb.b.resize((size_t)2 * C);
  for (int c = 0; c < C; ++c) {
    const long ng = (long)(c / 16) * 32 + (c % 16), nb = ng + 16;
    std::memcpy(gb.w.data() + ng * per, g.w.data() + (long)c * per, per * sizeof(float));
    std::memcpy(gb.w.data() + nb * per, b.w.data() + (long)c * per, per * sizeof(float));
    gb.b[ng] = g.b[c];
    gb.b[nb] = b.b[c];
  }
  s.gamma_beta = pack(ctx, gb);
  return s;
}

}  // namespace

// power of two that brings max|w| to ~2^10: the fp16 remainders of the packed tcgen05 weights stay in the normal range
float weight_prescale(const float* w, size_t n) {
  float mx = 0.f;
  for (size_t i = 0; i < n; ++i) { float a = std::fabs(w[i]); if (a > mx && std::isfinite(a)) mx = a; }
  if (!(mx > 0.f)) return 1.f;
  int k = (int)std::floor(std::log2(1024.0 / (double)mx));
  if (k < -8) k = -8;
  if (k > 24) k = 24;
  return std::ldexp(1.f, k);
}

// test entry: a 3x3 / 3x3x3 conv in phase form (input nearest-upsampled by 2^shift in (h, w))
ConvW pack_phase_conv_host(cs_ctx* ctx, const std::vector<float>& w_pt, const std::vector<float>* bias, int Cout, int Cin, int KD, int shift) {
  HostConv c;
  c.Cout = Cout; c.Cin = Cin; c.KD = KD; c.KH = 3; c.KW = 3;
  c.w = w_pt;
  if (bias) c.b = *bias;
  return pack(ctx, phase_conv(c, shift), shift);
}

// ------------------------------------------------------------------------------------------
// [Cout][Cin][taps] (PyTorch) -> device [taps][Cin][Cout] (+ bias) (+ tcgen05 operand)
// ------------------------------------------------------------------------------------------
ConvW pack_conv_host(cs_ctx* ctx, const std::vector<float>& w_pt, const std::vector<float>* bias, int Cout, int Cin,
                     int KD, int KH, int KW, int phase_shift) {
  ConvW c;
  c.Cin = Cin; c.Cout = Cout; c.KD = KD; c.KH = KH; c.KW = KW; c.phase_shift = phase_shift;
  int taps = KD * KH * KW;
  CS_REQUIRE((long)w_pt.size() == (long)Cout * Cin * taps, CS_ERR_WEIGHTS, "pack_conv_host: size mismatch");
  std::vector<float> w((size_t)taps * Cin * Cout);
  for (int o = 0; o < Cout; ++o)
    for (int i = 0; i < Cin; ++i) {
      const float* src = w_pt.data() + ((long)o * Cin + i) * taps;
      for (int t = 0; t < taps; ++t) w[((size_t)t * Cin + i) * Cout + o] = src[t];
    }
  c.w32 = upload(ctx, w);
  if (bias) c.bias = upload(ctx, *bias);
  c.wmul = weight_prescale(w_pt.data(), w_pt.size());
  pack_tc(ctx, c, nullptr);
  return c;
}

// every tcgen05 conv of the five networks (and its Winograd form), in a fixed order
template <class F>
static void for_each_conv(Weights& W, F f) {
  auto one = [&](ConvW& w) { if (w.wtc || w.w32) f(w); if (w.wn) f(*w.wn); };
  one(W.f_first); one(W.f_down[0]); one(W.f_down[1]); one(W.f_second);
  for (auto& r : W.f_res) { one(r.conv1); one(r.conv2); }
  one(W.dm_compress);
  for (auto& c : W.hg_enc) one(c);
  for (auto& c : W.hg_dec) one(c);
  for (auto& c : W.hg_dec_ph) one(c);
  one(W.hg_final); one(W.dm_mask); one(W.dm_occlusion); one(W.dm_occ_y); one(W.w_third); one(W.w_fourth);
  for (auto& a : W.ad) { one(a.mask_conv); one(a.combined); one(a.wino); }
  for (auto& r : W.t_res) { one(r.conv1); one(r.conv2); }
  for (auto& r : W.r_gn1) { one(r.conv1); one(r.conv2); }
  for (auto& r : W.r_gn3) { one(r.conv1); one(r.conv2); }
  for (auto& r : W.r_res2) { one(r.conv1); one(r.conv2); }
  one(W.g_fc); one(W.g_img);
  for (auto& b : W.g_blocks) {
    for (SpadeNormW* s : {&b.norm_0, &b.norm_1, &b.norm_s}) { one(s->shared); one(s->shared_ph); one(s->gamma_beta); }
    one(b.conv_0); one(b.conv_1); one(b.conv_s);
  }
}

static void assign_conv_ids(cs_ctx* ctx) {
  int id = 0;
  for_each_conv(ctx->W, [&](ConvW& w) { w.id = id++; });
  // the per-identity convs are filled by set_identity: give them ids now (their ConvW objects are stable members)
  for (auto& a : ctx->W.ad) { if (a.combined.id < 0) a.combined.id = id++; if (a.wino.id < 0) a.wino.id = id++; }
  ctx->n_conv_ids = id;
}

// Activation-scale calibration.  The split-fp16 operand format has fp16's exponent range: values beyond 65504 saturate and
// the lo half of values below ~0.06 is subnormal (the conv's relative error grows from 4e-7 to 2e-5 at |x| ~ 1e-3,
// tools/scale_probe.py).  Weights get a power-of-two pre-scale at pack time (ConvW::wmul); activations get one per conv
// (ConvW::amul, applied by whoever writes the operand, divided out by the conv's epilogue) chosen from the largest |input|
// seen while a representative batch runs between calibrate_begin and calibrate_end: max * amul ~ 2^10, i.e. a factor 64 of
// head-room to saturation and full precision down to 1e-4 of the maximum.
void calibrate_begin(cs_ctx* ctx) {
  CS_REQUIRE(ctx->weights_loaded, CS_ERR_STATE, "cs_calibrate before cs_load_weights");
  if (!ctx->calib_tab) ctx->calib_tab = static_cast<unsigned*>(ctx->dmalloc(sizeof(unsigned) * (size_t)(ctx->n_conv_ids + 1)));
  CS_CUDA(cudaMemset(ctx->calib_tab, 0, sizeof(unsigned) * (size_t)(ctx->n_conv_ids + 1)));
  ctx->calib_on = true;
}

int calibrate_end(cs_ctx* ctx, float* maxima, int cap) {
  CS_REQUIRE(ctx->calib_on && ctx->calib_tab, CS_ERR_STATE, "cs_calibrate(end) without cs_calibrate(begin)");
  CS_CUDA(cudaDeviceSynchronize());
  std::vector<float> mx((size_t)ctx->n_conv_ids + 1, 0.f);
  CS_CUDA(cudaMemcpy(mx.data(), ctx->calib_tab, sizeof(float) * mx.size(), cudaMemcpyDeviceToHost));
  ctx->calib_on = false;
  int changed = 0;
  auto apply = [&](ConvW& w) {
    if (w.id < 0 || w.id >= ctx->n_conv_ids || w.amul_fixed) return;
    const float m = mx[w.id];
    if (!(m > 0.f) || !std::isfinite(m)) return;                 // conv not reached by the calibration batch: keep its scale
    int k = (int)std::floor(std::log2(1024.0 / (double)m));
    if (k < -14) k = -14;
    if (k > 14) k = 14;
    const float a = std::ldexp(1.f, k);
    if (a != w.amul) { w.amul = a; ++changed; }
  };
  for_each_conv(ctx->W, apply);
  for (auto& a : ctx->W.ad) { apply(a.combined); apply(a.wino); }
  if (maxima) for (int i = 0; i < cap && i < ctx->n_conv_ids; ++i) maxima[i] = mx[i];
  return changed;
}

void reset_activation_scales(cs_ctx* ctx) {
  for_each_conv(ctx->W, [](ConvW& w) { w.amul = 1.f; });
  for (auto& a : ctx->W.ad) { a.combined.amul = 1.f; a.wino.amul = 1.f; }
}

// ------------------------------------------------------------------------------------------
// load_weights
// ------------------------------------------------------------------------------------------
void load_weights(cs_ctx* ctx, const cs_tensor_desc* table, int n) {
  CS_REQUIRE(table != nullptr && n > 0, CS_ERR_INVALID, "load_weights: empty table");
  CS_REQUIRE(!ctx->weights_loaded, CS_ERR_STATE, "weights already loaded on this ctx");
  Table t;
  for (int i = 0; i < n; ++i) {
    CS_REQUIRE(table[i].name != nullptr, CS_ERR_INVALID, "load_weights: unnamed tensor");
    t.m[table[i].name] = &table[i];
  }
  Weights& W = ctx->W;

  // ---- F: appearance feature extractor (reference appearance_feature_extractor.py:14-36) ----
  t.net = "appearance_feature_extractor.";
  {
    HostConv c = read_conv(t, "first.conv", 64, 3, 1, 3, 3);
    fold_bn_post(c, read_bn(t, "first.norm", 64));
    W.f_first = pack(ctx, c);
    int ch[3] = {64, 128, 256};
    for (int i = 0; i < 2; ++i) {
      std::string p = "down_blocks." + std::to_string(i);
      HostConv d = read_conv(t, p + ".conv", ch[i + 1], ch[i], 1, 3, 3);
      fold_bn_post(d, read_bn(t, p + ".norm", ch[i + 1]));
      W.f_down[i] = pack(ctx, d);
    }
    HostConv s = read_conv(t, "second", 512, 256, 1, 1, 1);
    permute_out_vol(s);
    W.f_second = pack(ctx, s);
    for (int i = 0; i < 6; ++i) W.f_res[i] = read_resblock3d(ctx, t, "resblocks_3d.3dr" + std::to_string(i));
  }

  // ---- W: warping network + dense motion (warping_network.py:14-44, dense_motion.py:14-27) ----
  t.net = "warping_module.";
  {
    const std::string p = "dense_motion_network";
    static const int enc[5][2] = {{HG_IN, 64}, {64, 128}, {128, 256}, {256, 512}, {512, 1024}};
    static const int dec[5][2] = {{1024, 512}, {1024, 256}, {512, 128}, {256, 64}, {128, 32}};
    for (int i = 0; i < 5; ++i) {
      std::string q = p + ".hourglass.encoder.down_blocks." + std::to_string(i);
      HostConv c = read_conv(t, q + ".conv", enc[i][1], enc[i][0], 3, 3, 3);
      fold_bn_post(c, read_bn(t, q + ".norm", enc[i][1]));
      W.hg_enc[i] = pack(ctx, c);
    }
    for (int i = 0; i < 5; ++i) {
      std::string q = p + ".hourglass.decoder.up_blocks." + std::to_string(i);
      HostConv c = read_conv(t, q + ".conv", dec[i][1], dec[i][0], 3, 3, 3);
      fold_bn_post(c, read_bn(t, q + ".norm", dec[i][1]));
      W.hg_dec[i] = pack(ctx, c);
      // the conv reads its input nearest-upsampled (1,2,2) (util.py:142-143): phase form on the LOW-resolution operand --
      // 2 x 2 instead of 3 x 3 in-plane taps (2.25x fewer MMAs) and no upsampled operand
      W.hg_dec_ph[i] = pack(ctx, phase_conv(c, 1), 1);
    }
    HostConv fin = read_conv(t, p + ".hourglass.decoder.conv", HG_OUT, HG_OUT, 3, 3, 3);
    fold_bn_post(fin, read_bn(t, p + ".hourglass.decoder.norm", HG_OUT));
    W.hg_final = pack(ctx, fin);
    W.dm_mask = pack(ctx, read_conv(t, p + ".mask", NUM_KP + 1, HG_OUT, 7, 7, 7));
    pack_conv7(ctx, W.dm_mask);
    HostConv cmp = read_conv(t, p + ".compress", 4, 32, 1, 1, 1);
    fold_bn_post(cmp, read_bn(t, p + ".norm", 4));
    W.dm_compress = pack(ctx, cmp);
    // occlusion: Conv2d(142*16 -> 1, 7x7) on prediction.view(B, 142*16, h, w), channel = c*16 + d
    // == Conv3d(142 -> 1, kernel (16,7,7), padding (0,3,3)) on the [B,142,16,h,w] prediction
    W.dm_occlusion = pack(ctx, read_conv(t, p + ".occlusion", 1, HG_OUT, 16, 7, 7));
    {  // per-tap projection weights of the occlusion conv: rows z*64 + (kh*7+kw), a 1x1x1 conv with depth-dependent weights
      HostConv oc = read_conv(t, p + ".occlusion", 1, HG_OUT, 16, 7, 7);      // [1][142][16*49]
      ConvW& y = W.dm_occ_y;
      y.Cin = HG_OUT; y.Cout = 64; y.KD = y.KH = y.KW = 1;
      y.nblk = (HG_OUT + 31) / 32; y.BN = 64; y.zrows = 64; y.Cout_p = 64 * 16;
      const long rowlen = (long)y.nblk * 64;
      y.wmul = weight_prescale(oc.w.data(), oc.w.size());
      // same accumulator plan and truncation pre-compensation as pack_tc gives a 1x1x1 conv of this shape (tc_ptx.cuh)
      const int npass = ctx->tc_passes >= 1 && ctx->tc_passes <= 3 ? ctx->tc_passes : 3;
      const tc::TcPlan plan = tc::tc_make_plan(y.BN, y.nblk, npass, ctx->tc_single_chain, ctx->tc_sets, ctx->tc_dbuf != 0, 0);
      y.plan_nsets = plan.nsets; y.plan_chunk = plan.chunk; y.plan_nacc = plan.nacc; y.plan_npass = npass; y.plan_thin = plan.thin;
      y.plan_kappa = (float)ctx->tc_poscomp * 1e-10f;
      const int last_ks = ((HG_OUT - (y.nblk - 1) * 32) + 15) / 16;
      std::vector<__nv_bfloat16> hw((size_t)y.Cout_p * rowlen, __float2bfloat16(0.f));
      for (int z = 0; z < 16; ++z)
        for (int tp = 0; tp < 49; ++tp)
          for (int ci = 0; ci < HG_OUT; ++ci) {
            const int rem = tc::tc_remaining_events(plan.nsets, plan.chunk, npass, y.nblk, y.nblk, last_ks, ci >> 5, (ci >> 4) & 1);
            const float v = oc.w[(long)ci * (16 * 49) + z * 49 + tp] * y.wmul * (1.0f + y.plan_kappa * (float)rem);
            __nv_bfloat16 hi, lo;
            split_operand(v, hi, lo);
            const long o = ((long)z * 64 + tp) * rowlen + (ci >> 5) * 64 + (ci & 31);
            hw[o] = hi; hw[o + 32] = lo;
          }
      y.wtc = static_cast<__nv_bfloat16*>(ctx->dmalloc(hw.size() * sizeof(__nv_bfloat16)));
      CS_CUDA(cudaMemcpy(y.wtc, hw.data(), hw.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
    }
    HostConv third = read_conv(t, "third.conv", 256, 512, 1, 3, 3);
    fold_bn_post(third, read_bn(t, "third.norm", 256));
    permute_in_vol(third);
    W.w_third = pack(ctx, third);
    pack_wino_static(ctx, W.w_third);
    W.w_fourth = pack(ctx, read_conv(t, "fourth", 256, 256, 1, 1, 1));
  }

  // ---- swap: transfer_model2 (adaptive_modulate.py:485-521) ----
  t.net = "transfer.";
  for (int i = 0; i < 14; ++i) {
    std::string p = "BottleNeck_2d." + std::to_string(i / 2) + (i % 2 ? ".conv2" : ".conv1");
    AdaptiveConvW& a = W.ad[i];
    HostConv base = read_conv(t, p, 512, 512, 1, 3, 3, false);
    permute_in_vol(base);
    permute_out_vol(base);
    {  // [tap][Cin][Cout] fp32 master
      std::vector<float> w((size_t)9 * 512 * 512);
      for (int o = 0; o < 512; ++o)
        for (int ci = 0; ci < 512; ++ci)
          for (int tp = 0; tp < 9; ++tp) w[((size_t)tp * 512 + ci) * 512 + o] = base.w[((size_t)o * 512 + ci) * 9 + tp];
      a.w_base = upload(ctx, w);
    }
    a.bias_param = upload(ctx, permute_vec_vol(t.f32(p + ".bias_param", 512)));
    a.fc0_w = upload(ctx, t.f32(p + ".style_fc.0.weight", 512 * 512), 512 * 512);
    a.fc0_b = upload(ctx, t.f32(p + ".style_fc.0.bias", 512), 512);
    {  // rows of fc2 (one per modulated input channel) in internal channel order
      const float* w2 = t.f32(p + ".style_fc.2.weight", 512 * 512);
      std::vector<float> w((size_t)512 * 512);
      for (int r = 0; r < 512; ++r) std::memcpy(w.data() + (size_t)r * 512, w2 + (size_t)vol_ref(r) * 512, 512 * sizeof(float));
      a.fc2_w = upload(ctx, w);
      a.fc2_b = upload(ctx, permute_vec_vol(t.f32(p + ".style_fc.2.bias", 512)));
    }
    HostConv mc = read_conv(t, p + ".mask_conv.0", 1, 512, 1, 3, 3);
    permute_in_vol(mc);
    a.mask_conv = pack(ctx, mc);
    // per-identity combined conv: filled by set_identity
    a.combined.Cin = 512; a.combined.Cout = 1024; a.combined.KD = 1; a.combined.KH = 3; a.combined.KW = 3;
    // [W | W*s*demod]: the demodulated half has filters of norm <= 1, so 2^14 keeps it far from fp16 saturation
    a.combined.wmul = std::fmin(weight_prescale(base.w.data(), base.w.size()), 16384.f);
    a.combined.w32 = static_cast<float*>(ctx->dmalloc((size_t)9 * 512 * 1024 * sizeof(float)));
    a.combined.bias = static_cast<float*>(ctx->dmalloc(1024 * sizeof(float)));
    a.style = static_cast<float*>(ctx->dmalloc(512 * sizeof(float)));
    a.demod = static_cast<float*>(ctx->dmalloc(512 * sizeof(float)));
  }
  for (int i = 0; i < 6; ++i) W.t_res[i] = read_resblock3d(ctx, t, "resblocks_3d.3dr" + std::to_string(i));

  // ---- refine: G3d (adaptive_modulate.py:700-720) ----
  t.net = "refine.";
  for (int i = 0; i < 3; ++i) W.r_gn1[i] = read_gn_resblock(ctx, t, "resblocks1." + std::to_string(i));
  for (int i = 0; i < 3; ++i) W.r_gn3[i] = read_gn_resblock(ctx, t, "resblocks3." + std::to_string(i));
  for (int i = 0; i < 3; ++i) {
    std::string p = "resblocks2." + std::to_string(i);            // ResBlock2d, reference util.py:105-128
    ResBlock2dW& r = W.r_res2[i];
    HostAffine b1 = read_bn(t, p + ".norm1", 512);
    HostAffine b1p;
    b1p.scale = permute_vec_vol(b1.scale.data());
    b1p.shift = permute_vec_vol(b1.shift.data());
    r.bn1 = upload_affine(ctx, b1p);
    HostConv c1 = read_conv(t, p + ".conv1", 512, 512, 1, 3, 3);
    fold_bn_post(c1, read_bn(t, p + ".norm2", 512));
    permute_in_vol(c1); permute_out_vol(c1);
    r.conv1 = pack(ctx, c1);
    HostConv c2 = read_conv(t, p + ".conv2", 512, 512, 1, 3, 3);
    permute_in_vol(c2); permute_out_vol(c2);
    r.conv2 = pack(ctx, c2);
    pack_wino_static(ctx, r.conv1);
    pack_wino_static(ctx, r.conv2);
  }

  // ---- G: SPADE decoder (spade_generator.py:13-39) ----
  t.net = "spade_generator.";
  W.g_fc = pack(ctx, read_conv(t, "fc", 512, 256, 1, 3, 3));
  pack_wino_static(ctx, W.g_fc);
  {
    static const char* names[8] = {"G_middle_0", "G_middle_1", "G_middle_2", "G_middle_3", "G_middle_4", "G_middle_5",
                                   "up_0", "up_1"};
    static const int fio[8][2] = {{512, 512}, {512, 512}, {512, 512}, {512, 512}, {512, 512}, {512, 512}, {512, 256}, {256, 64}};
    for (int i = 0; i < 8; ++i) {
      SpadeBlockW& b = W.g_blocks[i];
      std::string p = names[i];
      b.fin = fio[i][0]; b.fout = fio[i][1]; b.fmid = b.fin < b.fout ? b.fin : b.fout;
      b.learned_shortcut = b.fin != b.fout;
      b.conv_0 = pack(ctx, read_sn_conv(t, p + ".conv_0", b.fmid, b.fin, 3, true));
      b.conv_1 = pack(ctx, read_sn_conv(t, p + ".conv_1", b.fout, b.fmid, 3, true));
      pack_wino_static(ctx, b.conv_0);                       // 512 -> 512 / 256, 256 -> 256 (Cout % 256 == 0)
      pack_wino_static(ctx, b.conv_1);
      const int pshift = i == 6 ? 1 : (i == 7 ? 2 : 0);          // up_0 / up_1 read seg nearest-upsampled x2 / x4 (util.py:297)
      b.norm_0 = read_spade(ctx, t, p + ".norm_0", b.fin, pshift);
      b.norm_1 = read_spade(ctx, t, p + ".norm_1", b.fmid, pshift);
      if (b.learned_shortcut) {
        b.conv_s = pack(ctx, read_sn_conv(t, p + ".conv_s", b.fout, b.fin, 1, false));
        b.norm_s = read_spade(ctx, t, p + ".norm_s", b.fin, pshift);
      }
    }
  }
  W.g_img = pack(ctx, read_conv(t, "conv_img.0", 12, 64, 1, 3, 3));

  CS_CUDA(cudaDeviceSynchronize());
  assign_conv_ids(ctx);
  ctx->weights_loaded = true;
}

// ------------------------------------------------------------------------------------------
// per-identity state (adaptive_modulate.py:148-170)
// ------------------------------------------------------------------------------------------
// y[r] = act(dot(W[r,:], x) + b[r]);  one warp per row
__global__ void __launch_bounds__(256) gemv512_kernel(const float* __restrict__ Wm, const float* __restrict__ b,
                                                     const float* __restrict__ x, float* __restrict__ y, int rows, int cols,
                                                     float slope, int lrelu) {
  int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float acc = 0.f;
  for (int c = lane; c < cols; c += 32) acc = fmaf(Wm[(long)row * cols + c], x[c], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    float v = acc + b[row];
    if (lrelu) v = v > 0.f ? v : v * slope;
    y[row] = v;
  }
}

// demod[co] = rsqrt(sum_{tap,ci} (w[tap][ci][co] * s[ci])^2 + 1e-8); block = 64 couts x 4 k-slices
__global__ void __launch_bounds__(256) demod_kernel(const float* __restrict__ w, const float* __restrict__ s,
                                                   float* __restrict__ demod) {
  __shared__ float red[4][64];
  int co = blockIdx.x * 64 + (threadIdx.x & 63);
  int slice = threadIdx.x >> 6;
  float acc = 0.f;
  for (int k = slice; k < 9 * 512; k += 4) {
    float v = w[(long)k * 512 + co] * s[k & 511];
    acc = fmaf(v, v, acc);
  }
  red[slice][threadIdx.x & 63] = acc;
  __syncthreads();
  if (slice == 0) {
    float t = (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
    demod[co] = rsqrtf(t + 1e-8f);
  }
}

// combined[tap][ci][0:512] = w ; combined[tap][ci][512:1024] = (w * s[ci]) * demod[co]
__global__ void __launch_bounds__(256) combine_kernel(const float* __restrict__ w, const float* __restrict__ s,
                                                     const float* __restrict__ demod, const float* __restrict__ bias_param,
                                                     float* __restrict__ comb, float* __restrict__ comb_bias) {
  long total = 9L * 512 * 512;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int co = (int)(i & 511);
    long k = i >> 9;
    float v = w[i];
    comb[k * 1024 + co] = v;
    comb[k * 1024 + 512 + co] = (v * s[k & 511]) * demod[co];
  }
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 512) { comb_bias[t] = 0.f; comb_bias[512 + t] = bias_param[t]; }
}

void set_identity(cs_ctx* ctx, const float* id_dev, cudaStream_t stream) {
  CS_REQUIRE(ctx->weights_loaded, CS_ERR_STATE, "cs_set_identity before cs_load_weights");
  CS_REQUIRE(id_dev != nullptr, CS_ERR_INVALID, "cs_set_identity: null identity");
  float* hidden = reinterpret_cast<float*>(ctx->stats_scratch);   // 512 floats of the ctx scratch
  for (int i = 0; i < 14; ++i) {
    AdaptiveConvW& a = ctx->W.ad[i];
    gemv512_kernel<<<64, 256, 0, stream>>>(a.fc0_w, a.fc0_b, id_dev, hidden, 512, 512, 0.2f, 1);
    check_launch("style_fc0");
    gemv512_kernel<<<64, 256, 0, stream>>>(a.fc2_w, a.fc2_b, hidden, a.style, 512, 512, 0.f, 0);
    check_launch("style_fc2");
    demod_kernel<<<8, 256, 0, stream>>>(a.w_base, a.style, a.demod);
    check_launch("demod");
    combine_kernel<<<148 * 8, 256, 0, stream>>>(a.w_base, a.style, a.demod, a.bias_param, a.combined.w32, a.combined.bias);
    check_launch("combine");
    ctx->launches += 4;
    pack_tc(ctx, a.combined, stream);
    pack_wino(ctx, a, stream);
  }
  ctx->identity_set = true;
}

}  // namespace cs
