"""CPU oracle for the CanonSwap per-frame generator hot path  --  TEST INFRASTRUCTURE ONLY.

A functional, plain-PyTorch fp32 restatement of the reference algorithm for the path
`src/can_swap_pipeline_e2e.py:242-263` (F -> W.warp -> swap -> refine -> W.forward -> G).
It takes the reference's own flat state_dicts (keys as in `canonswap_b200/spec.py`) and
computes with `torch.nn.functional` ops on CPU.  Nothing in the product path may import this
file: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` do.

PARITY PINNING.  The reference ships no golden vectors, KATs or weights for this path
(SURVEY.md section 8c).  The oracle is therefore pinned against outputs of the reference
modules themselves, executed in the authoring container from /root/reference with the same
seeded synthetic state_dicts: `tests/test_oracle_vs_reference.py` (live, when /root/reference
exists) and `tests/golden/*.npz` produced by `tests/golden/make_golden.py` (committed, checked
everywhere).

Each function cites the reference file:line it follows.  All tensors are NCHW / NCDHW fp32
as in the reference.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5     # nn.BatchNorm default, reference util.py:91,116,140,158,182,201,253
IN_EPS = 1e-5     # nn.InstanceNorm2d default, reference util.py:286
GN_EPS = 1e-5     # nn.GroupNorm default, reference util.py:521
DEMOD_EPS = 1e-8  # reference adaptive_modulate.py:85


def _bn(x, sd, p):
    """eval-mode BatchNorm{2,3}d (running statistics)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def _conv2d(x, sd, p, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), padding=padding)


def _conv3d(x, sd, p, padding=0):
    return F.conv3d(x, sd[p + ".weight"], sd.get(p + ".bias"), padding=padding)


def resblock3d(x, sd, p):
    """reference util.py:94-102  BN-ReLU-conv3 x2 + skip."""
    out = F.relu(_bn(x, sd, p + ".norm1"))
    out = _conv3d(out, sd, p + ".conv1", 1)
    out = F.relu(_bn(out, sd, p + ".norm2"))
    out = _conv3d(out, sd, p + ".conv2", 1)
    return out + x


def resblock2d(x, sd, p, slope=0.01):
    """reference util.py:120-128  BN-LeakyReLU(0.01)-conv3 x2 + skip."""
    out = F.leaky_relu(_bn(x, sd, p + ".norm1"), slope)
    out = _conv2d(out, sd, p + ".conv1", 1)
    out = F.leaky_relu(_bn(out, sd, p + ".norm2"), slope)
    out = _conv2d(out, sd, p + ".conv2", 1)
    return out + x


def gn_resblock3d(x, sd, p, slope=0.01):
    """reference util.py:528-544  conv-GN-lrelu-conv-GN-(+id)-lrelu (ResBlock3D_stage3_leak, 32->32)."""
    out = _conv3d(x, sd, p + ".conv1", 1)
    out = F.leaky_relu(F.group_norm(out, 32, sd[p + ".gn1.weight"], sd[p + ".gn1.bias"], GN_EPS), slope)
    out = _conv3d(out, sd, p + ".conv2", 1)
    out = F.group_norm(out, 32, sd[p + ".gn2.weight"], sd[p + ".gn2.bias"], GN_EPS)
    return F.leaky_relu(out + x, slope)


# ---------------------------------------------------------------------------------------------
# F : appearance feature extractor
# ---------------------------------------------------------------------------------------------
def appearance_feature_extractor(sd, x):
    """reference appearance_feature_extractor.py:38-48.  x [B,3,H,W] -> [B,32,16,H/4,W/4]."""
    out = F.relu(_bn(_conv2d(x, sd, "first.conv", 1), sd, "first.norm"))            # util.py:207-211
    for i in range(2):                                                               # util.py:161-166
        p = f"down_blocks.{i}"
        out = F.avg_pool2d(F.relu(_bn(_conv2d(out, sd, p + ".conv", 1), sd, p + ".norm")), 2)
    out = _conv2d(out, sd, "second")
    bs, c, h, w = out.shape
    f_s = out.view(bs, 32, 16, h, w)
    for i in range(6):
        f_s = resblock3d(f_s, sd, f"resblocks_3d.3dr{i}")
    return f_s


# ---------------------------------------------------------------------------------------------
# W : dense motion + warp
# ---------------------------------------------------------------------------------------------
def make_coordinate_grid(d, h, w, dtype=torch.float32, device=None):
    """reference util.py:41-58: identity grid in [-1,1], last dim ordered (x, y, z)."""
    x = 2 * (torch.arange(w, dtype=dtype, device=device) / (w - 1)) - 1
    y = 2 * (torch.arange(h, dtype=dtype, device=device) / (h - 1)) - 1
    z = 2 * (torch.arange(d, dtype=dtype, device=device) / (d - 1)) - 1
    zz, yy, xx = torch.meshgrid(z, y, x, indexing="ij")
    return torch.stack([xx, yy, zz], dim=-1)          # [d,h,w,3]


def kp2gaussian(kp, grid, var=0.01):
    """reference util.py:17-38.  kp [B,K,3], grid [d,h,w,3] -> [B,K,d,h,w]."""
    diff = grid[None, None] - kp[:, :, None, None, None, :]
    return torch.exp(-0.5 * (diff ** 2).sum(-1) / var)


def hourglass(sd, p, x):
    """reference util.py:214-279 (Encoder/Decoder/Hourglass)."""
    outs = [x]
    for i in range(5):                                                  # DownBlock3d util.py:185-190
        q = f"{p}.encoder.down_blocks.{i}"
        o = F.relu(_bn(_conv3d(outs[-1], sd, q + ".conv", 1), sd, q + ".norm"))
        outs.append(F.avg_pool3d(o, (1, 2, 2)))
    out = outs.pop()
    for i in range(5):                                                  # UpBlock3d util.py:142-147
        q = f"{p}.decoder.up_blocks.{i}"
        out = F.interpolate(out, scale_factor=(1, 2, 2))
        out = F.relu(_bn(_conv3d(out, sd, q + ".conv", 1), sd, q + ".norm"))
        out = torch.cat([out, outs.pop()], dim=1)                       # util.py:259-260
    out = F.relu(_bn(_conv3d(out, sd, p + ".decoder.conv", 1), sd, p + ".decoder.norm"))
    return out


def dense_motion(sd, feature, kp_driving, kp_source, p="dense_motion_network"):
    """reference dense_motion.py:67-104.  Returns dict(mask, deformation, occlusion_map, hourglass_in, prediction)."""
    bs, _, d, h, w = feature.shape
    K = kp_source.shape[1]
    f = F.relu(_bn(_conv3d(feature, sd, p + ".compress"), sd, p + ".norm"))        # :70-72  [B,4,d,h,w]
    grid = make_coordinate_grid(d, h, w, kp_source.dtype, feature.device)          # :31
    # sparse motions :29-43
    d2s = grid[None, None] - kp_driving.view(bs, K, 1, 1, 1, 3) + kp_source.view(bs, K, 1, 1, 1, 3)
    motions = torch.cat([grid[None, None].expand(bs, 1, d, h, w, 3), d2s], dim=1)  # [B,K+1,d,h,w,3]
    # deformed features :45-53
    f_rep = f[:, None].expand(bs, K + 1, 4, d, h, w).reshape(bs * (K + 1), 4, d, h, w)
    deformed = F.grid_sample(f_rep, motions.reshape(bs * (K + 1), d, h, w, 3), align_corners=False)
    deformed = deformed.view(bs, K + 1, 4, d, h, w)
    # heatmaps :55-65
    heat = kp2gaussian(kp_driving, grid) - kp2gaussian(kp_source, grid)
    heat = torch.cat([torch.zeros(bs, 1, d, h, w, dtype=heat.dtype, device=heat.device), heat], dim=1)[:, :, None]
    inp = torch.cat([heat, deformed], dim=2).view(bs, (K + 1) * 5, d, h, w)        # :83-84
    pred = hourglass(sd, p + ".hourglass", inp)                                    # :86
    logits = _conv3d(pred, sd, p + ".mask", 3)                                     # :88
    mask = F.softmax(logits, dim=1)                                                # :89
    deformation = (motions.permute(0, 1, 5, 2, 3, 4) * mask[:, :, None]).sum(dim=1)  # :91-93
    deformation = deformation.permute(0, 2, 3, 4, 1)                               # [B,d,h,w,3]
    occ = torch.sigmoid(_conv2d(pred.reshape(bs, -1, h, w), sd, p + ".occlusion", 3))  # :98-102
    return {"mask": mask, "mask_logits": logits, "deformation": deformation, "occlusion_map": occ,
            "hourglass_in": inp, "prediction": pred, "compressed": f}


def warp(sd, feature_3d, kp_source, kp_driving):
    """reference warping_network.py:49-62  -> (out [B,32,16,h,w], occlusion_map [B,1,h,w])."""
    dm = dense_motion(sd, feature_3d, kp_driving=kp_driving, kp_source=kp_source)
    out = F.grid_sample(feature_3d, dm["deformation"], align_corners=False)        # :46-47
    return out, dm["occlusion_map"]


def warp_out(sd, out, occlusion_map=None):
    """reference warping_network.py:64-71."""
    bs, c, d, h, w = out.shape
    out = out.reshape(bs, c * d, h, w)
    out = F.leaky_relu(_bn(_conv2d(out, sd, "third.conv", 1), sd, "third.norm"), 0.01)  # SameBlock2d lrelu
    out = _conv2d(out, sd, "fourth")
    if occlusion_map is not None:
        out = out * occlusion_map
    return out


def warping_forward(sd, feature_3d, kp_driving, kp_source):
    """reference warping_network.py:83-111  -> dict(occlusion_map, deformation, out)."""
    dm = dense_motion(sd, feature_3d, kp_driving=kp_driving, kp_source=kp_source)
    out = F.grid_sample(feature_3d, dm["deformation"], align_corners=False)
    out = warp_out(sd, out, dm["occlusion_map"])
    return {"occlusion_map": dm["occlusion_map"], "deformation": dm["deformation"], "out": out}


# ---------------------------------------------------------------------------------------------
# swap : transfer_model2 (canonical-space identity modulation)
# ---------------------------------------------------------------------------------------------
def adaptive_conv(sd, p, x, latent):
    """reference adaptive_modulate.py:128-193 (AdaptiveSharedWeightConv2d.forward) -> (out, mask)."""
    w = sd[p + ".weight"]                                                  # [O,I,3,3]
    out_std = F.conv2d(x, w, None, padding=1)                              # :139-145
    s = F.linear(latent, sd[p + ".style_fc.0.weight"], sd[p + ".style_fc.0.bias"])
    s = F.linear(F.leaky_relu(s, 0.2), sd[p + ".style_fc.2.weight"], sd[p + ".style_fc.2.bias"])  # :148
    outs = []
    for n in range(x.shape[0]):                                            # groups=N conv, :150-167
        wm = w * s[n].view(1, -1, 1, 1)
        demod = torch.rsqrt((wm ** 2).sum(dim=(1, 2, 3), keepdim=True) + DEMOD_EPS)
        outs.append(F.conv2d(x[n:n + 1], wm * demod, None, padding=1))
    out_mod = torch.cat(outs, 0) + sd[p + ".bias_param"].view(1, -1, 1, 1)  # :169-170
    mask = torch.sigmoid(_conv2d(x, sd, p + ".mask_conv.0", 1))            # :176
    return mask * out_mod + (1 - mask) * out_std, mask                     # :186


def swap_module(sd, x, dlatents, return_mask=False):
    """reference adaptive_modulate.py:522-554 (transfer_model2.forward)."""
    bs, c, d, h, w = x.shape
    x = x.reshape(bs, c * d, h, w)
    masks = []
    for i in range(7):                                                     # ResnetBlock_Adaptive2D :337-349
        p = f"BottleNeck_2d.{i}"
        y, m1 = adaptive_conv(sd, p + ".conv1", x, dlatents)
        y = F.relu(y)
        y, m2 = adaptive_conv(sd, p + ".conv2", y, dlatents)
        x = x + y
        masks.append((m1 + m2) / 2)
    x = x.view(bs, c, d, h, w)
    for i in range(6):
        x = resblock3d(x, sd, f"resblocks_3d.3dr{i}")
    return (x, masks) if return_mask else x


# ---------------------------------------------------------------------------------------------
# refine : G3d
# ---------------------------------------------------------------------------------------------
def refine_module(sd, x):
    """reference adaptive_modulate.py:721-733."""
    for i in range(3):
        x = gn_resblock3d(x, sd, f"resblocks1.{i}")
    bs, c, d, h, w = x.shape
    x = x.reshape(bs, c * d, h, w)
    for i in range(3):
        x = resblock2d(x, sd, f"resblocks2.{i}")
    x = x.view(bs, c, d, h, w)
    for i in range(3):
        x = gn_resblock3d(x, sd, f"resblocks3.{i}")
    return x


# ---------------------------------------------------------------------------------------------
# G : SPADE decoder
# ---------------------------------------------------------------------------------------------
def sn_weight(sd, p):
    """eval-mode torch.nn.utils.spectral_norm: W / (u . (W_mat v)), no power iteration
    (reference util.py:318-322; SURVEY.md section 5 'Checkpoint')."""
    w = sd[p + ".weight_orig"]
    sigma = torch.dot(sd[p + ".weight_u"], torch.mv(w.reshape(w.shape[0], -1), sd[p + ".weight_v"]))
    return w / sigma


def spade(sd, p, x, seg):
    """reference util.py:295-302."""
    normalized = F.instance_norm(x, eps=IN_EPS)
    seg = F.interpolate(seg, size=x.shape[2:], mode="nearest")
    actv = F.relu(_conv2d(seg, sd, p + ".mlp_shared.0", 1))
    gamma = _conv2d(actv, sd, p + ".mlp_gamma", 1)
    beta = _conv2d(actv, sd, p + ".mlp_beta", 1)
    return normalized * (1 + gamma) + beta


def spade_resblock(sd, p, x, seg, learned_shortcut):
    """reference util.py:329-344."""
    if learned_shortcut:
        x_s = F.conv2d(spade(sd, p + ".norm_s", x, seg), sn_weight(sd, p + ".conv_s"), None)
    else:
        x_s = x
    dx = F.conv2d(F.leaky_relu(spade(sd, p + ".norm_0", x, seg), 0.2), sn_weight(sd, p + ".conv_0"),
                  sd[p + ".conv_0.bias"], padding=1)
    dx = F.conv2d(F.leaky_relu(spade(sd, p + ".norm_1", dx, seg), 0.2), sn_weight(sd, p + ".conv_1"),
                  sd[p + ".conv_1.bias"], padding=1)
    return x_s + dx


def spade_decoder(sd, feature, return_logits=False):
    """reference spade_generator.py:41-59.  feature [B,256,h,w] -> [B,3,8h,8w] in [0,1]."""
    seg = feature
    x = _conv2d(feature, sd, "fc", 1)
    for i in range(6):
        x = spade_resblock(sd, f"G_middle_{i}", x, seg, False)
    x = F.interpolate(x, scale_factor=2)            # nn.Upsample(scale_factor=2) nearest
    x = spade_resblock(sd, "up_0", x, seg, True)
    x = F.interpolate(x, scale_factor=2)
    x = spade_resblock(sd, "up_1", x, seg, True)
    logits = F.pixel_shuffle(_conv2d(F.leaky_relu(x, 0.2), sd, "conv_img.0", 1), 2)
    out = torch.sigmoid(logits)
    return (out, logits) if return_logits else out


# ---------------------------------------------------------------------------------------------
# wrapper glue and the per-frame loop body
# ---------------------------------------------------------------------------------------------
def prepare_videos(imgs_u8):
    """reference can_swap_e2e.py:147-163.  [T,H,W,3] u8 -> [T,1,3,H,W] fp32 in [0,1]."""
    y = imgs_u8.to(torch.float32) / 255.0
    return y.clamp(0, 1).permute(0, 3, 1, 2)[:, None].contiguous()


def parse_output(out):
    """reference can_swap_e2e.py:314-322.  [B,3,H,W] fp32 -> [B,H,W,3] u8 (truncating)."""
    o = out.permute(0, 2, 3, 1).clamp(0, 1)
    return (o * 255).clamp(0, 255).to(torch.uint8)


def frame(weights, I_s, x_t, x_can, source_id, debug_decodes=False):
    """The LOOP C body, reference can_swap_pipeline_e2e.py:242-263, returning every stage.

    weights: dict net-name -> flat state_dict (keys of canonswap_b200.spec.NETS)
    I_s [B,3,H,W] in [0,1];  x_t, x_can [B,21,3];  source_id [1 or B,512].
    """
    sdF, sdW, sdG = (weights["appearance_feature_extractor"], weights["warping_module"],
                     weights["spade_generator"])
    sdT, sdR = weights["transfer"], weights["refine"]
    B = I_s.shape[0]
    if source_id.shape[0] != B:
        source_id = source_id.expand(B, -1)
    r = {}
    with torch.no_grad():
        r["f_s"] = appearance_feature_extractor(sdF, I_s)                        # :242
        r["f_can"], r["occ_can"] = warp(sdW, r["f_s"], x_t, x_can)               # :244
        if debug_decodes:
            r["rec_can"] = spade_decoder(sdG, warp_out(sdW, r["f_can"], r["occ_can"]))      # :248
        r["f_swap"] = swap_module(sdT, r["f_can"], source_id)                    # :253
        if debug_decodes:
            r["swap_can"] = spade_decoder(sdG, warp_out(sdW, r["f_swap"], r["occ_can"]))    # :257
        r["f_refine"] = refine_module(sdR, r["f_swap"])                          # :262
        wf = warping_forward(sdW, r["f_refine"], kp_driving=x_t, kp_source=x_can)  # :263 -> can_swap_e2e.py:298
        r["occ"], r["deformation"], r["warp_out"] = wf["occlusion_map"], wf["deformation"], wf["out"]
        r["out"], r["logits"] = spade_decoder(sdG, wf["out"], return_logits=True)
    return r


def frame_v2i(weights, I, kp_source, kp_driving):
    """The per-frame body of the video-to-image pipeline, reference can_swap_pipeline_v2i.py:308-309:
    out = warp_decode(extract_feature_3d(I), kp_source, kp_driving)  (can_swap_e2e.py:286-308)."""
    sdF, sdW, sdG = (weights["appearance_feature_extractor"], weights["warping_module"], weights["spade_generator"])
    with torch.no_grad():
        f = appearance_feature_extractor(sdF, I)
        wf = warping_forward(sdW, f, kp_driving=kp_driving, kp_source=kp_source)
        return spade_decoder(sdG, wf["out"])


V2I_ANIMATE_HW = (256, 256)      # the size the reference resizes the swapped canonical image to (can_swap_pipeline_v2i.py:294)


def v2i_source_state(weights, I_s, driving_id):
    """The once-per-source part of the video-to-image pipeline: execute_face_canonical (reference
    src/can_swap_pipeline_v2i.py:86-98) and the `i == 0` block of its frame loop (:285-304).
    I_s [1,3,H,W] in [0,1], driving_id [1,512].  Returns the frame-invariant state the loop body reads."""
    sdW, sdG, sdM = weights["warping_module"], weights["spade_generator"], weights["motion_extractor"]
    with torch.no_grad():
        info = motion_extractor(sdM, I_s)                                          # :87 get_kp_info(I_s)
        kp_s = info["kp"].reshape(1, -1, 3)                                        # :88
        f_s = appearance_feature_extractor(weights["appearance_feature_extractor"], I_s)   # :89
        x_s = transform_keypoint(info)                                             # :90
        x_d = info["scale"] * kp_s                                                 # :94 scale_new * x_c_s
        f_s_can, occ = warp(sdW, f_s, x_s, x_d)                                    # :97
        f_can_swap = swap_module(weights["transfer"], f_s_can, driving_id)         # :286
        swap_can = spade_decoder(sdG, warp_out(sdW, f_can_swap, occ))              # :289 conv_decode
        swap_can_lr = F.interpolate(swap_can, size=V2I_ANIMATE_HW, mode="bilinear", align_corners=False)   # :294
        info2 = motion_extractor(sdM, swap_can_lr)                                 # :297
        x_swap = transform_keypoint(info2)                                         # :298
        deg = [headpose_pred_to_degree(info[k]) for k in ("pitch", "yaw", "roll")]
        R_swap = get_rotation_matrix(deg[0], deg[1], deg[2])                       # :301 (the SOURCE's rotation)
        t_swap = info["t"].clone()
        t_swap[..., 2] = 0                                                         # :303
        f2 = appearance_feature_extractor(weights["appearance_feature_extractor"], swap_can_lr)   # :308 (frame-invariant)
    return {"swap_can": swap_can, "swap_can_lr": swap_can_lr, "feature": f2, "x_swap": x_swap, "kp_swap": info2["kp"].reshape(1, -1, 3),
            "R_swap": R_swap, "t_swap": t_swap, "scale_swap": info["scale"], "f_s_can": f_s_can, "occ": occ}


def v2i_frames(weights, st, delta_t):
    """The per-frame body of the v2i loop (reference can_swap_pipeline_v2i.py:305-312) for a batch of driving expressions
    delta_t [B,21,3]: x_t_2 = scale_swap * (kp_swap @ R_swap + delta_t) + t_swap; out = warp_decode(F(swap_can_256), x_swap, x_t_2)."""
    B = delta_t.shape[0]
    with torch.no_grad():
        x_t_2 = st["scale_swap"] * (st["kp_swap"] @ st["R_swap"] + delta_t) + st["t_swap"]        # :305
        wf = warping_forward(weights["warping_module"], st["feature"].expand(B, -1, -1, -1, -1),
                             kp_driving=x_t_2, kp_source=st["x_swap"].expand(B, -1, -1))        # :309 -> can_swap_e2e.py:298
        return spade_decoder(weights["spade_generator"], wf["out"]), x_t_2


# ------------------------------------------------------------------------------------------
# motion extractor M + keypoint transform
# ------------------------------------------------------------------------------------------
LN_EPS = 1e-6     # reference convnextv2.py:27,76,79,93 / util.py:378
MOTION_DEPTHS = (3, 3, 9, 3)


def _ln_cf(x, w, b):
    """LayerNorm over the channel axis of an NCHW tensor, reference util.py:391-396 (channels_first branch)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + LN_EPS)
    return w[:, None, None] * x + b[:, None, None]


def convnext_block(sd, p, x):
    """reference convnextv2.py:34-47: dwconv7 -> LN -> Linear -> GELU -> GRN (util.py:365-368) -> Linear -> + input."""
    dim = x.shape[1]
    y = F.conv2d(x, sd[p + ".dwconv.weight"], sd[p + ".dwconv.bias"], padding=3, groups=dim)
    y = y.permute(0, 2, 3, 1)
    y = F.layer_norm(y, (dim,), sd[p + ".norm.weight"], sd[p + ".norm.bias"], LN_EPS)
    y = F.linear(y, sd[p + ".pwconv1.weight"], sd[p + ".pwconv1.bias"])
    y = F.gelu(y)
    gx = torch.norm(y, p=2, dim=(1, 2), keepdim=True)
    nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
    y = sd[p + ".grn.gamma"] * (y * nx) + sd[p + ".grn.beta"] + y
    y = F.linear(y, sd[p + ".pwconv2.weight"], sd[p + ".pwconv2.bias"])
    return x + y.permute(0, 3, 1, 2)


def motion_extractor(sd, x):
    """MotionExtractor.forward -> ConvNeXtV2.forward, reference motion_extractor.py:33-35, convnextv2.py:110-144.
    x [B,3,H,W] in [0,1] -> dict(pitch, yaw, roll [B,66], t [B,3], exp [B,63], scale [B,1], kp [B,63])."""
    p = "detector"
    x = F.conv2d(x, sd[f"{p}.downsample_layers.0.0.weight"], sd[f"{p}.downsample_layers.0.0.bias"], stride=4)
    x = _ln_cf(x, sd[f"{p}.downsample_layers.0.1.weight"], sd[f"{p}.downsample_layers.0.1.bias"])
    for i in range(4):
        if i > 0:
            x = _ln_cf(x, sd[f"{p}.downsample_layers.{i}.0.weight"], sd[f"{p}.downsample_layers.{i}.0.bias"])
            x = F.conv2d(x, sd[f"{p}.downsample_layers.{i}.1.weight"], sd[f"{p}.downsample_layers.{i}.1.bias"], stride=2)
        for j in range(MOTION_DEPTHS[i]):
            x = convnext_block(sd, f"{p}.stages.{i}.{j}", x)
    f = F.layer_norm(x.mean([-2, -1]), (x.shape[1],), sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], LN_EPS)
    return {k: F.linear(f, sd[f"{p}.fc_{k}.weight"], sd[f"{p}.fc_{k}.bias"])
            for k in ("pitch", "yaw", "roll", "t", "exp", "scale", "kp")}


def headpose_pred_to_degree(pred):
    """reference src/utils/camera.py:14-29."""
    if pred.ndim > 1 and pred.shape[1] == 66:
        idx = torch.arange(66, dtype=torch.float32)
        return torch.sum(F.softmax(pred, dim=1) * idx, dim=1) * 3 - 97.5
    return pred


def get_rotation_matrix(pitch_, yaw_, roll_):
    """reference src/utils/camera.py:32-73 (degrees in, R = (Rz Ry Rx)^T out)."""
    import math
    x, y, z = (a.reshape(-1, 1) / 180 * math.pi for a in (pitch_, yaw_, roll_))
    bs = x.shape[0]
    one, zero = torch.ones(bs, 1), torch.zeros(bs, 1)
    rx = torch.cat([one, zero, zero, zero, torch.cos(x), -torch.sin(x), zero, torch.sin(x), torch.cos(x)], 1).reshape(bs, 3, 3)
    ry = torch.cat([torch.cos(y), zero, torch.sin(y), zero, one, zero, -torch.sin(y), zero, torch.cos(y)], 1).reshape(bs, 3, 3)
    rz = torch.cat([torch.cos(z), -torch.sin(z), zero, torch.sin(z), torch.cos(z), zero, zero, zero, one], 1).reshape(bs, 3, 3)
    return (rz @ ry @ rx).permute(0, 2, 1)


def transform_keypoint(kp_info):
    """can_swapper.transform_keypoint, reference src/can_swap_e2e.py:226-254: s * (kp @ R + exp) + t_xy."""
    kp = kp_info["kp"]
    bs = kp.shape[0]
    pitch = headpose_pred_to_degree(kp_info["pitch"])
    yaw = headpose_pred_to_degree(kp_info["yaw"])
    roll = headpose_pred_to_degree(kp_info["roll"])
    rot = get_rotation_matrix(pitch, yaw, roll)
    out = kp.reshape(bs, -1, 3) @ rot + kp_info["exp"].reshape(bs, -1, 3)
    out = out * kp_info["scale"][..., None]
    out[:, :, 0:2] = out[:, :, 0:2] + kp_info["t"][:, None, 0:2]
    return out


def motion_keypoints(sd, I):
    """What LOOP C consumes from M per frame (make_motion_template, reference src/can_swap_pipeline_e2e.py:112-125, and
    :231-243): x_t = transform_keypoint(get_kp_info(I)) and x_can = scale * kp."""
    info = motion_extractor(sd, I)
    x_t = transform_keypoint(info)
    bs = I.shape[0]
    x_can = info["scale"][..., None] * info["kp"].reshape(bs, -1, 3)
    deg = torch.stack([headpose_pred_to_degree(info[k]) for k in ("pitch", "yaw", "roll")], 1)
    return {"info": info, "x_t": x_t, "x_can": x_can, "R": get_rotation_matrix(deg[:, 0], deg[:, 1], deg[:, 2]), "deg": deg}
