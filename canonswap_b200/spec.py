"""Parameter spec of the five generator networks on the CanonSwap per-frame hot path.

This is the repo's own statement of WHICH tensors (name, shape) the reference checkpoint
`combined_weights.pth` carries for the hot-path networks, keyed exactly like the reference
`state_dict()`s so that `load_cpk` (reference `src/can_swap_e2e.py:87-100`) works unchanged:

  'appearance_feature_extractor'  reference `src/modules/appearance_feature_extractor.py:14-36`
  'warping_module'                reference `src/modules/warping_network.py:14-44`, `dense_motion.py:14-27`
  'spade_generator'               reference `src/modules/spade_generator.py:13-39`
  'transfer'                      reference `src/modules/adaptive_modulate.py:485-521` (transfer_model2)
  'refine'                        reference `src/modules/adaptive_modulate.py:700-720` (G3d)

Hyper-parameters are the ones in reference `src/config/models.yaml:1-30` (with
`spade_generator_params.upscale` forced to 2, `src/can_swap_e2e.py:62`).
`tests/test_spec_vs_reference.py` checks every key/shape against the reference modules when
`/root/reference` is present.
"""
from __future__ import annotations

from collections import OrderedDict

NUM_KP = 21
RESHAPE_C = 32          # reshape_channel
RESHAPE_D = 16          # reshape_depth
LATENT = 512

NETS = ("appearance_feature_extractor", "warping_module", "spade_generator", "transfer", "refine")


def _conv(sd, p, cout, cin, *k, bias=True):
    sd[p + ".weight"] = (cout, cin, *k)
    if bias:
        sd[p + ".bias"] = (cout,)


def _bn(sd, p, c):
    sd[p + ".weight"] = (c,)
    sd[p + ".bias"] = (c,)
    sd[p + ".running_mean"] = (c,)
    sd[p + ".running_var"] = (c,)
    sd[p + ".num_batches_tracked"] = ()


def _resblock3d(sd, p, c=RESHAPE_C):
    # reference util.py:85-92 registers conv1, conv2, norm1, norm2 in this order
    _conv(sd, p + ".conv1", c, c, 3, 3, 3)
    _conv(sd, p + ".conv2", c, c, 3, 3, 3)
    _bn(sd, p + ".norm1", c)
    _bn(sd, p + ".norm2", c)


def appearance_feature_extractor_spec():
    sd = OrderedDict()
    _conv(sd, "first.conv", 64, 3, 3, 3)
    _bn(sd, "first.norm", 64)
    _conv(sd, "down_blocks.0.conv", 128, 64, 3, 3)
    _bn(sd, "down_blocks.0.norm", 128)
    _conv(sd, "down_blocks.1.conv", 256, 128, 3, 3)
    _bn(sd, "down_blocks.1.norm", 256)
    _conv(sd, "second", 512, 256, 1, 1)
    for i in range(6):
        _resblock3d(sd, f"resblocks_3d.3dr{i}")
    return sd


# hourglass: block_expansion 32, max_features 1024, 5 blocks, in_features (21+1)*(4+1)=110
HG_IN = (NUM_KP + 1) * 5
HG_ENC = [(HG_IN, 64), (64, 128), (128, 256), (256, 512), (512, 1024)]
HG_DEC = [(1024, 512), (1024, 256), (512, 128), (256, 64), (128, 32)]
HG_OUT = 32 + HG_IN   # 142


def warping_module_spec():
    sd = OrderedDict()
    p = "dense_motion_network"
    for i, (ci, co) in enumerate(HG_ENC):
        _conv(sd, f"{p}.hourglass.encoder.down_blocks.{i}.conv", co, ci, 3, 3, 3)
        _bn(sd, f"{p}.hourglass.encoder.down_blocks.{i}.norm", co)
    for i, (ci, co) in enumerate(HG_DEC):
        _conv(sd, f"{p}.hourglass.decoder.up_blocks.{i}.conv", co, ci, 3, 3, 3)
        _bn(sd, f"{p}.hourglass.decoder.up_blocks.{i}.norm", co)
    _conv(sd, f"{p}.hourglass.decoder.conv", HG_OUT, HG_OUT, 3, 3, 3)
    _bn(sd, f"{p}.hourglass.decoder.norm", HG_OUT)
    _conv(sd, f"{p}.mask", NUM_KP + 1, HG_OUT, 7, 7, 7)
    _conv(sd, f"{p}.compress", 4, RESHAPE_C, 1, 1, 1)
    _bn(sd, f"{p}.norm", 4)
    _conv(sd, f"{p}.occlusion", 1, HG_OUT * RESHAPE_D, 7, 7)
    _conv(sd, "third.conv", 256, 512, 3, 3)
    _bn(sd, "third.norm", 256)
    _conv(sd, "fourth", 256, 256, 1, 1)
    return sd


def _sn_conv(sd, p, cout, cin, k, bias=True):
    # torch.nn.utils.spectral_norm (hook flavour): bias, weight_orig, weight_u, weight_v
    if bias:
        sd[p + ".bias"] = (cout,)
    sd[p + ".weight_orig"] = (cout, cin, k, k)
    sd[p + ".weight_u"] = (cout,)
    sd[p + ".weight_v"] = (cin * k * k,)


def _spade(sd, p, norm_nc, label_nc=256, nhidden=128):
    _conv(sd, p + ".mlp_shared.0", nhidden, label_nc, 3, 3)
    _conv(sd, p + ".mlp_gamma", norm_nc, nhidden, 3, 3)
    _conv(sd, p + ".mlp_beta", norm_nc, nhidden, 3, 3)


def _spade_resblock(sd, p, fin, fout):
    fmid = min(fin, fout)
    _sn_conv(sd, p + ".conv_0", fmid, fin, 3)
    _sn_conv(sd, p + ".conv_1", fout, fmid, 3)
    if fin != fout:
        _sn_conv(sd, p + ".conv_s", fout, fin, 1, bias=False)
    _spade(sd, p + ".norm_0", fin)
    _spade(sd, p + ".norm_1", fmid)
    if fin != fout:
        _spade(sd, p + ".norm_s", fin)


SPADE_BLOCKS = [(f"G_middle_{i}", 512, 512) for i in range(6)] + [("up_0", 512, 256), ("up_1", 256, 64)]


def spade_generator_spec():
    sd = OrderedDict()
    _conv(sd, "fc", 512, 256, 3, 3)
    for name, fin, fout in SPADE_BLOCKS:
        _spade_resblock(sd, name, fin, fout)
    _conv(sd, "conv_img.0", 12, 64, 3, 3)
    return sd


def transfer_spec():
    sd = OrderedDict()
    for i in range(7):
        for c in ("conv1", "conv2"):
            p = f"BottleNeck_2d.{i}.{c}"
            sd[p + ".weight"] = (512, 512, 3, 3)
            sd[p + ".bias_param"] = (512,)
            sd[p + ".style_fc.0.weight"] = (512, LATENT)
            sd[p + ".style_fc.0.bias"] = (512,)
            sd[p + ".style_fc.2.weight"] = (512, 512)
            sd[p + ".style_fc.2.bias"] = (512,)
            _conv(sd, p + ".mask_conv.0", 1, 512, 3, 3)
    for i in range(6):
        _resblock3d(sd, f"resblocks_3d.3dr{i}")
    return sd


def _gn_resblock(sd, p, c=RESHAPE_C):
    _conv(sd, p + ".conv1", c, c, 3, 3, 3)
    sd[p + ".gn1.weight"] = (c,)
    sd[p + ".gn1.bias"] = (c,)
    _conv(sd, p + ".conv2", c, c, 3, 3, 3)
    sd[p + ".gn2.weight"] = (c,)
    sd[p + ".gn2.bias"] = (c,)


def refine_spec():
    sd = OrderedDict()
    for i in range(3):
        _gn_resblock(sd, f"resblocks1.{i}")
    for i in range(3):
        p = f"resblocks2.{i}"
        _conv(sd, p + ".conv1", 512, 512, 3, 3)
        _conv(sd, p + ".conv2", 512, 512, 3, 3)
        _bn(sd, p + ".norm1", 512)
        _bn(sd, p + ".norm2", 512)
    for i in range(3):
        _gn_resblock(sd, f"resblocks3.{i}")
    return sd


# motion extractor M: ConvNeXtV2-tiny, reference src/modules/convnextv2.py:48-144 wrapped as
# MotionExtractor.detector (src/modules/motion_extractor.py:18-24); checkpoint key 'motion_extractor' (can_swap_e2e.py:94)
MOTION_NET = "motion_extractor"
MOTION_DEPTHS = (3, 3, 9, 3)
MOTION_DIMS = (96, 192, 384, 768)
NUM_BINS = 66
# heads in registration order (convnextv2.py:95-103) -> output width
MOTION_HEADS = (("fc_kp", 3 * NUM_KP), ("fc_scale", 1), ("fc_pitch", NUM_BINS), ("fc_yaw", NUM_BINS), ("fc_roll", NUM_BINS),
                ("fc_t", 3), ("fc_exp", 3 * NUM_KP))


def motion_extractor_spec():
    sd = OrderedDict()
    p = "detector"
    d = MOTION_DIMS
    _conv(sd, f"{p}.downsample_layers.0.0", d[0], 3, 4, 4)
    sd[f"{p}.downsample_layers.0.1.weight"] = (d[0],)
    sd[f"{p}.downsample_layers.0.1.bias"] = (d[0],)
    for i in range(3):
        sd[f"{p}.downsample_layers.{i + 1}.0.weight"] = (d[i],)
        sd[f"{p}.downsample_layers.{i + 1}.0.bias"] = (d[i],)
        _conv(sd, f"{p}.downsample_layers.{i + 1}.1", d[i + 1], d[i], 2, 2)
    for i in range(4):
        for j in range(MOTION_DEPTHS[i]):
            q = f"{p}.stages.{i}.{j}"
            _conv(sd, q + ".dwconv", d[i], 1, 7, 7)
            sd[q + ".norm.weight"] = (d[i],)
            sd[q + ".norm.bias"] = (d[i],)
            sd[q + ".pwconv1.weight"] = (4 * d[i], d[i])
            sd[q + ".pwconv1.bias"] = (4 * d[i],)
            sd[q + ".grn.gamma"] = (1, 1, 1, 4 * d[i])
            sd[q + ".grn.beta"] = (1, 1, 1, 4 * d[i])
            sd[q + ".pwconv2.weight"] = (d[i], 4 * d[i])
            sd[q + ".pwconv2.bias"] = (d[i],)
    sd[f"{p}.norm.weight"] = (d[3],)
    sd[f"{p}.norm.bias"] = (d[3],)
    for name, n in MOTION_HEADS:
        sd[f"{p}.{name}.weight"] = (n, d[3])
        sd[f"{p}.{name}.bias"] = (n,)
    return sd


def net_spec(net: str):
    return {
        "motion_extractor": motion_extractor_spec,
        "appearance_feature_extractor": appearance_feature_extractor_spec,
        "warping_module": warping_module_spec,
        "spade_generator": spade_generator_spec,
        "transfer": transfer_spec,
        "refine": refine_spec,
    }[net]()


def all_specs():
    return OrderedDict((n, net_spec(n)) for n in NETS)
