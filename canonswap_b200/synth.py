"""Seeded synthetic weights and inputs for the generator hot path.

The reference's pretrained `combined_weights.pth` is a download (reference readme.md:25) and
is not available offline, and `nn.Module` default init is a degenerate fixture (uniform
softmax mask, occlusion == 0.5, saturated image).  This module emits
*calibrated* state_dicts in the exact `combined_weights.pth` layout
(`{'appearance_feature_extractor','warping_module','spade_generator','transfer','refine'}`,
reference can_swap_e2e.py:93-98) with non-trivial BN statistics, a non-uniform motion mask,
a spread occlusion map and un-saturated image logits, so that per-stage parity at 1e-3 is a
meaningful check.  Everything is derived from fixed gains and a torch CPU generator, so the
same seed gives the same bytes on every box of this image.

Synthetic inputs follow SURVEY.md section 8d.
"""
from __future__ import annotations

import math
import re
from collections import OrderedDict

import torch

from . import spec

WEIGHT_SEED = 4321
INPUT_SEED = 1234

# (regex on the key, gain) -- first match wins; conv weights are N(0, gain / sqrt(fan_in))
_GAINS = [
    (r"resblocks_3d\.3dr\d\.conv2\.weight$", 0.45),
    (r"resblocks[13]\.\d\.conv2\.weight$", 1.0),
    (r"resblocks2\.\d\.conv2\.weight$", 0.45),
    (r"BottleNeck_2d\.\d\.conv2\.weight$", 0.45),
    (r"BottleNeck_2d\.\d\.conv1\.weight$", 1.0),
    (r"mask_conv\.0\.weight$", 1.5),
    (r"dense_motion_network\.mask\.weight$", 2.6),
    (r"dense_motion_network\.occlusion\.weight$", 1.6),
    (r"dense_motion_network\.compress\.weight$", 1.4),
    (r"mlp_gamma\.weight$", 0.5),
    (r"mlp_beta\.weight$", 0.5),
    (r"conv_img\.0\.weight$", 0.7),
    (r"style_fc\.\d\.weight$", 1.0),
    (r"second\.weight$", 1.0),
    (r"fourth\.weight$", 1.2),
    (r"\.weight$", 1.35),
]
_SN_GAIN = 2.2      # effective spectral norm of the SPADE convs (sigma folded through weight_u)


def _gain(key):
    for pat, g in _GAINS:
        if re.search(pat, key):
            return g
    return 1.0


def _normal(shape, std, g):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std


def _uniform(shape, lo, hi, g):
    return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo


def synth_state_dict(net: str, seed: int = WEIGHT_SEED) -> "OrderedDict[str, torch.Tensor]":
    """One network's state_dict, keys/shapes/dtypes exactly as the reference module's."""
    g = torch.Generator().manual_seed(seed + 17 * spec.NETS.index(net))
    sp = spec.net_spec(net)
    sd = OrderedDict()
    for key, shape in sp.items():
        leaf = key.rsplit(".", 1)[-1]
        parent = key.rsplit(".", 1)[0]
        is_norm = re.search(r"(norm\d?|gn\d)$", parent) is not None and "mlp" not in key
        if leaf == "num_batches_tracked":
            t = torch.tensor(1000, dtype=torch.int64)
        elif leaf == "running_mean":
            t = _normal(shape, 0.2, g)
        elif leaf == "running_var":
            t = _uniform(shape, 0.6, 1.4, g)
        elif is_norm and leaf == "weight":
            t = _uniform(shape, 0.8, 1.2, g)
        elif is_norm and leaf == "bias":
            t = _normal(shape, 0.1, g)
        elif leaf == "weight_orig":
            fan_in = math.prod(shape[1:])
            t = _normal(shape, 1.0 / math.sqrt(fan_in), g)
        elif leaf in ("weight_u", "weight_v"):
            t = None        # filled below from weight_orig
        elif leaf == "weight":
            fan_in = math.prod(shape[1:])
            t = _normal(shape, _gain(key) / math.sqrt(fan_in), g)
        elif leaf == "bias_param":
            t = _normal(shape, 0.1, g)
        elif leaf == "bias":
            if re.search(r"style_fc\.2\.bias$", key):
                t = 1.0 + _normal(shape, 0.25, g)          # modulation scales centred on 1
            elif re.search(r"occlusion\.bias$", key):
                t = 1.2 + _normal(shape, 0.1, g)
            elif re.search(r"mask_conv\.0\.bias$", key):
                t = _normal(shape, 0.1, g)
            else:
                t = _normal(shape, 0.05, g)
        else:
            raise KeyError(f"no synthetic rule for {net}.{key}")
        sd[key] = t
    # spectral-norm triplets: a few deterministic power iterations, then fold the target
    # spectral norm into u (eval mode uses sigma = u . (W v) verbatim, no re-normalisation)
    for key in list(sd.keys()):
        if key.endswith(".weight_orig"):
            p = key[: -len(".weight_orig")]
            w = sd[key].reshape(sd[key].shape[0], -1).double()
            v = torch.randn(w.shape[1], generator=g, dtype=torch.float64)
            v /= v.norm()
            for _ in range(4):
                u = w @ v
                u /= u.norm()
                v = w.t() @ u
                v /= v.norm()
            u = w @ v
            u /= u.norm()
            sd[p + ".weight_u"] = (u / _SN_GAIN).float()
            sd[p + ".weight_v"] = v.float()
    return sd


def synth_motion_state_dict(seed: int = WEIGHT_SEED) -> "OrderedDict[str, torch.Tensor]":
    """Calibrated synthetic state_dict of the motion extractor (`combined_weights['motion_extractor']`, reference
    can_swap_e2e.py:94): LayerNorm / GRN parameters away from their identity init, residual branches damped so the 18
    blocks keep O(1) activations, heads scaled so that the keypoints land in (-1, 1), scale near 1 and the head-pose
    bins give non-trivial angles."""
    g = torch.Generator().manual_seed(seed + 17 * 7)
    sd = OrderedDict()
    for key, shape in spec.motion_extractor_spec().items():
        leaf = key.rsplit(".", 1)[-1]
        if re.search(r"(\.norm|downsample_layers\.0\.1|downsample_layers\.[123]\.0)\.weight$", key):
            t = _uniform(shape, 0.8, 1.2, g)
        elif re.search(r"(\.norm|downsample_layers\.0\.1|downsample_layers\.[123]\.0)\.bias$", key):
            t = _normal(shape, 0.1, g)
        elif leaf == "gamma":
            t = _normal(shape, 0.3, g)
        elif leaf == "beta":
            t = _normal(shape, 0.1, g)
        elif leaf == "weight":
            fan_in = math.prod(shape[1:])
            gain = 1.0
            if "pwconv2" in key:
                gain = 0.5
            elif "dwconv" in key:
                gain = 1.2
            elif "fc_kp" in key:
                gain = 0.25
            elif "fc_exp" in key:
                gain = 0.03
            elif "fc_t" in key:
                gain = 0.1
            elif "fc_scale" in key:
                gain = 0.1
            elif re.search(r"fc_(pitch|yaw|roll)", key):
                gain = 1.5
            t = _normal(shape, gain / math.sqrt(fan_in), g)
        elif leaf == "bias":
            if "fc_scale" in key:
                t = 1.2 + _normal(shape, 0.02, g)
            else:
                t = _normal(shape, 0.05, g)
        else:
            raise KeyError(f"no synthetic rule for motion_extractor.{key}")
        sd[key] = t
    return sd


def synth_weights(seed: int = WEIGHT_SEED, with_motion: bool = False):
    """The full `combined_weights.pth`-layout dict for the five hot-path networks (+ the motion extractor on request)."""
    w = OrderedDict((n, synth_state_dict(n, seed)) for n in spec.NETS)
    if with_motion:
        w[spec.MOTION_NET] = synth_motion_state_dict(seed)
    return w


def synth_inputs(T: int, net_hw: int, seed: int = INPUT_SEED, u8: bool = False):
    """Synthetic clip per SURVEY.md section 8d.

    Returns dict(frames [T,3,h,w] fp32 in [0,1) (or [T,h,w,3] u8 when u8=True),
                 x_can [T,21,3], x_t [T,21,3], source_id [1,512]).
    Frames are smooth random fields plus noise so that the extractor sees structure.
    """
    g = torch.Generator().manual_seed(seed)
    lo = torch.rand(T, 3, net_hw // 16, net_hw // 16, generator=g)
    frames = torch.nn.functional.interpolate(lo, size=(net_hw, net_hw), mode="bilinear", align_corners=False)
    frames = (0.8 * frames + 0.2 * torch.rand(T, 3, net_hw, net_hw, generator=g)).clamp(0, 1)
    x_can = (0.3 * torch.randn(T, spec.NUM_KP, 3, generator=g)).clamp(-0.9, 0.9)
    x_t = x_can + 0.05 * torch.randn(T, spec.NUM_KP, 3, generator=g)
    source_id = torch.nn.functional.normalize(torch.randn(1, spec.LATENT, generator=g), p=2, dim=1)
    frames_u8 = (frames * 255).to(torch.uint8)
    if u8:
        return {"frames": frames_u8.permute(0, 2, 3, 1).contiguous(), "x_can": x_can, "x_t": x_t,
                "source_id": source_id}
    # keep fp32 frames exactly representable as u8/255 so u8 and fp32 ingest agree bit-for-bit
    return {"frames": frames_u8.to(torch.float32) / 255.0, "x_can": x_can, "x_t": x_t,
            "source_id": source_id}
