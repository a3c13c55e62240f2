"""Development probe (GPU box): motion-extractor head / keypoint errors vs the CPU oracle and their effect on the image."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from canonswap_b200 import synth, spec, _lib
from canonswap_b200.engine import Engine
from oracle import canonswap_oracle as O

net = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
W = synth.synth_weights(with_motion=True)
inp = synth.synth_inputs(B, net)
mk = O.motion_keypoints(W["motion_extractor"], inp["frames"])
eng = Engine(W, net_hw=(net, net), max_batch=B, device=0)
eng.set_identity(inp["source_id"].cuda())
heads = eng.motion(inp["frames"].cuda())
d = eng.motion_dict(heads)
for k in ("kp", "scale", "pitch", "yaw", "roll", "t", "exp"):
    print(k, f"max|d|={(d[k].cpu() - mk['info'][k]).abs().max().item():.3e} range={mk['info'][k].abs().max().item():.2f}")
kp = eng.keypoints(heads)
print("x_s", (kp["x_s"].cpu() - mk["x_t"]).abs().max().item(), "x_can", (kp["x_can"].cpu() - mk["x_can"]).abs().max().item(),
      "deg", (kp["deg"].cpu() - mk["deg"]).abs().max().item())
ref = O.frame(W, inp["frames"], mk["x_t"], mk["x_can"], inp["source_id"])["out"]
o = torch.empty(B, 3, 2 * net, 2 * net, device="cuda")
eng.frame(inp["frames"].cuda(), mk["x_t"].cuda(), mk["x_can"].cuda(), out_f32=o)
print("image, oracle keypoints :", (o.cpu() - ref).abs().max().item())
eng.frame(inp["frames"].cuda(), out_f32=o, motion=True)
print("image, device keypoints :", (o.cpu() - ref).abs().max().item())
# sensitivity: oracle keypoints perturbed by 1e-5
eng.frame(inp["frames"].cuda(), (mk["x_t"] + 1e-5).cuda(), mk["x_can"].cuda(), out_f32=o)
print("image, x_t + 1e-5       :", (o.cpu() - ref).abs().max().item())
