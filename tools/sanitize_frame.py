"""One small frame batch through every entry point, for compute-sanitizer (GPU box):
    compute-sanitizer --tool memcheck python tools/sanitize_frame.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from canonswap_b200 import synth
from canonswap_b200.modules import can_swapper
W = synth.synth_weights(with_motion=True)
inp = synth.synth_inputs(2, 128)
sw = can_swapper(weights=W, device_id=0, max_batch=2)
sw.set_source_identity(inp["source_id"].cuda())
fr, xt, xc = inp["frames"].cuda(), inp["x_t"].cuda(), inp["x_can"].cuda()
u8, _ = sw.swap_frames(fr, xt, xc)
u8m, _ = sw.swap_frames(fr)                                   # motion extractor inside
eng = sw.engine((128, 128), 2)
f = eng.appearance(fr)
wf = eng.warp_forward(f, kp_driving=xt, kp_source=xc)
img = eng.spade(wf["out"])
mask = eng.parse_mask(torch.randn(2, 19, 64, 64, device="cuda"), (256, 256))
u8v, _ = eng.frame(f[:1].contiguous(), xt, xc, v2i_feature=True)
torch.cuda.synchronize()
print("ok", u8.float().mean().item(), u8m.float().mean().item(), img.mean().item(), mask.mean().item(), u8v.float().mean().item())
