"""One profiled step of the hot path for ncu (GPU box):
    ncu --profile-from-start off ... python tools/profile_step.py [batch]
warm-up 2 steps, then cudaProfilerStart .. one cs_frame step (B frames, net 256 -> 512 px) .. Stop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from canonswap_b200 import synth
from canonswap_b200.modules import can_swapper

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
NET = 256
W = synth.synth_weights()
clip = synth.synth_inputs(B, NET, u8=True)
sw = can_swapper(weights=W, device_id=0, max_batch=B)
sw.set_source_identity(clip["source_id"].cuda())
eng = sw.engine((NET, NET), B)
fr, xt, xc = clip["frames"].cuda(), clip["x_t"].cuda(), clip["x_can"].cuda()
out = torch.empty(B, 2 * NET, 2 * NET, 3, dtype=torch.uint8, device="cuda")
for _ in range(2):
    eng.frame(fr, xt, xc, out_u8=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.frame(fr, xt, xc, out_u8=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step, batch", B)
