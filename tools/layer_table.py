"""Per-launch timing table of one cs_frame step (GPU box): python tools/layer_table.py [batch] [reps] > gpurun_out/layers.csv
Every launch is bracketed by CUDA events on the launching stream (cs_profile); the table averages `reps` steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from canonswap_b200 import synth
from canonswap_b200.modules import can_swapper

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
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
acc = None
for _ in range(REPS):
    eng.profile(True)
    eng.frame(fr, xt, xc, out_u8=out)
    rows = eng.profile_dump()
    eng.profile(False)
    if acc is None:
        acc = rows
    else:
        for a, r in zip(acc, rows):
            a["ms"] += r["ms"]
print("idx,family,ms,gflop,tflops,gbs,desc")
tot = 0.0
for a in acc:
    ms = a["ms"] / REPS
    tot += ms
    tf = a["flops"] / ms / 1e9 if ms > 0 else 0.0
    gb = a["bytes"] / ms / 1e6 if ms > 0 else 0.0
    print(f'{a["idx"]},{a["family"]},{ms:.4f},{a["flops"]/1e9:.2f},{tf:.1f},{gb:.0f},{a["desc"]}')
print(f"# total {tot:.3f} ms over {len(acc)} launches, batch {B}", file=sys.stderr)
