"""Per-launch timing of the motion extractor (GPU box): python tools/motion_table.py [batch] > gpurun_out/motion_layers.csv"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from canonswap_b200 import synth, spec
from canonswap_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
W = synth.synth_weights(with_motion=True)
inp = synth.synth_inputs(B, 256)
eng = Engine(W, net_hw=(256, 256), max_batch=B, device=0)
x = inp["frames"].cuda()
for _ in range(3):
    eng.motion(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    h = eng.motion(x)
    eng.keypoints(h)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"# motion extractor + keypoints: {ms:.3f} ms per batch of {B} = {B / ms * 1e3:.0f} frames/s ({11.6 * B / ms:.1f} TFLOP/s-eq)", file=sys.stderr)
eng.profile(True)
eng.motion(x)
rows = eng.profile_dump()
eng.profile(False)
print("idx,family,ms,gflop,tflops,gbs,desc")
agg = {}
for a in rows:
    ms = a["ms"]
    print(f'{a["idx"]},{a["family"]},{ms:.4f},{a["flops"]/1e9:.2f},{a["flops"]/ms/1e9 if ms > 0 else 0:.1f},{a["bytes"]/ms/1e6 if ms > 0 else 0:.0f},{a["desc"]}')
    k = a["desc"].split(" M=")[0]
    agg[k] = agg.get(k, 0.0) + ms
print("# " + ", ".join(f"{k}: {v:.3f} ms" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])), file=sys.stderr)
