"""Development probe (GPU box): where does the end-to-end error come from?  Runs the stages on the GPU feeding each its OWN
previous output (errors compound as in cs_frame), compares every intermediate with the oracle's, and prints for the worst
image pixels the errors of the tensors they were decoded from.   python tools/err_trace.py NET [seed=S] [wseed=S] [B=n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from canonswap_b200 import synth
from canonswap_b200.engine import Engine
from oracle import canonswap_oracle as O

net = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kv = dict(a.split("=", 1) for a in sys.argv[2:] if "=" in a)
seed, wseed, B = int(kv.get("seed", 1234)), int(kv.get("wseed", synth.WEIGHT_SEED)), int(kv.get("B", 1))
W = synth.synth_weights(wseed)
inp = synth.synth_inputs(B, net, seed=seed)
ref = O.frame(W, inp["frames"], inp["x_t"], inp["x_can"], inp["source_id"])
cu = {k: v.cuda() for k, v in inp.items()}
eng = Engine(W, net_hw=(net, net), max_batch=B, device=0)
eng.set_identity(cu["source_id"])


def rep(name, a, b):
    e = (a.float().cpu().double() - b.double())
    print(f"  {name:10s} max={e.abs().max().item():.2e} rms={e.pow(2).mean().sqrt().item():.2e} range={b.abs().max().item():.1f}", flush=True)
    return e


f_s = eng.appearance(cu["frames"]); rep("f_s", f_s, ref["f_s"])
f_can, occ_can, _ = eng.warp(f_s, cu["x_t"], cu["x_can"], want_deformation=True); rep("f_can", f_can, ref["f_can"])
f_swap = eng.swap(f_can); rep("f_swap", f_swap, ref["f_swap"])
f_ref = eng.refine(f_swap); rep("f_refine", f_ref, ref["f_refine"])
wf = eng.warp_forward(f_ref, kp_driving=cu["x_t"], kp_source=cu["x_can"])
e_def = rep("deform", wf["deformation"], ref["deformation"]); e_occ = rep("occ", wf["occlusion_map"], ref["occ"])
e_wo = rep("warp_out", wf["out"], ref["warp_out"])
img = eng.spade(wf["out"]); e_img = rep("image", img, ref["out"])
# the same decode from the ORACLE's warp_out: the decoder's own error
rep("image|G", eng.spade(ref["warp_out"].cuda()), ref["out"])
# and the second warp + decode from the ORACLE's refine output
wf2 = eng.warp_forward(ref["f_refine"].cuda(), kp_driving=cu["x_t"], kp_source=cu["x_can"])
rep("image|W+G", eng.spade(wf2["out"]), ref["out"])
rep("image|R+W+G", eng.spade(eng.warp_forward(eng.refine(ref["f_swap"].cuda()), kp_driving=cu["x_t"], kp_source=cu["x_can"])["out"]), ref["out"])
rep("image|S+R+W+G", eng.spade(eng.warp_forward(eng.refine(eng.swap(ref["f_can"].cuda())), kp_driving=cu["x_t"], kp_source=cu["x_can"])["out"]), ref["out"])
ea = e_img.abs().amax(1)                      # [B,H,W]
flat = ea.flatten()
top = torch.topk(flat, 8).indices
Himg = ea.shape[1]
print("worst image pixels (b, y, x): err | warp_out err at (y/8, x/8) max over channels | occ err | deformation err max over depth")
for t in top.tolist():
    b, r = divmod(t, Himg * Himg); y, x = divmod(r, Himg)
    yy, xx = y // 8, x // 8
    print(f"  ({b},{y},{x}): {flat[t].item():.2e} | {e_wo[b, :, yy, xx].abs().max().item():.2e} (nbhd {e_wo[b, :, max(0,yy-1):yy+2, max(0,xx-1):xx+2].abs().max().item():.2e})"
          f" | {e_occ[b, 0, yy, xx].abs().item():.2e} | {e_def[b, :, yy, xx].abs().max().item():.2e}")
eng.close()
