"""Development probe (GPU box): per-stage max / rms error of the CUDA path against the CPU oracle (each stage fed the
ORACLE's input, so errors do not compound) and end to end, for a list of option sets.

    python tools/stage_err.py NET [seed=S] [wseed=S] [B=n] [opts=13:0,8:0 ...]      (each opts= argument is one variant)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from canonswap_b200 import synth, _lib
from canonswap_b200.engine import Engine
from oracle import canonswap_oracle as O

net = int(sys.argv[1]) if len(sys.argv) > 1 else 128
kv = dict(a.split("=", 1) for a in sys.argv[2:] if "=" in a and not a.startswith("opts="))
seed, wseed, B = int(kv.get("seed", 1234)), int(kv.get("wseed", synth.WEIGHT_SEED)), int(kv.get("B", 1))
variants = [a[5:] for a in sys.argv[2:] if a.startswith("opts=")] or [""]
W = synth.synth_weights(wseed)
inp = synth.synth_inputs(B, net, seed=seed)
ref = O.frame(W, inp["frames"], inp["x_t"], inp["x_can"], inp["source_id"])
cu = {k: v.cuda() for k, v in inp.items()}


def d(a, b):
    a, b = a.float().cpu().double(), b.float().cpu().double()
    e = a - b
    return e.abs().max().item(), e.pow(2).mean().sqrt().item(), b.abs().max().item()


for var in variants:
    opts = {int(k): int(v) for k, v in (x.split(":") for x in var.split(",") if x)}
    pre = {k: v for k, v in opts.items() if k in (_lib.CS_OPT_TC_CHAIN_MAX, _lib.CS_OPT_TC_BN_MAX)}
    eng = Engine(W, net_hw=(net, net), max_batch=B, device=0, options=pre)
    for k, v in opts.items():
        if k not in pre:
            eng.set_option(k, v)
    eng.set_identity(cu["source_id"])
    r = {}
    r["f_s"] = d(eng.appearance(cu["frames"]), ref["f_s"])
    out, occ, _ = eng.warp(ref["f_s"].cuda(), cu["x_t"], cu["x_can"], want_deformation=True)
    r["f_can"] = d(out, ref["f_can"]); r["occ_can"] = d(occ, ref["occ_can"])
    r["f_swap"] = d(eng.swap(ref["f_can"].cuda()), ref["f_swap"])
    r["f_refine"] = d(eng.refine(ref["f_swap"].cuda()), ref["f_refine"])
    wf = eng.warp_forward(ref["f_refine"].cuda(), kp_driving=cu["x_t"], kp_source=cu["x_can"])
    r["deform"] = d(wf["deformation"], ref["deformation"]); r["warp_out"] = d(wf["out"], ref["warp_out"])
    r["spade"] = d(eng.spade(ref["warp_out"].cuda()), ref["out"])
    o = torch.empty(B, 3, 2 * net, 2 * net, device="cuda")
    eng.frame(cu["frames"], cu["x_t"], cu["x_can"], out_f32=o)
    r["E2E"] = d(o, ref["out"])
    print(f"net={net} seed={seed} wseed={wseed} opts=[{var}]")
    for k, v in r.items():
        print(f"   {k:9s} max={v[0]:.2e} rms={v[1]:.2e} range={v[2]:.1f}  max/range={v[0] / max(v[2], 1e-30):.1e}", flush=True)
    eng.close()
