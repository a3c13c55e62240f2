"""Development probe (GPU box): per-stage and end-to-end max|d| of the CUDA path against the CPU oracle,
for the conv variants (fp32 SIMT, tcgen05 with 3 / 2 / 1 passes, single accumulator set)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from canonswap_b200 import synth, _lib
from canonswap_b200.engine import Engine
from oracle import canonswap_oracle as O

net = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = 2 if net <= 128 else 1
W = synth.synth_weights()
inp = synth.synth_inputs(B, net, seed=int([a[5:] for a in sys.argv[2:] if a.startswith('seed=')][0]) if any(a.startswith('seed=') for a in sys.argv[2:]) else 1234)
ref = O.frame(W, inp["frames"], inp["x_t"], inp["x_can"], inp["source_id"], debug_decodes=False)
cu = {k: v.cuda() for k, v in inp.items()}
pre = {}
for a in sys.argv[2:]:
    if a.startswith("pre="):
        for kv in a[4:].split(","):
            kk, vv = kv.split(":")
            pre[int(kk)] = int(vv)
seed = 1234
for a in sys.argv[2:]:
    if a.startswith("seed="):
        seed = int(a[5:])
eng = Engine(W, net_hw=(net, net), max_batch=B, device=0, options=pre)
eng.set_identity(cu["source_id"])


def d(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return (a - b).abs().max().item(), b.abs().max().item()


def run(tag):
    r = {}
    r["f_s"] = d(eng.appearance(cu["frames"]), ref["f_s"])
    out, occ, _ = eng.warp(ref["f_s"].cuda(), cu["x_t"], cu["x_can"], want_deformation=True)
    r["f_can"] = d(out, ref["f_can"]); r["occ_can"] = d(occ, ref["occ_can"])
    r["f_swap"] = d(eng.swap(ref["f_can"].cuda()), ref["f_swap"])
    r["f_refine"] = d(eng.refine(ref["f_swap"].cuda()), ref["f_refine"])
    wf = eng.warp_forward(ref["f_refine"].cuda(), kp_driving=cu["x_t"], kp_source=cu["x_can"])
    r["deformation"] = d(wf["deformation"], ref["deformation"]); r["warp_out"] = d(wf["out"], ref["warp_out"])
    r["spade"] = d(eng.spade(ref["warp_out"].cuda()), ref["out"])
    o = torch.empty(B, 3, 2 * net, 2 * net, device="cuda")
    eng.frame(cu["frames"], cu["x_t"], cu["x_can"], out_f32=o)
    r["E2E"] = d(o, ref["out"])
    print(tag, " ".join(f"{k}={v[0]:.2e}/{v[1]:.1f}" for k, v in r.items()), flush=True)


if "nosimt" not in sys.argv:
    eng.set_option(_lib.CS_OPT_CONV_IMPL, 1); run("simt     ")
eng.set_option(_lib.CS_OPT_CONV_IMPL, 0)
comps = (100, 120, 140, 170)
for a in sys.argv[2:]:
    if a.startswith("comps="):
        comps = tuple(int(x) for x in a[6:].split(","))
for comp in comps:
    eng.set_option(_lib.CS_OPT_TC_COMP, comp); run(f"tc comp={comp}")
eng.set_option(_lib.CS_OPT_TC_COMP, 170)
if "pair" in sys.argv:
    eng.set_option(_lib.CS_OPT_TC_PAIR, 1); run("tc pair  ")
