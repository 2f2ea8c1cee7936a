"""Dev helper (GPU): fused InteractionNetwork training step (mgb_in_edge_fwd + mgb_in_edge_bwd) against the fp64 oracle and against
the row-wise kernels, per gradient.  usage: python tools/dev_in_bwd.py [nodes_per_sample] [e_scale] [time]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import graph as OG, restatement as R
from magnet_b200 import functional as MF, graph as MG, synthetic as S
from magnet_b200.magnet_gnn import InteractionNetwork
dev = "cuda"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 700
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
g = S._gen(61)
B = 2
pos = 2 * torch.rand(B * N, 2, generator=g) - 1
batch = torch.arange(B).repeat_interleave(N)
e = OG.radius_graph(pos, 0.12 * (700 / N) ** 0.5, batch, loop=True)
ei = torch.stack([e[1], e[0]])
layer = InteractionNetwork(128, 128, 128, 128, 4, 128).to(dev)
sd = S.seeded_state_dict({k: tuple(v.shape) for k, v in layer.state_dict().items()}, 5)
layer.load_state_dict(sd, strict=True)
x = torch.randn(B * N, 128, generator=g)
ef = torch.randn(ei.shape[1], 128, generator=g)
gy = torch.randn(B * N, 128, generator=g) * 1e-3
print("N", B * N, "E", ei.shape[1])

xd, efd, eid, gyd = x.to(dev), ef.to(dev), ei.to(dev), gy.to(dev)
plan = MG.plan_for(eid, B * N)

def run(fused):
    MF.set_fused_training(fused)
    layer.zero_grad(set_to_none=True)
    xg, eg = xd.detach().requires_grad_(), efd.detach().requires_grad_()
    y, _ = layer(xg, eid, eg, e_scale=scale, return_e=False, plan=plan)
    y.backward(gyd)
    torch.cuda.synchronize()
    out = {"y": y.detach(), "dx": xg.grad, "de": eg.grad}
    out.update({k: p.grad.clone() for k, p in layer.named_parameters()})
    return out

def err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

x64, e64 = x.double().requires_grad_(), ef.double().requires_grad_()
sd64 = {"l." + k: v.double().requires_grad_() for k, v in sd.items()}
y64, _ = R.interaction_network(sd64, "l", x64, ei, e64 * scale)
(y64 * gy.double()).sum().backward()
ref = {"y": y64.detach(), "dx": x64.grad, "de": e64.grad}
ref.update({k: sd64["l." + k].grad for k in sd})
row = run(False)
fus = run(True)
for k in ref:
    print(f"{k:28s} fused-vs-fp64 {err(fus[k], ref[k]):9.2e}   rowwise-vs-fp64 {err(row[k], ref[k]):9.2e}   fused-vs-rowwise {err(fus[k], row[k]):9.2e}")
if len(sys.argv) > 3:
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for fused in (False, True):
        run(fused); torch.cuda.synchronize(); ev0.record()
        for _ in range(3): run(fused)
        ev1.record(); torch.cuda.synchronize()
        print("fused" if fused else "rowwise", ev0.elapsed_time(ev1) / 3, "ms per fwd+bwd")
    import ctypes
    from magnet_b200 import _lib
    L = _lib.lib()
    L.mgb_profile_enable(1)
    for _ in range(3): run(True)
    torch.cuda.synchronize()
    L.mgb_profile_enable(0)
    for kid, name in ((5, "in_edge_fwd"), (6, "in_edge_bwd (both passes)")):
        t, c = ctypes.c_double(0), ctypes.c_int64(0)
        L.mgb_profile_collect(kid, ctypes.byref(t), ctypes.byref(c))
        print(name, t.value / max(c.value, 1), "ms per launch,", c.value, "launches")
