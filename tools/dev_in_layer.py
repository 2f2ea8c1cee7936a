"""Dev helper (GPU): time one InteractionNetwork forward (fused edge kernel vs the unfused row-wise kernels) on a large
MAgNet-style graph.  usage: python tools/dev_in_layer.py [nodes_per_sample] [samples] [radius] [reps]"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from magnet_b200 import synthetic as S, graph as MG, functional as MF, _lib
from magnet_b200.magnet_gnn import InteractionNetwork

n_per = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
r = float(sys.argv[3]) if len(sys.argv) > 3 else 0.02
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device("cuda", 0)
g = S._gen(900)
pos = (2 * torch.rand(B * n_per, 2, generator=g) - 1).to(dev)
seg = MG.uniform_segments(B, n_per, dev)
ei = MG.radius_graph(pos, r, loop=True, ptr=seg.gptr, swap_rows=True)
N, E = B * n_per, ei.shape[1]
plan = MG.plan_for(ei, N)
layer = InteractionNetwork(128, 128, 128, 128, 4, 128).to(dev)
layer.load_state_dict(S.seeded_state_dict({k: tuple(v.shape) for k, v in layer.state_dict().items()}, 5))
x = torch.randn(N, 128, generator=g).to(dev)
ef = torch.randn(E, 128, generator=g).to(dev)
L = _lib.lib()
print(f"N={N} E={E} mean degree {E / N:.1f}")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for prec in ("fp32_tc", "bf16"):
    MF.set_precision(prec)
    with torch.no_grad():
        for fused in (True, False):
            if not fused:
                fus = MF.in_edge_fusable
                MF.in_edge_fusable = lambda *a, **k: False
            for _ in range(2):
                y, _ = layer(x, ei, ef, plan=plan, return_e=False)
            torch.cuda.synchronize()
            L.mgb_profile_enable(1)
            ev0.record()
            for _ in range(reps):
                y, _ = layer(x, ei, ef, plan=plan, return_e=False)
            ev1.record()
            torch.cuda.synchronize()
            L.mgb_profile_enable(0)
            t, c = ctypes.c_double(0), ctypes.c_int64(0)
            L.mgb_profile_collect(5, ctypes.byref(t), ctypes.byref(c))
            ms = ev0.elapsed_time(ev1) / reps
            kms = t.value / max(c.value, 1)
            print(f"{prec} fused={fused}: layer {ms:.3f} ms = {E / ms / 1e6:.2f} G edges/s; edge kernel {kms:.3f} ms"
                  + (f" = {E * 229376 / kms / 1e9:.0f} TFLOP/s (reference FLOPs), {E / kms / 1e6:.2f} G edges/s" if c.value else ""))
            if not fused:
                MF.in_edge_fusable = fus
                yu = y
            else:
                yf = y
        print("  fused vs unfused max rel diff", float((yf - yu).abs().max() / yu.abs().max()))
