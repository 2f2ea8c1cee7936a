"""Dev helper (GPU): per-tile role timeline of the backward edge kernel (CTA 0, first 16 tiles).
Build:  MGB_VARIANT=tl MGB_NVCC_EXTRA=-DMGB_TIMELINE python -m magnet_b200.build
Run:    MGB_VARIANT=tl python tools/dev_timeline.py [precision]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from magnet_b200 import _lib, functional as MF, graph as MG, synthetic as S
from magnet_b200.mpnn import GNN_Layer

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32_tc"
MF.set_precision(prec)
dev = "cuda"
L = _lib.lib()
B, N = 32, 4096
g = S._gen(1)
mesh = S.mesh("uniform", N, 2, g)
pos = mesh.repeat(B, 1).to(dev)
seg = MG.uniform_segments(B, N, dev)
ei = MG.radius_graph(pos, 0.09, loop=False, ptr=seg.gptr)
plan = MG.plan_for(ei, B * N)
batch = torch.arange(B, device=dev).repeat_interleave(N)
layer = GNN_Layer(128, 128, 128, 10, 1).to(dev)
x = torch.randn(B * N, 128, device=dev, requires_grad=True)
u = torch.randn(B * N, 10, device=dev)
var = torch.rand(B * N, 1, device=dev)
p2 = pos[:, :1].repeat(1, 2).contiguous()
for _ in range(2):
    y = layer(x, u, p2, var, ei, batch, plan=plan, segments=seg)
    y.backward(torch.ones_like(y))
tl = torch.zeros(5 * 16 * 4, dtype=torch.int64, device=dev)
L.mgb_debug_set_timeline.argtypes = [ctypes.c_void_p]
L.mgb_debug_set_timeline(ctypes.c_void_p(tl.data_ptr()))
y = layer(x, u, p2, var, ei, batch, plan=plan, segments=seg)
y.backward(torch.ones_like(y))
torch.cuda.synchronize()
t = tl.cpu().reshape(5, 16, 4)
t0 = int(t[t > 0].min())
names = ["producer(start,done)", "mma(full,dz_full,d2_empty)", "epi1(pref,d1_full,dz_empty,done)", "epi2a(start,d2_full,done)", "epi2b(start,d2_full,done)"]
for it in range(2, 10):
    print(f"--- tile {it}")
    for r in range(5):
        print(f"  {names[r]:36s}", [int(v) - t0 if v > 0 else None for v in t[r, it]])
