"""Dev helper (GPU): role timeline of the fused InteractionNetwork backward passes (CTA 0, tiles 1-2 of each pass).
Build:  MGB_VARIANT=tl MGB_NVCC_EXTRA=-DMGB_TIMELINE python -m magnet_b200.build
Run:    MGB_VARIANT=tl python tools/dev_ib_timeline.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from magnet_b200 import _lib, functional as MF, graph as MG, synthetic as S
from magnet_b200.magnet_gnn import InteractionNetwork
dev = torch.device("cuda", 0)
g = S._gen(900)
B, n_per, r = 8, 32768, 0.02
pos = (2 * torch.rand(B * n_per, 2, generator=g) - 1).to(dev)
seg = MG.uniform_segments(B, n_per, dev)
ei = MG.radius_graph(pos, r, loop=True, ptr=seg.gptr, swap_rows=True)
N, E = B * n_per, ei.shape[1]
plan = MG.plan_for(ei, N)
layer = InteractionNetwork(128, 128, 128, 128, 4, 128).to(dev)
layer.load_state_dict(S.seeded_state_dict({k: tuple(v.shape) for k, v in layer.state_dict().items()}, 5))
x = torch.randn(N, 128, generator=g).to(dev)
ef = torch.randn(E, 128, generator=g).to(dev)
gy = torch.randn(N, 128, generator=g).to(dev)
L = _lib.lib()
def step():
    xi, ei_ = x.detach().requires_grad_(), ef.detach().requires_grad_()
    y, _ = layer(xi, ei, ei_, plan=plan, return_e=False)
    y.backward(gy)
for _ in range(2): step()
tl = torch.zeros(2 * 4 * 3 * 32, dtype=torch.int64, device=dev)
L.mgb_debug_set_ib_timeline.argtypes = [ctypes.c_void_p]
L.mgb_debug_set_ib_timeline(ctypes.c_void_p(tl.data_ptr()))
step(); torch.cuda.synchronize()
t = tl.cpu().reshape(2, 4, 3, 32)
print("E", E)
for p in range(2):
    for it in (1, 2):
        base = int(t[p, it, 0, 0])
        f = lambda v: f"{int(v) - base:6d}" if v > 0 else "     -"
        print(f"pass {'AB'[p]} tile {it}  (cycles since the epilogue started waiting for layer 0)")
        print("  epilogue : " + " ".join(f(v) for v in t[p, it, 0] if v > 0))
        print("  mma      : " + " ".join(f(v) for v in t[p, it, 1] if v > 0))
        print("  producer : " + " ".join(f(v) for v in t[p, it, 2] if v > 0))
        if t[p, it + 1, 0, 0] > 0: print("  tile period:", int(t[p, it + 1, 0, 0] - t[p, it, 0, 0]))
