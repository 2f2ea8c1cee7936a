"""Dev helper (GPU): role timeline of the fused InteractionNetwork edge kernel (CTA 0, first 4 tile pairs).
Build:  MGB_VARIANT=tl MGB_NVCC_EXTRA=-DMGB_TIMELINE python -m magnet_b200.build
Run:    MGB_VARIANT=tl python tools/dev_ie_timeline.py [precision]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from magnet_b200 import _lib, functional as MF, graph as MG, synthetic as S
from magnet_b200.magnet_gnn import InteractionNetwork

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32_tc"
MF.set_precision(prec)
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
L = _lib.lib()
with torch.no_grad():
    for _ in range(2):
        layer(x, ei, ef, plan=plan, return_e=False)
    tl = torch.zeros(4 * 4 * 5 * 2 * 4, dtype=torch.int64, device=dev)
    L.mgb_debug_set_ie_timeline.argtypes = [ctypes.c_void_p]
    L.mgb_debug_set_ie_timeline(ctypes.c_void_p(tl.data_ptr()))
    layer(x, ei, ef, plan=plan, return_e=False)
    torch.cuda.synchronize()
t = tl.cpu().reshape(4, 4, 5, 2, 4)
t0 = int(t[t > 0].min())
f = lambda v: f"{int(v) - t0:7d}" if v > 0 else "      -"
for it in range(1, 3):
    for t_ in range(2):
        print(f"pair {it} tile {t_} producer: pq_done {f(t[3, it, 0, t_, 0])} slot_free {f(t[3, it, 0, t_, 1])} x_full {f(t[3, it, 0, t_, 2])}")
    for l in range(5):
        print(f"pair {it} layer {l}: w_issue {f(t[0, it, l, 0, 3])} w_ready {f(t[0, it, l, 0, 0])}")
        for t_ in range(2):
            print(f"    tile {t_}: mma operand_ready {f(t[0, it, l, t_, 1])} issued {f(t[0, it, l, t_, 2])} | epi0 start {f(t[1, it, l, t_, 0])} done {f(t[1, it, l, t_, 1])}"
                  f" | epi1 start {f(t[2, it, l, t_, 0])} done {f(t[2, it, l, t_, 1])}")
