"""Dev helper (GPU): role timeline of the fused node-update kernel of GNN_Layer (node_update_tc.cu; CTA 0, first 7 row tiles).
Build:  MGB_VARIANT=tl MGB_NVCC_EXTRA=-DMGB_TIMELINE python -m magnet_b200.build
Run:    MGB_VARIANT=tl python tools/dev_nu_timeline.py [precision]"""
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
x = torch.randn(B * N, 128, device=dev)
u = torch.randn(B * N, 10, device=dev)
var = torch.rand(B * N, 1, device=dev)
p2 = pos[:, :1].repeat(1, 2).contiguous()
with torch.no_grad():
    for _ in range(3):
        layer(x, u, p2, var, ei, batch, plan=plan, segments=seg)
    L.mgb_profile_enable(1)
    for _ in range(10):
        layer(x, u, p2, var, ei, batch, plan=plan, segments=seg)
    torch.cuda.synchronize()
    L.mgb_profile_enable(0)
    tt, cc = ctypes.c_double(0), ctypes.c_int64(0)
    L.mgb_profile_collect(2, ctypes.byref(tt), ctypes.byref(cc))
    print(f"node-level launches of a forward layer (PQ Linear + node update): {tt.value / 10 * 1e3:.1f} us in {cc.value // 10} launches")
    tl = torch.zeros(3 * 8 * 4, dtype=torch.int64, device=dev)
    L.mgb_debug_set_nu_timeline.argtypes = [ctypes.c_void_p]
    L.mgb_debug_set_nu_timeline(ctypes.c_void_p(tl.data_ptr()))
    layer(x, u, p2, var, ei, batch, plan=plan, segments=seg)
    torch.cuda.synchronize()
t = tl.cpu().reshape(3, 8, 4)
t0 = int(t[t > 0].min())
f = lambda v: f"{int(v) - t0:7d}" if v > 0 else "      -"
for it in range(7):
    print(f"tile {it}: prod x_conv {f(t[0, it, 0])} x_full {f(t[0, it, 1])} a_conv {f(t[0, it, 2])} a_full {f(t[0, it, 3])} | mma x_ready {f(t[1, it, 0])} A_issued {f(t[1, it, 1])}"
          f" h_ready {f(t[1, it, 2])} B_issued {f(t[1, it, 3])} | epi A_done {f(t[2, it, 0])} h_written {f(t[2, it, 1])} B_done {f(t[2, it, 2])} stored {f(t[2, it, 3])}")

# ---- the P|Q Linear of the same forward (the only linear_tc_kernel launch of a forward layer) ----
with torch.no_grad():
    tl = torch.zeros(5 * 8 * 4, dtype=torch.int64, device=dev)
    L.mgb_debug_set_nu_timeline(ctypes.c_void_p(0))
    L.mgb_debug_set_lt_timeline.argtypes = [ctypes.c_void_p]
    L.mgb_debug_set_lt_timeline(ctypes.c_void_p(tl.data_ptr()))
    layer(x, u, p2, var, ei, batch, plan=plan, segments=seg)
    torch.cuda.synchronize()
t = tl.cpu().reshape(5, 8, 4)[:3]
t0 = int(t[t > 0].min())
for it in range(7):
    print(f"PQ tile {it}: prod converted {f(t[0, it, 0])} stage_free {f(t[0, it, 1])} full {f(t[0, it, 2])} | mma ready {f(t[1, it, 0])} issued {f(t[1, it, 1])}"
          f" | epi wait {f(t[2, it, 0])} acc_ready {f(t[2, it, 1])} first_ld {f(t[2, it, 2])} done {f(t[2, it, 3])}")
