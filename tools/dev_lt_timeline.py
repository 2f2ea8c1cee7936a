"""Dev helper (GPU): role timeline of linear_tc_kernel (CTA 0, first 8 row tiles) for a [N,128] x [128,128] Linear with
Swish, residual and the pre-activation copy (the update_net_2 shape of GNN_Layer).
Build:  MGB_VARIANT=tl MGB_NVCC_EXTRA=-DMGB_TIMELINE python -m magnet_b200.build
Run:    MGB_VARIANT=tl python tools/dev_lt_timeline.py [rows] [k]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from magnet_b200 import _lib, functional as MF

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(3)
x = torch.randn(rows, 128, generator=g).to(dev)
res = torch.randn(rows, 128, generator=g).to(dev)
W = (torch.randn(128, 128, generator=g) / 11).to(dev)
b = torch.randn(128, generator=g).to(dev)
MF._LINEAR_TC_PRECISION = 1
packed = MF._tc_weight_images(W, None, 1)
L = _lib.lib()
L.mgb_debug_set_lt_timeline.argtypes = [ctypes.c_void_p]
for use_res, want_pre, act in ((1, 1, 2),):
    r = res if use_res else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.mgb_debug_set_lt_timeline(ctypes.c_void_p(0))
    for _ in range(3):
        MF._linear_forward(x, W, b, act, r, packed, bool(want_pre))
    ev0.record()
    for _ in range(10):
        MF._linear_forward(x, W, b, act, r, packed, bool(want_pre))
    ev1.record()
    torch.cuda.synchronize()
    print(f"linear {rows} x 128 -> 128, act {act} residual {use_res} y_pre {want_pre}: {ev0.elapsed_time(ev1) / 10 * 1e3:.1f} us")
    tl = torch.zeros(5 * 8 * 4, dtype=torch.int64, device=dev)
    L.mgb_debug_set_lt_timeline(ctypes.c_void_p(tl.data_ptr()))
    MF._linear_forward(x, W, b, act, r, packed, bool(want_pre))
    torch.cuda.synchronize()
    t = tl.cpu().reshape(5, 8, 4)[:3]
    t0 = int(t[t > 0].min())
    f = lambda v: f"{int(v) - t0:7d}" if v > 0 else "      -"
    for it in range(7):
        print(f"tile {it}: prod converted {f(t[0, it, 0])} stage_free {f(t[0, it, 1])} full {f(t[0, it, 2])} | mma ready {f(t[1, it, 0])} issued {f(t[1, it, 1])}"
              f" | epi wait {f(t[2, it, 0])} acc_ready {f(t[2, it, 1])} first_ld {f(t[2, it, 2])} done {f(t[2, it, 3])}")

# ---- weight gradient of the same Linear (mgb_linear_tc_bwd: wgrad_tc_kernel, then the data gradient) ----
dy = torch.randn(rows, 128, generator=g).to(dev)
y_pre = torch.randn(rows, 128, generator=g).to(dev)
dx, dw, db = torch.empty_like(x), torch.empty_like(W), torch.empty_like(b)
for K in (128, 256):
    xk = torch.randn(rows, K, generator=g).to(dev)
    Wk = (torch.randn(128, K, generator=g) / 11).to(dev)
    pk = MF._tc_weight_images(Wk, None, 1)
    dxk, dwk = torch.empty_like(xk), torch.empty_like(Wk)
    ws_bytes = L.mgb_linear_tc_bwd_workspace(rows, K, 128)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    call = lambda: _lib.check(L.mgb_linear_tc_bwd(_lib.ptr(dy), _lib.ptr(y_pre), 2, _lib.ptr(xk), rows, K, 128, _lib.ptr(pk), _lib.ptr(dxk), _lib.ptr(dwk),
                                                  _lib.ptr(db), 0, 1, _lib.ptr(ws), ws_bytes, _lib.stream()), "linear_tc_bwd")
    L.mgb_debug_set_lt_timeline(ctypes.c_void_p(0))
    for _ in range(3):
        call()
    L.mgb_profile_enable(1)
    for _ in range(10):
        call()
    torch.cuda.synchronize()
    L.mgb_profile_enable(0)
    for kid, name in ((2, "dgrad linear"), (3, "wgrad")):
        tt, cc = ctypes.c_double(0), ctypes.c_int64(0)
        L.mgb_profile_collect(kid, ctypes.byref(tt), ctypes.byref(cc))
        print(f"K = {K} {name}: {tt.value / max(cc.value, 1) * 1e3:.1f} us x {cc.value}")
    tl = torch.zeros(5 * 8 * 4, dtype=torch.int64, device=dev)
    L.mgb_debug_set_lt_timeline(ctypes.c_void_p(tl.data_ptr()))
    call()
    torch.cuda.synchronize()
    t = tl.cpu().reshape(5, 8, 4)
    w = t[3:5]
    t0 = int(w[w > 0].min())
    f = lambda v: f"{int(v) - t0:7d}" if v > 0 else "      -"
    for it in range(7):
        print(f"wgrad tile {it}: prod Y'conv {f(t[3, it, 0])} Y'full {f(t[3, it, 1])} X0full {f(t[3, it, 2])} Xlast {f(t[3, it, 3])} | mma Y'ready {f(t[4, it, 0])} X0ready {f(t[4, it, 1])} Xnready {f(t[4, it, 2])} issued {f(t[4, it, 3])}")
