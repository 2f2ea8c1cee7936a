"""Dev helper (GPU): calls mgb_in_edge_bwd directly and compares its intermediates (dz2 in the workspace, dz0, dP, weight / vector
gradients) with an fp64 torch evaluation of the same chain.  usage: python tools/dev_in_bwd_dbg.py [nodes] [e_scale]
The MGB_IB_DEBUG=1..7 dumps of intermediate quantities need the developer build:
  MGB_VARIANT=dbg MGB_NVCC_EXTRA=-DMGB_IB_DEBUG python -m magnet_b200.build;  MGB_VARIANT=dbg MGB_IB_DEBUG=2 python tools/dev_in_bwd_dbg.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import graph as OG
from magnet_b200 import functional as MF, graph as MG, synthetic as S, _lib
from magnet_b200.magnet_gnn import InteractionNetwork
dev = "cuda"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 700
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
g = S._gen(61)
B = 2
pos = 2 * torch.rand(B * N, 2, generator=g) - 1
batch = torch.arange(B).repeat_interleave(N)
e_ = OG.radius_graph(pos, 0.12 * (700 / N) ** 0.5, batch, loop=True)
ei = torch.stack([e_[1], e_[0]]).to(dev)
layer = InteractionNetwork(128, 128, 128, 128, 4, 128).to(dev)
layer.load_state_dict(S.seeded_state_dict({k: tuple(v.shape) for k, v in layer.state_dict().items()}, 5), strict=True)
x = torch.randn(B * N, 128, generator=g).to(dev)
ef = torch.randn(ei.shape[1], 128, generator=g).to(dev)
dagg = (torch.randn(B * N, 128, generator=g) * 1e-3).to(dev)
Nn, E = B * N, ei.shape[1]
plan = MG.plan_for(ei, Nn)
lin = layer.edge_fn[0].linears()
ln = layer.edge_fn[1]
with torch.no_grad():
    packs = MF.in_edge_pack(x, lin, ln)
    W0 = lin[0].weight
    p = MF.linear_act(x, W0[:, :128], lin[0].bias, "none", owner=W0)
    q = MF.linear_act(x, W0[:, 128:256], torch.zeros_like(lin[0].bias), "none", owner=W0)
    pq = torch.cat([p, q], 1).contiguous()
    L = _lib.lib()
    dpq = torch.empty(Nn, 256, device=dev); dz0 = torch.empty(E, 128, device=dev)
    dW = torch.empty(4, 128, 128, device=dev); db = torch.empty(4, 128, device=dev)
    dgam = torch.empty(128, device=dev); dbet = torch.empty(128, device=dev)
    ws = _lib.workspace(L.mgb_in_edge_bwd_workspace(E), x.device)
    _lib.check(L.mgb_in_edge_bwd(_lib.ptr(dagg), _lib.ptr(ef), scale, _lib.ptr(plan.perm), _lib.ptr(pq), _lib.ptr(plan.rowptr),
                                 _lib.ptr(plan.dst), _lib.ptr(plan.src), Nn, E, _lib.ptr(packs[2]), 3, _lib.ptr(dpq), _lib.ptr(dz0),
                                 _lib.ptr(dW), _lib.ptr(db), _lib.ptr(dgam), _lib.ptr(dbet), _lib.ptr(ws), ws.numel(), _lib.stream()), "bwd")
    torch.cuda.synchronize()
    tiles = (E + 127) // 128
    dz2_k = ws[: tiles * 128 * 128 * 4].view(torch.float32).reshape(tiles * 128, 128)[:E]
    # ---- fp64 evaluation of the same chain, aggregation order
    D = torch.float64
    perm, dst, src = plan.perm.long(), plan.dst.long(), plan.src.long()
    Ws = [l.weight.double() for l in lin]; bs = [l.bias.double() for l in lin]
    eo = ef.double()[perm] * scale
    z0 = pq.double()[dst, :128] + pq.double()[src, 128:] + eo @ Ws[0][:, 256:].T
    h0 = z0.clamp_min(0); z1 = h0 @ Ws[1].T + bs[1]; h1 = z1.clamp_min(0); z2 = h1 @ Ws[2].T + bs[2]; h2 = z2.clamp_min(0)
    z3 = h2 @ Ws[3].T + bs[3]; h3 = z3.clamp_min(0); y = h3 @ Ws[4].T + bs[4]
    mean = y.mean(1, keepdim=True); var = ((y - mean) ** 2).mean(1, keepdim=True); rstd = (var + 1e-5).rsqrt(); xh = (y - mean) * rstd
    deg = (plan.rowptr[1:] - plan.rowptr[:-1]).double()
    dm = dagg.double()[dst] / deg[dst][:, None]
    gg = dm * ln.weight.double()
    dy = rstd * (gg - gg.mean(1, keepdim=True) - xh * (gg * xh).mean(1, keepdim=True))
    dz3 = (dy @ Ws[4]) * (z3 > 0); dz2 = (dz3 @ Ws[3]) * (z2 > 0); dz1 = (dz2 @ Ws[2]) * (z1 > 0); dz0r = (dz1 @ Ws[1]) * (z0 > 0)
    gmax = (dagg.double().abs().amax(1) / deg.clamp_min(1))[deg > 0].max()
    import math
    gs = 2.0 ** (-math.frexp(float(gmax))[1])
    def err(a, b): return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-300))
    def rows_bad(a, b, tol=1e-4):
        sc = b.abs().max()
        return int(((a.double() - b).abs().amax(1) > tol * sc).sum())
    print("E", E, "tiles", tiles, "gs", gs)
    dbg = int(os.environ.get("MGB_IB_DEBUG", "0"))
    if dbg:
        ref = {1: y, 2: dy * gs, 3: dz3 * gs, 4: h0, 5: h1, 6: h2, 7: h3}[dbg]
        print("debug quantity", dbg, err(dz2_k, ref), "bad rows", rows_bad(dz2_k, ref), "of", E)
        bad = ((dz2_k.double() - ref).abs().amax(1) > 1e-4 * ref.abs().max()).nonzero().flatten()
        print("  bad positions (mod 128):", (bad % 128)[:40].tolist())
        selfloop = (dst == src)
        isbad = torch.zeros(E, dtype=torch.bool, device=dev); isbad[bad] = True
        print("  self loops:", int(selfloop.sum()), "bad among self loops:", int((isbad & selfloop).sum()), " bad among others:", int((isbad & ~selfloop).sum()))
        if dbg == 4:
            pqd = pq.double(); ee = eo @ Ws[0][:, 256:].T
            for name, cand in (("P[dst]+Q[src]", pqd[dst, :128] + pqd[src, 128:]), ("P[src]+Q[dst]", pqd[src, :128] + pqd[dst, 128:]),
                               ("P[dst]+Q[dst]", pqd[dst, :128] + pqd[dst, 128:]), ("P[src]+Q[src]", pqd[src, :128] + pqd[src, 128:]),
                               ("P[dst] only", pqd[dst, :128]), ("Q[src] only", pqd[src, 128:]), ("none", 0 * ee)):
                print("   h0 vs relu(%s + We e): %.3e   without e: %.3e" % (name, err(dz2_k, (cand + ee).clamp_min(0)), err(dz2_k, (cand + 0 * ee).clamp_min(0))))
        if dbg == 99:
            for pp in bad[:12].tolist():
                dist = (ref - dz2_k[pp].double()[None]).abs().amax(1)
                qq = int(dist.argmin())
                print("   kernel row", pp, "(dst %d src %d coo %d)" % (int(dst[pp]), int(src[pp]), int(perm[pp])), "best ref row", qq,
                      "(dst %d src %d coo %d)" % (int(dst[qq]), int(src[qq]), int(perm[qq])), "dist", float(dist[qq]), "own dist", float(dist[pp]))
        segstart = (plan.rowptr[:-1]).long()
        print("  segment starts:", segstart[:12].tolist(), " deg:", deg[:12].tolist())
        sys.exit(0)
    print("dz2 (scaled)", err(dz2_k, dz2 * gs), "bad rows", rows_bad(dz2_k, dz2 * gs), "of", E)
    bad = ((dz2_k.double() - dz2 * gs).abs().amax(1) > 1e-4 * (dz2 * gs).abs().max()).nonzero().flatten()
    print("  bad positions (mod 128) sample:", (bad % 128)[:40].tolist(), " tiles:", (bad // 128).unique()[:20].tolist())
    dz0_ref = torch.empty_like(dz0r); dz0_ref[perm] = dz0r * scale
    print("dz0", err(dz0, dz0_ref), "bad rows", rows_bad(dz0, dz0_ref))
    dP = torch.zeros(Nn, 128, dtype=D, device=dev).index_add_(0, dst, dz0r)
    print("dP", err(dpq[:, :128], dP))
    print("dgamma", err(dgam, (dm * xh).sum(0)), "dbeta", err(dbet, dm.sum(0)))
    print("db4", err(db[3], dy.sum(0)), "db3", err(db[2], dz3.sum(0)), "db2", err(db[1], dz2.sum(0)), "db1", err(db[0], dz1.sum(0)))
    print("dW4", err(dW[3], dy.T @ h3), "dW3", err(dW[2], dz3.T @ h2), "dW2", err(dW[1], dz2.T @ h1), "dW1", err(dW[0], dz1.T @ h0))
