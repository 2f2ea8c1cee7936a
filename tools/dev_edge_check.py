"""Dev helper (GPU): GNN_Layer fwd/bwd error vs the fp64 oracle and kernel timings for each precision mode."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import graph as OG, restatement as R
from magnet_b200 import functional as MF, graph as MG, synthetic as S
from magnet_b200.mpnn import GNN_Layer


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max())


def main():
    dev = "cuda"
    B, N, r = 4, 4096, 0.09
    g = S._gen(77)
    pos = torch.rand(B * N, 2, generator=g)
    batch = torch.arange(B).repeat_interleave(N)
    ei = OG.radius_graph(pos, r, batch, loop=False, threads=8)
    shapes = {"message_net_1.0.weight": (128, 269), "message_net_1.0.bias": (128,), "message_net_2.0.weight": (128, 128),
              "message_net_2.0.bias": (128,), "update_net_1.0.weight": (128, 257), "update_net_1.0.bias": (128,),
              "update_net_2.0.weight": (128, 128), "update_net_2.0.bias": (128,)}
    sd = S.seeded_state_dict(shapes, 3)
    x, u, var = torch.randn(B * N, 128, generator=g), torch.randn(B * N, 10, generator=g), torch.rand(B * N, 1, generator=g)
    gy = torch.randn(B * N, 128, generator=g)
    x64, u64, p64 = (t.double().requires_grad_() for t in (x, u, pos))
    sd64 = {k: v.double().requires_grad_() for k, v in sd.items()}
    y64 = R.gnn_layer(sd64, "", x64, u64, p64, var.double(), ei, batch)
    y64.backward(gy.double())
    print(f"E={ei.shape[1]}")
    for prec in ("fp32", "fp32_tc", "bf16"):
        MF.set_precision(prec)
        layer = GNN_Layer(128, 128, 128, 10, 1).to(dev)
        layer.load_state_dict(sd)
        xg, ug, pg = (t.to(dev).requires_grad_() for t in (x, u, pos))
        try:
            y = layer(xg, ug, pg, var.to(dev), ei.to(dev), batch.to(dev))
            y.backward(gy.to(dev))
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(prec, "FAILED", e)
            continue
        errs = {"y": rel(y, y64), "dx": rel(xg.grad, x64.grad), "du": rel(ug.grad, u64.grad), "dpos": rel(pg.grad, p64.grad)}
        for k, p in layer.named_parameters():
            errs["d" + k.replace("_net_", "").replace(".0.", ".")] = rel(p.grad, sd64[k].grad)
        print(prec, " ".join(f"{k}={v:.2e}" for k, v in errs.items()))


if __name__ == "__main__":
    main()
