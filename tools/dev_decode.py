"""Dev helper (GPU): where the fused INR decoder spends its time — plain 5-layer chain on the same row count, fused decode with a
precomputed neighbour table, fused decode with the in-kernel search."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from magnet_b200 import functional as MF, graph as MG
dev = torch.device("cuda", 0)
Lr, Q, T, k = 1 << 18, 1 << 20, 10, 4
m = bench.magnet_model(dev)
lr_coords, hr_coords, enc, x_lr, t = bench.decode_inputs(Lr, Q, T, 500, dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timeit(fn, reps=3):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ev0.record()
    for _ in range(reps): fn()
    ev1.record(); torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps
lin = m.projector.linears()
with torch.no_grad():
    z = torch.randn(Q * T // 4, 128, device=dev)
    print("chain on %d rows: %.3f ms" % (z.shape[0], timeit(lambda: MF.mlp_chain(z, lin, "relu", cache_owner=m.projector))))
    del z
    lrc, hrc = lr_coords.reshape(Lr, 2), hr_coords.reshape(Q, 2)
    idx = MG.knn_indices(lrc, hrc, k, MG.uniform_ptr(1, Lr, dev), MG.uniform_ptr(1, Q, dev))
    print("knn alone: %.3f ms" % timeit(lambda: MG.knn_indices(lrc, hrc, k, MG.uniform_ptr(1, Lr, dev), MG.uniform_ptr(1, Q, dev))))
    args = (x_lr.reshape(1, T, Lr), enc.reshape(Lr, 128), lrc, hrc, t[:, :T], m.proj_head.weight, m.proj_head.bias, lin, 1, Lr, Q, k, "area")
    print("fused, precomputed idx: %.3f ms" % timeit(lambda: MF.inr_decode_fused(*args, cache_owner=m.projector, idx=idx)))
    print("fused, in-kernel search: %.3f ms" % timeit(lambda: MF.inr_decode_fused(*args, cache_owner=m.projector)))
