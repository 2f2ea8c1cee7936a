"""Dev helper (GPU): one InteractionNetwork training step (forward + backward through autograd) at the bench size, for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from magnet_b200 import graph as MG, functional as MF
from magnet_b200.magnet_gnn import InteractionNetwork
dev = torch.device("cuda", 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
pos, x, g = bench.in_workload(bench.IN_SAMPLES, bench.IN_NODES_PER_SAMPLE, 900, dev)
pos, x = pos.to(dev), x.to(dev)
seg = MG.uniform_segments(bench.IN_SAMPLES, bench.IN_NODES_PER_SAMPLE, dev)
ei = MG.radius_graph(pos, 0.08, loop=True, ptr=seg.gptr, swap_rows=True)
N, E = x.shape[0], ei.shape[1]
plan = MG.plan_for(ei, N)
ef = torch.randn(E, 128, generator=g).to(dev)
layer = InteractionNetwork(128, 128, 128, 128, 4, 128).to(dev)
layer.load_state_dict(bench.in_state_dict())
gy = torch.randn(N, 128, generator=g).to(dev)
for _ in range(reps):
    xi, ei_ = x.detach().requires_grad_(), ef.detach().requires_grad_()
    y, _ = layer(xi, ei, ei_, plan=plan, return_e=False)
    y.backward(gy)
    layer.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print("E", E)
