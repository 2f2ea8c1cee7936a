"""CPU tests of the oracle itself: third-party restatement vs independent implementations, the
restatement vs the committed golden vectors (generated from the reference's unmodified files),
and — when /root/reference is present — vs those files directly."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import graph as OG
from oracle import restatement as R
from magnet_b200 import synthetic as S


def _sorted_pairs(ei):
    a = ei.t().tolist()
    return sorted(map(tuple, a))


@pytest.mark.parametrize("d,r", [(2, 0.11), (1, 0.02)])
def test_radius_matches_kdtree_without_truncation(d, r):
    from scipy.spatial import cKDTree
    g = S._gen(1)
    x = torch.rand(300, d, generator=g)
    ei = OG.radius_graph(x, r, None, loop=False, max_num_neighbors=1000)
    tree = cKDTree(x.double().numpy())
    pairs = {(j, i) for i, nb in enumerate(tree.query_ball_point(x.double().numpy(), r)) for j in nb if j != i}
    # fp32 vs fp64 can only differ on pairs at the boundary; the mesh has none within 1e-6
    assert set(_sorted_pairs(ei)) == pairs
    assert torch.equal(ei, OG.radius_graph(x, r, None, loop=False, max_num_neighbors=1000, fma=False))


def test_radius_truncation_is_first_by_index():
    g = S._gen(2)
    x = torch.rand(200, 2, generator=g)
    r, cap = 0.4, 8
    ei = OG.radius_graph(x, r, None, loop=True, max_num_neighbors=cap)
    xn = x.numpy()
    r2 = np.float32(r * r)
    for c in range(200):
        want = []
        for j in range(200):
            dx, dy = np.float32(xn[j, 0] - xn[c, 0]), np.float32(xn[j, 1] - xn[c, 1])
            if np.float32(np.float32(dx * dx) + np.float32(dy * dy)) < r2:
                want.append(j)
            if len(want) == cap:
                break
        got = ei[0][ei[1] == c].tolist()
        assert got == want
    # loop=False: cap+1 hits, then the centre is dropped (may keep cap+1 neighbours)
    ei2 = OG.radius_graph(x, r, None, loop=False, max_num_neighbors=cap)
    deg = torch.bincount(ei2[1], minlength=200)
    assert int(deg.max()) <= cap + 1 and not bool((ei2[0] == ei2[1]).any())


def test_knn_order_and_ties():
    x = torch.tensor([[0.0, 0.0], [1.0, 0.0], [1.0, 0.0], [-1.0, 0.0], [3.0, 0.0]])
    y = torch.tensor([[0.0, 0.0]])
    ai = OG.knn(x, y, 4)
    assert ai[1].tolist() == [0, 1, 2, 3]          # equal distances keep the lower index first
    g = S._gen(3)
    xs, ys = torch.rand(400, 2, generator=g), torch.rand(50, 2, generator=g)
    ai = OG.knn(xs, ys, 5)
    d = torch.cdist(ys.double(), xs.double())
    assert torch.equal(ai[1].reshape(50, 5), d.topk(5, largest=False).indices)


def test_scatter_mean_and_instance_norm():
    g = S._gen(4)
    src = torch.randn(500, 16, generator=g)
    idx = torch.randint(0, 40, (500,), generator=g)
    onehot = torch.zeros(40, 500).scatter_(0, idx[None], 1.0)
    want = onehot @ src / onehot.sum(1, keepdim=True).clamp(min=1)
    assert rel_err(R.scatter_mean(src, idx, 40), want) < 1e-6
    x = torch.randn(90, 8, generator=g)
    batch = torch.arange(3).repeat_interleave(30)
    want = torch.cat([torch.nn.functional.instance_norm(x[b * 30:(b + 1) * 30].T[None])[0].T for b in range(3)])
    assert rel_err(R.instance_norm(x, batch), want) < 1e-5


def test_restatement_matches_golden_graphs(golden):
    for name, c in golden("radius_graph.pt").items():
        ei = OG.radius_graph(c["x"], c["r"], c["batch"], loop=c["loop"])
        assert torch.equal(ei, c["edge_index"]), name
    for name, c in golden("knn.pt").items():
        assert torch.equal(OG.knn(c["x"], c["y"], c["k"], c["batch_x"], c["batch_y"]), c["assign_index"]), name


def _gnn_layer_shapes(tw, dp):
    return {"message_net_1.0.weight": (128, 256 + tw + dp + 1), "message_net_1.0.bias": (128,),
            "message_net_2.0.weight": (128, 128), "message_net_2.0.bias": (128,),
            "update_net_1.0.weight": (128, 257), "update_net_1.0.bias": (128,),
            "update_net_2.0.weight": (128, 128), "update_net_2.0.bias": (128,)}


def test_restatement_matches_golden_gnn_layer(golden):
    for name, c in golden("gnn_layer.pt").items():
        dp = c["pos"].shape[1]
        sd = S.seeded_state_dict(_gnn_layer_shapes(c["time_window"], dp), c["seed"])
        y = R.gnn_layer(sd, "", c["x"], c["u"], c["pos"], c["variables"], c["edge_index"], c["batch"])
        assert rel_err(y, c["y"]) < 1e-6, name


def test_restatement_matches_golden_models(golden):
    from magnet_b200.mpnn import MPNN_2d
    from magnet_b200.magnet_gnn import MAgNetGNN
    from oracle.reference_loader import HParams
    c = golden("mpnn.pt")["mpnn_2d"]
    sd = S.seeded_state_dict({k: tuple(v.shape) for k, v in MPNN_2d(HParams(c["hparams"])).state_dict().items()}, c["seed"])
    b = c["batch"]
    u = b["u"].permute(0, 2, 1)
    g = R.mpnn_build_graph(u[:, :10], b["t"], b["x"], [9, 9], 4, True)
    assert torch.equal(g["edge_index"], c["edge_index"])
    y = R.mpnn_forward(sd, g, b["x"][0, -1], b["t"][0, -1], b["t"][0][1] - b["t"][0][0], 10, 5, True)
    assert rel_err(y, c["y"]) < 1e-6
    m = golden("magnet_gnn.pt")["forward"]
    sd = S.seeded_state_dict({k: tuple(v.shape) for k, v in MAgNetGNN(HParams(m["hparams"])).state_dict().items()}, m["seed"])
    b = m["batch"]
    out = R.magnet_forward(sd, b["lr_frames"][:, :10], b["coords_lr"], b["coords_hr"], b["t"][:, :20],
                           b["hr_points"][:, 9], 0.08, 4, 5)
    for got, key in zip(out, ("out_hr", "out_lr", "hr_points")):
        assert rel_err(got, m[key]) < 1e-6, key


@pytest.mark.reference
def test_restatement_matches_unmodified_reference():
    from oracle import reference_loader as rl
    ref = rl.load()
    m = ref.mpnn_2d.MPNN_2d(rl.mpnn_2d_hparams()).eval()
    sd = S.seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 5)
    m.load_state_dict(sd, strict=True)
    b = S.graph_batch(B=2, N=144, nt=30, d=2, kind="uniform", seed=5)
    u = b["u"].permute(0, 2, 1)
    graph = m._build_graph(u[:, :10], b["t"], b["x"], steps=[9, 9])
    with torch.no_grad():
        y = m.forward(graph, b["x"][0, -1], b["t"][0, -1], b["t"][0][1] - b["t"][0][0])
    g = R.mpnn_build_graph(u[:, :10], b["t"], b["x"], [9, 9], 4, True)
    assert torch.equal(g["edge_index"], graph.edge_index)
    y2 = R.mpnn_forward(sd, g, b["x"][0, -1], b["t"][0, -1], b["t"][0][1] - b["t"][0][0], 10, 5, True)
    assert rel_err(y2, y) < 1e-6
    m.validation_step(b, 0)
    ro = R.mpnn_rollout(sd, b, 10, 5, 4, True)
    assert rel_err(torch.nn.functional.l1_loss(ro, u[:, 10:]), m.logged["val_loss"]) < 1e-5
