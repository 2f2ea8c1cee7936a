"""GPU parity: dense stages, GNN_Layer forward/backward and the MPNN models vs the oracle.
fp32 contract: max |a-b| / max |b| <= 1e-5 (BASELINE.json north_star)."""
import pytest
import torch

from conftest import rel_err
from oracle import graph as OG
from oracle import restatement as R
from oracle.reference_loader import HParams, mpnn_2d_hparams, mpnn_hparams
from magnet_b200 import functional as MF, graph as MG, synthetic as S
from magnet_b200.mpnn import GNN_Layer, MPNN, MPNN_2d

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5


def _layer_shapes(tw, dp):
    return {"message_net_1.0.weight": (128, 256 + tw + dp + 1), "message_net_1.0.bias": (128,),
            "message_net_2.0.weight": (128, 128), "message_net_2.0.bias": (128,),
            "update_net_1.0.weight": (128, 257), "update_net_1.0.bias": (128,),
            "update_net_2.0.weight": (128, 128), "update_net_2.0.bias": (128,)}


@pytest.mark.parametrize("rows,fin,fout,act", [(1000, 13, 128, "swish"), (777, 128, 128, "relu"), (300, 257, 128, "none"),
                                               (129, 128, 10, "none"), (5, 132, 128, "relu"), (2048, 128, 1, "none")])
def test_linear_act_fwd_bwd(rows, fin, fout, act):
    g = S._gen(rows)
    x = torch.randn(rows, fin, generator=g)
    W = torch.randn(fout, fin, generator=g) / fin ** 0.5
    b = torch.randn(fout, generator=g)
    if act == "relu":
        # ReLU' is discontinuous at 0: a pre-activation within rounding distance of the kink flips the mask of any fp32
        # implementation (the reference's included), so rows that have one are redrawn until none is closer than 1e-3
        for _ in range(50):
            near = ((x.double() @ W.double().T + b.double()).abs() < 1e-3).any(1)
            if not near.any():
                break
            x[near] = torch.randn(int(near.sum()), fin, generator=g)
        assert not near.any()
    x, W, b = x.to(DEV).requires_grad_(), W.to(DEV).requires_grad_(), b.to(DEV).requires_grad_()
    res = torch.randn(rows, fout, generator=g).to(DEV).requires_grad_()
    y = MF.linear_act(x, W, b, act, res)
    gy = torch.randn(rows, fout, generator=g).to(DEV)
    y.backward(gy)
    xd, Wd, bd, rd = (t.detach().double().cpu().requires_grad_() for t in (x, W, b, res))
    z = torch.nn.functional.linear(xd, Wd, bd)
    z = {"swish": lambda v: v * torch.sigmoid(v), "relu": torch.relu, "none": lambda v: v}[act](z) + rd
    z.backward(gy.double().cpu())
    assert rel_err(y, z) < TOL
    for got, want in ((x.grad, xd.grad), (W.grad, Wd.grad), (b.grad, bd.grad), (res.grad, rd.grad)):
        assert rel_err(got, want) < TOL


@pytest.mark.parametrize("rows,fin,fout,act", [(777, 128, 128, "relu"), (129, 128, 10, "none"), (2048, 128, 1, "none"),
                                               (1000, 256, 128, "swish"), (300, 128, 256, "none")])
def test_linear_act_tensor_core_path(rows, fin, fout, act):
    """tcgen05 Linear (fp16 hi/lo split, the default for 128/256-wide inputs): same call, fp32-grade results."""
    g = S._gen(rows + 1)
    x = torch.randn(rows, fin, generator=g).to(DEV)
    W = (torch.randn(fout, fin, generator=g) / fin ** 0.5).to(DEV)
    b = torch.randn(fout, generator=g).to(DEV)
    res = torch.randn(rows, fout, generator=g).to(DEV)
    old = MF.set_linear_tc(True)
    try:
        assert MF._tc_weight_images(W) is not None
        with torch.no_grad():
            y = MF.linear_act(x, W, b, act, res)
            y2 = MF.linear_act(x, W, b, act, res)          # cached weight images
    finally:
        MF.set_linear_tc(old)
    z = torch.nn.functional.linear(x.double().cpu(), W.double().cpu(), b.double().cpu())
    z = {"swish": lambda v: v * torch.sigmoid(v), "relu": torch.relu, "none": lambda v: v}[act](z) + res.double().cpu()
    assert rel_err(y, z) < 1e-6 and torch.equal(y, y2)


@pytest.mark.parametrize("rows,n_layers,n_out", [(1000, 5, 1), (129, 4, 128), (5, 1, 10), (2048, 4, 10), (257, 2, 128)])
def test_mlp_chain_one_launch(rows, n_layers, n_out):
    """A whole MLP (Linear(128,128)+ReLU ... Linear(128,n_out)) in one launch against an fp64 evaluation of the same layers;
    odd tile counts leave the second tile slot of the last pair empty."""
    g = S._gen(rows + n_layers)
    lins = [torch.nn.Linear(128, 128 if i + 1 < n_layers else n_out).to(DEV) for i in range(n_layers)]
    for l in lins:
        l.weight.data = (torch.randn(l.weight.shape, generator=g) / 128 ** 0.5).to(DEV)
        l.bias.data = (0.1 * torch.randn(l.bias.shape, generator=g)).to(DEV)
    x = torch.randn(rows, 128, generator=g).to(DEV)
    holder = torch.nn.Module()
    with torch.no_grad():
        y = MF.mlp_chain(x, lins, "relu", cache_owner=holder)
        y2 = MF.mlp_chain(x, lins, "relu", cache_owner=holder)       # cached weight images
        yr = MF.mlp_chain(-x, lins, "relu", in_act="relu")           # ReLU applied to the input on load
    assert y is not None and y.shape == (rows, n_out) and torch.equal(y, y2)
    z, zr = x.double().cpu(), torch.relu(-x.double().cpu())
    for i, l in enumerate(lins):
        W, b = l.weight.double().cpu(), l.bias.double().cpu()
        z, zr = z @ W.T + b, zr @ W.T + b
        if i + 1 < n_layers:
            z, zr = torch.relu(z), torch.relu(zr)
    assert rel_err(y, z) < 2e-6 and rel_err(yr, zr) < 2e-6


def test_linear_act_tensor_core_range_guard():
    """fp16 tops out at 65504: an input beyond 32768 is reported by the next call instead of turning into infinities quietly."""
    x = torch.full((256, 128), 1.0, device=DEV)
    x[3, 5] = 1e5
    W = torch.eye(128, device=DEV)
    b = torch.zeros(128, device=DEV)
    old = MF.set_linear_tc(True)
    try:
        with torch.no_grad():
            MF.linear_act(x, W, b, "none")
            torch.cuda.synchronize()
            with pytest.raises(RuntimeError, match="fp16 range"):
                MF.linear_act(torch.ones(256, 128, device=DEV), W, b, "none")
            y = MF.linear_act(torch.ones(256, 128, device=DEV), W, b, "none")       # the flag is cleared by the report
            assert torch.equal(y, torch.ones(256, 128, device=DEV))
    finally:
        MF.set_linear_tc(old)
    with torch.no_grad():
        MF.set_linear_tc(False)
        y = MF.linear_act(x, W, b, "none")                                           # the fp32 GEMM takes any magnitude
        MF.set_linear_tc(old)
    assert float(y[3, 5]) == 1e5


def test_layernorm_fwd_bwd():
    g = S._gen(9)
    x = (3 * torch.randn(1001, 128, generator=g) + 1).to(DEV).requires_grad_()
    gm = (1 + 0.1 * torch.randn(128, generator=g)).to(DEV).requires_grad_()
    bt = (0.1 * torch.randn(128, generator=g)).to(DEV).requires_grad_()
    y = MF.layer_norm(x, gm, bt)
    gy = torch.randn(1001, 128, generator=g).to(DEV)
    y.backward(gy)
    xd, gd, bd = (t.detach().double().cpu().requires_grad_() for t in (x, gm, bt))
    z = torch.nn.functional.layer_norm(xd, (128,), gd, bd, 1e-5)
    z.backward(gy.double().cpu())
    assert rel_err(y, z) < TOL
    for got, want in ((x.grad, xd.grad), (gm.grad, gd.grad), (bt.grad, bd.grad)):
        assert rel_err(got, want) < TOL


def test_instance_norm_ragged():
    g = S._gen(10)
    sizes = [300, 1, 0, 1000, 129]
    x = torch.randn(sum(sizes), 128, generator=g)
    batch = torch.cat([torch.full((n,), b) for b, n in enumerate(sizes)]).long()
    gptr = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int64, device=DEV)
    seg = MG.GraphSegments(gptr, len(sizes), max(sizes))
    y = MF.instance_norm(x.to(DEV), seg)
    assert rel_err(y, R.instance_norm(x.double(), batch)) < TOL


def _run_layer(c, tw, dp, seed):
    sd = S.seeded_state_dict(_layer_shapes(tw, dp), seed)
    layer = GNN_Layer(128, 128, 128, tw, 1, pos_dim=dp).to(DEV)
    layer.load_state_dict(sd, strict=True)
    x = c["x"].to(DEV).requires_grad_()
    u = c["u"].to(DEV).requires_grad_()
    pos = c["pos"].to(DEV).requires_grad_()
    y = layer(x, u, pos, c["variables"].to(DEV), c["edge_index"].to(DEV), c["batch"].to(DEV))
    return layer, sd, x, u, pos, y


@pytest.fixture(params=["fp32", "fp32_tc", "bf16"])
def precision(request):
    old = MF.set_precision(request.param)
    yield request.param
    MF.set_precision(old)


def _tol(precision):
    return 1e-2 if precision == "bf16" else TOL


def _gtol(precision):
    """Gradient tolerance.  North-star contract: messages, aggregates and predictions within 1e-5 (fp32) / 1e-2 (bf16).
    Forward quantities meet 1e-5 in both fp32 modes (measured <= 6e-7); gradients are checked at 1e-5 in the FFMA mode and
    at 5e-5 in the tensor-core mode, where every contraction carries the 2^-17 bf16 hi/lo split error (measured <= 1.6e-5,
    the largest on bias gradients, which are cancelling column sums)."""
    return {"fp32": TOL, "fp32_tc": 5e-5, "bf16": 3e-2}[precision]


def test_gnn_layer_golden(golden, precision):
    TOL, GT = _tol(precision), _gtol(precision)
    for name, c in golden("gnn_layer.pt").items():
        tw, dp = c["time_window"], c["pos"].shape[1]
        layer, sd, x, u, pos, y = _run_layer(c, tw, dp, c["seed"])
        assert rel_err(y, c["y"]) < TOL, name
        y.backward(c["grad_y"].to(DEV))
        assert rel_err(x.grad, c["grad_x"]) < GT, name
        assert rel_err(u.grad, c["grad_u"]) < GT, name
        assert rel_err(pos.grad, c["grad_pos"]) < GT, name
        for k, p in layer.named_parameters():
            assert rel_err(p.grad, c["grads"][k]) < GT, (name, k)


@pytest.mark.parametrize("B,N,r,trunc", [(4, 4096, 0.03, False), (2, 2048, 0.12, True)])
def test_gnn_layer_vs_oracle_fp64(B, N, r, trunc, precision):
    """config-2 shaped layer (64x64 irregular-uniform mesh); fp64 oracle as the arbiter."""
    TOL, GT = _tol(precision), _gtol(precision)
    g = S._gen(40 + B)
    pos = torch.rand(B * N, 2, generator=g)
    batch = torch.arange(B).repeat_interleave(N)
    ei = OG.radius_graph(pos, r, batch, loop=False, threads=8)
    if trunc:
        assert int(torch.bincount(ei[1]).max()) >= 32
    c = dict(x=torch.randn(B * N, 128, generator=g), u=torch.randn(B * N, 10, generator=g), pos=pos,
             variables=torch.rand(B * N, 1, generator=g), edge_index=ei, batch=batch)
    layer, sd, x, u, posd, y = _run_layer(c, 10, 2, 3)
    sd64 = R.cast_sd(sd, torch.float64)
    x64, u64, p64 = (t.double().requires_grad_() for t in (c["x"], c["u"], c["pos"]))
    y64 = R.gnn_layer(sd64, "", x64, u64, p64, c["variables"].double(), ei, batch)
    assert rel_err(y, y64) < TOL
    gy = torch.randn(y64.shape, generator=g)
    y.backward(gy.to(DEV))
    y64.backward(gy.double())
    assert rel_err(x.grad, x64.grad) < GT
    assert rel_err(u.grad, u64.grad) < GT
    assert rel_err(posd.grad, p64.grad) < GT
    # determinism: no atomics anywhere -> bit-identical reruns
    layer2, _, x2, u2, pos2, y2 = _run_layer(c, 10, 2, 3)
    assert torch.equal(y, y2)
    y2.backward(gy.to(DEV))
    assert torch.equal(x.grad, x2.grad)
    for (k, p), (_, q) in zip(layer.named_parameters(), layer2.named_parameters()):
        assert torch.equal(p.grad, q.grad), k


def test_gnn_layer_isolated_nodes_and_arbitrary_edge_index(precision):
    """edge_index the layer did not build itself: unsorted, duplicated, with isolated nodes."""
    TOL = _tol(precision)
    g = S._gen(50)
    N, E = 500, 3000
    ei = torch.randint(0, N - 50, (2, E), generator=g)       # the last 50 nodes receive nothing
    ei[1, :400] = 3                                          # one destination with 400 incoming edges
    c = dict(x=torch.randn(N, 128, generator=g), u=torch.randn(N, 10, generator=g), pos=torch.rand(N, 2, generator=g),
             variables=torch.rand(N, 1, generator=g), edge_index=ei, batch=torch.zeros(N, dtype=torch.long))
    layer, sd, x, u, pos, y = _run_layer(c, 10, 2, 4)
    y64 = R.gnn_layer(R.cast_sd(sd, torch.float64), "", c["x"].double(), c["u"].double(), c["pos"].double(),
                      c["variables"].double(), ei, c["batch"])
    assert rel_err(y, y64) < TOL


def test_mpnn_2d_golden(golden):
    c = golden("mpnn.pt")["mpnn_2d"]
    m = MPNN_2d(HParams(c["hparams"])).to(DEV)
    sd = S.seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, c["seed"])
    m.load_state_dict(sd, strict=True)
    b = {k: v.to(DEV) for k, v in c["batch"].items()}
    u = b["u"].permute(0, 2, 1)
    graph = m._build_graph(u[:, :10], b["t"], b["x"], [9, 9])
    assert torch.equal(graph.edge_index.cpu(), c["edge_index"])
    assert torch.equal(graph.pos.cpu(), c["pos"])
    with torch.no_grad():
        y = m.forward(graph, b["x"][0, -1], b["t"][0, -1], b["t"][0][1] - b["t"][0][0])
    assert rel_err(y, c["y"]) < TOL
    m.eval()
    with torch.no_grad():
        m.validation_step(b, 0)
    assert abs(float(m.logged["val_loss"]) - float(c["val_loss"])) <= 2e-5 * abs(float(c["val_loss"]))
    m.train()
    loss = m.training_step(b, 0)
    loss.backward()
    assert abs(float(loss) - float(c["train_loss"])) <= 2e-5 * abs(float(c["train_loss"]))
    assert rel_err(m.embedding_mlp[0].weight.grad, c["grad_embedding0"]) < 5e-5
    for k, p in m.named_parameters():
        assert abs(float(p.grad.norm()) - float(c["grad_norms"][k])) <= 1e-4 * float(c["grad_norms"][k]) + 1e-9, k


def test_mpnn_1d_golden(golden):
    c = golden("mpnn.pt")["mpnn_1d"]
    m = MPNN(HParams(c["hparams"])).to(DEV)
    sd = S.seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, c["seed"])
    m.load_state_dict(sd, strict=True)
    b = {k: v.to(DEV) for k, v in c["batch"].items()}
    u = b["u"].permute(0, 2, 1)
    x = b["x"].squeeze(-1)
    graph = m._build_graph(u[:, :25], b["t"], x, [0] * 4)
    assert torch.equal(graph.edge_index.cpu(), c["edge_index"])
    with torch.no_grad():
        y = m.forward(graph, x[0, -1], b["t"][0, -1], b["t"][0][1] - b["t"][0][0])
    assert rel_err(y, c["y"]) < TOL


@pytest.mark.parametrize("weight_decay", [0.0, 1e-2])
def test_flat_adam_matches_torch_adam(weight_decay):
    """FlatAdam (one launch over flat buffers) against torch.optim.Adam + StepLR as the reference configures them
    (models/mpnn_2d.py:205-213), fed identical gradients: the update rule itself."""
    from magnet_b200.optim import FlatAdam
    g = S._gen(5)
    shapes = [(128, 269), (128,), (7, 3), (1,), (128, 128), (5,)]
    pa = [torch.nn.Parameter(torch.randn(sh, generator=g).to(DEV)) for sh in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = torch.optim.Adam(pa, lr=1e-3, weight_decay=weight_decay)
    ob = FlatAdam(pb, lr=1e-3, weight_decay=weight_decay)
    sa, sb = (torch.optim.lr_scheduler.StepLR(o, step_size=2, gamma=0.3) for o in (oa, ob))
    for step in range(6):
        grads = [torch.randn(sh, generator=g).to(DEV) * 10.0 ** (step - 3) for sh in shapes]
        ob.zero_grad()
        for p, q, gr in zip(pa, pb, grads):
            p.grad = gr.clone()
            q.grad += gr                                  # accumulates into the flat buffer, as autograd does
        oa.step(), ob.step()
        sa.step(), sb.step()
    for p, q in zip(pa, pb):
        assert rel_err(q, p) < 1e-6
        assert ob.flat_param.data_ptr() <= q.data_ptr() < ob.flat_param.data_ptr() + 4 * ob.flat_param.numel()


def test_flat_adam_updates_reach_the_kernels():
    """The step writes through raw pointers: the version counters must move so that the cached packed weights follow."""
    from magnet_b200.optim import FlatAdam
    g = S._gen(6)
    B, N = 2, 300
    pos = torch.rand(B * N, 2, generator=g).to(DEV)
    batch = torch.arange(B).repeat_interleave(N).to(DEV)
    ei = MG.radius_graph(pos, 0.12, batch, loop=False)
    layer = GNN_Layer(128, 128, 128, 10, 1).to(DEV)
    layer.load_state_dict(S.seeded_state_dict(_layer_shapes(10, 2), 11))
    x = torch.randn(B * N, 128, generator=g).to(DEV)
    u = torch.randn(B * N, 10, generator=g).to(DEV)
    var = torch.rand(B * N, 1, generator=g).to(DEV)
    opt = FlatAdam(layer.parameters(), lr=1e-2)
    y0 = layer(x, u, pos, var, ei, batch)
    opt.zero_grad()
    (y0 * torch.randn(y0.shape, generator=g).to(DEV)).sum().backward()
    opt.step()
    with torch.no_grad():
        y1 = layer(x, u, pos, var, ei, batch)
        fresh = GNN_Layer(128, 128, 128, 10, 1).to(DEV)
        fresh.load_state_dict(layer.state_dict())
        y2 = fresh(x, u, pos, var, ei, batch)
    assert not torch.equal(y0, y1) and torch.equal(y1, y2)


def test_flat_adam_checkpoint_interchanges_with_torch_adam():
    """state_dict / load_state_dict use torch.optim.Adam's per-parameter layout: a run resumed from a
    checkpoint continues bit-for-bit like the uninterrupted run, and a torch-Adam checkpoint loads into FlatAdam."""
    from magnet_b200.optim import FlatAdam
    g = S._gen(15)
    shapes = [(128, 128), (128,), (9, 5), (3,)]
    base = [torch.randn(sh, generator=g).to(DEV) for sh in shapes]
    grads = [[torch.randn(sh, generator=g).to(DEV) for sh in shapes] for _ in range(6)]

    def feed(opt, params, gs):
        opt.zero_grad()
        for p, gr in zip(params, gs):
            if p.grad is None:
                p.grad = gr.clone()
            else:
                p.grad += gr
        opt.step()

    pa = [torch.nn.Parameter(b.clone()) for b in base]
    oa = FlatAdam(pa, lr=1e-3, weight_decay=1e-2)
    for k in range(6):
        feed(oa, pa, grads[k])
    # interrupted run: 3 steps, checkpoint, fresh optimizer + parameters, 3 more steps
    pb = [torch.nn.Parameter(b.clone()) for b in base]
    ob = FlatAdam(pb, lr=1e-3, weight_decay=1e-2)
    for k in range(3):
        feed(ob, pb, grads[k])
    ck_opt, ck_par = ob.state_dict(), [p.detach().clone() for p in pb]
    assert set(ck_opt) == {"state", "param_groups"} and set(ck_opt["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    pc = [torch.nn.Parameter(b.clone()) for b in ck_par]
    oc = FlatAdam(pc, lr=1e-3, weight_decay=1e-2)
    oc.load_state_dict(ck_opt)
    for k in range(3, 6):
        feed(oc, pc, grads[k])
    for p, q in zip(pa, pc):
        assert torch.equal(p, q)
    # the same checkpoint drives torch's Adam to the same parameters (update rule parity is test_flat_adam_matches_torch_adam)
    pd = [torch.nn.Parameter(b.clone()) for b in ck_par]
    od = torch.optim.Adam(pd, lr=1e-3, weight_decay=1e-2)
    od.load_state_dict(ck_opt)
    for k in range(3, 6):
        feed(od, pd, grads[k])
    for p, q in zip(pa, pd):
        assert rel_err(q, p) < 1e-6
    # and torch's checkpoint loads into FlatAdam
    pe = [torch.nn.Parameter(b.clone()) for b in ck_par]
    oe = FlatAdam(pe, lr=1e-3, weight_decay=1e-2)
    oe.load_state_dict(od.state_dict())
    assert oe.steps == 6


# ---- temporal-bundling decoder + Euler update in one launch per direction (csrc/decoder.cu) -----------------------------------
@pytest.mark.parametrize("tw,swish", [(10, True), (10, False), (16, True), (20, True), (25, True), (50, True)])
def test_bundling_decoder_vs_torch_fp64(tw, swish):
    """out = u[:, -1:] + cumsum(dt) * Conv1d(8,1,k2)(Swish(Conv1d(1,8,k1,stride)(h[:, None]))) (models/mpnn_2d.py:138-162,196-200;
    no Swish for the 1-D tw = 10 decoder, models/mpnn.py:139-142): values and every gradient against an fp64 torch evaluation,
    all five decoder shapes of the reference, a node count that is no multiple of the warps per block."""
    from magnet_b200.mpnn import _CONV
    k1, s1, k2 = _CONV[tw]
    g = torch.Generator().manual_seed(300 + tw)
    N = 1003
    conv1, conv2 = torch.nn.Conv1d(1, 8, k1, stride=s1), torch.nn.Conv1d(8, 1, k2)
    h = torch.randn(N, 128, generator=g)
    u = torch.randn(N, tw, generator=g)
    gy = torch.randn(N, tw, generator=g)
    dt = torch.tensor(0.0375)
    c1, c2 = conv1.double(), conv2.double()
    h64, u64 = h.double().requires_grad_(), u.double().requires_grad_()
    z = c1(h64[:, None])
    z = z * torch.sigmoid(z) if swish else z
    diff = c2(z).squeeze(1)
    dts = torch.cumsum(torch.ones(1, tw, dtype=torch.float64) * dt.double(), 1)
    want = u64[:, -1:].expand(-1, tw) + dts * diff
    want.backward(gy.double())
    d1, d2 = torch.nn.Conv1d(1, 8, k1, stride=s1).to(DEV), torch.nn.Conv1d(8, 1, k2).to(DEV)
    d1.load_state_dict({k: v.float() for k, v in c1.state_dict().items()})
    d2.load_state_dict({k: v.float() for k, v in c2.state_dict().items()})
    hg, ug = h.to(DEV).requires_grad_(), u.to(DEV).requires_grad_()
    got = MF.bundling_decoder(hg, ug, d1, d2, dt.to(DEV), swish)
    assert got.shape == (N, tw) and rel_err(got, want) < 1e-6
    got.backward(gy.to(DEV))
    assert rel_err(hg.grad, h64.grad) < 1e-5 and rel_err(ug.grad, u64.grad) < 1e-6
    for a, b in ((d1.weight, c1.weight), (d1.bias, c1.bias), (d2.weight, c2.weight), (d2.bias, c2.bias)):
        assert rel_err(a.grad, b.grad) < 1e-5, (tuple(a.shape), rel_err(a.grad, b.grad))


# ---- Linears with a handful of inputs (csrc/small_linear.cu): Encoder / embedding first layers -----------------------------------
@pytest.mark.parametrize("K,act", [(1, "none"), (12, "relu"), (13, "swish"), (16, "relu")])
def test_small_k_linear_vs_fp64(K, act):
    """y = act(x W^T + b) for K <= 16 inputs and 128 outputs (models/magnet_gnn.py:20-35 first Linears: 13 / 12 features;
    models/mpnn_2d.py:130-131): the streaming kernels behind mgb_linear_fwd / mgb_linear_bwd, values and all gradients."""
    g = torch.Generator().manual_seed(400 + K)
    rows = 4099
    x, W, b = torch.randn(rows, K, generator=g), torch.randn(128, K, generator=g) / K ** 0.5, torch.randn(128, generator=g)
    gy = torch.randn(rows, 128, generator=g)
    x64, W64, b64 = (v.double().requires_grad_() for v in (x, W, b))
    z = x64 @ W64.T + b64
    want = {"none": z, "relu": z.clamp_min(0), "swish": z * torch.sigmoid(z)}[act]
    want.backward(gy.double())
    xg, Wg, bg = (v.to(DEV).requires_grad_() for v in (x, W, b))
    got = MF.linear_act(xg, Wg, bg, act)
    assert rel_err(got, want) < 1e-6
    got.backward(gy.to(DEV))
    assert rel_err(xg.grad, x64.grad) < 1e-5 and rel_err(Wg.grad, W64.grad) < 1e-5 and rel_err(bg.grad, b64.grad) < 1e-5
    with torch.no_grad():                              # inference path: no saved pre-activation
        assert torch.equal(MF.linear_act(xg, Wg, bg, act), got)


def test_bundling_decoder_and_small_linear_on_empty_input():
    """Zero rows: the streaming kernels launch nothing and the parameter gradients are zero."""
    d1, d2 = torch.nn.Conv1d(1, 8, 16, stride=6).to(DEV), torch.nn.Conv1d(8, 1, 10).to(DEV)
    h = torch.zeros(0, 128, device=DEV, requires_grad=True)
    out = MF.bundling_decoder(h, torch.zeros(0, 10, device=DEV), d1, d2, torch.tensor(0.1, device=DEV), True)
    assert out.shape == (0, 10)
    out.sum().backward()
    assert float(d1.weight.grad.abs().sum()) == 0.0 and float(d2.bias.grad.abs().sum()) == 0.0
    W, b = torch.randn(128, 12, device=DEV, requires_grad=True), torch.randn(128, device=DEV, requires_grad=True)
    y = MF.linear_act(torch.zeros(0, 12, device=DEV, requires_grad=True), W, b, "relu")
    assert y.shape == (0, 128)
    y.sum().backward()
    assert float(W.grad.abs().sum()) == 0.0 and float(b.grad.abs().sum()) == 0.0


@pytest.mark.parametrize("rows,nv,precision,tol", [(1, 1, 1, TOL), (127, 1, 1, TOL), (128, 1, 1, TOL), (129, 1, 1, TOL), (777, 3, 1, TOL),
                                                   (148 * 128 + 65, 1, 1, TOL), (40000, 0, 1, TOL), (5000, 1, 2, 1e-2)])
def test_fused_node_update_vs_fp64(rows, nv, precision, tol):
    """GNN_Layer.update (models/mpnn_2d.py:81-90) without the InstanceNorm as one launch (mgb_gnn_node_update_fwd): both
    Linears, the Swish between them and the residual, against an fp64 evaluation; row counts around the 128-row tile, more
    tiles than CTAs, no / several equation variables.  Rerun: bit-identical."""
    from magnet_b200 import _lib
    L = _lib.lib()
    g = S._gen(rows + nv)
    x = torch.randn(rows, 128, generator=g)
    agg = torch.randn(rows, 128, generator=g) * 0.5
    var = torch.rand(rows, max(nv, 1), generator=g)[:, :nv].contiguous()
    W3 = torch.randn(128, 256 + nv, generator=g) / 16
    W4 = torch.randn(128, 128, generator=g) / 11
    b3, b4 = torch.randn(128, generator=g) * 0.1, torch.randn(128, generator=g) * 0.1
    d = {k: v.to(DEV) for k, v in dict(x=x, agg=agg, var=var, W3=W3, W4=W4, b3=b3, b4=b4).items()}
    packed = torch.empty(L.mgb_gnn_node_update_packed_floats(), device=DEV)
    _lib.check(L.mgb_gnn_node_update_pack(_lib.ptr(d["W3"]), _lib.ptr(d["W4"]), nv, _lib.ptr(packed), _lib.stream()), "pack")
    outs = []
    for _ in range(2):
        y1, y2, out = (torch.full((rows, 128), float("nan"), device=DEV) for _ in range(3))
        _lib.check(L.mgb_gnn_node_update_fwd(_lib.ptr(d["x"]), _lib.ptr(d["agg"]), _lib.ptr(d["var"]) if nv else None, nv, rows,
                                             _lib.ptr(packed), _lib.ptr(d["b3"]), _lib.ptr(d["b4"]), _lib.ptr(y1), _lib.ptr(y2),
                                             _lib.ptr(out), precision, _lib.stream()), "node_update_fwd")
        outs.append((y1, y2, out))
    sw = lambda v: v * torch.sigmoid(v)
    z1 = torch.cat([x, agg, var], 1).double() @ W3.double().T + b3.double()
    z2 = sw(z1) @ W4.double().T + b4.double()
    want = x.double() + sw(z2)
    for got, ref in zip(outs[0], (z1, z2, want)):
        assert rel_err(got, ref) < tol
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
