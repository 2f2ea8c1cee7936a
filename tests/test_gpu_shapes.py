"""GPU parity at BASELINE.json's own shapes (VERDICT r1 "next" item 1): MAgNet[GNN] at C3's training shape (B=32, L=Nq=256,
concentrated meshes, r=0.08), one sample at test resolution 256 (L=Nq=32,768), C1 (1-D E1, time_slice 25, batch 16), and the
bf16 arithmetic mode on the MAgNet path.  Golden vectors: tests/golden/magnet_shapes.pt, produced by oracle/gen_golden.py from
the UNMODIFIED reference files (1-D: the same file with its three hard-coded widths parametrised).
fp32 contract 1e-5, bf16 contract 1e-2 (max-norm relative, conftest.rel_err); graphs bit-exact (SHA-256 of edge_index)."""
import hashlib

import pytest
import torch

from conftest import rel_err
from oracle.reference_loader import HParams
from magnet_b200 import functional as MF, synthetic as S
from magnet_b200.magnet_gnn import MAgNetGNN

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5


def _sha(t):
    return hashlib.sha256(t.cpu().contiguous().numpy().tobytes()).hexdigest()


def _model(c):
    m = MAgNetGNN(HParams(c["hparams"])).to(DEV).eval()
    m.load_state_dict(S.seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, c["seed"]), strict=True)
    return m


def _batch(c):
    return {k: v.to(DEV) for k, v in S.implicit_batch(**c["batch_args"]).items()}


def test_magnet_c3_training_shape(golden):
    c = golden("magnet_shapes.pt")["c3_train"]
    m, b = _model(c), _batch(c)
    inp, hr_last, tt = b["lr_frames"][:, :10], b["hr_points"][:, 9], b["t"][:, :20]
    with torch.no_grad():
        u = inp.permute(0, 3, 1, 2).reshape(32, 256, -1)
        _, ei1, _ = m._build_graph(u, b["coords_lr"], tt[:, :10])
        assert ei1.shape[1] == c["ei1_edges"] and _sha(ei1) == c["ei1_sha"]          # bit-exact stage-1 graph
        out_hr, out_lr, hr_points = m.forward(inp, b["coords_lr"], b["coords_hr"], tt, hr_last)
        allc = torch.cat([b["coords_lr"], b["coords_hr"]], 1)
        allf = torch.cat([u, hr_points.permute(0, 2, 1, 3).reshape(32, 256, -1)], 1)
        _, ei3, _ = m._build_graph(allf, allc, tt[:, :10])
        assert ei3.shape[1] == c["ei3_edges"] and _sha(ei3) == c["ei3_sha"]          # bit-exact stage-3 graph (LR u HR points)
        assert rel_err(hr_points, c["hr_points"]) < TOL, rel_err(hr_points, c["hr_points"])
        assert rel_err(out_hr, c["out_hr"]) < TOL and rel_err(out_lr, c["out_lr"]) < TOL
        m.validation_step(b, 0)
    assert abs(float(m.logged["val_loss"]) - float(c["val_loss"])) <= 2e-5 * abs(float(c["val_loss"]))
    assert abs(float(m.logged["val_mae_loss"]) - float(c["val_mae_loss"])) <= 2e-5 * abs(float(c["val_mae_loss"]))


def test_magnet_c3_test_resolution_256(golden):
    """One sample with L = Nq = 32,768: E = 1.05 M (stage 1) / ~2 M (stage 3) edges, in-degrees far above the 32-cap."""
    c = golden("magnet_shapes.pt")["c3_res256"]
    m, b = _model(c), _batch(c)
    inp, hr_last, tt = b["lr_frames"][:, :10], b["hr_points"][:, 9], b["t"][:, :20]
    s = c["stride"]
    with torch.no_grad():
        u = inp.permute(0, 3, 1, 2).reshape(1, 32768, -1)
        _, ei1, _ = m._build_graph(u, b["coords_lr"], tt[:, :10])
        assert ei1.shape[1] == c["ei1_edges"] and _sha(ei1) == c["ei1_sha"]
        out_hr, out_lr, hr_points = m.forward(inp, b["coords_lr"], b["coords_hr"], tt, hr_last)
    for name, got in (("out_hr", out_hr), ("out_lr", out_lr), ("hr_points", hr_points)):
        scale = c["sums"][name][2]                                  # max |reference| over the FULL tensor
        sub = got[:, :, ::s].cpu()
        err = float((sub.double() - c[name].double()).abs().max()) / scale
        assert err < TOL, (name, err)
        tot, atot, _ = c["sums"][name]                             # whole-tensor sums: every point takes part
        assert abs(float(got.double().sum()) - tot) <= 1e-5 * atot, name
        assert abs(float(got.double().abs().sum()) - atot) <= 1e-5 * atot, name


def test_magnet_c1_one_dimensional(golden):
    """hparams.dim = 1 (the F8 extension) at BASELINE configs[0]'s shape: graph + features bit-exact, predictions 1e-5,
    losses, gradient norms."""
    c = golden("magnet_shapes.pt")["c1_1d"]
    m, b = _model(c), _batch(c)
    inp, hr_last, tt = b["lr_frames"][:, :25], b["hr_points"][:, 24], b["t"][:, :50]
    with torch.no_grad():
        u = inp.permute(0, 3, 1, 2).reshape(16, 25, -1)
        nf, ei, ef = m._build_graph(u, b["coords_lr"], tt[:, :25])
        assert torch.equal(ei.cpu(), c["edge_index"])
        assert torch.equal(nf.cpu(), c["node_features"]) and torch.equal(ef.cpu(), c["edge_features"])
        out_hr, out_lr, hr_points = m.forward(inp, b["coords_lr"], b["coords_hr"], tt, hr_last)
        assert rel_err(hr_points, c["hr_points"]) < TOL
        assert rel_err(out_hr, c["out_hr"]) < TOL and rel_err(out_lr, c["out_lr"]) < TOL
        m.validation_step(b, 0)
    assert abs(float(m.logged["val_loss"]) - float(c["val_loss"])) <= 2e-5 * abs(float(c["val_loss"]))
    m.train()
    loss = m.training_step(b, 0)
    loss.backward()
    assert abs(float(loss.detach()) - float(c["train_loss"])) <= 2e-5 * abs(float(c["train_loss"]))
    bad = [(k, float(p.grad.norm()), float(c["grad_norms"][k])) for k, p in m.named_parameters()
           if not abs(float(p.grad.norm()) - float(c["grad_norms"][k])) <= 1e-3 * float(c["grad_norms"][k]) + 1e-9]
    assert not bad, bad[:8]       # L1 losses back-propagate sign(pred - target): norms agree to ~1e-4 (see test_gpu_magnet.py)


def test_magnet_bf16_mode(golden):
    """set_precision('bf16'): bf16 tensor-core operands on every 128-wide contraction of the MAgNet path (1e-2 contract)."""
    c = golden("magnet_shapes.pt")["c3_train"]
    m, b = _model(c), _batch(c)
    inp, hr_last, tt = b["lr_frames"][:, :10], b["hr_points"][:, 9], b["t"][:, :20]
    old = MF.set_precision("bf16")
    try:
        with torch.no_grad():
            out_hr, out_lr, hr_points = m.forward(inp, b["coords_lr"], b["coords_hr"], tt, hr_last)
    finally:
        MF.set_precision(old)
    assert rel_err(hr_points, c["hr_points"]) < 1e-2, rel_err(hr_points, c["hr_points"])
    assert rel_err(out_hr, c["out_hr"]) < 1e-2 and rel_err(out_lr, c["out_lr"]) < 1e-2
    assert rel_err(out_hr, c["out_hr"]) > 1e-7          # the mode really changed the arithmetic
