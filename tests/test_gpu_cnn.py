"""GPU parity of the MAgNet[CNN]_2d drop-in (models/magnet_cnn_2d.py:142-478 reuses the graph stage of MAgNet[GNN]):
golden vectors from the UNMODIFIED reference file (oracle/gen_golden.py `cnn`, tests/golden/magnet_cnn_2d.pt)."""
import pytest
import torch

from conftest import rel_err
from oracle.reference_loader import HParams
from magnet_b200 import synthetic as S
from magnet_b200.magnet_cnn import MAgNetCNN_2d

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5


def _model(c):
    m = MAgNetCNN_2d(HParams(c["hparams"])).to(DEV)
    m.load_state_dict(S.seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, c["seed"]), strict=True)
    return m


@pytest.fixture(autouse=True)
def _exact_fp32_convolutions():
    """The EDSR trunk is ordinary cuDNN (outside the hot path): PyTorch's default lets cuDNN use TF32 for fp32 convolutions
    (1.5e-4 against the reference's CPU run); the parity test switches that off so that the 1e-5 contract is visible."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def test_magnet_cnn_forward_and_losses_golden(golden):
    c = golden("magnet_cnn_2d.pt")
    m = _model(c).eval()
    b = {k: v.to(DEV) for k, v in c["batch"].items()}
    inp, hr_last, tt = b["lr_frames"][:, :10], b["hr_points"][:, 9], b["t"][:, :20]
    with torch.no_grad():
        feat = m.feature_encoding(inp)
        assert rel_err(feat, c["feat"]) < TOL                      # EDSR trunk (cuDNN here, CPU convolutions in the reference)
        z = m.continuous_decoder(inp, feat, b["cells"], b["coords"], tt)
        assert rel_err(z, c["z"]) < TOL, rel_err(z, c["z"])
        out_hr, out_lr, hr_points = m.forward(inp, b["coords"], b["cells"], tt, hr_last)
        assert out_lr.shape == c["out_lr"].shape
        assert rel_err(hr_points, c["hr_points"]) < TOL and rel_err(out_hr, c["out_hr"]) < TOL and rel_err(out_lr, c["out_lr"]) < TOL
        m.validation_step(b, 0)
    assert abs(float(m.logged["val_loss"]) - float(c["val_loss"])) <= 2e-5 * abs(float(c["val_loss"]))
    m.train()
    loss = m.training_step(b, 0)            # graph stage: fused InteractionNetwork forward + recompute backward
    loss.backward()
    assert abs(float(loss.detach()) - float(c["train_loss"])) <= 2e-5 * abs(float(c["train_loss"]))
    bad = [(k, float(p.grad.norm()), float(c["grad_norms"][k])) for k, p in m.named_parameters()
           if not abs(float(p.grad.norm()) - float(c["grad_norms"][k])) <= 2e-3 * float(c["grad_norms"][k]) + 1e-9]
    assert not bad, bad[:8]      # L1 losses back-propagate sign(pred - target): norms agree to ~1e-3 (see test_gpu_magnet.py)
