"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the
synthetic generators are deterministic, the drop-in modules keep the reference's state_dict
layout, and the product never imports the oracle."""
import ctypes
import os
import re

import pytest
import torch

from magnet_b200 import _lib, synthetic as S
from oracle.reference_loader import HParams, magnet_gnn_hparams, mpnn_2d_hparams, mpnn_hparams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    protos = _lib.parse_header()
    assert len(protos) >= 25
    l = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(l, name), name
    assert _lib.lib().mgb_abi_version() == 1
    text = open(_lib.HEADER).read()
    declared = set(re.findall(r"\b(mgb_\w+)\s*\(", text))
    assert declared == set(protos), declared ^ set(protos)


def test_error_convention_without_gpu():
    L = _lib.lib()
    # bad argument -> -1 and a message; no compute is attempted
    rc = L.mgb_knn(None, 0, None, 0, 3, None, None, 1, 4, None, None, None, 0, None)
    assert rc == -1 and "1-D and 2-D" in _lib.last_error()
    rc = L.mgb_radius_graph_search(None, 10, 5, None, 1, 0.1, 32, 0, None, None, None, None, 0, None)
    assert rc == -1


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "magnet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src or f.endswith(".md"), f


def test_cpu_tensors_fail_loudly():
    from magnet_b200 import graph as MG
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MG.radius_graph(torch.rand(10, 2), 0.1)


def test_new_entry_points_reject_bad_arguments_without_gpu():
    L = _lib.lib()
    # tensor-core Linear: only 128/256 inputs and <= 256 outputs, at most two weight tiles (128 x 128 floats per tile: two 16-bit
    # images); the shapes mgb_linear_tc_bwd covers (128 outputs) carry the W^T blocks in tensor-memory order behind the images
    assert L.mgb_linear_tc_packed_floats(128, 128) == 2 * 128 * 128 and L.mgb_linear_tc_packed_floats(128, 1) == 128 * 128
    assert L.mgb_linear_tc_packed_floats(256, 128) == 4 * 128 * 128 and L.mgb_linear_tc_packed_floats(128, 256) == 2 * 128 * 128
    assert L.mgb_gnn_node_update_packed_floats() == 3 * 128 * 128 + 128 * 4 and L.mgb_f16_range_check() in (0, 1)
    assert L.mgb_linear_tc_packed_floats(13, 128) == 0 and L.mgb_linear_tc_packed_floats(256, 256) == 0
    rc = L.mgb_linear_tc_fwd(None, 10, 13, 128, None, None, 0, None, None, None, 3, None)
    assert rc == -1 and "unsupported shape" in _lib.last_error()
    rc = L.mgb_linear_tc_fwd(None, 10, 128, 128, None, None, 0, None, None, None, 7, None)
    assert rc == -1 and "precision" in _lib.last_error()
    # flat Adam: the step counter starts at 1
    rc = L.mgb_adam_step(None, None, None, None, 16, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0, 1.0, None)
    assert rc == -1 and "step >= 1" in _lib.last_error()


def test_flat_adam_and_linear_paths_have_no_cpu_fallback():
    from magnet_b200 import functional as MF
    from magnet_b200.optim import FlatAdam
    with pytest.raises(RuntimeError, match="CUDA"):
        FlatAdam([torch.nn.Parameter(torch.zeros(4))])
    with pytest.raises(RuntimeError):
        MF.linear_act(torch.zeros(4, 128), torch.zeros(128, 128), torch.zeros(128))
    assert MF.set_linear_tc(False) in (True, False)
    MF.set_linear_tc(True)


def test_synthetic_is_deterministic_and_shaped_like_the_reference_batches():
    a = S.graph_batch(B=3, N=64, nt=20, seed=4)
    b = S.graph_batch(B=3, N=64, nt=20, seed=4)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert a["u"].shape == (3, 64, 20) and a["x"].shape == (3, 64, 2) and a["t"].shape == (3, 20)
    c = S.implicit_batch(B=2, L=32, Nq=16, nt=20, seed=1)
    assert c["lr_frames"].shape == (2, 20, 1, 32) and c["hr_points"].shape == (2, 20, 16, 1)
    assert c["coords_lr"].shape == (2, 32, 2) and c["coords_hr"].shape == (2, 16, 2)
    assert float(torch.cat([c["coords_lr"], c["coords_hr"]], 1).abs().max()) <= 1.0 + 1e-6


def test_state_dict_layout_matches_reference_keys(golden):
    from magnet_b200.mpnn import MPNN, MPNN_2d
    m = MPNN_2d(mpnn_2d_hparams())
    keys = set(m.state_dict())
    assert sum(p.numel() for p in m.parameters()) == 521_561          # SURVEY §2.2 probe count
    for l in range(5):
        for net in ("message_net_1", "message_net_2", "update_net_1", "update_net_2"):
            assert f"gnn_layers.{l}.{net}.0.weight" in keys and f"gnn_layers.{l}.{net}.0.bias" in keys
    assert {"embedding_mlp.0.weight", "embedding_mlp.2.bias", "output_mlp.0.weight", "output_mlp.2.bias"} <= keys
    assert m.gnn_layers[0].message_net_1[0].weight.shape == (128, 269)
    m1 = MPNN(mpnn_hparams(time_window=25))
    assert m1.gnn_layers[0].message_net_1[0].weight.shape == (128, 283)


@pytest.mark.reference
def test_state_dicts_load_strictly_both_ways():
    from oracle import reference_loader as rl
    from magnet_b200.mpnn import MPNN, MPNN_2d
    ref = rl.load()
    for ours, theirs, hp in ((MPNN_2d, ref.mpnn_2d.MPNN_2d, mpnn_2d_hparams()),
                             (MPNN, ref.mpnn.MPNN, mpnn_hparams(time_window=25)),
                             (MPNN, ref.mpnn.MPNN, mpnn_hparams(time_window=10))):
        a, b = ours(hp), theirs(hp)
        a.load_state_dict(b.state_dict(), strict=True)
        b.load_state_dict(a.state_dict(), strict=True)


@pytest.mark.reference
def test_magnet_cnn_state_dict_loads_strictly_both_ways(golden):
    """MAgNet[CNN]_2d drop-in: same parameter names and shapes as models/magnet_cnn_2d.py."""
    import importlib
    from oracle import reference_loader as rl
    from magnet_b200.magnet_cnn import MAgNetCNN_2d
    rl.load()
    theirs = importlib.import_module("models.magnet_cnn_2d").MAgNetCNN_2d
    hp = rl.HParams(golden("magnet_cnn_2d.pt")["hparams"])
    a, b = MAgNetCNN_2d(hp), theirs(hp)
    a.load_state_dict(b.state_dict(), strict=True)
    b.load_state_dict(a.state_dict(), strict=True)


def test_node_update_entry_rejects_bad_arguments_without_gpu():
    """mgb_gnn_node_update_fwd / pack (GNN_Layer.update, models/mpnn_2d.py:81-90): argument checks run before any launch."""
    L = _lib.lib()
    rc = L.mgb_gnn_node_update_fwd(None, None, None, 7, 100, None, None, None, None, None, None, 1, None)
    assert rc == -1 and "var columns" in _lib.last_error()
    rc = L.mgb_gnn_node_update_fwd(None, None, None, 1, 100, None, None, None, None, None, None, 5, None)
    assert rc == -1 and "precision" in _lib.last_error()
    rc = L.mgb_gnn_node_update_fwd(None, None, None, 1, 100, None, None, None, None, None, None, 1, None)
    assert rc == -1 and "var is NULL" in _lib.last_error()
    assert L.mgb_gnn_node_update_fwd(None, None, None, 0, 0, None, None, None, None, None, None, 1, None) == 0      # no rows: nothing to do
    rc = L.mgb_gnn_node_update_pack(None, None, 9, None, None)
    assert rc == -1 and "var columns" in _lib.last_error()
