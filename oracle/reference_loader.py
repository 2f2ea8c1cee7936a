"""ORACLE (test infrastructure only) — import the reference's UNMODIFIED model files.

Works only where /root/reference exists (the build container); the GPU box has no
reference tree, so nothing that runs there (``-m gpu`` tests, smoke(), bench.py) may call
this.  It is used by ``oracle/gen_golden.py`` to produce ``tests/golden/*.pt`` and by the
CPU tests that validate ``oracle/restatement.py`` against the real files.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MAGNET_REFERENCE_ROOT", "/root/reference")
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # bytecode of the same files (oracle/stage_ref.py)
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "thirdparty")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "mpnn_2d.py"))


_EXT = ".pyc.bin"


def staged() -> bool:
    """oracle/_ref/ holds the byte-compiled reference files (travels to the GPU box, where /root/reference does not exist)."""
    return os.path.isfile(os.path.join(STAGED_ROOT, "models", "mpnn_2d" + _EXT))


class _StagedFinder:
    """meta-path finder for the staged bytecode: `models.mpnn_2d` -> oracle/_ref/models/mpnn_2d.pyc.bin (sourceless)."""

    @staticmethod
    def find_spec(fullname, path=None, target=None):
        import importlib.machinery
        import importlib.util
        rel = os.path.join(STAGED_ROOT, *fullname.split("."))
        if os.path.isfile(rel + _EXT):
            loader = importlib.machinery.SourcelessFileLoader(fullname, rel + _EXT)
            return importlib.util.spec_from_file_location(fullname, rel + _EXT, loader=loader)
        if os.path.isdir(rel) and fullname.split(".")[0] == "models":
            spec = importlib.machinery.ModuleSpec(fullname, None, is_package=True)
            spec.submodule_search_locations = [rel]
            return spec
        return None


def load(prefer_staged: bool = False):
    """Returns a namespace with the reference modules: .mpnn, .mpnn_2d, .magnet_gnn, .mlp (+ .root, .kind).
    Source tree when present (this container), else the staged bytecode of the very same files (GPU box)."""
    if available() and not (prefer_staged and staged()):
        root = REFERENCE_ROOT
    elif staged():
        root = STAGED_ROOT
    else:
        raise RuntimeError(f"reference not found: neither {REFERENCE_ROOT} nor staged bytecode under {STAGED_ROOT} "
                           "(python -m oracle.stage_ref)")
    for p in (_REPO, _STUBS) + ((root,) if root == REFERENCE_ROOT else ()):
        if p not in sys.path:
            sys.path.insert(0, p) if p != root else sys.path.append(p)
    if root == STAGED_ROOT and not any(f is _StagedFinder for f in sys.meta_path):
        sys.meta_path.append(_StagedFinder)
    ns = types.SimpleNamespace(root=root, kind="source" if root == REFERENCE_ROOT else "bytecode")
    ns.mlp = importlib.import_module("models.backbones.mlp")
    ns.mpnn = importlib.import_module("models.mpnn")
    ns.mpnn_2d = importlib.import_module("models.mpnn_2d")
    ns.magnet_gnn = importlib.import_module("models.magnet_gnn")
    for m in (ns.mpnn, ns.mpnn_2d, ns.magnet_gnn):
        assert m.__file__.startswith(root), m.__file__
    return ns


class HParams(dict):
    """Stand-in for the hydra `cfg.model.params` object (attribute access; run.py:50)."""
    __getattr__ = dict.__getitem__


def mpnn_2d_hparams(**over):
    # configs/model/mpnn_2d.yaml:4-14
    hp = dict(hidden_features=128, hidden_layer=5, time_window=10, teacher_forcing=False, neighbors=4,
              factor=0.3, step_size=50, loss="l1", lr=1e-3, weight_decay=0)
    hp.update(over)
    return HParams(hp)


def mpnn_hparams(**over):
    # configs/model/mpnn.yaml
    hp = dict(hidden_features=128, hidden_layer=5, time_window=16, teacher_forcing=False, neighbors=3,
              factor=0.3, step_size=50, loss="l1", lr=1e-3, weight_decay=0)
    hp.update(over)
    return HParams(hp)


def magnet_gnn_hparams(**over):
    # configs/model/magnet_gnn.yaml:4-20 (scripts override time_slice=10 for 2-D)
    hp = dict(time_slice=10, latent_dim=128, num_message_passing_steps=5, mlp_layers=4, mlp_hidden=128,
              radius=0.08, n_chan=128, teacher_forcing=True, codec_neighbors=4, noise=0,
              interpolation="area", factor=0.3, step_size=50, loss="l1", lr=1e-3, weight_decay=0)
    hp.update(over)
    return HParams(hp)
