"""Optional third-party surface.  When pytorch_lightning / torch_geometric are installed the
drop-in models derive from / return the real classes (so run.py's Trainer accepts them); when they
are not (this image), minimal stand-ins keep the modules importable and testable."""
import torch
from torch import nn

try:  # pragma: no cover - not installed in the build image
    from pytorch_lightning import LightningModule  # type: ignore
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    HAVE_LIGHTNING = False

    class LightningModule(nn.Module):
        """Just enough of pl.LightningModule for the models' own code paths."""

        def __init__(self, *args, **kwargs):
            super().__init__()
            self.logged = {}

        def save_hyperparameters(self, *args, **kwargs):
            pass

        def log(self, name, value, *args, **kwargs):
            self.logged[name] = value.detach() if torch.is_tensor(value) else value

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

try:  # pragma: no cover
    from torch_geometric.data import Data  # type: ignore
except Exception:  # noqa: BLE001

    class Data:
        """Attribute bag with the fields the reference reads (models/mpnn_2d.py:247-249)."""

        def __init__(self, x=None, edge_index=None, pos=None, batch=None, **kwargs):
            self.x, self.edge_index, self.pos, self.batch = x, edge_index, pos, batch
            for k, v in kwargs.items():
                setattr(self, k, v)
