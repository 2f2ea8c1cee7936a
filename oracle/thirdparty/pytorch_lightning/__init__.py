"""Restated subset of pytorch_lightning==1.5.9 (requirements.txt:6): the LightningModule
surface the reference models use (save_hyperparameters, log, device). Oracle only."""
import torch
from torch import nn


class LightningModule(nn.Module):
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


def seed_everything(seed, workers=False):
    torch.manual_seed(seed)
    return seed
