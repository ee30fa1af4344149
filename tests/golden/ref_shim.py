"""Import shim for the UNMODIFIED reference at /root/reference (container only).

Test infrastructure: used solely by tests/golden/make_golden.py to generate the
committed fixtures.  Nothing on the GPU box imports this (the reference does not
travel).  Recipe follows SURVEY.md §8(c): a stub `lightning` package plus
MagicMocks for the data/plot/logging packages that are not installed.
"""
import sys
import types
from unittest.mock import MagicMock

import torch
from torch import nn

REFERENCE_ROOT = "/root/reference"


def install():
    if "lightning" not in sys.modules:
        lightning = types.ModuleType("lightning")

        class LightningModule(nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass

            @property
            def device(self):
                try:
                    return next(self.parameters()).device
                except StopIteration:
                    return torch.device("cpu")

        lightning.LightningModule = LightningModule
        lightning.Trainer = MagicMock()
        sys.modules["lightning"] = lightning
        for sub in ("lightning.pytorch", "lightning.pytorch.callbacks",
                    "lightning.pytorch.loggers", "lightning.pytorch.utilities"):
            sys.modules[sub] = MagicMock()
    for name in ("awkward", "fastjet", "vector", "uproot", "h5py", "seaborn",
                 "matplotlib", "matplotlib.pyplot", "matplotlib.lines",
                 "mlflow", "comet_ml", "pytorch_lightning", "wandb"):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = MagicMock()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # import order matters: models before the dataloader
    import multimodal_particles.models  # noqa: F401
    return sys.modules["multimodal_particles"]
