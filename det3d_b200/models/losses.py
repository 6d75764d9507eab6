"""Loss registry entries so that stock configs build unchanged.

Losses are training-only and out of scope: these placeholders accept the
config's keyword arguments and refuse to be called."""
from torch import nn

from .registry import LOSSES


class _TrainingOnly(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.cfg = kwargs

    def forward(self, *args, **kwargs):
        raise NotImplementedError("%s is training-only; det3d_b200 implements the inference hot path"
                                  % type(self).__name__)


for _name in ("SigmoidFocalLoss", "WeightedSmoothL1Loss", "WeightedSoftmaxClassificationLoss",
              "CrossEntropyLoss", "SmoothL1Loss", "WeightedL2LocalizationLoss", "SoftmaxFocalLoss",
              "WeightedSigmoidClassificationLoss", "BootstrappedSigmoidClassificationLoss", "GHMCLoss",
              "GHMRLoss", "MSELoss", "FocalLoss", "BalancedL1Loss", "IoULoss", "AccuracyLoss"):
    LOSSES.register_module(type(_name, (_TrainingOnly,), {}))
