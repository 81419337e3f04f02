"""Drop-in replacement of the reference's ``model`` package (model/__init__.py:1-3).

``from pesr_b200.model import *`` yields the same public names the reference's star-import gives
train.py / test.py: Generator, Discriminator, VGG, FocalLoss and the building blocks, plus the
``nn`` / ``torch`` / ``F`` / ``models`` names the reference leaks into its callers.
"""
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import torch.nn.functional as F  # noqa: F401

from .basic import BasicBlock, Conv, MeanShift, ResBlock, Upsampler  # noqa: F401
from .focal_loss import FocalLoss  # noqa: F401
from .pesr import Discriminator, Generator  # noqa: F401
from .vgg import VGG  # noqa: F401

try:  # torchvision is only needed for the name the reference leaks (`models`); VGG itself does not use it
    import torchvision.models as models  # noqa: F401
except Exception:  # pragma: no cover
    models = None
