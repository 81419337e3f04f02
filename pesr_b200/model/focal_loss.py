"""FocalLoss with the reference's constructor and call surface (model/focal_loss.py:4-13).

w = (1 - pt)^gamma, loss = mean(w * BCEWithLogits(x, t)).  The reference was written for torch 0.4,
whose Python BCE let the gradient flow through ``w``; torch >= 1.0 refuses that call.  The default
here reproduces the torch-0.4 gradient; ``detach_weight=True`` gives the detached variant."""
import torch.nn as nn

from ..losses import _GanLoss, _uniform_target


class FocalLoss(nn.Module):
    def __init__(self, gamma, detach_weight=False):
        nn.Module.__init__(self)
        self.gamma = gamma
        self.detach_weight = detach_weight

    def forward(self, x, t):
        # x may be an autograd expression such as pred_fake - pred_real (train.py:251); gradients flow
        # back through it unchanged.
        return _GanLoss.apply(x, None, 1.0, 0.0, _uniform_target(t), 2 if self.detach_weight else 1,
                              float(self.gamma))
