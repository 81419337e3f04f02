"""Parameter containers with the reference's building-block names (model/basic.py:4-60).

In the reference these classes *are* the computation (nn.Conv2d -> cuDNN).  Here they only own the
fp32 parameters, with the reference's shapes, initialisation order and state_dict keys; the arithmetic
is scheduled by the owning network (Generator / Discriminator) onto the sm_100a kernels, which fuse
bias, activation, residual scaling and PixelShuffle into the convolution epilogues.
"""
import torch
import torch.nn as nn


class Conv(nn.Conv2d):
    """Same-padded convolution parameters (model/basic.py:4-7)."""

    def __init__(self, in_planes, out_planes, kernel_size, stride=1, bias=True):
        nn.Conv2d.__init__(self, in_planes, out_planes, kernel_size, padding=kernel_size // 2, stride=stride,
                           bias=bias)

    def forward(self, x):
        from ..functional import conv2d_same
        return conv2d_same(x, self.weight, self.bias, stride=self.stride[0])


class MeanShift(nn.Conv2d):
    """1x1 colour-space affine (model/basic.py:9-17): W = I/std, b = sign*range*mean/std.

    The reference sets ``self.requires_grad = False`` on the *module*, which freezes nothing: weight and
    bias stay trainable and are updated by Adam.  That behaviour is kept.
    """

    def __init__(self, rgb_range, rgb_mean, rgb_std, sign=-1):
        nn.Conv2d.__init__(self, 3, 3, kernel_size=1)
        std = torch.tensor(rgb_std, dtype=torch.float32)
        mean = torch.tensor(rgb_mean, dtype=torch.float32)
        with torch.no_grad():
            self.weight.copy_((torch.eye(3) / std.view(3, 1)).view(3, 3, 1, 1))
            self.bias.copy_(sign * rgb_range * mean / std)
        self.requires_grad = False  # no-op attribute, as in the reference

    def forward(self, x):
        from ..functional import mean_shift
        return mean_shift(x, self.weight, self.bias)


class BasicBlock(nn.Sequential):
    """conv [+ BatchNorm] [+ activation] parameter group (model/basic.py:19-31)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, bias=False, bn=True, act=nn.ReLU(True),
                 sn=True):
        if sn:
            # model/basic.py:25 calls an undefined `spectral_norm`; the reference raises NameError here too.
            raise NameError("name 'spectral_norm' is not defined")
        layers = [Conv(in_channels, out_channels, kernel_size, stride, bias)]
        if bn:
            layers.append(nn.BatchNorm2d(out_channels))
        if act is not None:
            layers.append(act)
        nn.Sequential.__init__(self, *layers)


class ResBlock(nn.Module):
    """x + res_scale * conv(relu(conv(x))) parameter group (model/basic.py:33-52)."""

    def __init__(self, n_feats, kernel_size, bias=True, bn=False, act=nn.ReLU(True), res_scale=1):
        nn.Module.__init__(self)
        if bn:
            raise NotImplementedError("pesr_b200: ResBlock(bn=True) is not on the PESR path (model/pesr.py:13)")
        self.body = nn.Sequential(Conv(n_feats, n_feats, kernel_size, bias=bias), act,
                                  Conv(n_feats, n_feats, kernel_size, bias=bias))
        self.res_scale = res_scale

    def forward(self, x):
        from ..functional import conv2d_same
        c1, c2 = self.body[0], self.body[2]
        t = torch.relu(conv2d_same(x, c1.weight, c1.bias))
        return conv2d_same(t, c2.weight, c2.bias) * self.res_scale + x


class Upsampler(nn.Sequential):
    """conv -> PixelShuffle(2) -> conv -> PixelShuffle(2) -> conv(->3) parameter group (model/basic.py:54-60)."""

    def __init__(self, n_feats):
        nn.Sequential.__init__(self, Conv(n_feats, 4 * n_feats, 3), nn.PixelShuffle(2),
                               Conv(n_feats, 4 * n_feats, 3), nn.PixelShuffle(2), Conv(n_feats, 3, 3))
