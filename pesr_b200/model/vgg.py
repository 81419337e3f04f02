"""VGG19 perceptual-feature extractor with the reference's surface (model/vgg.py:5-28):
``VGG()(sr, hr) -> (features(sr), features(hr) detached)``, state_dict keys ``vgg.{idx}.*`` and
``sub_mean.*``."""
import torch
import torch.nn as nn

from .basic import MeanShift

_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512]


def _vgg19_features_35():
    """The first 35 modules of torchvision's vgg19().features (16 convs, 15 ReLUs, 4 max-pools), built
    directly with torchvision's initialisation (kaiming_normal fan_out / zero bias)."""
    layers, cin = [], 3
    for v in _CFG:
        if v == 'M':
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            conv = nn.Conv2d(cin, v, kernel_size=3, padding=1)
            nn.init.kaiming_normal_(conv.weight, mode='fan_out', nonlinearity='relu')
            nn.init.constant_(conv.bias, 0)
            layers += [conv, nn.ReLU(inplace=True)]
            cin = v
    return layers[:35]


class _VGGFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, sr, hr):
        need = ctx.needs_input_grad[1]
        f_sr, f_hr, state = engine.forward(sr, hr, save=need)
        ctx.engine, ctx.state = engine, state
        ctx.mark_non_differentiable(f_hr)
        return f_sr, f_hr

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dfeat, _unused):
        if ctx.state is None:
            raise RuntimeError("pesr_b200.VGG: backward called twice through one forward")
        dsr = ctx.engine.backward(ctx.state, dfeat)
        ctx.state = None
        return None, dsr, None


class VGG(nn.Module):
    def __init__(self, pretrained=True, dtype=torch.float16, split_precision=False):
        nn.Module.__init__(self)
        self.vgg = nn.Sequential(*_vgg19_features_35())   # conv5_4, before its ReLU (model/vgg.py:10)
        if pretrained:
            # model/vgg.py:8 `models.vgg19(pretrained=True)`: needs torchvision's ImageNet checkpoint
            try:
                import torchvision.models as models
                src = models.vgg19(weights=models.VGG19_Weights.IMAGENET1K_V1).features
                self.vgg.load_state_dict({k: v for k, v in src.state_dict().items() if int(k.split('.')[0]) < 35})
            except Exception as e:  # no network / no cached checkpoint
                raise RuntimeError("pesr_b200.VGG: could not load torchvision's pretrained VGG19 weights "
                                   f"({e}); pass pretrained=False for random-init weights") from e
        rgb_range = 255
        vgg_mean = (0.485, 0.456, 0.406)
        vgg_std = (0.229 * rgb_range, 0.224 * rgb_range, 0.225 * rgb_range)
        self.sub_mean = MeanShift(rgb_range, vgg_mean, vgg_std)
        # The reference writes `self.vgg.requires_grad = False`, a no-op; no optimiser owns these weights
        # (train.py:123-126), so they are frozen in effect.  Freeze them for real: results are identical and the
        # unused weight gradients are not computed.
        for p in self.parameters():
            p.requires_grad_(False)
        self._compute_dtype = dtype
        # split_precision: fp16 hi+lo operands, three tensor-core passes per conv, fp32 activations (engine_v_split.py)
        self._split = bool(split_precision)
        self._engine = None

    def engine(self):
        if self._engine is None:
            if self._split:
                from ..engine_v_split import SplitVGGEngine
                self._engine = SplitVGGEngine(self)
            else:
                from ..engine_v import VGGEngine
                self._engine = VGGEngine(self, self._compute_dtype)
        return self._engine

    def forward(self, sr, hr):
        if not sr.is_cuda:
            raise RuntimeError(f"pesr_b200.VGG: input is on {sr.device}; the B200 path has no CPU fallback")
        if not torch.is_grad_enabled():
            f_sr, f_hr, _ = self.engine().forward(sr, hr, save=False)
            return f_sr, f_hr
        return _VGGFn.apply(self.engine(), sr, hr)
