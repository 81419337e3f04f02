"""Generator and Discriminator with the reference's constructor, call and state_dict surface
(model/pesr.py:3-81), executed by the sm_100a kernel schedules in pesr_b200.engine_g / engine_d."""
import torch
import torch.nn as nn

from .basic import BasicBlock, Conv, MeanShift, ResBlock, Upsampler  # noqa: F401


def _require_cuda(x, who):
    if not x.is_cuda:
        raise RuntimeError(f"pesr_b200.{who}: input is on {x.device}; the B200 path has no CPU fallback "
                           "(use the oracle under oracle/ for CPU reference results)")


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, *params):
        need = any(ctx.needs_input_grad)
        sr, state = engine.forward(x, train=need)
        ctx.engine, ctx.state = engine, state
        return sr

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dsr):
        eng = ctx.engine
        if ctx.state is None:
            raise RuntimeError("pesr_b200.Generator: backward called twice through one forward (activations are "
                               "released after the first backward; retain_graph is not supported)")
        grads, dlr = eng.backward(ctx.state, dsr, need_input_grad=ctx.needs_input_grad[1])
        ctx.state = None        # hands the plan (the saved activations) back to the engine's pool
        out = [None, dlr]
        for i, p in enumerate(eng.param_list):
            out.append(grads.get(p) if ctx.needs_input_grad[2 + i] else None)
        return tuple(out)


class Generator(nn.Module):
    """EDSR-style x4 SR network (model/pesr.py:3-38): sub_mean, embed, `depth` ResBlocks + conv with a
    global skip, Upsampler, add_mean.  ``opt`` keys: depth, num_channels, res_scale."""

    def __init__(self, opt, dtype=torch.float16, split_precision=False):
        nn.Module.__init__(self)
        self.n_resblock = opt['depth']
        self.n_feats = opt['num_channels']
        self.res_scale = opt['res_scale']
        if self.n_feats % 64 != 0:
            raise ValueError("pesr_b200.Generator: num_channels must be a multiple of 64 (tensor-core K block)")
        rgb_mean = (0.4488, 0.4371, 0.4040)  # DIV2K800, model/pesr.py:13
        rgb_std = (1.0, 1.0, 1.0)
        act = nn.ReLU(True)
        # construction order == the reference's, so a given torch seed yields the same initial weights
        blocks = [ResBlock(self.n_feats, 3, act=act, res_scale=self.res_scale) for _ in range(self.n_resblock)]
        blocks.append(Conv(self.n_feats, self.n_feats, 3))
        self.sub_mean = MeanShift(255, rgb_mean, rgb_std)
        self.embed = Conv(3, self.n_feats, 3)
        self.body = nn.Sequential(*blocks)
        self.upsample = Upsampler(self.n_feats)
        self.add_mean = MeanShift(255, rgb_mean, rgb_std, 1)
        self._compute_dtype = dtype
        # split_precision: fp16 hi+lo operands, three tensor-core passes per convolution (pesr_b200/engine_g_split.py):
        # fp32-grade forward AND gradients at ~3.5x the cost; the headline path is the 16-bit schedule
        self._split = bool(split_precision)
        self._engine = None

    def engine(self):
        if self._engine is None:
            if self._split:
                from ..engine_g_split import SplitGeneratorEngine
                self._engine = SplitGeneratorEngine(self)
            else:
                from ..engine_g import GeneratorEngine
                self._engine = GeneratorEngine(self, self._compute_dtype)
            self._engine.param_list = list(self.parameters())
        return self._engine

    def forward(self, x):
        _require_cuda(x, "Generator")
        eng = self.engine()
        if not torch.is_grad_enabled():
            # inference (test.py:106, the validation loop of train.py:281-295): no activations are kept.  (Inside an
            # autograd.Function ctx.needs_input_grad is True for the parameters even under no_grad, so the decision is
            # taken here.)
            return eng.forward(x, train=False)[0]
        return _GeneratorFn.apply(eng, x, *eng.param_list)


class _DiscriminatorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, *params):
        logits, state = engine.forward(x, save=any(ctx.needs_input_grad))
        ctx.engine, ctx.state = engine, state
        return logits

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dlogits):
        eng = ctx.engine
        if ctx.state is None:
            raise RuntimeError("pesr_b200.Discriminator: backward called twice through one forward")
        need_params = any(ctx.needs_input_grad[2:])
        grads, dxs = eng.backward(ctx.state, dlogits, need_param_grads=need_params,
                                  need_input_grad=ctx.needs_input_grad[1])
        ctx.state = None
        out = [None, dxs[0]]
        for i, p in enumerate(eng.param_list):
            out.append(grads.get(p) if ctx.needs_input_grad[2 + i] else None)
        return tuple(out)


class _DiscriminatorPairFn(torch.autograd.Function):
    """D(a), D(b) as two separate train-mode calls (separate BatchNorm statistics, running statistics updated in call
    order, exactly as two module calls) executed as ONE batch of 2N images: every layer is one launch instead of two
    (the Discriminator's launches are latency-bound at batch 16), the parameter gradients of both calls come out of one
    backward pass already summed, and data-parallel training all-reduces the Discriminator's 80 M gradients once per
    optimiser step (train.py:205-216 back-propagates both logits before optim_D.step())."""

    @staticmethod
    def forward(ctx, engine, a, b, *params):
        (la, lb), state = engine.forward([a, b], save=any(ctx.needs_input_grad))
        ctx.engine, ctx.state = engine, state
        return la, lb

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dla, dlb):
        eng = ctx.engine
        if ctx.state is None:
            raise RuntimeError("pesr_b200.Discriminator: backward called twice through one forward")
        need_params = any(ctx.needs_input_grad[3:])
        grads, dxs = eng.backward(ctx.state, [dla, dlb], need_param_grads=need_params,
                                  need_input_grad=[ctx.needs_input_grad[1], ctx.needs_input_grad[2]])
        ctx.state = None
        out = [None, dxs[0], dxs[1]]
        for i, p in enumerate(eng.param_list):
            out.append(grads.get(p) if ctx.needs_input_grad[3 + i] else None)
        return tuple(out)


class Discriminator(nn.Module):
    """model/pesr.py:40-81: eight conv-BN-LeakyReLU blocks (64 -> 512 channels, stride 1/2 alternating), NCHW
    flatten, Linear(512*(patch/4)^2 -> 1024), LeakyReLU, Linear(1024 -> 1).  ``opt`` keys: patch_size,
    spectral_norm (True raises NameError exactly as the reference does, model/basic.py:25)."""

    def __init__(self, opt, dtype=torch.float16, split_precision=False):
        nn.Module.__init__(self)
        act = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        self.patch_size = opt['patch_size']
        sn = opt['spectral_norm']
        in_channels, out_channels, depth = 3, 64, 7
        blocks = [BasicBlock(3, out_channels, 3, bn=True, act=act, sn=sn)]
        for i in range(depth):
            in_channels = out_channels
            if i % 2 == 1:
                stride = 1
                out_channels *= 2
            else:
                stride = 2
            blocks.append(BasicBlock(in_channels, out_channels, 3, stride=stride, bn=True, act=act, sn=sn))
        self.features = nn.Sequential(*blocks)
        side = self.patch_size * 4 // (2 ** ((depth + 1) // 2))
        self.classifier = nn.Sequential(nn.Linear(out_channels * side ** 2, 1024), act, nn.Linear(1024, 1))
        self._compute_dtype = dtype
        # split_precision: fp16 hi+lo operands, three tensor-core passes per layer, fp32 activations (engine_d_split.py)
        self._split = bool(split_precision)
        self._engine = None

    def engine(self):
        if self._engine is None:
            if self._split:
                from ..engine_d_split import SplitDiscriminatorEngine
                self._engine = SplitDiscriminatorEngine(self)
            else:
                from ..engine_d import DiscriminatorEngine
                self._engine = DiscriminatorEngine(self, self._compute_dtype)
            self._engine.param_list = list(self.parameters())
        return self._engine

    def forward(self, x):
        _require_cuda(x, "Discriminator")
        eng = self.engine()
        if not torch.is_grad_enabled():
            return eng.forward(x, save=False)[0]
        return _DiscriminatorFn.apply(eng, x, *eng.param_list)

    def forward_pair(self, a, b):
        """(D(a), D(b)), identical to two calls, with one gradient pass for the parameters (see _DiscriminatorPairFn)."""
        _require_cuda(a, "Discriminator")
        _require_cuda(b, "Discriminator")
        eng = self.engine()
        if not torch.is_grad_enabled():
            la, lb = eng.forward([a, b], save=False)[0]
            return la, lb
        return _DiscriminatorPairFn.apply(eng, a, b, *eng.param_list)
