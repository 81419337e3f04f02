"""GPU parity of the inference path (test.py:101-112, utils.py:13-25) against the oracle."""
import os

import pytest
import torch

from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu


def test_x8_ensemble_blend_and_uint8_match_oracle():
    from oracle import pesr_oracle as O
    from pesr_b200 import infer
    from pesr_b200.model import Generator
    opt = {'depth': 2, 'num_channels': 64, 'res_scale': 0.1}
    sd_a, sd_b = O.init_generator(opt, 1), O.init_generator(opt, 2)
    Ga, Gb = Generator(opt), Generator(opt)
    Ga.load_state_dict(sd_a), Gb.load_state_dict(sd_b)
    Ga, Gb = Ga.cuda().eval(), Gb.cuda().eval()
    g = torch.Generator().manual_seed(3)
    img_u8 = (torch.rand(13, 9, 3, generator=g) * 255).to(torch.uint8)          # HWC, h != w so transposes matter
    x = infer.imgs_to_tensor(img_u8.numpy())
    assert torch.equal(x.cpu(), img_u8.permute(2, 0, 1)[None].float())

    fa = lambda t: O.generator_forward({k: v.double() for k, v in sd_a.items()}, t.double(), 2, 0.1)  # noqa: E731
    fb = lambda t: O.generator_forward({k: v.double() for k, v in sd_b.items()}, t.double(), 2, 0.1)  # noqa: E731
    for alpha in (1.0, 0.5, 0.0):
        out32, out8 = infer.super_resolve(Ga, x, alpha=alpha, model_psnr=Gb)
        ref = O.infer(fa, x.cpu(), alpha, fb)
        assert rel_l2(out32.cpu(), ref) < 1e-3
        ref8 = O.tensors_to_img_u8(ref.float())
        diff = (out8.cpu().int() - torch.from_numpy(ref8.copy()).int()).abs()
        assert int(diff.max()) <= 1 and float((diff > 0).float().mean()) < 0.02     # only .5 rounding ties may differ
    assert rel_l2(infer.x8_forward(x, Gb).cpu(), O.x8_forward(fb, x.cpu())) < 1e-3


def test_blend_kernel_is_bit_exact_on_given_generator_outputs():
    """The fused kernel against the reference's own arithmetic (golden from tests/golden/make_golden.py): the
    clip/round of utils.py:15-17 incl. numpy's round-half-even, and the x8 inverse-transform order."""
    from pesr_b200 import ops
    from pesr_b200._lib import check, lib
    gd = torch.load(os.path.join(GOLDEN, "x8_round.pt"), weights_only=False)
    a = gd["round_in"].cuda().contiguous()                                       # [1,3,1,8]
    out8 = torch.empty(1, 8, 3, device="cuda", dtype=torch.uint8)
    check(lib.pesr_blend_x8_to_u8(a.data_ptr(), 0, 1, 8, 1.0, 0, 0, out8.data_ptr(), torch.cuda.current_stream().cuda_stream))
    assert torch.equal(out8.cpu(), gd["round_out"])
    wt, img = gd["wt"].cuda(), gd["img"].cuda()

    def model(t):
        return torch.nn.functional.conv2d(torch.nn.functional.interpolate(t, scale_factor=2), wt, padding=1)
    from pesr_b200 import infer
    got = infer.x8_forward(img, model)
    assert rel_l2(got.cpu(), gd["out"]) < 1e-6
    del ops
