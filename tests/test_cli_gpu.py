"""The drop-in scripts end to end on tiny synthetic configurations: train.py (both phases of the reference,
train.py:164-176 and train.py:184-276) and test.py's inference path on a generated checkpoint."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, cwd):
    r = subprocess.run([sys.executable] + args, cwd=cwd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout


def test_train_py_pretrain_then_gan(tmp_path):
    common = ["--synthetic", "--num_channels", "64", "--num_blocks", "2", "--batch_size", "4", "--patch_size", "16",
              "--num_epochs", "1", "--max_iters", "2", "--check_point", str(tmp_path / "ck")]
    out = _run([os.path.join(ROOT, "train.py"), "--phase", "pretrain"] + common, ROOT)
    assert "Finish train [1/1]" in out and "Saved snapshot model." in out
    assert "Finish valid [1/1]. Best PSNR:" in out and "Saved new best model." in out
    ckpt = tmp_path / "ck" / "pretrain" / "model_1.pt"
    assert ckpt.exists() and (tmp_path / "ck" / "pretrain" / "best_model.pt").exists()
    out = _run([os.path.join(ROOT, "train.py"), "--phase", "train", "--vgg_random", "--pretrained_model", str(ckpt)] + common,
               ROOT)
    assert "Finish train [1/1]" in out and "Total G" in out and "Finish valid [1/1]. PSNR:" in out
    assert (tmp_path / "ck" / "train" / "model_1.pt").exists()
    # full-state checkpoint (G, D incl. BatchNorm running statistics, both Adam states, epoch, RNG) and --resume
    import torch
    state = tmp_path / "ck" / "train" / "state_1.pt"
    st = torch.load(state, map_location="cpu", weights_only=False)
    assert st["epoch"] == 1 and {"G", "D", "optim_G", "optim_D", "rng", "best_psnr"} <= set(st)
    assert any("running_mean" in k for k in st["D"]) and len(st["optim_D"]["state"]) == len(list(st["D"])) - 3 * 8
    common2 = [a if a != "1" or common[i - 1] != "--num_epochs" else "2" for i, a in enumerate(common)]
    out = _run([os.path.join(ROOT, "train.py"), "--phase", "train", "--vgg_random", "--resume", str(state)] + common2, ROOT)
    assert "Resumed from" in out and "Epoch [2/2]" in out and "Epoch [1/2]" not in out
    assert (tmp_path / "ck" / "train" / "state_2.pt").exists()
    assert torch.load(tmp_path / "ck" / "train" / "state_2.pt", map_location="cpu", weights_only=False)["epoch"] == 2


def test_train_py_with_cuda_graph(tmp_path):
    """--cuda_graph true: the GAN step captured once and replayed (pesr_b200/graph.py); validation in between runs eagerly on
    the replayed parameters."""
    common = ["--synthetic", "--vgg_random", "--num_channels", "64", "--num_blocks", "2", "--batch_size", "4", "--patch_size", "16",
              "--num_epochs", "2", "--max_iters", "3", "--cuda_graph", "true", "--check_point", str(tmp_path / "ck")]
    out = _run([os.path.join(ROOT, "train.py"), "--phase", "train"] + common, ROOT)
    assert "Finish train [2/2]" in out and "Finish valid [2/2]. PSNR:" in out
    import re
    psnr = [float(x) for x in re.findall(r"PSNR: ([0-9.]+)dB", out)]
    assert len(psnr) == 2 and all(0 < p < 100 for p in psnr)


def test_test_py_writes_x4_images(tmp_path):
    """test.py end to end (test.py:101-112): two random LR PNGs, random tiny checkpoints, alpha blend with the x8
    self-ensemble of the PSNR model; the outputs are x4 PNGs."""
    import numpy as np
    import torch
    from PIL import Image
    from pesr_b200.model import Generator
    lr_dir = tmp_path / "data" / "origin" / "test" / "Tiny" / "LR"
    lr_dir.mkdir(parents=True)
    rng = np.random.RandomState(0)
    for name, (h, w) in (("a.png", (24, 20)), ("b.png", (17, 33))):
        Image.fromarray(rng.randint(0, 256, size=(h, w, 3), dtype=np.uint8)).save(lr_dir / name)
    opt = {'num_channels': 64, 'depth': 2, 'res_scale': 0.1}
    for i, name in enumerate(("perc.pt", "psnr.pt")):
        torch.manual_seed(i)
        torch.save(Generator(opt).state_dict(), tmp_path / name)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "test.py"), "--dataset", "Tiny", "--num_channels", "64",
                        "--num_blocks", "2", "--perceptual_model", str(tmp_path / "perc.pt"), "--psnr_model",
                        str(tmp_path / "psnr.pt"), "--alpha", "0.5", "--save_path", str(tmp_path / "out")],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for name, (h, w) in (("a.png", (24, 20)), ("b.png", (17, 33))):
        out = np.asarray(Image.open(tmp_path / "out" / "Tiny" / name))
        assert out.shape == (4 * h, 4 * w, 3) and out.dtype == np.uint8
