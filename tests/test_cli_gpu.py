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
    ckpt = tmp_path / "ck" / "pretrain" / "model_1.pt"
    assert ckpt.exists()
    out = _run([os.path.join(ROOT, "train.py"), "--phase", "train", "--vgg_random", "--pretrained_model", str(ckpt)] + common,
               ROOT)
    assert "Finish train [1/1]" in out and "Total G" in out
    assert (tmp_path / "ck" / "train" / "model_1.pt").exists()
