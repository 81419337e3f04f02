#!/usr/bin/env python
"""Inference driver with the reference's command line (test.py:13-33): x4 super-resolution of every PNG under
data/origin/test/<dataset>/LR with the perceptual model, optionally blended (alpha != 1) with the x8 self-ensemble
of the PSNR model; writes PNGs to <save_path>/<dataset>."""
import argparse
import glob
import os

import torch

from pesr_b200 import infer
from pesr_b200.model import Generator

parser = argparse.ArgumentParser(description='SR benchmark')
parser.add_argument('--dataset', type=str, default='Set5', help='test dataset')
parser.add_argument('--perceptual_model', type=str, default='check_point/PESR/train/PERC_model.pt', help='perceptual model name')
parser.add_argument('--psnr_model', type=str, default='check_point/PESR/pretrain/PSNR_model.pt', help='pretrained (l1 loss) model name')
parser.add_argument('--num_channels', type=int, default=256)
parser.add_argument('--num_blocks', type=int, default=32)
parser.add_argument('--res_scale', type=float, default=0.1)
parser.add_argument('--alpha', type=float, default=1, help='PSNR-perceptual tradeoff')
parser.add_argument('--save_path', type=str, default='results')


def _read(path):
    try:
        import imageio
        return imageio.imread(path)
    except ImportError:
        import numpy as np
        from PIL import Image
        return np.asarray(Image.open(path).convert('RGB')).copy()


def _write(path, img):
    try:
        import imageio
        imageio.imwrite(path, img)
    except ImportError:
        from PIL import Image
        Image.fromarray(img).save(path)


def main():
    args = parser.parse_args()
    print('-------YOUR SETTINGS_________')
    for arg in vars(args):
        print("%20s: %s" % (str(arg), str(getattr(args, arg))))
    lr_paths = glob.glob(os.path.join('data/origin/test/', args.dataset, 'LR', '*.png'))
    print('Loading model...')
    opt = {'num_channels': args.num_channels, 'depth': args.num_blocks, 'res_scale': args.res_scale}
    model = Generator(opt).cuda()
    model.load_state_dict(torch.load(args.perceptual_model, map_location='cpu'))
    print("Number of parameters:", sum(p.nelement() for p in model.parameters()))
    model_psnr = None
    if args.alpha != 1:
        model_psnr = Generator(opt).cuda()
        model_psnr.load_state_dict(torch.load(args.psnr_model, map_location='cpu'))
    save_path = os.path.join(args.save_path, args.dataset)
    os.makedirs(save_path, exist_ok=True)
    print('Start testing')
    for i, lr_path in enumerate(lr_paths):
        inp = infer.imgs_to_tensor(_read(lr_path))
        _, out8 = infer.super_resolve(model, inp, alpha=args.alpha, model_psnr=model_psnr)
        print('Tested %d img(s)' % (i + 1))
        _write(os.path.join(save_path, os.path.basename(lr_path)), out8.cpu().numpy())
    print('Finish')


if __name__ == '__main__':
    main()
