#!/usr/bin/env python
"""Training driver with the reference's command line (train.py:21-77) on the pesr_b200 modules.

    python train.py --phase pretrain --patch_size 48 ...          # L1 pretraining  (train.py:155-181)
    python train.py --phase train    --pretrained_model ...       # GAN fine-tuning (train.py:184-276)
    torchrun --nproc-per-node 8 train.py ...                      # one process per GPU (replaces nn.DataParallel)

Additions to the reference's flags: --synthetic (random tensors instead of data/, for machines without the
dataset), --max_iters, --log_every, --vgg_random.  The dataset / tensorboard / validation plumbing of the reference is out of the
hot path's scope; a minimal PNG-folder dataset with the reference's directory layout is
provided so that the script is usable end to end.
"""
import argparse
import glob
import os
import random

import torch
import torch.distributed as dist

from pesr_b200 import steps
from pesr_b200.model import VGG, Discriminator, Generator
from pesr_b200.optim import Adam
from pesr_b200.parallel import DataParallel


def _bool(x):
    return str(x).lower() == 'true'


parser = argparse.ArgumentParser(description='PIRM 2018')
parser.add_argument('--scale', type=int, default=4, help='interpolation scale')
parser.add_argument('--train_dataset', type=str, default='DIV2K', help='training dataset')
parser.add_argument('--valid_dataset', type=str, default='PIRM', help='validation dataset')
parser.add_argument('--num_valids', type=int, default=10, help='number of images for validation')
parser.add_argument('--num_channels', type=int, default=256, help='number of resnet channel')
parser.add_argument('--num_blocks', type=int, default=32, help='number of resnet blocks')
parser.add_argument('--res_scale', type=float, default=0.1)
parser.add_argument('--phase', type=str, default='train', help='phase: pretrain or train')
parser.add_argument('--pretrained_model', type=str, default='', help='pretrained model for the train phase')
parser.add_argument('--batch_size', type=int, default=16, help='batch size used for training (per GPU)')
parser.add_argument('--learning_rate', type=float, default=5e-5, help='learning rate used for training')
parser.add_argument('--lr_step', type=int, default=120, help='steps to decay learning rate')
parser.add_argument('--num_epochs', type=int, default=200, help='number of training epochs')
parser.add_argument('--num_repeats', type=int, default=20, help='number of repeats for each image per epoch')
parser.add_argument('--patch_size', type=int, default=24, help='input patch size')
parser.add_argument('--check_point', type=str, default='check_point/my_model', help='path to save log and model')
parser.add_argument('--snapshot_every', type=int, default=10, help='snapshot freq')
parser.add_argument('--gan_type', type=str, default='RSGAN')
parser.add_argument('--GP', type=_bool, default=False, help='gradient penalty')
parser.add_argument('--spectral_norm', type=_bool, default=False, help='spectral normalization')
parser.add_argument('--focal_loss', type=_bool, default=True)
parser.add_argument('--fl_gamma', type=float, default=1, help='focal loss gamma')
parser.add_argument('--alpha_vgg', type=float, default=50)
parser.add_argument('--alpha_gan', type=float, default=1)
parser.add_argument('--alpha_tv', type=float, default=1e-6)
parser.add_argument('--alpha_l1', type=float, default=0)
parser.add_argument('--synthetic', action='store_true', help='random patches instead of data/origin/train/<dataset>')
parser.add_argument('--max_iters', type=int, default=0, help='stop an epoch after this many iterations (0 = full epoch)')
parser.add_argument('--vgg_random', action='store_true', help='random-init VGG19 instead of torchvision\'s ImageNet '
                    'checkpoint (machines without network access)')
parser.add_argument('--log_every', type=int, default=50)


def read_png(path):
    try:
        import imageio
        return torch.from_numpy(imageio.imread(path))
    except ImportError:
        import numpy as np
        from PIL import Image
        return torch.from_numpy(np.asarray(Image.open(path).convert('RGB')).copy())


class PatchSource:
    """Aligned random LR/HR crops with the 8-way flip/transpose augmentation of data.py:64-126, produced on the GPU
    from device-resident uint8 images (or random tensors with --synthetic)."""

    def __init__(self, args, device):
        self.args, self.device, self.images = args, device, None
        if not args.synthetic:
            root = os.path.join('data/origin/train', args.train_dataset)
            lr = sorted(glob.glob(os.path.join(root, 'LR', '*.png')))
            hr = sorted(glob.glob(os.path.join(root, 'HR', '*.png')))
            if not hr or len(lr) != len(hr):
                raise Exception('No images found (use --synthetic to train on random tensors)')
            self.images = [(read_png(a).to(device), read_png(b).to(device)) for a, b in zip(lr, hr)]
        self.per_epoch = (800 if self.images is None else len(self.images)) * args.num_repeats

    def batch(self):
        a, dev = self.args, self.device
        p, s, b = a.patch_size, a.scale, a.batch_size
        if self.images is None:
            return torch.rand(b, 3, p, p, device=dev) * 255, torch.rand(b, 3, p * s, p * s, device=dev) * 255
        lrs, hrs = [], []
        for _ in range(b):
            lr, hr = random.choice(self.images)
            y, x = random.randrange(lr.shape[0] - p + 1), random.randrange(lr.shape[1] - p + 1)
            l = lr[y:y + p, x:x + p].permute(2, 0, 1).float()
            h = hr[y * s:(y + p) * s, x * s:(x + p) * s].permute(2, 0, 1).float()
            k = random.randrange(8)
            if k & 1:
                l, h = l.flip(2), h.flip(2)
            if k & 2:
                l, h = l.flip(1), h.flip(1)
            if k & 4:
                l, h = l.transpose(1, 2), h.transpose(1, 2)
            lrs.append(l)
            hrs.append(h)
        return torch.stack(lrs).contiguous(), torch.stack(hrs).contiguous()


def main():
    args = parser.parse_args()
    if args.GP:
        raise NotImplementedError('--GP true (gradient penalty, train.py:216-226) needs second-order conv backward; '
                                  'it is off by default in the reference and outside the B200 hot path')
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    if rank == 0:
        print('Loading model using %d GPU(s)' % world)
    opt = {'patch_size': args.patch_size, 'num_channels': args.num_channels, 'depth': args.num_blocks,
           'res_scale': args.res_scale, 'spectral_norm': args.spectral_norm}
    G = Generator(opt)
    if args.pretrained_model != '':
        print('Fetching pretrained model', args.pretrained_model)
        G.load_state_dict(torch.load(args.pretrained_model, map_location='cpu'))
    G = DataParallel(G.to(device))
    optim_G = Adam([p for p in G.parameters() if p.requires_grad], betas=(0.9, 0.999), lr=args.learning_rate)
    sched_G = torch.optim.lr_scheduler.StepLR(optim_G, step_size=args.lr_step, gamma=0.5)
    gan = args.phase != 'pretrain'
    if gan:
        D = DataParallel(Discriminator(opt).to(device))
        vgg = VGG(pretrained=not args.vgg_random).to(device)
        optim_D = Adam(D.parameters(), betas=(0.9, 0.999), lr=args.learning_rate)
        sched_D = torch.optim.lr_scheduler.StepLR(optim_D, step_size=args.lr_step, gamma=0.5)
        cfg = dict(alpha_l1=args.alpha_l1, alpha_vgg=args.alpha_vgg, alpha_gan=args.alpha_gan, alpha_tv=args.alpha_tv,
                   fl_gamma=args.fl_gamma, gan_type=args.gan_type, focal_loss=args.focal_loss,
                   target_real=torch.ones(args.batch_size, 1, device=device),
                   target_fake=torch.zeros(args.batch_size, 1, device=device))
    check_point = os.path.join(args.check_point, args.phase)
    if rank == 0:
        os.makedirs(check_point, exist_ok=True)
    data = PatchSource(args, device)
    iters = data.per_epoch // (args.batch_size * world)
    if args.max_iters:
        iters = min(iters, args.max_iters)
    for epoch in range(1, args.num_epochs + 1):
        cur_lr = optim_G.param_groups[0]['lr']
        if rank == 0:
            print('Model {}. Epoch [{}/{}]. Learning rate: {}'.format(check_point, epoch, args.num_epochs, cur_lr))
        running = torch.zeros(5 if gan else 1, device=device)
        for it in range(iters):
            lr, hr = data.batch()
            if gan:
                running += steps.gan_step(G, D, vgg, optim_G, optim_D, lr, hr, cfg, ddp_g=G if world > 1 else None,
                                          ddp_d=D if world > 1 else None)
            else:
                running += steps.pretrain_step(G, optim_G, lr, hr, ddp=G if world > 1 else None).detach()
            if rank == 0 and args.log_every and (it + 1) % args.log_every == 0:
                print('  iter %d/%d  %s' % (it + 1, iters, (running / (it + 1)).tolist()))   # one sync per log line
        avr = (running / max(iters, 1)).tolist()
        if rank == 0:
            if gan:
                print('Finish train [%d/%d]. L1: %.2f. VGG: %.2f. G: %.2f. TV: %.2f. Total G: %.2f. D: %.2f'
                      % (epoch, args.num_epochs, avr[0], avr[1], avr[2], avr[3], sum(avr[0:4]), avr[4]))
            else:
                print('Finish train [%d/%d]. Loss: %.2f' % (epoch, args.num_epochs, avr[0]))
            if epoch % args.snapshot_every == 0 or epoch == args.num_epochs:
                model_path = os.path.join(check_point, 'model_{}.pt'.format(epoch))
                torch.save(G.module.state_dict(), model_path)                     # train.py:303,309
                print('Saved snapshot model.')
        sched_G.step()
        if gan:
            sched_D.step()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
