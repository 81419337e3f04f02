#!/usr/bin/env python
"""Training driver with the reference's command line (train.py:21-77) on the pesr_b200 modules.

    python train.py --phase pretrain --patch_size 48 ...          # L1 pretraining  (train.py:155-181)
    python train.py --phase train    --pretrained_model ...       # GAN fine-tuning (train.py:184-276)
    torchrun --nproc-per-node 8 train.py ...                      # one process per GPU (replaces nn.DataParallel)

Additions to the reference's flags: --synthetic (random tensors instead of data/, for machines without the
dataset), --max_iters, --log_every, --vgg_random, --resume.  Training patches come from a device-resident uint8 image
cache through one gather kernel per batch (pesr_b200.data.PatchSource, data.py:64-126); validation (train.py:281-295)
computes the Y-channel PSNR on the device (pesr_b200.utils, utils.py:27-41); besides the reference's G-only snapshots
(model_<epoch>.pt, best_model.pt) a full resumable state is written.  tensorboard logging is not reproduced.
"""
import argparse
import glob
import os
import random

import torch
import torch.distributed as dist

from pesr_b200 import steps
from pesr_b200.data import PatchSource
from pesr_b200.infer import imgs_to_tensor
from pesr_b200.model import VGG, Discriminator, Generator
from pesr_b200.optim import Adam
from pesr_b200.parallel import DataParallel
from pesr_b200.utils import PSNRMeter


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
parser.add_argument('--cuda_graph', type=_bool, default=False, help='replay the training step as one CUDA graph (single GPU): '
                    'host cost 0.04 ms instead of ~8 ms per step')
parser.add_argument('--resume', type=str, default='', help='full training state written by this script '
                    '(<check_point>/<phase>/state_<epoch>.pt): networks, both Adam states, epoch, RNG')


def read_png(path):
    try:
        import imageio
        return torch.from_numpy(imageio.imread(path))
    except ImportError:
        import numpy as np
        from PIL import Image
        return torch.from_numpy(np.asarray(Image.open(path).convert('RGB')).copy())


def load_pairs(root, device, limit=None):
    """data.py:30-53: sorted LR/HR PNG pairs of one dataset directory, kept on the device as uint8 HWC."""
    lr = sorted(glob.glob(os.path.join(root, 'LR', '*.png')))
    hr = sorted(glob.glob(os.path.join(root, 'HR', '*.png')))
    if not hr or len(lr) != len(hr):
        raise Exception('No images found in %s (use --synthetic to train on random tensors)' % root)
    if limit:
        lr, hr = lr[:limit], hr[:limit]
    return [(read_png(a).to(device), read_png(b).to(device)) for a, b in zip(lr, hr)]


def rng_state():
    import numpy as np
    return {'python': random.getstate(), 'numpy': np.random.get_state(), 'torch': torch.get_rng_state(),
            'cuda': torch.cuda.get_rng_state()}


def set_rng_state(st):
    import numpy as np
    random.setstate(st['python'])
    np.random.set_state(st['numpy'])
    torch.set_rng_state(st['torch'])
    torch.cuda.set_rng_state(st['cuda'])


def save_full_state(path, epoch, best_psnr, G, optim_G, D=None, optim_D=None):
    """Everything a resume needs (the reference only ever saves G, train.py:297-310): both networks incl. BatchNorm
    running statistics, both Adam states, the epoch counter (which determines the learning rate), the best validation
    PSNR and the random-number generator states."""
    st = {'epoch': epoch, 'best_psnr': best_psnr, 'G': G.module.state_dict(), 'optim_G': optim_G.state_dict(),
          'rng': rng_state()}
    if D is not None:
        st.update(D=D.module.state_dict(), optim_D=optim_D.state_dict())
    tmp = path + '.tmp'
    torch.save(st, tmp)
    os.replace(tmp, path)


def main():
    args = parser.parse_args()
    if args.GP and args.cuda_graph:
        raise SystemExit('--GP true (gradient penalty, train.py:216-226) runs its double backward through ATen autograd '
                         '(pesr_b200/gp.py) and is not captured: use it without --cuda_graph')
    if args.num_channels % 64 != 0:
        raise SystemExit('--num_channels must be a multiple of 64 (tensor-core K block)')
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        from pesr_b200.parallel import nccl_env_defaults
        nccl_env_defaults()
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
    graphed = args.cuda_graph and world == 1
    optim_G = Adam([p for p in G.parameters() if p.requires_grad], betas=(0.9, 0.999), lr=args.learning_rate,
                   capturable=graphed)
    gan = args.phase != 'pretrain'
    D = optim_D = None
    if gan:
        D = DataParallel(Discriminator(opt).to(device))
        vgg = VGG(pretrained=not args.vgg_random).to(device)
        optim_D = Adam(D.parameters(), betas=(0.9, 0.999), lr=args.learning_rate, capturable=graphed)
        cfg = dict(alpha_l1=args.alpha_l1, alpha_vgg=args.alpha_vgg, alpha_gan=args.alpha_gan, alpha_tv=args.alpha_tv,
                   fl_gamma=args.fl_gamma, gan_type=args.gan_type, focal_loss=args.focal_loss,
                   target_real=torch.ones(args.batch_size, 1, device=device),
                   target_fake=torch.zeros(args.batch_size, 1, device=device), GP=args.GP)
    check_point = os.path.join(args.check_point, args.phase)
    if rank == 0:
        os.makedirs(check_point, exist_ok=True)
    images = None if args.synthetic else load_pairs(os.path.join('data/origin/train', args.train_dataset), device)
    data = PatchSource(images, args.patch_size, args.scale, args.batch_size, device, num_repeats=args.num_repeats)
    # validation set (train.py:99-104): whole images, batch 1
    if args.synthetic:
        vside = 2 * args.patch_size
        gen = torch.Generator().manual_seed(1234)
        val = [((torch.rand(vside, vside, 3, generator=gen) * 255).to(torch.uint8).to(device),
                (torch.rand(vside * args.scale, vside * args.scale, 3, generator=gen) * 255).to(torch.uint8).to(device))
               for _ in range(min(args.num_valids, 2))]
    else:
        val = load_pairs(os.path.join('data/origin/valid', args.valid_dataset), device, limit=args.num_valids)
    iters = data.per_epoch // (args.batch_size * world)
    if args.max_iters:
        iters = min(iters, args.max_iters)
    best_psnr, start_epoch = 0.0, 1
    if args.resume:
        st = torch.load(args.resume, map_location='cpu', weights_only=False)
        G.module.load_state_dict(st['G'])
        optim_G.load_state_dict(st['optim_G'])
        if gan and 'D' in st:
            D.module.load_state_dict(st['D'])
            optim_D.load_state_dict(st['optim_D'])
        best_psnr, start_epoch = st['best_psnr'], st['epoch'] + 1
        set_rng_state(st['rng'])
        if rank == 0:
            print('Resumed from %s at epoch %d' % (args.resume, start_epoch))
    if gan:
        def step_fn(lr, hr):
            return steps.gan_step(G, D, vgg, optim_G, optim_D, lr, hr, cfg, ddp_g=G if world > 1 else None,
                                  ddp_d=D if world > 1 else None)
    else:
        def step_fn(lr, hr):
            return steps.pretrain_step(G, optim_G, lr, hr, ddp=G if world > 1 else None).detach().reshape(1)
    gstep = None
    for epoch in range(start_epoch, args.num_epochs + 1):
        # train.py:156,185-186 call StepLR.step() at the START of every epoch.  Under the pinned torch 0.4 (README.md:22)
        # the scheduler's counter starts at -1, so epoch e (1-based) runs at lr * 0.5 ** ((e - 1) // lr_step): the first
        # halving takes effect in epoch lr_step + 1.  (The same script on torch >= 1.1 would halve one epoch earlier.)
        # The rate is set explicitly so that it does not depend on the installed torch's scheduler semantics.
        cur_lr = args.learning_rate * 0.5 ** ((epoch - 1) // args.lr_step)
        for o in (optim_G, optim_D):
            if o is not None:
                for g in o.param_groups:
                    g['lr'] = cur_lr
        if rank == 0:
            print('Model {}. Epoch [{}/{}]. Learning rate: {}'.format(check_point, epoch, args.num_epochs, cur_lr))
        running = torch.zeros(5 if gan else 1, device=device)
        for it in range(iters):
            lr, hr = data.batch()
            if graphed:
                if gstep is None:       # captured once (after the learning rate of this epoch is set); replayed afterwards
                    from pesr_b200.graph import GraphedStep
                    gstep = GraphedStep(step_fn, (lr, hr), modules=[m.module for m in (G, D) if m is not None] + ([vgg] if gan else []),
                                        optimizers=[o for o in (optim_G, optim_D) if o is not None], warmup=2)
                running += gstep(lr, hr)
            else:
                running += step_fn(lr, hr)
            if rank == 0 and args.log_every and (it + 1) % args.log_every == 0:
                print('  iter %d/%d  %s' % (it + 1, iters, (running / (it + 1)).tolist()))   # one sync per log line
        avr = (running / max(iters, 1)).tolist()
        if rank == 0:
            if gan:
                print('Finish train [%d/%d]. L1: %.2f. VGG: %.2f. G: %.2f. TV: %.2f. Total G: %.2f. D: %.2f'
                      % (epoch, args.num_epochs, avr[0], avr[1], avr[2], avr[3], sum(avr[0:4]), avr[4]))
            else:
                print('Finish train [%d/%d]. Loss: %.2f' % (epoch, args.num_epochs, avr[0]))
        # ---- validation (train.py:281-295): G stays in train mode in the reference (it has no BatchNorm / dropout);
        # Y-channel PSNR is accumulated on the device, one read-back per epoch
        if rank == 0:
            print('Validating...')
            meter = PSNRMeter(device)
            with torch.no_grad():
                for lr_u8, hr_u8 in val:
                    sr = G(imgs_to_tensor(lr_u8))
                    h = min(sr.shape[2], hr_u8.shape[0])
                    w = min(sr.shape[3], hr_u8.shape[1])
                    meter.update(imgs_to_tensor(hr_u8)[:, :, :h, :w].contiguous(), sr[:, :, :h, :w].contiguous())
            val_psnr = meter.value()
            if not gan:
                print('Finish valid [%d/%d]. Best PSNR: %.4fdB. Cur PSNR: %.4fdB' % (epoch, args.num_epochs, best_psnr, val_psnr))
                if best_psnr < val_psnr:
                    best_psnr = val_psnr
                    torch.save(G.module.state_dict(), os.path.join(check_point, 'best_model.pt'))    # train.py:297-303
                    print('Saved new best model.')
            else:
                print('Finish valid [%d/%d]. PSNR: %.4fdB' % (epoch, args.num_epochs, val_psnr))
            if epoch % args.snapshot_every == 0 or epoch == args.num_epochs:
                torch.save(G.module.state_dict(), os.path.join(check_point, 'model_{}.pt'.format(epoch)))   # train.py:306-310
                print('Saved snapshot model.')
                save_full_state(os.path.join(check_point, 'state_{}.pt'.format(epoch)), epoch, best_psnr, G, optim_G,
                                D, optim_D)
            print('')
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
