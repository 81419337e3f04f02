"""Gradient penalty of the Discriminator phase (reference: train.py:216-226, `--GP true`, off by default).

    u ~ U(0,1) per sample;  x = hr*u + sr*(1-u)  (a new leaf);  g = d(sum_n D(x)_n)/dx  with create_graph
    GP = 10 * mean_n (||g_n||_2 - 1)^2,  added to total_D_loss before its backward

The penalty is a function of D's INPUT GRADIENT, so its parameter gradient is a second derivative through every layer
of D (convolutions, train-mode BatchNorm, LeakyReLU, the two Linear layers).  The B200 kernel schedule of the
Discriminator (engine_d.py) implements the first derivative only; this optional branch therefore evaluates D(x) for the
penalty with ATen operators on the SAME parameters and lets autograd build the double-backward graph.  It is exact
(same function, same parameters, BatchNorm running statistics updated by this third train-mode call exactly as in the
reference), it is NOT a hand-written path, and it is labelled as such in DESIGN.md: the headline step (GP off, the
reference's default) never touches it.
"""
import torch
import torch.nn.functional as F


def discriminator_aten(D, x):
    """model/pesr.py:77-81 on ATen operators over D's own parameters / buffers (train or eval mode as D is)."""
    for block in D.features:
        conv, bn = block[0], block[1]
        x = F.conv2d(x, conv.weight, conv.bias, stride=conv.stride, padding=conv.padding)
        x = F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, training=D.training, momentum=bn.momentum,
                         eps=bn.eps)
        if D.training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        x = F.leaky_relu(x, 0.2)
    fc1, fc2 = D.classifier[0], D.classifier[2]
    x = F.leaky_relu(F.linear(x.reshape(x.shape[0], -1), fc1.weight, fc1.bias), 0.2)
    return F.linear(x, fc2.weight, fc2.bias)


def gradient_penalty(D, hr, sr, u=None, weight=10.0):
    """train.py:216-226.  `u`: the per-sample mixing factors [N,1,1,1] (drawn here when None, as the reference does).
    Returns the scalar penalty with a graph reaching D's parameters."""
    module = getattr(D, "module", D)
    n = hr.shape[0]
    if u is None:
        u = torch.rand(n, 1, 1, 1, device=hr.device, dtype=hr.dtype)
    x_both = (hr.detach() * u + sr.detach() * (1 - u)).requires_grad_(True)      # `Variable(x_both, requires_grad=True)`: a leaf
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False     # the reference (torch 0.4) is plain fp32
    try:
        with torch.enable_grad():
            pred = discriminator_aten(module, x_both)
            grad = torch.autograd.grad(outputs=pred, inputs=x_both, grad_outputs=torch.ones_like(pred), retain_graph=True,
                                       create_graph=True, only_inputs=True)[0]
            return weight * ((grad.norm(2, 1).norm(2, 1).norm(2, 1) - 1) ** 2).mean()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
