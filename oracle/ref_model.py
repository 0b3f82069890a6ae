"""TEST INFRASTRUCTURE ONLY -- functional CPU restatement of the reference ResNet training step.

The arithmetic of the reference lives in torch (unpinned in its requirements.txt:1; here torch 2.11): the
restatement therefore spells the reference's model code as explicit calls of the same torch functionals on
a plain ``state_dict`` (no nn.Module, none of this repo's model classes), which makes it an independent
check of both the re-authored model classes and the CUDA pipeline.

Follows, line by line:
  ResNet.features/forward          models/resnet.py:196-213   -> ``forward``
  stem (imagenet / cifar)          models/resnet.py:226-230, 328-332
  Bottleneck.forward               models/resnet.py:141-165   -> ``_bottleneck``
  BasicBlock.forward               models/resnet.py:98-118    -> ``_basic``
  downsample                       models/resnet.py:173-181
  nn.BatchNorm2d train semantics   (eps 1e-5, momentum 0.1, biased var to normalise, unbiased into running_var)
  CrossEntropyLoss                 utils/cross_entropy.py:14-67
  Trainer._step + OptimRegime.step trainer.py:106-177, utils/optim.py:254-264, utils/regularization.py:127-131,
                                   torch.optim.SGD (momentum, first step m = g)   -> ``sgd_step``
``quant`` emulates the storage precision of the CUDA path: a straight-through bf16 round placed where the
kernels store bf16 (conv outputs, BN/activation outputs, pooled features, their incoming gradients) --
the "T2" oracle of SURVEY.md section 8c.
"""
import re

import torch
import torch.nn.functional as F


class _STRound(torch.autograd.Function):
    """y = bf16(x) in forward; gradient rounded to bf16 in backward (storage emulation)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _q(x, quant):
    return _STRound.apply(x) if quant else x


def _bn(x, sd, prefix, training, buffers_out, quant):
    w, b = sd[prefix + '.weight'], sd[prefix + '.bias']
    rm, rv = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
    if training:
        n = x.numel() // x.size(1)
        mean = x.mean((0, 2, 3))
        var = x.var((0, 2, 3), unbiased=False)
        if buffers_out is not None:
            buffers_out[prefix + '.running_mean'] = 0.9 * rm + 0.1 * mean.detach().to(rm.dtype)
            buffers_out[prefix + '.running_var'] = 0.9 * rv + 0.1 * (var.detach() * n / max(n - 1, 1)).to(rv.dtype)
            buffers_out[prefix + '.num_batches_tracked'] = sd[prefix + '.num_batches_tracked'] + 1
    else:
        mean, var = rm.to(x.dtype), rv.to(x.dtype)
    inv = torch.rsqrt(var + 1e-5)
    return (x - mean[None, :, None, None]) * (inv * w)[None, :, None, None] + b[None, :, None, None]


def _conv(x, sd, name, stride, padding, quant):
    w = sd[name + '.weight']
    return _q(F.conv2d(x, w, None, stride=stride, padding=padding, groups=x.size(1) // w.size(1)), quant)


def _block_names(sd, layer):
    idx = sorted({int(m.group(1)) for k in sd for m in [re.match(r'%s\.(\d+)\.' % layer, k)] if m})
    return ['%s.%d' % (layer, i) for i in idx]


def _skip(x, sd, p, stride, training, bufs, quant):
    if p + '.downsample.0.weight' in sd:
        z = _conv(x, sd, p + '.downsample.0', stride, 0, quant)
        return _bn(z, sd, p + '.downsample.1', training, bufs, quant)
    return x


def _bottleneck(x, sd, p, stride, training, bufs, quant):
    out = _q(F.relu(_bn(_conv(x, sd, p + '.conv1', 1, 0, quant), sd, p + '.bn1', training, bufs, quant)), quant)
    out = _q(F.relu(_bn(_conv(out, sd, p + '.conv2', stride, 1, quant), sd, p + '.bn2', training, bufs, quant)), quant)
    out = _bn(_conv(out, sd, p + '.conv3', 1, 0, quant), sd, p + '.bn3', training, bufs, quant)
    return _q(F.relu(out + _skip(x, sd, p, stride, training, bufs, quant)), quant)


def _basic(x, sd, p, stride, training, bufs, quant):
    out = _q(F.relu(_bn(_conv(x, sd, p + '.conv1', stride, 1, quant), sd, p + '.bn1', training, bufs, quant)), quant)
    out = _bn(_conv(out, sd, p + '.conv2', 1, 1, quant), sd, p + '.bn2', training, bufs, quant)
    return _q(F.relu(out + _skip(x, sd, p, stride, training, bufs, quant)), quant)


def forward(sd, x, training=True, buffers_out=None, quant=False):
    """logits of a reference-layout ResNet ``state_dict`` (cifar or imagenet variant, basic or bottleneck)."""
    x = _q(x, quant)
    imagenet = sd['conv1.weight'].shape[-1] == 7
    if imagenet:
        x = _conv(x, sd, 'conv1', 2, 3, quant)
        x = _q(F.relu(_bn(x, sd, 'bn1', training, buffers_out, quant)), quant)
        x = F.max_pool2d(x, 3, 2, 1)
    else:
        x = _conv(x, sd, 'conv1', 1, 1, quant)
        x = _q(F.relu(_bn(x, sd, 'bn1', training, buffers_out, quant)), quant)
    for li, layer in enumerate(('layer1', 'layer2', 'layer3', 'layer4')):
        for bi, p in enumerate(_block_names(sd, layer)):
            stride = 2 if (bi == 0 and li > 0) else 1
            block = _bottleneck if (p + '.conv3.weight') in sd else _basic
            x = block(x, sd, p, stride, training, buffers_out, quant)
    x = _q(x.mean((2, 3)), quant)
    return F.linear(x, sd['fc.weight'], sd['fc.bias'])


def cross_entropy(logits, target, smooth_eps=0.0):
    """mean over the batch of -((1-eps-eps/C) lsm[t] + (eps/C) sum_c lsm[c])  (utils/cross_entropy.py:48-52);
    with eps = 0 this is F.cross_entropy (:20-24)."""
    lsm = F.log_softmax(logits, dim=-1)
    n_cls = logits.size(-1)
    u = smooth_eps / n_cls
    picked = lsm.gather(-1, target.view(-1, 1)).squeeze(-1)
    return (-((1.0 - u - smooth_eps) * picked + u * lsm.sum(-1))).mean()


def is_decayed(name):
    """membership of the reference's WeightDecay filter (models/resnet.py:34-40): not a bias, not in a BN."""
    if name.endswith('bias'):
        return False
    return not re.search(r'(^|\.)bn\d*\.|downsample\.1\.', name)


def param_names(sd):
    return [k for k in sd if not (k.endswith('running_mean') or k.endswith('running_var')
                                  or k.endswith('num_batches_tracked'))]


def loss_and_grads(sd, x, y, smooth_eps=0.0, quant=False, training=True):
    """One forward/backward: returns logits, loss, {param: grad}, updated BN buffers."""
    names = param_names(sd)
    work = {k: (v.detach().clone().requires_grad_(True) if k in names else v) for k, v in sd.items()}
    bufs = {}
    logits = forward(work, x, training=training, buffers_out=bufs, quant=quant)
    loss = cross_entropy(logits, y, smooth_eps)
    grads = torch.autograd.grad(loss, [work[k] for k in names])
    return logits.detach(), loss.detach(), dict(zip(names, grads)), bufs


def sgd_step(sd, grads, momentum_buf, lr, momentum=0.9, weight_decay=1e-4, loss_scale=1.0):
    """unscale -> WeightDecay.pre_step (decayed set only) -> SGD with momentum (first step: m = g)."""
    new_sd, new_m = dict(sd), {}
    for k, g in grads.items():
        g = g / loss_scale
        if is_decayed(k):
            g = g + weight_decay * sd[k]
        m = g.clone() if momentum_buf.get(k) is None else momentum * momentum_buf[k] + g
        new_m[k] = m
        new_sd[k] = sd[k] - lr * m
    return new_sd, new_m


def train_steps(sd, x, y, steps, lr=0.1, momentum=0.9, weight_decay=1e-4, quant=False, smooth_eps=0.0):
    """``steps`` full training steps on one resident batch; returns final state, momentum and the loss trace."""
    mom, losses = {}, []
    for _ in range(steps):
        _, loss, grads, bufs = loss_and_grads(sd, x, y, smooth_eps=smooth_eps, quant=quant)
        sd, mom = sgd_step(sd, grads, mom, lr, momentum, weight_decay)
        sd.update(bufs)
        losses.append(float(loss))
    return sd, mom, losses
