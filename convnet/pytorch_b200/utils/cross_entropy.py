"""Cross-entropy with label smoothing and soft targets; interface of utils/cross_entropy.py:14-85.

Logits produced by a taped forward of the B200 engine carry a ``_b200_head`` tag (engine._HeadHandle);
``CrossEntropyLoss`` then runs the fused softmax-CE kernel (csrc/loss.cu, forward + backward, label smoothing
included) through ``engine._FusedCE``.  Everything else -- CPU (config C1), soft targets, class weights,
``reduction != 'mean'``, ``ignore_index >= 0`` -- uses the torch definition below, which is also the semantic
specification of that kernel.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .misc import onehot


def _is_long(t):
    return t.dtype == torch.long


def cross_entropy(inputs, target, weight=None, ignore_index=-100, reduction='mean',
                  smooth_eps=None, smooth_dist=None, from_logits=True):
    eps = smooth_eps or 0
    if _is_long(target) and eps == 0:  # plain negative log-likelihood
        fn = F.cross_entropy if from_logits else F.nll_loss
        return fn(inputs, target, weight, ignore_index=ignore_index, reduction=reduction)

    lsm = F.log_softmax(inputs, dim=-1) if from_logits else inputs
    n_cls = inputs.size(-1)
    ignored = target.eq(ignore_index) if (_is_long(target) and ignore_index >= 0) else None

    if eps > 0 and smooth_dist is not None:
        if _is_long(target):
            target = onehot(target, n_cls).type_as(inputs)
        if smooth_dist.dim() < target.dim():
            smooth_dist = smooth_dist.unsqueeze(0)
        target = torch.lerp(target, smooth_dist.expand_as(target), eps)
    if weight is not None:
        lsm = lsm * weight.unsqueeze(0)

    if _is_long(target):
        uniform = eps / n_cls
        picked = lsm.gather(-1, target.unsqueeze(-1)).squeeze(-1)
        loss = -((1.0 - uniform - eps) * picked + uniform * lsm.sum(-1))
    else:
        loss = -(target * lsm).sum(-1)
    if ignored is not None:
        loss = loss.masked_fill(ignored, 0)

    if reduction == 'sum':
        return loss.sum()
    if reduction == 'mean':
        if ignored is None:
            return loss.mean()
        return loss.sum() / float(loss.size(0) - int(ignored.sum()))
    return loss


class CrossEntropyLoss(nn.CrossEntropyLoss):
    """nn.CrossEntropyLoss accepting distributions as targets and a label-smoothing coefficient."""

    def __init__(self, weight=None, ignore_index=-100, reduction='mean', smooth_eps=None, smooth_dist=None,
                 from_logits=True):
        super(CrossEntropyLoss, self).__init__(weight=weight, ignore_index=ignore_index, reduction=reduction)
        self.smooth_eps = smooth_eps
        self.smooth_dist = smooth_dist
        self.from_logits = from_logits

    def forward(self, input, target, smooth_dist=None):
        dist = self.smooth_dist if smooth_dist is None else smooth_dist
        head = getattr(input, '_b200_head', None)
        if head is not None and _is_long(target) and target.is_cuda and self.weight is None and dist is None \
                and self.reduction == 'mean' and self.ignore_index < 0 and self.from_logits:
            return head.loss(input, target, self.smooth_eps or 0.0)
        return cross_entropy(input, target, weight=self.weight, ignore_index=self.ignore_index,
                             reduction=self.reduction, smooth_eps=self.smooth_eps, smooth_dist=dist,
                             from_logits=self.from_logits)
