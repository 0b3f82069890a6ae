"""Optimizer-step regularizers selected by name from a regime (reference: utils/regularization.py).

In scope (the ones the five north-star configs use): ``WeightDecay`` / ``L2Regularization``
(utils/regularization.py:117-148), ``GradSmooth`` (:198-224) and ``GradClip``.  Each is a plain torch
implementation operating on ``p`` / ``p.grad`` -- that is what runs on CPU (config C1).  When the model has
been converted for the B200 path, ``OptimRegime.step`` recognises these classes and folds them into the
fused arena kernels instead of calling the per-tensor hooks (engine.B200Optimizer).
"""
import logging

import torch
from torch.nn.utils import clip_grad_norm_

from .param_filter import FilterParameters, is_not_bn, is_not_bias


def _param_grad_norm(parameters):
    total = 0.0
    for p in parameters:
        total += float(p.grad.data.norm(2)) ** 2
    return total ** 0.5


class Regularizer(object):
    """Base class: hooks called around forward/backward/step; ``filter`` selects the parameters."""

    def __init__(self, model, value=0, filter={}, log=False):
        self._model = model
        self._named_parameters = list(FilterParameters(model, **filter).named_parameters())
        self.value = value
        self.log = log
        if log:
            logging.debug('Applying regularization to parameters: %s', [n for n, _ in self._named_parameters])

    def named_parameters(self):
        return iter(self._named_parameters)

    def parameters(self):
        return (p for _, p in self._named_parameters)

    def pre_step(self):
        pass

    def post_step(self):
        pass

    def pre_forward(self):
        pass

    def pre_backward(self):
        pass


class RegularizerList(Regularizer):
    """Sequence of regularizers; items are instances or ``(cls, kwargs)`` pairs."""

    def __init__(self, model, regularization_list):
        super(RegularizerList, self).__init__(model)
        self.regularization_list = []
        for item in regularization_list:
            if not isinstance(item, Regularizer):
                cls, kwargs = item
                item = cls(model=model, **kwargs)
            self.regularization_list.append(item)

    def _each(self, hook):
        for reg in self.regularization_list:
            getattr(reg, hook)()

    def pre_step(self):
        self._each('pre_step')

    def post_step(self):
        self._each('post_step')

    def pre_forward(self):
        self._each('pre_forward')

    def pre_backward(self):
        self._each('pre_backward')


class L2Regularization(Regularizer):
    """g += value * p before the step (pre_op) and/or p -= value * p after it (post_op)."""

    def __init__(self, model, value=0, filter={'parameter_name': is_not_bias, 'module': is_not_bn},
                 pre_op=True, post_op=False, **kwargs):
        super(L2Regularization, self).__init__(model, value, filter=filter, **kwargs)
        self.pre_op = pre_op
        self.post_op = post_op

    def pre_step(self):
        if not self.pre_op:
            return
        with torch.no_grad():
            for _, p in self._named_parameters:
                if p.grad is not None:
                    p.grad.add_(p, alpha=self.value)

    def post_step(self):
        if not self.post_op:
            return
        with torch.no_grad():
            for _, p in self._named_parameters:
                p.add_(p, alpha=-self.value)


class WeightDecay(L2Regularization):
    pass


class GradClip(Regularizer):
    def __init__(self, model, value=float('inf'), norm=2, filter={}, log=False):
        super(GradClip, self).__init__(model, value, filter=filter, log=log)
        self.norm = norm

    def pre_step(self):
        if self.value > 0:
            clip_grad_norm_(list(self.parameters()), self.value, self.norm)


class GradSmooth(Regularizer):
    """Rescale the gradient so that its norm follows an exponential moving average of past norms."""

    def __init__(self, model, value=True, momentum=0.9, filter={}, log=False):
        super(GradSmooth, self).__init__(model, value=value, filter=filter, log=log)
        self.momentum = momentum
        self.running_norm = None
        self.enabled = value
        self.counter = 0

    def pre_step(self):
        params = [p for p in self.parameters() if p.grad is not None]
        norm = _param_grad_norm(params)
        if self.running_norm is None:
            self.running_norm = norm
        else:
            self.running_norm = self.momentum * self.running_norm + (1 - self.momentum) * norm
            if self.enabled:
                coef = self.running_norm / (norm + 1e-6)
                for p in params:
                    p.grad.data.mul_(coef)
        self.counter += 1
