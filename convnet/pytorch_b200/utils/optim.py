"""Regime-driven optimizer wrapper: the reference's ``OptimRegime`` contract (utils/optim.py:89-284).

``OptimRegime(model, regime, defaults={}, filter=None, use_float_copy=False, log=True)`` exposes
``update(epoch, steps)``, ``zero_grad()``, ``pre_forward()``, ``pre_backward()``, ``step()``, ``get_lr()``,
``state_dict()`` / ``load_state_dict()``.  Regime phases may set ``optimizer`` (a name looked up in
``_OPTIMIZERS``), any param-group hyper-parameter (``lr``, ``momentum``, ...), ``regularizer`` (dicts with a
``name`` resolved in utils.regularization), and ``lr_scheduler``.

Two execution modes:
  * plain torch (CPU, config C1, or any non-converted model): ``torch.optim`` + per-tensor regularizer hooks,
    optional fp32 master copy for low-precision models (``ModuleFloatShadow``), as in the reference;
  * B200 (model converted by engine.convert_b200): ``SGD`` resolves to ``B200SGD`` whose ``step`` is ONE
    fused kernel over the parameter arenas, with loss-scale division, WeightDecay, GradSmooth / clipping
    folded in (see csrc/optim.cu).  Unknown regularizers still run through their torch hooks on the arena
    views.
"""
import logging
from copy import deepcopy
from math import floor

import torch
import torch.nn as nn
from torch.optim.lr_scheduler import _LRScheduler

from . import regularization
from .param_filter import FilterParameters
from .regime import Regime

_OPTIMIZERS = {name: obj for name, obj in torch.optim.__dict__.items()}
_LRSCHEDULERS = {name: obj for name, obj in torch.optim.lr_scheduler.__dict__.items()}


def cosine_anneal_lr(lr0, lrT, T, t0=0):
    return f"lambda t: {{'lr': {lrT} + {(lr0 - lrT)} * (1 + math.cos(math.pi * (t - {t0}) / {T - t0})) / 2}}"


def linear_scale_lr(lr0, lrT, T, t0=0):
    rate = (lrT - lr0) / T
    return f"lambda t: {{'lr': max({lr0} + (t - {t0}) * {rate}, 0)}}"


class _EmptySchedule(_LRScheduler):
    """Scheduler that never changes anything (the default of every regime in scope)."""

    def __init__(self, optimizer, last_epoch=-1):
        self.optimizer = optimizer
        self.base_lrs = [g['lr'] for g in optimizer.param_groups]
        self.last_epoch = 0

    def step(self, epoch=None):
        pass

    def get_lr(self):
        return [g['lr'] for g in self.optimizer.param_groups]


def copy_params(param_target, param_src):
    with torch.no_grad():
        for src, dst in zip(param_src, param_target):
            dst.copy_(src)


def copy_params_grad(param_target, param_src):
    for src, dst in zip(param_src, param_target):
        if src.grad is None:
            continue
        if dst.grad is None:
            dst.grad = src.grad.detach().to(dtype=dst.dtype).clone()
        else:
            dst.grad.detach().copy_(src.grad)


class ModuleFloatShadow(nn.Module):
    """fp32 deep copy of a low-precision module; the optimizer steps the copy (utils/optim.py:57-86)."""

    def __init__(self, module):
        super(ModuleFloatShadow, self).__init__()
        self.original_module = module
        self.float_module = deepcopy(module).to(dtype=torch.float)

    def parameters(self, *a, **k):
        return self.float_module.parameters(*a, **k)

    def named_parameters(self, *a, **k):
        return self.float_module.named_parameters(*a, **k)

    def modules(self, *a, **k):
        return self.float_module.modules(*a, **k)

    def named_modules(self, *a, **k):
        return self.float_module.named_modules(*a, **k)

    def original_parameters(self, *a, **k):
        return self.original_module.parameters(*a, **k)

    def original_named_parameters(self, *a, **k):
        return self.original_module.named_parameters(*a, **k)


class B200SGD(torch.optim.Optimizer):
    """SGD with momentum whose ``step`` is one fused kernel over the B200 parameter arenas.

    Semantics of torch.optim.SGD (momentum, dampening, weight_decay on all params, no nesterov); momentum
    buffers live in a flat fp32 arena and are exposed per parameter in ``state[p]['momentum_buffer']`` so
    ``state_dict()`` has the stock SGD layout (checkpoints interchange with the reference)."""

    def __init__(self, params, runtime=None, lr=0.0, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
        if runtime is None:
            raise ValueError('B200SGD needs the model runtime (engine.convert_b200)')
        if nesterov:
            raise NotImplementedError('nesterov momentum is not implemented in the fused kernel')
        super(B200SGD, self).__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening,
                                                   weight_decay=weight_decay, nesterov=False))
        self.rt = runtime
        arena = runtime.arena
        covered = {id(p) for g in self.param_groups for p in g['params']}
        if covered != {id(s.param) for s in arena.slots}:
            raise ValueError('B200SGD must own exactly the parameters of the converted model')
        self.m32 = torch.zeros_like(arena.p32)
        self._have_momentum = False
        # per-step extras set by OptimRegime / Trainer
        self.inv_scale = 1.0
        self.extra_wd = (0.0, 0)      # (value, arena prefix length) from a folded WeightDecay regularizer
        self.coef_dev = None          # device scalar multiplied into the gradient (clip / GradSmooth)
        # True: the kernel also clears the gradient arena (the zero_grad() of the next step, utils/optim.py:246-252
        # of the reference, folded into this pass).  Trainer switches it on; stand-alone users keep p.grad after step()
        self.fold_zero_grad = False

    def _bind_state(self):
        arena = self.rt.arena
        for s in arena.slots:
            self.state[s.param]['momentum_buffer'] = arena.logical_view(self.m32, s)

    @torch.no_grad()
    def step(self, closure=None):
        from .. import ops
        if len(self.param_groups) != 1:
            raise NotImplementedError('B200SGD supports a single parameter group')
        g = self.param_groups[0]
        arena = self.rt.arena
        wd_val, wd_count = self.extra_wd
        if g['weight_decay'] != 0:
            if wd_val != 0 and wd_count != arena.total:
                raise NotImplementedError('optimizer weight_decay together with a partial WeightDecay regularizer')
            wd_val, wd_count = wd_val + g['weight_decay'], arena.total
        first = not self._have_momentum
        ops.fused_sgd(arena.p32, arena.g32, self.m32, arena.p16, arena.total, wd_count, g['lr'], g['momentum'],
                      g['dampening'], wd_val, self.inv_scale, self.coef_dev, first, zero_grad=self.fold_zero_grad)
        arena.grads_zero = bool(self.fold_zero_grad)
        arena.version += 1       # folded inference weights / transposed shadows derived from p16 are stale
        if g['momentum'] != 0 and first:
            self._have_momentum = True
            self._bind_state()
        self.coef_dev = None

    def load_state_dict(self, state_dict):
        super(B200SGD, self).load_state_dict(state_dict)
        arena = self.rt.arena
        loaded = False
        for s in arena.slots:
            buf = self.state.get(s.param, {}).get('momentum_buffer')
            if buf is not None:
                arena.logical_view(self.m32, s).copy_(buf)
                loaded = True
        if loaded:
            self._have_momentum = True
            self._bind_state()

    def zero_grad(self, set_to_none=False):
        self.rt.arena.zero_grad()
        self.rt.arena.rebind_grads()


class OptimRegime(Regime):
    def __init__(self, model, regime, defaults={}, filter=None, use_float_copy=False, log=True):
        super(OptimRegime, self).__init__(regime, defaults)
        self._b200 = getattr(model, '_b200', None)
        if self._b200 is not None and (filter is not None or use_float_copy):
            raise NotImplementedError('B200 models keep fp32 masters in the arena: no filter / float copy')
        if filter is not None:
            model = FilterParameters(model, **filter)
        if use_float_copy:
            model = ModuleFloatShadow(model)
            self._original_parameters = list(model.original_parameters())
        self.parameters = list(model.parameters())
        self.optimizer = self._fresh_sgd()
        self.regularizer = regularization.Regularizer(model)
        self.use_float_copy = use_float_copy
        self.lr_scheduler = _EmptySchedule(self.optimizer, last_epoch=-1)
        self.schedule_time_frame = 'epoch'
        self.log = log
        self._inv_scale = 1.0
        self._clip = None
        self._device_state = {}
        self.last_grad_norm = None
        self._fold_zero_grad = False

    # ------------------------------------------------------------------ construction helpers
    def _fresh_sgd(self):
        if self._b200 is not None:
            return B200SGD(self.parameters, runtime=self._b200, lr=0)
        return torch.optim.SGD(self.parameters, lr=0)

    def _optimizer_class(self, name):
        if self._b200 is not None and name == 'SGD':
            return B200SGD
        return _OPTIMIZERS[name]

    # ------------------------------------------------------------------ regime handling
    def update(self, epoch=None, train_steps=None, metrics=None):
        """Apply the regime phase active at (epoch, train_steps); True if anything changed."""
        updated = False
        if super(OptimRegime, self).update(epoch, train_steps):
            self.adjust(self.setting)
            updated = True
        if self.schedule_time_frame == 'epoch':
            time = int(floor(epoch)) + 1
        elif self.schedule_time_frame == 'step':
            time = train_steps + 1
        else:
            raise ValueError(self.schedule_time_frame)
        if not isinstance(self.lr_scheduler, _EmptySchedule) and time != self.lr_scheduler.last_epoch \
                and getattr(self.optimizer, '_step_count', 0) > 0:
            before = self.get_lr()[0]
            if isinstance(self.lr_scheduler, torch.optim.lr_scheduler.ReduceLROnPlateau):
                self.lr_scheduler.step(metrics)
            else:
                self.lr_scheduler.step()
            updated = True
            if before != self.get_lr()[0] and self.log:
                logging.debug('OPTIMIZER - lr scheduled = %s' % self.get_lr()[0])
        return updated

    def adjust(self, setting):
        """Reconfigure optimizer / hyper-parameters / regularizers / scheduler from a setting dict."""
        reset = setting.get('reset', False)
        if 'optimizer' in setting or reset:
            cls = self._optimizer_class(setting.get('optimizer', 'SGD'))
            if reset:
                self.optimizer = self._fresh_sgd()
            if not isinstance(self.optimizer, cls):
                if cls is B200SGD:
                    self.optimizer = B200SGD(self.optimizer.param_groups, runtime=self._b200)
                else:
                    self.optimizer = cls(self.optimizer.param_groups)
                if self.log:
                    logging.debug('OPTIMIZER - setting method = %s' % setting.get('optimizer'))
        for group in self.optimizer.param_groups:
            for key in list(group.keys()):
                if key in setting and setting[key] != group[key]:
                    if self.log:
                        logging.debug('OPTIMIZER - setting %s = %s' % (key, setting[key]))
                    group[key] = setting[key]
                    if key == 'lr':
                        group['initial_lr'] = group['lr']
                        base_lrs = [g['lr'] for g in self.optimizer.param_groups]
                        self.lr_scheduler.base_lrs = base_lrs
                        if hasattr(self.optimizer, 'base_lrs'):
                            self.optimizer.base_lrs = base_lrs

        if 'regularizer' in setting:
            spec = deepcopy(setting['regularizer'])
            if not isinstance(spec, (list, tuple)):
                spec = (spec,)
            built = []
            for reg in spec:
                if isinstance(reg, dict):
                    name = reg.pop('name')
                    built.append((regularization.__dict__[name], reg))
                elif isinstance(reg, regularization.Regularizer):
                    built.append(reg)
                else:
                    built.append(reg(self.regularizer._model))
            previous = self.regularizer
            self.regularizer = regularization.RegularizerList(self.regularizer._model, built)
            self._carry_regularizer_state(previous, self.regularizer)

        if 'lr_scheduler' in setting:
            cfg = setting['lr_scheduler']
            if isinstance(cfg, _LRScheduler):
                self.lr_scheduler = cfg
            elif isinstance(cfg, dict):
                cfg = dict(cfg)
                name = cfg.pop('name')
                self.schedule_time_frame = cfg.pop('time_frame', 'epoch')
                cfg['last_epoch'] = self.lr_scheduler.last_epoch
                self.lr_scheduler = _LRSCHEDULERS[name](self.optimizer, **cfg)
            elif cfg is None:
                self.lr_scheduler = _EmptySchedule(self.optimizer, last_epoch=self.lr_scheduler.last_epoch)
            else:
                raise NotImplementedError(cfg)

    def _carry_regularizer_state(self, old, new):
        """The reference rebuilds the RegularizerList on every adjust(), which resets
        GradSmooth.running_norm whenever the setting changes.  We keep that observable behaviour for the torch
        path; the device-side smoothing state is keyed by position and reset the same way."""
        self._device_state = {}

    # ------------------------------------------------------------------ per-step API
    def zero_grad(self):
        if self._b200 is not None:
            self._b200.arena.zero_grad()
            self._b200.arena.rebind_grads()
            return
        self.optimizer.zero_grad()
        if self.use_float_copy:
            for p in self._original_parameters:
                if p.grad is not None:
                    p.grad.detach().zero_()

    def pre_forward(self):
        self.regularizer.pre_forward()

    def pre_backward(self):
        self.regularizer.pre_backward()

    # hooks used by Trainer on the B200 path instead of per-tensor loops (trainer.py:165-172)
    def set_grad_unscale(self, loss_scale, world_size=1):
        self._inv_scale = 1.0 / (float(loss_scale) * float(world_size))

    def request_clip(self, max_norm):
        self._clip = float(max_norm)

    def fold_zero_grad(self, enabled=True):
        """B200: let the fused SGD kernel clear the gradient arena in its own pass (the next zero_grad() is then free).
        p.grad reads zero after step(); Trainer enables it, nothing else does."""
        self._fold_zero_grad = bool(enabled)

    def step(self, *args, **kwargs):
        if self._b200 is not None and isinstance(self.optimizer, B200SGD):
            self.optimizer.fold_zero_grad = self._fold_zero_grad
            return self._step_b200()
        if self._b200 is not None:
            # any other torch optimizer on the arena views: the unscale (loss scale x world size) and the clip that
            # Trainer delegates to this class (set_grad_unscale / request_clip) are applied here, as in the
            # non-foldable branch of _step_b200 (reference: trainer.py:165-172)
            if self._inv_scale != 1.0:
                self._b200.arena.g32.mul_(self._inv_scale)
            if self._clip is not None and self._clip > 0:
                self.last_grad_norm = torch.nn.utils.clip_grad_norm_(self.parameters, self._clip)
            self._clip = None
        if self.use_float_copy:
            copy_params_grad(self.parameters, self._original_parameters)
        self.regularizer.pre_step()
        self.optimizer.step(*args, **kwargs)
        self.regularizer.post_step()
        if self.use_float_copy:
            copy_params(self._original_parameters, self.parameters)
        if self._b200 is not None:
            self._b200.arena.sync_shadow()

    def _step_b200(self):
        """Fold unscale / clip / GradSmooth / WeightDecay into the fused arena kernel when the regularizer
        list has the foldable shape; otherwise apply the unscale with a torch op and run the hooks."""
        from .. import ops
        rt, opt = self._b200, self.optimizer
        arena = rt.arena
        regs = getattr(self.regularizer, 'regularization_list', [])
        wd, smooth, others = None, None, []
        seen_wd = False
        foldable = True
        for r in regs:
            if type(r) in (regularization.WeightDecay, regularization.L2Regularization) and r.pre_op \
                    and not r.post_op and wd is None:
                count = self._wd_prefix(r, arena)
                if count is None:
                    foldable = False
                wd, seen_wd = (r.value, count), True
            elif type(r) is regularization.GradSmooth and smooth is None and not seen_wd and not r._named_parameters == []:
                smooth = r
                if len(r._named_parameters) != len(arena.slots):
                    foldable = False
            else:
                others.append(r)
        if others:
            foldable = False
        if not foldable:
            if self._inv_scale != 1.0:
                arena.g32.mul_(self._inv_scale)
            if self._clip is not None and self._clip > 0:
                self.last_grad_norm = torch.nn.utils.clip_grad_norm_(self.parameters, self._clip)
            opt.inv_scale, opt.extra_wd, opt.coef_dev = 1.0, (0.0, 0), None
            self.regularizer.pre_step()
            opt.step()
            self.regularizer.post_step()
            self._clip = None
            return
        st = self._device_state
        if 'buf' not in st:
            st['buf'] = torch.zeros(8, device=arena.p32.device, dtype=torch.float32)
            st['ws'] = torch.empty(1024, device=arena.p32.device, dtype=torch.float32)
        buf = st['buf']  # [0]=sumsq [1]=coef [2]=norm [3]=clip coef scratch [4:6]=GradSmooth state
        coef = None
        if (self._clip is not None and self._clip > 0) or smooth is not None:
            ops.sumsq(arena.g32, arena.total, buf[0:1], st['ws'])
            if self._clip is not None and self._clip > 0:
                ops.grad_coef(buf[0:1], self._inv_scale, 0, self._clip, 0.0, None, buf[1:2], buf[2:3])
                self.last_grad_norm = buf[2]
                coef = buf[1:2]
                if smooth is not None:
                    raise NotImplementedError('grad clipping together with GradSmooth')
            else:
                ops.grad_coef(buf[0:1], self._inv_scale, 1, 0.0, smooth.momentum, buf[4:6], buf[1:2], buf[2:3])
                smooth.counter += 1
                coef = buf[1:2] if smooth.enabled else None
        opt.inv_scale = self._inv_scale
        opt.extra_wd = wd if wd is not None else (0.0, 0)
        opt.coef_dev = coef
        opt.step()
        self._clip = None

    @staticmethod
    def _wd_prefix(reg, arena):
        """Length of the arena prefix equal to the regularizer's parameter set, or None if it is not a prefix."""
        ids = {id(p) for _, p in reg._named_parameters}
        for end in (arena.group_end[0], arena.group_end[1], arena.group_end[2]):
            prefix = {id(s.param) for s in arena.slots if s.offset < end}
            if prefix == ids:
                return end
        return None

    # ------------------------------------------------------------------ state / introspection
    def __getstate__(self):
        return {'optimizer_state': self.optimizer.__getstate__(), 'regime': self.regime}

    def __setstate__(self, state):
        self.regime = state.get('regime')
        self.optimizer.__setstate__(state.get('optimizer_state'))

    def state_dict(self):
        return self.optimizer.state_dict()

    def load_state_dict(self, state_dict):
        self.optimizer.load_state_dict(state_dict)

    def get_value(self, key):
        return [group[key] for group in self.optimizer.param_groups]

    def get_lr(self):
        return self.get_value('lr')

    @property
    def state(self):
        return self.optimizer.state


class MultiOptimRegime(OptimRegime):
    def __init__(self, *optim_regime_list, log=True):
        self.optim_regime_list = list(optim_regime_list)
        assert all(isinstance(o, OptimRegime) for o in self.optim_regime_list)
        self.log = log

    def update(self, epoch=None, train_steps=None):
        flags = [o.update(epoch, train_steps) for o in self.optim_regime_list]
        return any(flags)

    def zero_grad(self):
        for o in self.optim_regime_list:
            o.zero_grad()

    def step(self):
        for o in self.optim_regime_list:
            o.step()

    def pre_forward(self):
        for o in self.optim_regime_list:
            o.pre_forward()

    def pre_backward(self):
        for o in self.optim_regime_list:
            o.pre_backward()

    def get_value(self, key):
        return [[g[key] for g in o.optimizer.param_groups] for o in self.optim_regime_list]

    def get_lr(self):
        return self.get_value('lr')
