"""Phase-based setting schedules ("regimes"): same grammar as the reference's utils/regime.py:13-96.

A regime is a list of dicts; each may carry ``epoch`` and/or ``step`` (phase start) and arbitrary setting
keys.  Settings accumulate from phase to phase.  Special keys: ``step_lambda`` / ``epoch_lambda`` (a callable
or a string evaluated to one, returning a dict of overrides), ``lr_decay_rate`` (+ ``lr_decay_steps``),
``execute`` / ``execute_once`` (callbacks).
"""
import math  # noqa: F401  (available to string lambdas, as in the reference)
from copy import deepcopy


def eval_func(f, x):
    if isinstance(f, str):
        f = eval(f)  # regimes are trusted configuration, exactly as in the reference (utils/regime.py:7-10)
    return f(x)


class Regime(object):
    def __init__(self, regime, defaults={}):
        self.regime = regime
        self.defaults = defaults
        self.reset(regime, defaults)

    def reset(self, regime=None, defaults=None):
        if regime is not None:
            self.regime = regime
        if defaults is not None:
            self.defaults = defaults
        self.current_regime_phase = None
        self.setting = self.defaults

    def _locate_initial_phase(self, epoch, steps, setting):
        for idx, phase in enumerate(self.regime):
            if epoch >= phase.get('epoch', 0) or steps >= phase.get('step', 0):
                self.current_regime_phase = idx
                return
            setting.update(phase)  # phases before the first active one still contribute their keys

    def update(self, epoch=None, train_steps=None):
        """Advance to the phase active at (epoch, train_steps); True iff the effective setting changed."""
        if self.regime is None:
            return False
        epoch = -1 if epoch is None else epoch
        steps = -1 if train_steps is None else train_steps
        setting = deepcopy(self.setting)
        if self.current_regime_phase is None:
            self._locate_initial_phase(epoch, steps, setting)
        nxt = self.current_regime_phase + 1
        if nxt < len(self.regime):
            phase = self.regime[nxt]
            if epoch >= phase.get('epoch', float('inf')) or steps >= phase.get('step', float('inf')):
                self.current_regime_phase = nxt
        setting.update(self.regime[self.current_regime_phase])

        if 'lr_decay_rate' in setting and 'lr' in setting:
            every = setting.pop('lr_decay_steps', 100)
            if steps % every == 0:
                setting['lr'] *= setting.pop('lr_decay_rate') ** (steps / every)
        elif 'step_lambda' in setting:
            setting.update(eval_func(setting.pop('step_lambda'), steps))
        elif 'epoch_lambda' in setting:
            setting.update(eval_func(setting.pop('epoch_lambda'), epoch))

        if 'execute' in setting:
            setting.pop('execute')()
        if 'execute_once' in setting:
            setting.pop('execute_once')()
            self.regime[self.current_regime_phase].pop('execute_once', None)

        if setting == self.setting:
            return False
        self.setting = setting
        return True

    def __repr__(self):
        return 'Current: %s\n Regime:%s' % (self.setting, self.regime)
