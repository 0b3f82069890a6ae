"""Select parameters by module / module-name / parameter-name predicates.

Same contract as the reference's utils/param_filter.py:5-76; used to build the weight-decay set
(models/resnet.py:34-40) and the "keep BN in float" set (main.py:240).
"""
import torch.nn as nn


def is_not_bias(name):
    return not name.endswith('bias')


def is_bn(module):
    return isinstance(module, nn.modules.batchnorm._BatchNorm)


def is_not_bn(module):
    return not is_bn(module)


def _walk(model, module_fn, module_name_fn, parameter_name_fn):
    seen = set()
    for mod_name, mod in model.named_modules():
        if module_fn is not None and not module_fn(mod):
            continue
        if module_name_fn is not None and not module_name_fn(mod_name):
            continue
        for p_name, p in mod.named_parameters(prefix=mod_name, recurse=False):
            if parameter_name_fn is not None and not parameter_name_fn(p_name):
                continue
            if id(p) in seen:
                continue
            seen.add(id(p))
            yield {'named_module': (mod_name, mod), 'named_parameter': (p_name, p)}


class FilterParameters(object):
    def __init__(self, source, module=None, module_name=None, parameter_name=None):
        if isinstance(source, FilterParameters):
            self._filtered_parameter_info = list(source.filter(module=module, module_name=module_name,
                                                               parameter_name=parameter_name))
        elif isinstance(source, nn.Module):
            self._filtered_parameter_info = list(_walk(source, module, module_name, parameter_name))
        else:
            raise TypeError('FilterParameters expects an nn.Module or another FilterParameters')

    def filter(self, module=None, module_name=None, parameter_name=None):
        for info in self._filtered_parameter_info:
            mod_name, mod = info['named_module']
            p_name, _ = info['named_parameter']
            if module is not None and not module(mod):
                continue
            if module_name is not None and not module_name(mod_name):
                continue
            if parameter_name is not None and not parameter_name(p_name):
                continue
            yield info

    def named_parameters(self):
        return (info['named_parameter'] for info in self._filtered_parameter_info)

    def parameters(self):
        return (p for _, p in self.named_parameters())

    def named_modules(self):
        return (info['named_module'] for info in self._filtered_parameter_info)

    def modules(self):
        return (m for _, m in self.named_modules())

    def to(self, *args, **kwargs):
        for m in self.modules():
            m.to(*args, **kwargs)


class FilterModules(FilterParameters):
    pass
