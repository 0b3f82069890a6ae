"""Small helpers mirroring the reference's utils/misc.py (dtype names, one-hot, seeding).

Difference from the reference (utils/misc.py:7-22): the dtype table also knows ``bfloat16`` / ``bf16``,
which is the compute type of the B200 path.
"""
import random

import numpy as np
import torch

_NAMES = ("float", "float32", "float64", "double", "float16", "half", "bfloat16", "uint8", "int8", "int16",
          "short", "int32", "int", "int64", "long")
torch_dtypes = {name: getattr(torch, name) for name in _NAMES}
torch_dtypes["bf16"] = torch.bfloat16


def is_low_precision(dtype_name):
    """True for the dtypes the reference treats as 'half' (main.py:239-240,250) plus bfloat16."""
    name = str(dtype_name)
    return "half" in name or "float16" in name or "bf16" in name


def onehot(indexes, N=None, ignore_index=None):
    """One-hot encode a LongTensor along a new last dimension (cf. utils/misc.py:25-39)."""
    if N is None:
        N = int(indexes.max()) + 1
    encoded = torch.zeros(*indexes.shape, N, dtype=torch.uint8, device=indexes.device)
    encoded.scatter_(-1, indexes.unsqueeze(-1), 1)
    if ignore_index is not None and ignore_index >= 0:
        encoded.masked_fill_(indexes.eq(ignore_index).unsqueeze(-1), 0)
    return encoded


def set_global_seeds(seed):
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)
