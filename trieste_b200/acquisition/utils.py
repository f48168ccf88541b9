"""``split_acquisition_function`` / ``split_acquisition_function_calls`` — mirrors trieste/acquisition/utils.py:31-109.

In the reference these wrappers bound the memory of one TensorFlow evaluation by cutting the leading (candidate) axis into
blocks.  Here the fused kernels already stream any batch through bounded scratch (``run_eval`` / ``run_eval_oz`` chunk at
~1 GB of K* digits), so the wrappers are not needed for memory; they exist so that callers which wrap their functions or
optimisers keep working, with the reference's splitting rule and error behaviour."""
from __future__ import annotations

import functools
import math

import numpy as np


def _concat(parts):
    if type(parts[0]).__module__.split(".")[0] == "torch":
        import torch

        return torch.cat(parts, dim=0)
    return np.concatenate([np.asarray(p) for p in parts], axis=0)


def split_acquisition_function(fn, split_size: int):
    """utils.py:31-84: call ``fn`` on blocks of at most ``split_size`` ELEMENTS of ``x`` along its first axis and stitch the
    results back together."""
    if split_size <= 0:
        raise ValueError(f"split_size must be positive, got {split_size}")

    @functools.wraps(fn, updated=())
    def wrapper(x):
        x = x if hasattr(x, "shape") else np.asarray(x)
        length = x.shape[0]
        if length == 0:
            return fn(x)
        elements_per_block = int(np.prod(x.shape)) / length
        blocks_per_batch = int(math.ceil(split_size / elements_per_block))
        if length <= blocks_per_batch:
            return fn(x)
        return _concat([fn(x[i : i + blocks_per_batch]) for i in range(0, length, blocks_per_batch)])

    # the fused entry points pass straight through: they never materialise more than one bounded chunk
    for name in ("value_and_gradient", "fused_argmax", "maximize_from"):
        if hasattr(fn, name):
            setattr(wrapper, name, getattr(fn, name))
    return wrapper


def split_acquisition_function_calls(optimizer, split_size: int):
    """utils.py:87-109: an optimiser whose acquisition-function evaluations are split as above."""
    if split_size <= 0:
        raise ValueError(f"split_size must be positive, got {split_size}")

    def split_optimizer(search_space, f):
        af, n = f if isinstance(f, tuple) else (f, 1)
        taf = split_acquisition_function(af, split_size)
        return optimizer(search_space, (taf, n) if isinstance(f, tuple) else taf)

    return split_optimizer
