"""Synthetic objectives used to build the BASELINE configs — NumPy restatements of
trieste/objectives/single_objectives.py (branin :83-107, ackley_5 :433-458 generalised to d dims,
hartmann_6 :476-501).  Input generators only; not on the per-candidate path."""
from __future__ import annotations

import math

import numpy as np


def branin(x):
    x = np.asarray(x, dtype=np.float64)
    x0 = x[..., :1] * 15.0 - 5.0
    x1 = x[..., 1:] * 15.0
    b = 5.1 / (4 * math.pi**2)
    c = 5 / math.pi
    t = 1 / (8 * math.pi)
    return (x1 - b * x0**2 + c * x0 - 6) ** 2 + 10 * (1 - t) * np.cos(x0) + 10


def scaled_branin(x):
    x = np.asarray(x, dtype=np.float64)
    x0 = x[..., :1] * 15.0 - 5.0
    x1 = x[..., 1:] * 15.0
    b = 5.1 / (4 * math.pi**2)
    c = 5 / math.pi
    t = 1 / (8 * math.pi)
    return (1 / 51.95) * ((x1 - b * x0**2 + c * x0 - 6) ** 2 + 10 * (1 - t) * np.cos(x0) - 44.81)


def ackley(x):
    x = np.asarray(x, dtype=np.float64)
    d = x.shape[-1]
    x = (x - 0.5) * (32.768 * 2.0)
    e1 = -0.2 * np.sqrt((1.0 / d) * np.square(x).sum(-1))
    e2 = (1.0 / d) * np.cos(2.0 * math.pi * x).sum(-1)
    return (-20.0 * np.exp(e1) - np.exp(e2) + 20.0 + math.e)[..., None]


_H6_A = np.array(
    [[10.0, 3.0, 17.0, 3.5, 1.7, 8.0], [0.05, 10.0, 17.0, 0.1, 8.0, 14.0], [3.0, 3.5, 1.7, 10.0, 17.0, 8.0], [17.0, 8.0, 0.05, 10.0, 0.1, 14.0]]
)
_H6_P = np.array(
    [
        [0.1312, 0.1696, 0.5569, 0.0124, 0.8283, 0.5886],
        [0.2329, 0.4135, 0.8307, 0.3736, 0.1004, 0.9991],
        [0.2348, 0.1451, 0.3522, 0.2883, 0.3047, 0.6650],
        [0.4047, 0.8828, 0.8732, 0.5743, 0.1091, 0.0381],
    ]
)


def hartmann_6(x):
    x = np.asarray(x, dtype=np.float64)
    a = np.array([1.0, 1.2, 3.0, 3.2])
    inner = -(_H6_A * (x[..., None, :] - _H6_P) ** 2).sum(-1)
    return -(a * np.exp(inner)).sum(-1, keepdims=True)
