"""Weights of the self-play network, `snetwork2` (DenseNet.jl:279-304): a bias-free residual MLP with a
policy head (+bias) and a sigmoid value head (+bias).  Arrays keep the Julia shapes and are stored
Fortran-ordered, so their bytes are exactly what `convert_back(net)` (DenseNet.jl:331-333) hands to
`agpu_set_weights`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class SNetwork2:
    base: np.ndarray          # (n, 2*VectorizedState)
    res: List[np.ndarray]     # k x (n, n)
    policy: np.ndarray        # (maxActions, n)
    policy_bias: np.ndarray   # (maxActions,)
    value: np.ndarray         # (1, n)
    value_bias: np.ndarray    # (1,)

    def __post_init__(self):
        f = lambda a: np.asfortranarray(np.asarray(a, dtype=np.float32))
        self.base, self.res = f(self.base), [f(r) for r in self.res]
        self.policy, self.policy_bias = f(self.policy), f(self.policy_bias).reshape(-1)
        self.value, self.value_bias = f(self.value).reshape(1, -1), f(self.value_bias).reshape(-1)

    @property
    def width(self): return self.base.shape[0]
    @property
    def blocks(self): return len(self.res)
    @property
    def nbytes(self): return sum(a.nbytes for a in [self.base, *self.res, self.policy, self.policy_bias, self.value, self.value_bias])


def ressimplesf(in_features: int, out_actions: int, n_filter: int, n_tower: int, seed: int = 0) -> SNetwork2:
    """Random-init net as `ressimplesf(in, out, fsize, n_filter, n_tower)` would build it (DenseNet.jl:193-198):
    Flux 0.12 `Dense` = Glorot-uniform weights U(±sqrt(6/(fan_in+fan_out))), no trunk biases, zero head biases.
    (The `feature` head exists only for training and is not part of `convert_back`.)"""
    rng = np.random.default_rng(seed)
    glorot = lambda o, i: (rng.uniform(-1.0, 1.0, size=(o, i)) * np.sqrt(6.0 / (o + i))).astype(np.float32)
    return SNetwork2(glorot(n_filter, in_features), [glorot(n_filter, n_filter) for _ in range(n_tower)], glorot(out_actions, n_filter),
                     np.zeros(out_actions, np.float32), glorot(1, n_filter), np.zeros(1, np.float32))
