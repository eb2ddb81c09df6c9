"""Weights of the self-play network, `snetwork2` (DenseNet.jl:279-304): a bias-free residual MLP with a
policy head (+bias) and a sigmoid value head (+bias) — and of the network that is trained, `networkf`
(DenseNet.jl:161-198), which adds a tanh `feature` head.  Arrays keep the Julia shapes and are stored
Fortran-ordered, so their bytes are exactly what `convert_back(net)` (DenseNet.jl:331-333) hands to
`agpu_set_weights` and what `Flux.params(net)` holds for `agpu_trainer_set_params`.

Checkpoints: the reference `JLD2.@save`s `to_cpu(net)` (selfplay.jl:86-98); here `save_network` / `load_network` write
the same arrays as raw little-endian fp32 blobs behind a small JSON header, which Julia glue can wrap in JLD2.
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass
from typing import List, Union

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


@dataclass
class NetworkF:
    """`networkf` (DenseNet.jl:161-166): base, res, policy, value, feature — the parameters train.jl updates."""
    base: np.ndarray          # (n, 2*VectorizedState)
    res: List[np.ndarray]     # k x (n, n)
    policy: np.ndarray        # (maxActions, n)
    policy_bias: np.ndarray   # (maxActions,)
    value: np.ndarray         # (1, n)
    value_bias: np.ndarray    # (1,)
    feature: np.ndarray       # (FeatureSize, n)
    feature_bias: np.ndarray  # (FeatureSize,)

    def __post_init__(self):
        f = lambda a: np.asfortranarray(np.asarray(a, dtype=np.float32))
        self.base, self.res = f(self.base), [f(r) for r in self.res]
        self.policy, self.policy_bias = f(self.policy), f(self.policy_bias).reshape(-1)
        self.value, self.value_bias = f(self.value).reshape(1, -1), f(self.value_bias).reshape(-1)
        self.feature, self.feature_bias = f(self.feature), f(self.feature_bias).reshape(-1)

    @property
    def width(self): return self.base.shape[0]
    @property
    def blocks(self): return len(self.res)
    @property
    def in_features(self): return self.base.shape[1]
    @property
    def actions(self): return self.policy.shape[0]
    @property
    def fsize(self): return self.feature.shape[0]

    def arrays(self):
        return [self.base, *self.res, self.policy, self.policy_bias, self.value, self.value_bias, self.feature, self.feature_bias]

    def copy(self) -> "NetworkF":   # deepcopy(trainingnet) (selfplay.jl:68)
        return NetworkF(self.base.copy(), [r.copy() for r in self.res], self.policy.copy(), self.policy_bias.copy(), self.value.copy(),
                        self.value_bias.copy(), self.feature.copy(), self.feature_bias.copy())


def convert_back(net: Union[NetworkF, SNetwork2]) -> SNetwork2:
    """convert_back(net) (DenseNet.jl:331-337): the actor the search evaluates — everything but the feature head."""
    return SNetwork2(net.base, list(net.res), net.policy, net.policy_bias, net.value, net.value_bias)


def ressimplesf_full(in_features: int, out_actions: int, fsize: int, n_filter: int, n_tower: int, seed: int = 0) -> NetworkF:
    """ressimplesf(in, out, fsize, n_filter, n_tower) (DenseNet.jl:193-198) with all three heads.  The draws for base, res, policy
    and value are those of `ressimplesf(in, out, n_filter, n_tower, seed)`, so `convert_back` of this equals that net."""
    rng = np.random.default_rng(seed)
    glorot = lambda o, i: (rng.uniform(-1.0, 1.0, size=(o, i)) * np.sqrt(6.0 / (o + i))).astype(np.float32)
    base, res = glorot(n_filter, in_features), [glorot(n_filter, n_filter) for _ in range(n_tower)]
    pol, val = glorot(out_actions, n_filter), glorot(1, n_filter)
    feat = glorot(fsize, n_filter)
    return NetworkF(base, res, pol, np.zeros(out_actions, np.float32), val, np.zeros(1, np.float32), feat, np.zeros(fsize, np.float32))


_MAGIC = b"AGPUNET1"


def save_network(path: str, net: Union[NetworkF, SNetwork2], meta: dict | None = None) -> None:
    """Checkpoint (selfplay.jl:86-98 `JLD2.@save ... reseau`): magic, u32 header length, JSON header {kind, arrays: [{name, shape}],
    meta}, then every array as little-endian fp32 in Julia column-major order."""
    kind = "networkf" if isinstance(net, NetworkF) else "snetwork2"
    names = ["base"] + [f"res{i}" for i in range(net.blocks)] + ["policy", "policy_bias", "value", "value_bias"]
    arrs = [net.base, *net.res, net.policy, net.policy_bias, net.value, net.value_bias]
    if kind == "networkf":
        names += ["feature", "feature_bias"]
        arrs += [net.feature, net.feature_bias]
    hdr = json.dumps(dict(kind=kind, arrays=[dict(name=n, shape=list(a.shape)) for n, a in zip(names, arrs)], meta=meta or {})).encode()
    with open(path, "wb") as f:
        f.write(_MAGIC + struct.pack("<I", len(hdr)) + hdr)
        for a in arrs:
            f.write(np.asarray(a, "<f4").tobytes(order="F"))


def load_network(path: str):
    """Inverse of save_network -> (NetworkF | SNetwork2, meta)."""
    with open(path, "rb") as f:
        if f.read(8) != _MAGIC:
            raise ValueError(f"{path}: not an alphagpu_b200 network checkpoint")
        (n,) = struct.unpack("<I", f.read(4))
        hdr = json.loads(f.read(n).decode())
        arrs = {}
        for a in hdr["arrays"]:
            cnt = int(np.prod(a["shape"]))
            buf = f.read(4 * cnt)
            if len(buf) != 4 * cnt:
                raise ValueError(f"{path}: truncated at array {a['name']}")
            arrs[a["name"]] = np.frombuffer(buf, "<f4").reshape(a["shape"], order="F").copy(order="F")
    k = sum(1 for a in arrs if a.startswith("res"))
    res = [arrs[f"res{i}"] for i in range(k)]
    if hdr["kind"] == "networkf":
        net = NetworkF(arrs["base"], res, arrs["policy"], arrs["policy_bias"], arrs["value"], arrs["value_bias"], arrs["feature"], arrs["feature_bias"])
    else:
        net = SNetwork2(arrs["base"], res, arrs["policy"], arrs["policy_bias"], arrs["value"], arrs["value_bias"])
    return net, hdr.get("meta", {})
