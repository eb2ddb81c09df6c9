"""Host-side mirror of train.jl: `lossTot`, `custom_train!` and `traininPipe` over the C ABI of include/alphagpu_train.h.

The reference trains a `networkf` with Flux/Zygote on one GPU (train.jl:47-126).  Here one `Trainer` owns the parameters,
Adam state and workspaces on its GPU; a step is `loss_grad` (forward + backward into one flat gradient) followed by `apply`
(ADAM + WeightDecay).  Data-parallel training (BASELINE config 4: "NCCL sample gather + train allreduce") splits every batch
over the ranks of the default `torch.distributed` group and all-reduces the flat gradient in place between the two calls —
the one collective of the training path.  There is no CPU fallback: without the CUDA library `Trainer()` raises.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Optional

import numpy as np

from . import _lib
from .densenet import NetworkF


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _check(h, rc):
    if rc != _lib.OK:
        msg = _lib.load().agpu_trainer_last_error(h)
        raise _lib.AlphaGPUError(rc, msg.decode() if msg else "")


class _DeviceArray:
    """A raw device pointer dressed in __cuda_array_interface__ so torch can alias it without a copy."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = dict(shape=(count,), typestr="<f4", data=(ptr, False), version=3, strides=None)


class Trainer:
    """trainingnet + opt (train.jl:47-51) on one GPU."""

    def __init__(self, in_features: int, width: int, blocks: int, actions: int, fsize: int, max_batch: int, *, device: int = 0, lr: float = 0.001,
                 beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 1e-4, feature_weight: float = 0.001):
        self.lib = _lib.load()
        self.inp, self.n, self.k, self.A, self.FS, self.max_batch, self.device = in_features, width, blocks, actions, fsize, max_batch, device
        cfg = _lib.TrainConfig(device, in_features, width, blocks, actions, fsize, max_batch, 0, lr, beta1, beta2, eps, weight_decay,
                               feature_weight, 0.0)
        h = C.c_void_p()
        _check(None, self.lib.agpu_trainer_create(C.byref(h), C.byref(cfg)))
        self.h = h
        self._grad_tensor = None

    @classmethod
    def for_network(cls, net: NetworkF, max_batch: int, **kw) -> "Trainer":
        tr = cls(net.in_features, net.width, net.blocks, net.actions, net.fsize, max_batch, **kw)
        tr.set_params(net)
        return tr

    def close(self):
        if getattr(self, "h", None):
            self._grad_tensor = None
            self.lib.agpu_trainer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters ----
    def _blank(self) -> NetworkF:
        z = lambda *s: np.zeros(s, np.float32, order="F")
        return NetworkF(z(self.n, self.inp), [z(self.n, self.n) for _ in range(self.k)], z(self.A, self.n), z(self.A), z(1, self.n), z(1),
                        z(self.FS, self.n), z(self.FS))

    @staticmethod
    def _ptrs(net: NetworkF):
        arr = (C.c_void_p * max(1, net.blocks))(*[r.ctypes.data for r in net.res])
        return arr, [_p(net.base), C.cast(arr, C.c_void_p), _p(net.policy), _p(net.policy_bias), _p(net.value), _p(net.value_bias),
                     _p(net.feature), _p(net.feature_bias)]

    def set_params(self, net: NetworkF, reset_optimizer: bool = True):
        assert (net.in_features, net.width, net.blocks, net.actions, net.fsize) == (self.inp, self.n, self.k, self.A, self.FS), "shape mismatch"
        keep, ptrs = self._ptrs(net)
        _check(self.h, self.lib.agpu_trainer_set_params(self.h, *ptrs, int(reset_optimizer)))

    def get_params(self) -> NetworkF:
        net = self._blank()
        keep, ptrs = self._ptrs(net)
        _check(self.h, self.lib.agpu_trainer_get_params(self.h, *ptrs))
        return net

    def get_grads(self) -> NetworkF:
        net = self._blank()
        keep, ptrs = self._ptrs(net)
        _check(self.h, self.lib.agpu_trainer_get_grads(self.h, *ptrs))
        return net

    def opt_state(self):
        _, count = self.grad_buffer()
        m, v, bp = np.zeros(count, np.float32), np.zeros(count, np.float32), np.zeros(2, np.float64)
        _check(self.h, self.lib.agpu_trainer_opt_state(self.h, _p(m), _p(v), _p(bp), 0))
        return m, v, bp

    def set_opt_state(self, m, v, bp):
        m, v, bp = np.ascontiguousarray(m, np.float32), np.ascontiguousarray(v, np.float32), np.ascontiguousarray(bp, np.float64)
        _check(self.h, self.lib.agpu_trainer_opt_state(self.h, _p(m), _p(v), _p(bp), 1))

    # ---- one batch ----
    def _batch(self, state, policy, value, fstate):
        state = np.ascontiguousarray(state, np.int8); policy = np.ascontiguousarray(policy, np.float32)
        value = np.ascontiguousarray(value, np.float32).reshape(-1); fstate = np.ascontiguousarray(fstate, np.int8)
        B = state.shape[0]
        if not (state.shape == (B, self.inp) and policy.shape == (B, self.A) and value.shape == (B,) and fstate.shape == (B, self.FS)):
            raise ValueError("batch arrays do not match the network: state (B,in) int8, policy (B,A) f32, value (B) f32, fstate (B,FS) int8")
        return B, (state, policy, value, fstate)

    def loss_grad(self, state, policy, value, fstate) -> np.ndarray:
        """gradient(ps) do lossTot(net, x, y) end (train.jl:133-136) -> [total, policy, value, feature] loss."""
        B, arrs = self._batch(state, policy, value, fstate)
        out = np.zeros(4, np.float32)
        _check(self.h, self.lib.agpu_trainer_loss_grad(self.h, *[_p(a) for a in arrs], B, _p(out)))
        return out

    def lossTot(self, state, policy, value, fstate) -> np.ndarray:
        """lossTot(net, x, y) (train.jl:12-15) without the gradient."""
        B, arrs = self._batch(state, policy, value, fstate)
        out = np.zeros(4, np.float32)
        _check(self.h, self.lib.agpu_trainer_loss(self.h, *[_p(a) for a in arrs], B, _p(out)))
        return out

    def apply(self, grad_scale: float = 1.0):
        """Flux.update!(opt, ps, gs) (train.jl:158)."""
        _check(self.h, self.lib.agpu_trainer_apply(self.h, float(grad_scale)))

    def step(self, state, policy, value, fstate) -> np.ndarray:
        B, arrs = self._batch(state, policy, value, fstate)
        out = np.zeros(4, np.float32)
        _check(self.h, self.lib.agpu_trainer_step(self.h, *[_p(a) for a in arrs], B, _p(out)))
        return out

    def last_ms(self):
        ms = np.zeros(2, np.float32)
        _check(self.h, self.lib.agpu_trainer_last_ms(self.h, _p(ms)))
        return float(ms[0]), float(ms[1])

    # ---- data parallel ----
    def grad_buffer(self):
        ptr, cnt = C.c_void_p(), C.c_int64()
        _check(self.h, self.lib.agpu_trainer_grad_buffer(self.h, C.byref(ptr), C.byref(cnt)))
        return int(ptr.value), int(cnt.value)

    def grad_tensor(self):
        """The flat device gradient as a torch tensor aliasing the library's buffer (for dist.all_reduce)."""
        if self._grad_tensor is None:
            import torch
            ptr, cnt = self.grad_buffer()
            self._grad_tensor = torch.as_tensor(_DeviceArray(ptr, cnt), device=f"cuda:{self.device}")
        return self._grad_tensor

    def step_dp(self, state, policy, value, fstate, group=None) -> np.ndarray:
        """One data-parallel step: this rank's shard of the batch, sum all-reduce of the flat gradient over NCCL, ADAM with the
        gradient scaled by 1/world (the mean over the global batch when shards are equal).  Returns the losses averaged over ranks."""
        import torch
        import torch.distributed as dist
        loss = self.loss_grad(state, policy, value, fstate)
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world > 1:
            g = self.grad_tensor()
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
            lt = torch.from_numpy(loss.astype(np.float64)).to(g.device)
            dist.all_reduce(lt, op=dist.ReduceOp.SUM, group=group)
            torch.cuda.synchronize(g.device)
            loss = (lt.cpu().numpy() / world).astype(np.float32)
        self.apply(1.0 / world)
        return loss


def dp_slice(n: int, rank: int, world: int) -> slice:
    """Rows of a batch of n samples that rank `rank` of `world` trains on (equal contiguous shards; n % world rows at the end of the
    batch are dropped so that every rank's mean has the same weight)."""
    per = n // world
    return slice(rank * per, (rank + 1) * per)


def traininPipe(batchsize: int, net: NetworkF, p, *, epoch: int = 1, lr: float = 0.001, seed: int = 0, device: int = 0,
                trainer: Optional[Trainer] = None, max_samples: int = 2_000_000, verbose: bool = True):
    """traininPipe(batchsize, net, p; epoch, lr) (train.jl:47-126).  Per epoch: q = sample(p.pool[1:L], min(2000000, L)) (with
    replacement, as StatsBase.sample does), L = div(length(q), batchsize) and the first L-1 batches are trained on (the loop breaks
    at cpt >= L, train.jl:81-84).  A fresh optimiser every call (train.jl:50).  Trains `net` in place (returns it) and reports
    (mean loss over the batches, seconds, samples/s of the last epoch).  Under torch.distributed every rank draws the same q
    (same seed) and trains on its `dp_slice` of each batch."""
    from .mcts_gpu import _world
    rank, world, _ = _world()                                   # (0, 1, None) without torch / outside a process group
    own = trainer is None
    per_rank = batchsize // world
    if own:
        trainer = Trainer(net.in_features, net.width, net.blocks, net.actions, net.fsize, per_rank, device=device, lr=lr)
    trainer.set_params(net, reset_optimizer=True)
    rng = np.random.default_rng(seed)
    report = dict(loss=float("nan"), seconds=0.0, samples_per_s=0.0, batches=0)
    for i in range(1, epoch + 1):
        L = p.length_buffer()
        q = rng.integers(0, L, size=min(max_samples, L)) if L > 0 else np.zeros(0, np.int64)
        nb = len(q) // batchsize
        if verbose and rank == 0:
            print(f"epoque: {i}\nbatch number: {nb}")
        t0 = time.perf_counter()
        tot, done = 0.0, 0
        for cpt in range(1, nb):                                     # cpt >= L breaks: L-1 batches
            idx = q[(cpt - 1) * batchsize: cpt * batchsize][dp_slice(batchsize, rank, world)]
            batch = (p.state[idx], p.policy[idx], p.value[idx], p.fstate[idx])
            loss = trainer.step_dp(*batch) if world > 1 else trainer.step(*batch)
            tot += float(loss[0])
            done += 1
        dt = time.perf_counter() - t0
        report = dict(loss=tot / max(1, done), seconds=dt, samples_per_s=done * batchsize / dt if dt > 0 else 0.0, batches=done)
        if verbose and rank == 0:
            print(f"total loss: {report['loss']}\ntraining time :{dt}")
    new = trainer.get_params()
    net.base, net.res, net.policy, net.policy_bias = new.base, new.res, new.policy, new.policy_bias
    net.value, net.value_bias, net.feature, net.feature_bias = new.value, new.value_bias, new.feature, new.feature_bias
    if own:
        trainer.close()
    return net, report
