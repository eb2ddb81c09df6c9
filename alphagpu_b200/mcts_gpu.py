"""Host-side mirror of `module mcts_gpu` (mcts_gpu.jl): `init`, `re_init`, `mcts_single`, `mcts`
(self-play and duel methods) and `duelnetwork`, over the C ABI of libalphagpu.so.  Names, argument
meaning and return values follow the reference; what differs is that the whole rollout loop — and for
`mcts` the whole ply loop — runs on the GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import check
from .densenet import SNetwork2
from .game import GameSpec


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class PoolSample:
    """SoA stand-in for `Game.PoolSample` / `Sample` (main4IARow.jl:29-47): a ring of `length` samples with
    fields state (Int8, 2·VS), policy (Float32, A), player, value, fstate (Int8, FS)."""

    def __init__(self, spec: GameSpec, length: int):
        self.spec, self.length, self.currentIndex, self.full = spec, length, 0, False
        self.state = np.zeros((length, 2 * spec.VectorizedState), np.int8)
        self.policy = np.zeros((length, spec.maxActions), np.float32)
        self.player = np.zeros(length, np.int8)
        self.value = np.zeros(length, np.float32)
        self.fstate = np.zeros((length, spec.FeatureSize), np.int8)

    def push_block(self, state, policy, player, value, fstate):
        """push_buffer + update_buffer (main4IARow.jl:49-75) for a block of finished samples, ring semantics kept."""
        n = state.shape[0]
        idx = (self.currentIndex + np.arange(n)) % self.length
        self.state[idx], self.policy[idx], self.player[idx], self.value[idx], self.fstate[idx] = state, policy, player, value, fstate
        if self.currentIndex + n >= self.length:
            self.full = True
        self.currentIndex = int((self.currentIndex + n) % self.length)

    def length_buffer(self):   # main4IARow.jl:77
        return self.length if self.full else self.currentIndex


class PinnedSamples:
    """Page-locked sample arrays (agpu_host_alloc) of a fixed capacity, as numpy views: the `out=` argument of selfplay().  With them
    agpu_selfplay streams the rows of a ply to the host while the next ply searches.  The views die with close()."""

    FIELDS = ("state", "policy", "player", "value", "fstate", "game", "ply")

    def __init__(self, lib, cap: int, VS: int, A: int, FS: int):
        self.lib, self.cap, self._ptrs = lib, cap, []
        shapes = dict(state=((cap, 2 * VS), np.int8), policy=((cap, A), np.float32), player=((cap,), np.int8), value=((cap,), np.float32),
                      fstate=((cap, FS), np.int8), game=((cap,), np.int32), ply=((cap,), np.int32))
        self.arrays = {}
        for k in self.FIELDS:
            shape, dt = shapes[k]
            nbytes = max(1, int(np.prod(shape)) * np.dtype(dt).itemsize)
            p = C.c_void_p()
            rc = lib.agpu_host_alloc(C.byref(p), nbytes)
            if rc != _lib.OK:
                self.close()
                raise _lib.AlphaGPUError(rc, "agpu_host_alloc failed")
            self._ptrs.append(p)
            buf = (C.c_char * nbytes).from_address(p.value)
            self.arrays[k] = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)

    def close(self):
        self.arrays = {}
        for p in self._ptrs:
            self.lib.agpu_host_free(p)
        self._ptrs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One `init(positions, visits)` (mcts_gpu.jl:350-357): the tree arrays for `ngames` games × `visits` nodes on one GPU."""

    def __init__(self, spec: GameSpec, visits: int, ngames: int, width: int, blocks: int, device: int = 0, nn_mode: int = _lib.NN_FP16_TC):
        self.lib = _lib.load()
        self.spec, self.visits, self.ngames = spec, visits, ngames
        self.width, self.blocks = width, blocks
        self.A, self.VS, self.FS = spec.maxActions, spec.VectorizedState, spec.FeatureSize
        cfg = _lib.Config(spec.game, spec.N, spec.Nvict, visits, ngames, width, blocks, device, nn_mode)
        h = C.c_void_p()
        check(None, self.lib.agpu_create(C.byref(h), C.byref(cfg)))
        self.h = h
        self.live = 0

    def close(self):
        if getattr(self, "_pinned", None) is not None:
            self._pinned.close()
            self._pinned = None
        if getattr(self, "h", None):
            self.lib.agpu_destroy(self.h)
            self.h = None

    def pinned_samples(self):
        """Page-locked sample arrays sized for this context (ngames x maxLengthGame rows), allocated once and reused: pass them as
        selfplay(out=...).  Their contents are overwritten by the next generation."""
        if getattr(self, "_pinned", None) is None:
            self._pinned = PinnedSamples(self.lib, self.ngames * self.spec.maxLengthGame, self.VS, self.A, self.FS)
        return self._pinned.arrays

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- network ----
    def set_weights(self, net: SNetwork2, slot: int = 0):
        # the C side trusts the shape the context was created with (width, blocks): a differently shaped actor would be read out of bounds
        n, k, inp = self.width, self.blocks, 2 * self.VS
        ok = (net.base.shape == (n, inp) and len(net.res) == k and all(r.shape == (n, n) for r in net.res) and net.policy.shape == (self.A, n)
              and net.policy_bias.shape == (self.A,) and net.value.shape == (1, n) and net.value_bias.shape == (1,))
        arrs = [net.base, *net.res, net.policy, net.policy_bias, net.value, net.value_bias]
        ok = ok and all(a.dtype == np.float32 and (a.ndim < 2 or a.flags.f_contiguous) for a in arrs)
        if not ok:
            raise ValueError(f"set_weights: the actor does not match this context ({n}x{k}, {inp} inputs, {self.A} actions, float32 column-major)")
        arr = (C.c_void_p * max(1, net.blocks))(*[r.ctypes.data for r in net.res])
        check(self.h, self.lib.agpu_set_weights(self.h, slot, _p(net.base), C.cast(arr, C.c_void_p), _p(net.policy), _p(net.policy_bias),
                                                _p(net.value), _p(net.value_bias)))

    def forward(self, x: np.ndarray, slot: int = 0):
        x = np.ascontiguousarray(x, np.float32)
        L = x.shape[0]
        logits, v = np.zeros((L, self.A), np.float32), np.zeros(L, np.float32)
        check(self.h, self.lib.agpu_forward(self.h, slot, _p(x), L, _p(logits), _p(v)))
        return logits, v

    # ---- plugin surface ----
    def Position(self, n: int = 1) -> np.ndarray:
        out = np.zeros(n, self.spec.position_dtype)
        check(self.h, self.lib.agpu_position_init(self.h, _p(out), n))
        return out

    def canPlay(self, pos: np.ndarray) -> np.ndarray:
        pos = np.ascontiguousarray(pos)
        out = np.zeros((pos.shape[0], self.A), np.uint8)
        check(self.h, self.lib.agpu_can_play(self.h, _p(pos), pos.shape[0], _p(out)))
        return out.astype(bool)

    def play(self, pos: np.ndarray, actions) -> np.ndarray:
        pos = np.ascontiguousarray(pos)
        act = np.ascontiguousarray(np.broadcast_to(np.asarray(actions, np.int32), pos.shape))
        out = np.zeros_like(pos)
        check(self.h, self.lib.agpu_play(self.h, _p(pos), _p(act), pos.shape[0], _p(out)))
        return out

    def isOver(self, pos: np.ndarray):
        pos = np.ascontiguousarray(pos)
        over, res = np.zeros(pos.shape[0], np.uint8), np.zeros(pos.shape[0], np.int8)
        check(self.h, self.lib.agpu_is_over(self.h, _p(pos), pos.shape[0], _p(over), _p(res)))
        return over.astype(bool), res

    def encode(self, pos: np.ndarray) -> np.ndarray:
        pos = np.ascontiguousarray(pos)
        out = np.zeros((pos.shape[0], 2 * self.VS), np.float32)
        check(self.h, self.lib.agpu_encode(self.h, _p(pos), pos.shape[0], _p(out)))
        return out

    # ---- mcts_single seam ----
    def re_init(self, positions: np.ndarray, uids: Optional[np.ndarray] = None):
        positions = np.ascontiguousarray(positions)
        u = None if uids is None else np.ascontiguousarray(uids, np.uint32)
        check(self.h, self.lib.agpu_reinit(self.h, _p(positions), positions.shape[0], _p(u)))
        self.live = positions.shape[0]

    def mcts_single(self, visits: int, L: Optional[int] = None, *, training=True, cpuct=2.0, noise=0.0, slot=0, prob=None, seed=0, ply=0):
        L = self.live if L is None else L
        pr = None if prob is None else np.ascontiguousarray(prob, np.float32)
        check(self.h, self.lib.agpu_search(self.h, L, slot, visits, int(training), cpuct, noise, _p(pr), seed, ply))

    def roots(self, L: Optional[int] = None):
        L = self.live if L is None else L
        pol, batch = np.zeros((L, self.A), np.float32), np.zeros((L, 2 * self.VS), np.float32)
        check(self.h, self.lib.agpu_get_roots(self.h, L, _p(pol), _p(batch)))
        return pol, batch

    def search_begin(self):
        check(self.h, self.lib.agpu_search_begin(self.h, self.live))

    def select(self, rollout: int, cpuct: float, *, last=False, prob=None, seed=0, ply=0):
        pr = None if prob is None else np.ascontiguousarray(prob, np.float32)
        check(self.h, self.lib.agpu_select(self.h, self.live, rollout, int(last), cpuct, _p(pr), seed, ply))

    def leaves(self):
        leaf, batch = np.zeros(self.live, np.int32), np.zeros((self.live, 2 * self.VS), np.float32)
        check(self.h, self.lib.agpu_get_leaves(self.h, self.live, _p(leaf), _p(batch)))
        return leaf, batch

    def eval(self, slot: int = 0, fetch: bool = True):
        logits = np.zeros((self.live, self.A), np.float32) if fetch else None
        v = np.zeros(self.live, np.float32) if fetch else None
        check(self.h, self.lib.agpu_eval(self.h, self.live, slot, _p(logits), _p(v)))
        return logits, v

    def expand_backup(self, prior=None, value=None, *, training=True, last=False):
        pr = None if prior is None else np.ascontiguousarray(prior, np.float32)
        vv = None if value is None else np.ascontiguousarray(value, np.float32)
        check(self.h, self.lib.agpu_expand_backup(self.h, self.live, int(training), int(last), _p(pr), _p(vv)))

    def tree(self):
        L, R, A = self.live, self.visits, self.A
        d = dict(nnodes=np.zeros(L, np.int32), parent=np.zeros((L, R), np.int32), action=np.zeros((L, R), np.int32),
                 child=np.zeros((L, R, A), np.int32), order=np.zeros((L, R, A), np.int32), nchild=np.zeros((L, R), np.int32),
                 expanded=np.zeros((L, R), np.int8), prior=np.zeros((L, R, A), np.float32), q=np.zeros((L, R, A), np.float32),
                 visits=np.zeros((L, R, A), np.float32), states=np.zeros((L, R), self.spec.position_dtype))
        td = _lib.TreeDump(*[d[k].ctypes.data for k in ("nnodes", "parent", "action", "child", "order", "nchild", "expanded", "prior", "q", "visits", "states")])
        check(self.h, self.lib.agpu_get_tree(self.h, L, C.byref(td)))
        return d

    # ---- whole loops ----
    def selfplay(self, visits: int, ngames: int, *, cpuct=2.0, noise=0.0, seed=0, uid_base=0, slot=0, want_samples=True, out=None):
        """agpu_selfplay.  `out`: optional preallocated (e.g. pinned) sample arrays keyed state/policy/player/value/fstate/game/ply."""
        res = np.zeros(3, np.int64)
        st = _lib.RunStats()
        if want_samples or out is not None:
            if out is None:
                cap = ngames * self.spec.maxLengthGame
                out = dict(state=np.empty((cap, 2 * self.VS), np.int8), policy=np.empty((cap, self.A), np.float32), player=np.empty(cap, np.int8),
                           value=np.empty(cap, np.float32), fstate=np.empty((cap, self.FS), np.int8), game=np.empty(cap, np.int32), ply=np.empty(cap, np.int32))
            cap = out["player"].shape[0]
            sc = _lib.Samples(cap, 0, *[out[k].ctypes.data for k in ("state", "policy", "player", "value", "fstate", "game", "ply")])
            rc = self.lib.agpu_selfplay(self.h, slot, visits, ngames, uid_base, cpuct, noise, seed, C.byref(sc), _p(res), C.byref(st))
            n = min(int(sc.count), cap)
            out = {k: v[:n] for k, v in out.items()}
        else:
            rc = self.lib.agpu_selfplay(self.h, slot, visits, ngames, uid_base, cpuct, noise, seed, None, _p(res), C.byref(st))
        check(self.h, rc, allow=(_lib.ERR_ILLEGAL_MOVE,))
        stats = {k: getattr(st, k) for k, _ in _lib.RunStats._fields_}
        return res, stats, out

    def duel(self, visits: int, ngames: int, *, slot_a=0, slot_b=1, cpuct=2.0, seed=0, uid_base=0):
        res = np.zeros(3, np.int64)
        st = _lib.RunStats()
        rc = self.lib.agpu_duel(self.h, slot_a, slot_b, visits, ngames, uid_base, cpuct, seed, _p(res), C.byref(st))
        check(self.h, rc, allow=(_lib.ERR_ILLEGAL_MOVE,))
        return res, {k: getattr(st, k) for k, _ in _lib.RunStats._fields_}

    # ---- measurement ----
    def profile(self, enable: bool):
        check(self.h, self.lib.agpu_profile(self.h, int(enable)))

    def kernel_times(self, reset: bool = False):
        kt = _lib.KernelTimes()
        check(self.h, self.lib.agpu_get_kernel_times(self.h, C.byref(kt), int(reset)))
        d = {name: dict(launches=int(kt.launches[i]), ms=float(kt.ms[i])) for i, name in enumerate(_lib.KERNEL_CLASSES)}
        d["nodes_traversed"], d["descents"] = int(kt.nodes_traversed), int(kt.descents)
        return d

    def layout(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(self.h, self.lib.agpu_layout_info(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(node_bytes=a.value, game_bytes=b.value, lanes_per_game=c.value)

    def debug_expf(self, x, sigmoid=False):
        x = np.ascontiguousarray(x, np.float32)
        y = np.zeros_like(x)
        check(self.h, self.lib.agpu_debug_expf(self.h, _p(x), x.size, _p(y), int(sigmoid)))
        return y


class MultiContext:
    """`ngpus` devices behind one call from one process (agpu_multi_*, include/alphagpu.h): one context, host thread and stream per
    device inside the library; games block-partitioned by uid, samples gathered into the caller's arrays in ascending uid blocks.
    The reference's caller (selfplay.jl:34) stays one process doing one call per generation."""

    def __init__(self, spec: GameSpec, visits: int, ngames: int, width: int, blocks: int, ngpus: int, devices=None, nn_mode: int = _lib.NN_FP16_TC):
        self.lib = _lib.load()
        self.spec, self.visits, self.ngames, self.ngpus = spec, visits, ngames, ngpus
        self.width, self.blocks = width, blocks
        self.A, self.VS, self.FS = spec.maxActions, spec.VectorizedState, spec.FeatureSize
        cfg = _lib.Config(spec.game, spec.N, spec.Nvict, visits, ngames, width, blocks, 0, nn_mode)
        dev = None if devices is None else (C.c_int32 * ngpus)(*devices)
        h = C.c_void_p()
        rc = self.lib.agpu_multi_create(C.byref(h), C.byref(cfg), ngpus, dev)
        if rc != _lib.OK:
            msg = self.lib.agpu_multi_last_error(None)
            raise _lib.AlphaGPUError(rc, msg.decode() if msg else "")
        self.h = h

    def close(self):
        if getattr(self, "_pinned", None) is not None:
            self._pinned.close()
            self._pinned = None
        if getattr(self, "h", None):
            self.lib.agpu_multi_destroy(self.h)
            self.h = None

    def pinned_samples(self):
        """Page-locked sample arrays sized for this context, allocated once and reused (see Context.pinned_samples)."""
        if getattr(self, "_pinned", None) is None:
            self._pinned = PinnedSamples(self.lib, self.ngames * self.spec.maxLengthGame, self.VS, self.A, self.FS)
        return self._pinned.arrays

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow=()):
        if rc != _lib.OK and rc not in allow:
            msg = self.lib.agpu_multi_last_error(self.h)
            raise _lib.AlphaGPUError(rc, msg.decode() if msg else "")

    def set_weights(self, net: SNetwork2, slot: int = 0):
        n, k, inp = self.width, self.blocks, 2 * self.VS
        if not (net.base.shape == (n, inp) and len(net.res) == k and net.policy.shape == (self.A, n) and net.value.shape == (1, n)):
            raise ValueError(f"set_weights: the actor does not match this context ({n}x{k}, {inp} inputs, {self.A} actions)")
        arr = (C.c_void_p * max(1, net.blocks))(*[r.ctypes.data for r in net.res])
        self._check(self.lib.agpu_multi_set_weights(self.h, slot, _p(net.base), C.cast(arr, C.c_void_p), _p(net.policy), _p(net.policy_bias),
                                                    _p(net.value), _p(net.value_bias)))

    def selfplay(self, visits: int, ngames: int, *, cpuct=2.0, noise=0.0, seed=0, uid_base=0, slot=0, want_samples=True, out=None):
        res = np.zeros(3, np.int64)
        st = _lib.RunStats()
        if want_samples or out is not None:
            if out is None:
                cap = ngames * self.spec.maxLengthGame
                out = dict(state=np.empty((cap, 2 * self.VS), np.int8), policy=np.empty((cap, self.A), np.float32), player=np.empty(cap, np.int8),
                           value=np.empty(cap, np.float32), fstate=np.empty((cap, self.FS), np.int8), game=np.empty(cap, np.int32), ply=np.empty(cap, np.int32))
            cap = out["player"].shape[0]
            sc = _lib.Samples(cap, 0, *[out[k].ctypes.data for k in ("state", "policy", "player", "value", "fstate", "game", "ply")])
            rc = self.lib.agpu_multi_selfplay(self.h, slot, visits, ngames, uid_base, cpuct, noise, seed, C.byref(sc), _p(res), C.byref(st))
            n = min(int(sc.count), cap)
            out = {k: v[:n] for k, v in out.items()}
        else:
            rc = self.lib.agpu_multi_selfplay(self.h, slot, visits, ngames, uid_base, cpuct, noise, seed, None, _p(res), C.byref(st))
        self._check(rc, allow=(_lib.ERR_ILLEGAL_MOVE,))
        return res, {k: getattr(st, k) for k, _ in _lib.RunStats._fields_}, out

    def duel(self, visits: int, ngames: int, *, slot_a=0, slot_b=1, cpuct=2.0, seed=0, uid_base=0):
        res = np.zeros(3, np.int64)
        st = _lib.RunStats()
        rc = self.lib.agpu_multi_duel(self.h, slot_a, slot_b, visits, ngames, uid_base, cpuct, seed, _p(res), C.byref(st))
        self._check(rc, allow=(_lib.ERR_ILLEGAL_MOVE,))
        return res, {k: getattr(st, k) for k, _ in _lib.RunStats._fields_}


# ------------------------------------------------------------------------------------------------
# The reference's public entry points
# ------------------------------------------------------------------------------------------------
def init(spec: GameSpec, visits: int, ngames: int, actor: SNetwork2, device: int = 0, nn_mode: int = _lib.NN_FP16_TC) -> Context:
    """init(positions, visits) (mcts_gpu.jl:350-357) + the actor's weights made resident."""
    ctx = Context(spec, visits, ngames, actor.width, actor.blocks, device, nn_mode)
    ctx.set_weights(actor, 0)
    return ctx


def _world():
    """(rank, world, backend) of the default torch.distributed group, (0, 1, None) outside one."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.get_rank(), dist.get_world_size(), dist.get_backend()
    except ImportError:
        pass
    return 0, 1, None


def _sum_over_ranks(values, device):
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(v) for v in t.tolist()]


def _fresh_seed(seed, device, backend):
    """seed=None: fresh randomness on every call, as the reference's unseeded RNGs give (the Julia glue passes rand(UInt64)); drawn on
    rank 0 and broadcast so that all ranks play one generation."""
    if seed is not None:
        return int(seed)
    import secrets
    seed = secrets.randbits(63)
    rank, world, _ = _world()
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([seed], dtype=torch.int64, device=f"cuda:{device}" if backend == "nccl" else None)
        dist.broadcast(t, src=0)
        seed = int(t.item())
    return seed


def _empty_run(spec: GameSpec):
    """What a rank with no games contributes (ngames < world): zero counters and empty sample blocks, so it still joins the collectives."""
    res = np.zeros(3, np.int64)
    stats = {k: 0 for k, _ in _lib.RunStats._fields_}
    out = dict(state=np.empty((0, 2 * spec.VectorizedState), np.int8), policy=np.empty((0, spec.maxActions), np.float32), player=np.empty(0, np.int8),
               value=np.empty(0, np.float32), fstate=np.empty((0, spec.FeatureSize), np.int8), game=np.empty(0, np.int32), ply=np.empty(0, np.int32))
    return res, stats, out


def mcts(actor: SNetwork2, visits: int, ngames: int, buffer: Optional[PoolSample], *, spec: GameSpec, cpuct=2.0, noise=None, seed=None,
         uid_base=0, device=0, nn_mode=_lib.NN_FP16_TC, ctx=None, ngpus: int = 1):
    """mcts(actor, visits, ngames, buffer; cpuct, noise) (mcts_gpu.jl:477-579): one generation of self-play; samples are
    pushed into `buffer`.  Returns (data, valid) like the reference plus the run statistics.

    Under torch.distributed (one process per GPU) the `ngames` games are block-partitioned over the ranks by uid — no collective on
    the search path — and the sample blocks are all-gathered afterwards, so every rank pushes the same samples in the same order
    (rank-major) into its buffer; results and counters are summed over ranks.

    `ngpus` > 1 (outside torch.distributed): one process drives that many devices through agpu_multi_selfplay — the library runs one
    host thread per device and gathers the samples itself (`ctx` may be a MultiContext to reuse)."""
    rank, world, backend = _world()
    if world == 1 and (ngpus > 1 or isinstance(ctx, MultiContext)):
        seed = _fresh_seed(seed, device, backend)
        own = ctx is None
        if own:
            ctx = MultiContext(spec, visits, ngames, actor.width, actor.blocks, ngpus, nn_mode=nn_mode)
        ctx.set_weights(actor, 0)
        noise = float(2.0 / spec.maxActions) if noise is None else noise
        # a context the caller keeps across generations brings page-locked sample arrays (the samples are copied into `buffer` right away)
        pinned = ctx.pinned_samples() if (buffer is not None and not own and ctx.ngames >= ngames) else None
        res, stats, out = ctx.selfplay(visits, ngames, cpuct=cpuct, noise=noise, seed=seed, uid_base=uid_base, want_samples=buffer is not None, out=pinned)
        if buffer is not None:
            buffer.push_block(out["state"], out["policy"], out["player"], out["value"], out["fstate"])
        if own:
            ctx.close()
        stats["results"] = res
        return dict(data=[], valid=stats["faults"] == 0, stats=stats)
    if world > 1:
        from .parallel import gather_samples, shard_games
        base, count = shard_games(ngames, rank, world)
        uid_base, ngames_local = uid_base + base, count
    else:
        ngames_local = ngames
    seed = _fresh_seed(seed, device, backend)
    own = ctx is None
    noise = float(2.0 / spec.maxActions) if noise is None else noise
    if ngames_local == 0:                                     # more ranks than games: nothing to play here, but the collectives below still run
        own = False
        res, stats, out = _empty_run(spec)
    else:
        if own:
            ctx = init(spec, visits, ngames_local, actor, device, nn_mode)
        else:
            ctx.set_weights(actor, 0)
        pinned = ctx.pinned_samples() if (buffer is not None and not own and ctx.ngames >= ngames_local) else None
        res, stats, out = ctx.selfplay(visits, ngames_local, cpuct=cpuct, noise=noise, seed=seed, uid_base=uid_base, want_samples=buffer is not None, out=pinned)
    if world > 1:
        dev = f"cuda:{device}" if backend == "nccl" else None
        if buffer is not None:
            out = gather_samples({k: out[k] for k in ("state", "policy", "player", "value", "fstate")}, device=dev, reuse_buffers=True)
        keys = ("sims", "positions", "plies", "total_length", "faults", "kernel_launches")
        tot = _sum_over_ranks(list(res) + [stats[k] for k in keys], dev)
        res = np.asarray(tot[:3], np.int64)
        stats.update(dict(zip(keys, tot[3:])))
    if buffer is not None:
        buffer.push_block(out["state"], out["policy"], out["player"], out["value"], out["fstate"])
    if own:
        ctx.close()
    stats["results"] = res
    return dict(data=[], valid=stats["faults"] == 0, stats=stats)


def mcts_duel(actor1: SNetwork2, actor2: SNetwork2, visits: int, ngames: int, *, spec: GameSpec, cpuct=2.0, seed=None, device=0,
              nn_mode=_lib.NN_FP16_TC, ctx: Optional[Context] = None):
    """mcts(actor1, actor2, visits, ngames; cpuct) (mcts_gpu.jl:581-651) -> [v, n, d].  Under torch.distributed the games are
    block-partitioned over the ranks and the three counts summed."""
    rank, world, backend = _world()
    uid_base = 0
    if world > 1:
        from .parallel import shard_games
        uid_base, ngames = shard_games(ngames, rank, world)
    seed = _fresh_seed(seed, device, backend)
    own = ctx is None
    if ngames == 0:                                           # more ranks than games
        res = np.zeros(3, np.int64)
    else:
        if own:
            ctx = Context(spec, visits, ngames, actor1.width, actor1.blocks, device, nn_mode)
        ctx.set_weights(actor1, 0)
        ctx.set_weights(actor2, 1)
        res, _ = ctx.duel(visits, ngames, cpuct=cpuct, seed=seed, uid_base=uid_base)
        if own:
            ctx.close()
    if world > 1:
        res = np.asarray(_sum_over_ranks(res, f"cuda:{device}" if backend == "nccl" else None), np.int64)
    return res


def duelnetwork(actor1: SNetwork2, actor2: SNetwork2, visits: int, ngames: int, *, spec: GameSpec, seed=None, device=0, nn_mode=_lib.NN_FP16_TC):
    """duelnetwork(actor1, actor2, visits, ngames) (mcts_gpu.jl:653-668): half the games with each net moving first."""
    seed = _fresh_seed(seed, device, _world()[2])
    h = ngames // 2
    v1, n1, d1 = mcts_duel(actor1, actor2, visits, h, spec=spec, seed=seed, device=device, nn_mode=nn_mode)
    d2, n2, v2 = mcts_duel(actor2, actor1, visits, h, spec=spec, seed=seed + 1, device=device, nn_mode=nn_mode)
    return int(v1 + v2), int(n1 + n2), int(d1 + d2)
