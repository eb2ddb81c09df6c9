"""CPU oracle — TEST INFRASTRUCTURE ONLY (parity unpinned, see oracle/oracle.cpp header).

ctypes front end for ``oracle/oracle.cpp``, the C++ restatement of the reference
hot path (``mcts_gpu.jl``, ``Bitboard.jl``, the game plugins, ``DenseNet.jl:294-304``).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this package; nothing under ``alphagpu_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "oracle.cpp")
_SRCS = [_SRC, os.path.join(_HERE, "train_oracle.cpp")]

CONNECT4, GOBANG, HEX, REVERSI8, REVERSI6 = 0, 1, 2, 3, 4

# -O3 with AVX2 where the host has it (the timed CPU baseline should be a fair one: 3x over -O2); never -ffast-math, never FMA
# contraction (no -mfma): every float operation stays one IEEE-rounded operation, results are identical for all flag sets.
CXXFLAGS = ["-O3", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math"]


def _host_has_avx2() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            return " avx2 " in f.read().replace("\n", " ")
    except OSError:
        return False


_AVX2 = _host_has_avx2()
_SO = os.path.join(_HERE, "_build", "liboracle_avx2.so" if _AVX2 else "liboracle.so")


def build_all(force: bool = False) -> None:
    """Both flavours, so that the prebuilt files serve whatever host the repository snapshot lands on."""
    global _SO, _AVX2
    keep = (_SO, _AVX2)
    for avx2 in (False, True):
        _AVX2, _SO = avx2, os.path.join(_HERE, "_build", "liboracle_avx2.so" if avx2 else "liboracle.so")
        build(force)
    _SO, _AVX2 = keep


def build(force: bool = False) -> str:
    """Compile oracle.cpp -> oracle/_build/liboracle[_avx2].so (g++, seconds)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in _SRCS):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["g++", *CXXFLAGS, *(["-mavx2"] if _AVX2 else []), "-o", _SO, *_SRCS])
    return _SO


# Julia isbits layouts: bitboard{2} = 48 B; Position = 104 B / 152 B (Reversi)
BB = np.dtype([("chunks", "<u8", (3,)), ("len", "<i8"), ("dims", "<i8", (2,))])
POS2 = np.dtype([("bplayer", BB), ("bopponent", BB), ("player", "i1"), ("aux", "i1"), ("pad", "i1", (6,))])
POS3 = np.dtype([("bplayer", BB), ("bopponent", BB), ("legalplay", BB), ("player", "i1"), ("pad", "i1", (7,))])
assert BB.itemsize == 48 and POS2.itemsize == 104 and POS3.itemsize == 152


class GameInfo(C.Structure):
    _fields_ = [("A", C.c_int32), ("VS", C.c_int32), ("FS", C.c_int32), ("maxLen", C.c_int32), ("pos_bytes", C.c_int32)]


class SamplesC(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("count", C.c_int64), ("state", C.c_void_p), ("policy", C.c_void_p),
                ("player", C.c_void_p), ("value", C.c_void_p), ("fstate", C.c_void_p), ("game", C.c_void_p), ("ply", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        try:
            build()
            _lib = C.CDLL(_SO)
        except OSError:
            build(force=True)
            _lib = C.CDLL(_SO)
        _lib.orc_uniform.restype = C.c_float
        _lib.orc_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        _lib.orc_net_create.restype = C.c_void_p
        _lib.orc_tree_create.restype = C.c_void_p
        _lib.orc_tree_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64]
        _lib.orc_trainer_create.restype = C.c_void_p
        _lib.orc_trainer_create.argtypes = [C.c_int] * 5 + [C.c_double] * 5 + [C.c_float]
        _lib.orc_trainer_count.restype = C.c_int64
        for f in ("orc_trainer_destroy", "orc_trainer_count", "orc_trainer_reset_opt"):
            getattr(_lib, f).argtypes = [C.c_void_p]
        _lib.orc_trainer_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.orc_trainer_set.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.orc_trainer_loss_grad.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_int]
        _lib.orc_trainer_apply.argtypes = [C.c_void_p, C.c_float]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Spec:
    """(game, N, Nvict) + the plugin constants VectorizedState, FeatureSize, maxActions, maxLengthGame."""

    def __init__(self, game: int, N: int = 0, Nvict: int = 0):
        self.game, self.N, self.Nvict = game, N, Nvict
        gi = GameInfo()
        if lib().orc_game_info(game, N, Nvict, C.byref(gi)) != 0:
            raise ValueError(f"bad game spec {(game, N, Nvict)}")
        self.A, self.VS, self.FS, self.maxLen, self.pos_bytes = gi.A, gi.VS, gi.FS, gi.maxLen, gi.pos_bytes
        self.pos_dtype = POS2 if self.pos_bytes == 104 else POS3

    @property
    def g(self):
        return (self.game, self.N, self.Nvict)

    # ---- plugin surface, batched over numpy arrays of wire positions ----
    def position(self, n: int = 1) -> np.ndarray:
        one = np.zeros(1, dtype=self.pos_dtype)
        lib().orc_position_init(*self.g, _p(one))
        return np.repeat(one, n)

    def can_play(self, pos: np.ndarray, action) -> np.ndarray:
        pos = np.ascontiguousarray(pos)
        act = np.ascontiguousarray(np.broadcast_to(np.asarray(action, dtype=np.int32), pos.shape))
        out = np.zeros(pos.shape[0], dtype=np.uint8)
        lib().orc_can_play(*self.g, _p(pos), _p(act), C.c_int64(pos.shape[0]), _p(out))
        return out.astype(bool)

    def legal(self, pos: np.ndarray) -> np.ndarray:
        pos = np.ascontiguousarray(pos)
        out = np.zeros((pos.shape[0], self.A), dtype=np.uint8)
        lib().orc_legal(*self.g, _p(pos), C.c_int64(pos.shape[0]), _p(out))
        return out.astype(bool)

    def play(self, pos: np.ndarray, action) -> np.ndarray:
        pos = np.ascontiguousarray(pos)
        act = np.ascontiguousarray(np.broadcast_to(np.asarray(action, dtype=np.int32), pos.shape))
        out = np.zeros_like(pos)
        lib().orc_play(*self.g, _p(pos), _p(act), C.c_int64(pos.shape[0]), _p(out))
        return out

    def is_over(self, pos: np.ndarray):
        pos = np.ascontiguousarray(pos)
        over = np.zeros(pos.shape[0], dtype=np.uint8)
        res = np.zeros(pos.shape[0], dtype=np.int8)
        lib().orc_is_over(*self.g, _p(pos), C.c_int64(pos.shape[0]), _p(over), _p(res))
        return over.astype(bool), res

    def encode(self, pos: np.ndarray) -> np.ndarray:
        pos = np.ascontiguousarray(pos)
        out = np.zeros((pos.shape[0], 2 * self.VS), dtype=np.float32)
        lib().orc_encode(*self.g, _p(pos), C.c_int64(pos.shape[0]), _p(out))
        return out

    def decode_fstate(self, pos: np.ndarray) -> np.ndarray:
        pos = np.ascontiguousarray(pos)
        out = np.zeros((pos.shape[0], self.VS), dtype=np.int8)
        lib().orc_decode_fstate(*self.g, _p(pos), C.c_int64(pos.shape[0]), _p(out))
        return out


def bb_op(op: str, bb: np.ndarray, n: int = 0) -> np.ndarray:
    code = {"<<": 0, ">>>": 1, "right": 2, "left": 3, "down": 4, "up": 5, "~": 6}[op]
    out = np.zeros(1, dtype=BB)
    src = np.ascontiguousarray(bb.reshape(1))
    lib().orc_bb_op(code, _p(src), C.c_int64(n), _p(out))
    return out[0]


def uniform(seed: int, uid: int, ply: int, rollout: int, depth: int) -> float:
    return float(lib().orc_uniform(seed, uid, ply, rollout, depth))


def philox(ctr: Sequence[int], k0: int, k1: int):
    c = (C.c_uint32 * 4)(*ctr)
    lib().orc_philox(c, C.c_uint32(k0), C.c_uint32(k1))
    return [int(x) for x in c]


class Net:
    """snetwork2 weights (DenseNet.jl:279-286). Arrays are numpy with the Julia shapes:
    base (n, in), res[k] (n, n), policy (A, n), policy_bias (A,), value (1, n), value_bias (1,).
    Stored Fortran-ordered so the bytes equal Julia's column-major arrays."""

    FP32, BF16, BF16_RESID, F16, F16_RESID = 0, 1, 2, 3, 4

    def __init__(self, base, res, policy, policy_bias, value, value_bias):
        f = lambda a: np.asfortranarray(np.asarray(a, dtype=np.float32))
        self.base, self.res = f(base), [f(r) for r in res]
        self.policy, self.policy_bias = f(policy), f(policy_bias).reshape(-1)
        self.value, self.value_bias = f(value).reshape(1, -1), f(value_bias).reshape(-1)
        self.n, self.inp = self.base.shape
        self.k, self.A = len(self.res), self.policy.shape[0]
        arr = (C.c_void_p * max(1, self.k))(*[r.ctypes.data for r in self.res])
        self._h = C.c_void_p(lib().orc_net_create(self.inp, self.n, self.k, self.A, _p(self.base), arr, _p(self.policy),
                                                  _p(self.policy_bias), _p(self.value), _p(self.value_bias)))

    def __del__(self):
        try:
            lib().orc_net_destroy(self._h)
        except Exception:
            pass

    def forward(self, x: np.ndarray, mode: int = 0, softmax: bool = False):
        """x: (L, in) rows = games (the transpose of the reference's (in, L) batch; same bytes). -> logits (L, A), v (L,)"""
        x = np.ascontiguousarray(x, dtype=np.float32)
        L = x.shape[0]
        logits = np.zeros((L, self.A), dtype=np.float32)
        v = np.zeros(L, dtype=np.float32)
        lib().orc_net_forward(self._h, _p(x), C.c_int64(L), _p(logits), _p(v), int(mode), int(softmax))
        return logits, v


def c_expf(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.zeros_like(x)
    lib().orc_expf(_p(x), C.c_int64(x.size), _p(y))
    return y


def sigmoid(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.zeros_like(x)
    lib().orc_sigmoid(_p(x), C.c_int64(x.size), _p(y))
    return y


def round_to(x: np.ndarray, fmt: str) -> np.ndarray:
    """round fp32 values to bf16 ('bf16') or saturating fp16 ('f16'), returned as fp32"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.zeros_like(x)
    lib().orc_round(_p(x), C.c_int64(x.size), _p(y), 1 if fmt == "f16" else 0)
    return y


def softmax(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32).copy()
    lib().orc_softmax(_p(x), x.shape[1], C.c_int64(x.shape[0]))
    return x


class Tree:
    """Tree arrays of mcts_gpu.jl:35-53 plus the search entry points."""

    def __init__(self, spec: Spec, R: int, L: int):
        self.spec, self.R, self.L = spec, R, L
        self._h = C.c_void_p(lib().orc_tree_create(*spec.g, R, L))
        if not self._h:
            raise ValueError("tree_create failed")

    def __del__(self):
        try:
            lib().orc_tree_destroy(self._h)
        except Exception:
            pass

    def reinit(self, positions: np.ndarray, uids: Optional[np.ndarray] = None):
        positions = np.ascontiguousarray(positions)
        self.live = positions.shape[0]
        u = None if uids is None else np.ascontiguousarray(uids, dtype=np.uint32)
        rc = lib().orc_tree_reinit(self._h, _p(positions), _p(u), C.c_int64(self.live))
        assert rc == 0

    def search_begin(self):
        lib().orc_search_begin(self._h, C.c_int64(self.live))

    def select(self, rollout: int, cpuct: float, prob: Optional[np.ndarray] = None, seed: int = 0, ply: int = 0):
        pr = None if prob is None else np.ascontiguousarray(prob, dtype=np.float32)
        self._keep = pr
        lib().orc_select(self._h, C.c_int64(self.live), rollout, C.c_float(cpuct), _p(pr), C.c_uint64(seed), C.c_uint32(ply))

    def leaf_batch(self):
        leaf = np.zeros(self.live, dtype=np.int32)
        batch = np.zeros((self.live, 2 * self.spec.VS), dtype=np.float32)
        lib().orc_get_leaf_batch(self._h, C.c_int64(self.live), _p(leaf), _p(batch))
        return leaf, batch

    def eval(self, net: Net, mode: int = 0):
        logits = np.zeros((self.live, self.spec.A), dtype=np.float32)
        lib().orc_eval(self._h, net._h, C.c_int64(self.live), mode, _p(logits))
        prior = np.zeros((self.live, self.spec.A), dtype=np.float32)
        v = np.zeros(self.live, dtype=np.float32)
        lib().orc_get_eval(self._h, C.c_int64(self.live), _p(prior), _p(v))
        return logits, prior, v

    def expand_backup(self, prior: Optional[np.ndarray], v: Optional[np.ndarray], training: bool):
        pr = None if prior is None else np.ascontiguousarray(prior, dtype=np.float32)
        vv = None if v is None else np.ascontiguousarray(v, dtype=np.float32)
        lib().orc_expand_backup(self._h, C.c_int64(self.live), _p(pr), _p(vv), int(training))

    def finish_search(self):
        lib().orc_finish_search(self._h, C.c_int64(self.live))

    def mcts_single(self, net: Optional[Net], visits: int, training: bool, cpuct: float, prob=None, inj_prior=None, inj_v=None,
                    seed: int = 0, ply: int = 0, nn_mode: int = 0):
        pr = None if prob is None else np.ascontiguousarray(prob, dtype=np.float32)
        ip = None if inj_prior is None else np.ascontiguousarray(inj_prior, dtype=np.float32)
        iv = None if inj_v is None else np.ascontiguousarray(inj_v, dtype=np.float32)
        rc = lib().orc_mcts_single(self._h, None if net is None else net._h, visits, C.c_int64(self.live), int(training), C.c_float(cpuct),
                                   _p(pr), _p(ip), _p(iv), C.c_uint64(seed), C.c_uint32(ply), nn_mode)
        assert rc == 0

    def roots(self):
        pol = np.zeros((self.live, self.spec.A), dtype=np.float32)
        batch = np.zeros((self.live, 2 * self.spec.VS), dtype=np.float32)
        lib().orc_get_roots(self._h, C.c_int64(self.live), _p(pol), _p(batch))
        return pol, batch

    def dump(self):
        L, R, A = self.live, self.R, self.spec.A
        d = dict(nnodes=np.zeros(L, np.int32), parent=np.zeros((L, R), np.int32), action=np.zeros((L, R), np.int32),
                 child=np.zeros((L, R, A), np.int32), order=np.zeros((L, R, A), np.int32), nchild=np.zeros((L, R), np.int32),
                 expanded=np.zeros((L, R), np.int8), prior=np.zeros((L, R, A), np.float32), q=np.zeros((L, R, A), np.float32),
                 visits=np.zeros((L, R, A), np.float32), policy=np.zeros((L, R, A), np.float32),
                 states=np.zeros((L, R), dtype=self.spec.pos_dtype))
        lib().orc_tree_dump(self._h, C.c_int64(L), *[_p(d[k]) for k in
                                                     ("nnodes", "parent", "action", "child", "order", "nchild", "expanded", "prior", "q", "visits", "policy", "states")])
        return d

    def poke(self, g: int, node: int, prior, q, visits, child_order):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        pr, qq, vv = f(prior), f(q), f(visits)
        co = np.ascontiguousarray(child_order, dtype=np.int32)
        lib().orc_tree_poke(self._h, C.c_int64(g), node, _p(pr), _p(qq), _p(vv), _p(co), len(co))

    def counters(self):
        out = np.zeros(4, dtype=np.int64)
        lib().orc_get_counters(self._h, _p(out))
        return dict(descents=int(out[0]), nodes_traversed=int(out[1]), newton_solves=int(out[2]), newton_iters=int(out[3]))


class Samples:
    """SoA samples in push order (main4IARow.jl:29-37): state i8 (2VS), policy f32 (A), player i8, value f32, fstate i8 (FS)."""

    def __init__(self, spec: Spec, capacity: int):
        self.spec, self.capacity, self.count = spec, capacity, 0
        self.state = np.zeros((capacity, 2 * spec.VS), np.int8)
        self.policy = np.zeros((capacity, spec.A), np.float32)
        self.player = np.zeros(capacity, np.int8)
        self.value = np.zeros(capacity, np.float32)
        self.fstate = np.zeros((capacity, spec.FS), np.int8)
        self.game = np.zeros(capacity, np.int32)
        self.ply = np.zeros(capacity, np.int32)

    def _c(self):
        return SamplesC(self.capacity, 0, *[a.ctypes.data for a in (self.state, self.policy, self.player, self.value, self.fstate, self.game, self.ply)])


def selfplay(spec: Spec, net: Net, visits: int, ngames: int, cpuct: float = 1.5, seed: int = 0, uid_base: int = 0, nn_mode: int = 0,
             samples: Optional[Samples] = None):
    """mcts(actor, visits, ngames, buffer) (mcts_gpu.jl:477-579). Returns (results [v,n,d], stats dict)."""
    res = np.zeros(3, np.int64)
    st = np.zeros(5, np.int64)
    sc = samples._c() if samples is not None else None
    rc = lib().orc_selfplay(*spec.g, net._h, visits, C.c_int64(ngames), C.c_uint32(uid_base), C.c_float(cpuct), C.c_uint64(seed), nn_mode,
                            C.byref(sc) if sc is not None else None, _p(res), _p(st))
    assert rc == 0
    if samples is not None:
        samples.count = int(sc.count)
    return res, dict(sims=int(st[0]), positions=int(st[1]), plies=int(st[2]), total_length=int(st[3]), faults=int(st[4]))


def duel(spec: Spec, net1: Net, net2: Net, visits: int, ngames: int, cpuct: float = 2.0, seed: int = 0, uid_base: int = 0, nn_mode: int = 0):
    """mcts(actor1, actor2, visits, ngames) (mcts_gpu.jl:581-651)."""
    res = np.zeros(3, np.int64)
    st = np.zeros(5, np.int64)
    rc = lib().orc_duel(*spec.g, net1._h, net2._h, visits, C.c_int64(ngames), C.c_uint32(uid_base), C.c_float(cpuct), C.c_uint64(seed), nn_mode,
                        _p(res), _p(st))
    assert rc == 0
    return res, dict(sims=int(st[0]), positions=int(st[1]), plies=int(st[2]), faults=int(st[4]))


def choose_move(pol: np.ndarray, rnd: int, u: float, duel_mode: bool = False) -> int:
    pol = np.ascontiguousarray(pol, dtype=np.float32)
    return int(lib().orc_choose_move(_p(pol), pol.shape[0], C.c_uint32(rnd), C.c_float(u), int(duel_mode)))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int):
    lib().orc_set_num_threads(int(n))


class Trainer:
    """One `networkf` + Optimiser(ADAM(lr), WeightDecay(wd)) (train.jl:12-15,47-51,128-162; oracle/train_oracle.cpp).

    Parameters travel as a dict of Julia column-major arrays: base (n,in), res [k x (n,n)], pol_w (A,n), pol_b (A), val_w (1,n),
    val_b (1), feat_w (FS,n), feat_b (FS) — numpy arrays in Fortran order or any array whose .T is the row-major transpose.
    """
    PARAMS, GRADS, M, V = 0, 1, 2, 3

    def __init__(self, inp: int, n: int, k: int, A: int, FS: int, lr=0.001, beta1=0.9, beta2=0.999, eps=1e-8, wd=1e-4, fweight=0.001):
        self.inp, self.n, self.k, self.A, self.FS, self.NH = inp, n, k, A, FS, A + 1 + FS
        self._h = C.c_void_p(lib().orc_trainer_create(inp, n, k, A, FS, lr, beta1, beta2, eps, wd, fweight))
        self.P = int(lib().orc_trainer_count(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_trainer_destroy(self._h)
            self._h = None

    # flat layout: base | res | heads packed (NH x n) column-major | head biases
    def pack(self, d) -> np.ndarray:
        cm = lambda a, r, c: np.asarray(a, np.float32).reshape(r, c, order="F") if np.asarray(a).ndim == 1 else np.asarray(a, np.float32)
        heads = np.concatenate([cm(d["pol_w"], self.A, self.n), cm(d["val_w"], 1, self.n), cm(d["feat_w"], self.FS, self.n)], axis=0)
        parts = [cm(d["base"], self.n, self.inp).ravel(order="F")] + [cm(w, self.n, self.n).ravel(order="F") for w in d["res"]]
        parts += [heads.ravel(order="F"), np.asarray(d["pol_b"], np.float32).ravel(), np.asarray(d["val_b"], np.float32).ravel(),
                  np.asarray(d["feat_b"], np.float32).ravel()]
        flat = np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)
        assert flat.size == self.P
        return flat

    def unpack(self, flat: np.ndarray):
        n, k, A, FS, NH, inp = self.n, self.k, self.A, self.FS, self.NH, self.inp
        o = 0
        base = flat[o:o + n * inp].reshape(n, inp, order="F"); o += n * inp
        res = []
        for _ in range(k):
            res.append(flat[o:o + n * n].reshape(n, n, order="F")); o += n * n
        heads = flat[o:o + NH * n].reshape(NH, n, order="F"); o += NH * n
        bias = flat[o:o + NH]
        return dict(base=base, res=res, pol_w=heads[:A], val_w=heads[A:A + 1], feat_w=heads[A + 1:], pol_b=bias[:A], val_b=bias[A:A + 1],
                    feat_b=bias[A + 1:])

    def set(self, which: int, flat: np.ndarray):
        flat = np.ascontiguousarray(flat, dtype=np.float32)
        assert flat.size == self.P
        assert lib().orc_trainer_set(self._h, which, _p(flat)) == 0

    def get(self, which: int) -> np.ndarray:
        out = np.zeros(self.P, np.float32)
        assert lib().orc_trainer_get(self._h, which, _p(out)) == 0
        return out

    def set_params(self, d, reset_optimizer=True):
        self.set(self.PARAMS, self.pack(d))
        if reset_optimizer:
            lib().orc_trainer_reset_opt(self._h)

    def loss_grad(self, state, policy, value, fstate, want_grad=True) -> np.ndarray:
        state = np.ascontiguousarray(state, np.int8); policy = np.ascontiguousarray(policy, np.float32)
        value = np.ascontiguousarray(value, np.float32).ravel(); fstate = np.ascontiguousarray(fstate, np.int8)
        B = state.shape[0]
        assert state.shape == (B, self.inp) and policy.shape == (B, self.A) and value.shape == (B,) and fstate.shape == (B, self.FS)
        out = np.zeros(4, np.float32)
        assert lib().orc_trainer_loss_grad(self._h, _p(state), _p(policy), _p(value), _p(fstate), B, _p(out), int(want_grad)) == 0
        return out

    def apply(self, gscale: float = 1.0):
        assert lib().orc_trainer_apply(self._h, gscale) == 0

    def step(self, state, policy, value, fstate) -> np.ndarray:
        out = self.loss_grad(state, policy, value, fstate)
        self.apply(1.0)
        return out
