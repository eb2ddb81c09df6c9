"""Regenerates tests/golden/selfplay_digests.json — SHA-256 digests of whole self-play generations played by the CPU oracle
(oracle/oracle.cpp, fp32 network, Philox streams keyed by game uid) on fixed seeds.

DERIVED vectors (the reference ships none and cannot run here): they freeze the oracle's behaviour at the commit that made them, so
that (a) a later change to the oracle that alters any sample bit is caught on CPU (tests/test_oracle_search.py) and (b) the CUDA path
is compared with a COMMITTED fixture, not only with the oracle built on the day (tests/test_gpu_parity.py).

    python tests/golden/make_selfplay_digests.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))                       # tests/ (helpers, conftest)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))      # repository root (oracle, alphagpu_b200)
import numpy as np  # noqa: E402

# (game, games, rollouts, width, blocks, net seed, game seed, uid_base, cpuct) — the cases of test_selfplay_fp32_end_to_end_bit_exact
CASES = [("ttt", 256, 16, 128, 2, 6, 77, 1000, 1.5), ("connect4", 192, 16, 128, 2, 6, 77, 1000, 1.5), ("hex5", 96, 12, 128, 2, 6, 77, 1000, 1.5),
         ("reversi6", 64, 8, 128, 2, 6, 77, 1000, 1.5), ("gobang5", 96, 12, 128, 2, 6, 77, 1000, 1.5)]
FIELDS = (("state", np.int8), ("policy", np.float32), ("player", np.int8), ("value", np.float32), ("fstate", np.int8), ("game", np.int32), ("ply", np.int32))


def digest(samples: dict, results) -> str:
    """SHA-256 over the sample arrays in FIELDS order (canonical dtypes, C order) followed by the [v, n, d] tally as int64."""
    h = hashlib.sha256()
    for name, dt in FIELDS:
        h.update(np.ascontiguousarray(samples[name], dtype=dt).tobytes())
    h.update(np.ascontiguousarray(results, dtype=np.int64).tobytes())
    return h.hexdigest()


def oracle_case(case):
    import oracle
    from conftest import GAME_SPECS
    from helpers import make_nets
    name, games, R, n, k, net_seed, seed, uid_base, cpuct = case
    ospec = oracle.Spec(*GAME_SPECS[name])
    _, onet = make_nets(GAME_SPECS[name], n, k, seed=net_seed)
    osmp = oracle.Samples(ospec, games * ospec.maxLen)
    res, st = oracle.selfplay(ospec, onet, R, games, cpuct=cpuct, seed=seed, uid_base=uid_base, samples=osmp)
    m = osmp.count
    return dict(digest=digest({f: getattr(osmp, f)[:m] for f, _ in FIELDS}, res), samples=int(m), results=[int(x) for x in res], positions=st["positions"])


if __name__ == "__main__":
    out = [dict(case=list(c), **oracle_case(c)) for c in CASES]
    with open(os.path.join(HERE, "selfplay_digests.json"), "w") as f:
        json.dump(out, f, indent=1)
    for o in out:
        print(o)
