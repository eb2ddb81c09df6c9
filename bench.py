#!/usr/bin/env python
"""bench.py — self-play MCTS sims/sec on BASELINE.json config 2 (Connect4, DenseNet 128x6, 64 rollouts, 32768 games per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one self-play generation: `mcts(actor, 64, 32768, buffer)` played to the end of every game (mcts_gpu.jl:477-579).
Each rank owns one GPU and an independent shard of games (uids rank*32768 ...): no collective on the search path, weak scaling.
`value` is timed on the device (CUDA events on the library's stream, max over ranks) with weights and trees resident in HBM;
`e2e` is the same metric through the public Python/C-ABI call with host buffers: weights copied host->device and all samples
copied device->host inside the timed region.  `--impl reference` times the reference's CPU path (the C++ oracle port, OpenMP
over games, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line: everything else that lands on file descriptor 1 (NCCL prints its version banner there at
# NCCL_DEBUG=VERSION/WARN/INFO, torchrun children inherit it) is sent to stderr, and the line is written to the saved descriptor.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


GAMES, ROLLOUT, WIDTH, BLOCKS, CPUCT = 32768, 64, 128, 6, 1.5
WORKLOAD = "connect4_densenet128x6_rollout64_games32768"
METRIC = "selfplay_mcts_sims_per_sec"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def kernel_src_sha16():
    """sha256 over alphagpu_b200/csrc/* (names + contents): the stamp profiles/ncu_dominant_kernel.json carries."""
    import glob
    import hashlib
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "alphagpu_b200", "csrc", "*"))):
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def survey_bytes_per_sim(A, S, VS, dbar):
    """B_sim = d̄·B_node + B_expand + B_io + 16·d̄ — the per-unit figure the roofline contract asks for
    (988 B at Connect4's d̄ = 3.66)."""
    b_node = 12 * A + 2 * A + 8
    b_expand = 2 * S + 8 + 4 * A + 10 * A
    b_io = 2 * (2 * VS * 2) + 2 * 4 * (A + 1)
    return dbar * b_node + b_expand + b_io + 16 * dbar


def algorithmic_bytes_per_sim(A, S, dbar):
    """SURVEY.md §8(d) / BASELINE.md §3, split by kernel.  A actions, S packed state bytes, dbar expanded nodes per descent."""
    b_node = 12 * A + 2 * A + 8
    select = dbar * b_node + (2 * S + 8 + 10 * A) + 4          # traversed nodes + child allocation (state r/w, header, zero-init) + leaf id
    expand = 4 * A + 4 * (A + 1) + 16 * dbar + 4               # prior write, logits+value read, q/visits RMW per ancestor, leaf id
    nn = S + 4 * (A + 1) + 4                                   # packed leaf state read, logits+value write, leaf id
    return {"select": select, "expand_backup": expand, "nn": nn}


# BASELINE.json configs 3-5 (large boards, DenseNet 512x8) at the shard one GPU plays when the configuration's games are spread over the
# GPUs it names; config 3 names 2/4/8 GPUs: its shard follows the GPU count of this run (a quarter of the games on one GPU).
EXTRA_CONFIGS = [
    dict(name="config3_hex7_densenet512x8_rollout64", game=("hex", 7, 0), width=512, blocks=8, rollout=64, total_games=65536, gpus=(2, 4, 8), S=34),
    dict(name="config4_gobang9_densenet512x8_rollout128", game=("gobang", 9, 5), width=512, blocks=8, rollout=128, total_games=131072, gpus=(8,), S=50),
    dict(name="config5_reversi8_densenet512x8_rollout64", game=("reversi8", 0, 0), width=512, blocks=8, rollout=64, total_games=262144, gpus=(8,), S=26),
]


def run_extra_config(cfg, ag, torch, parallel, rank, world, local_rank, nn_mode, sync, peaks_):
    """One generation of self-play of a large-board configuration at its per-GPU shard: device-timed value, end-to-end value with host
    buffers (weights H2D, samples D2H inside the timed region), rooflines of the search kernels (HBM) and of the tcgen05 chain (tensor)."""
    import numpy as np
    spec = ag.GameSpec.named(*cfg["game"])
    ngpu = world if world in cfg["gpus"] else (4 if len(cfg["gpus"]) > 1 else cfg["gpus"][0])
    games = cfg["total_games"] // ngpu
    R, n, k = cfg["rollout"], cfg["width"], cfg["blocks"]
    net = ag.ressimplesf(2 * spec.VectorizedState, spec.maxActions, n, k, seed=0)
    ctx = ag.Context(spec, R, games, n, k, device=local_rank, nn_mode=nn_mode)
    ctx.set_weights(net)
    uid_base = rank * games
    ctx.selfplay(R, games, cpuct=CPUCT, seed=100, uid_base=uid_base, want_samples=False)              # warm-up
    sync()
    res, st, _ = ctx.selfplay(R, games, cpuct=CPUCT, seed=0, uid_base=uid_base, want_samples=False)
    assert st["faults"] == 0 and int(res.sum()) == games
    cap = games * spec.maxLengthGame
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
    bufs = dict(state=pin((cap, 2 * spec.VectorizedState), torch.int8), policy=pin((cap, spec.maxActions), torch.float32), player=pin((cap,), torch.int8),
                value=pin((cap,), torch.float32), fstate=pin((cap, spec.FeatureSize), torch.int8), game=pin((cap,), torch.int32), ply=pin((cap,), torch.int32))
    sync()
    t0 = time.perf_counter()
    ctx.set_weights(net)
    _, est, smp = ctx.selfplay(R, games, cpuct=CPUCT, seed=0, uid_base=uid_base, out=bufs)
    sync()
    e2e_wall = time.perf_counter() - t0
    d2h = sum(v.nbytes for v in smp.values()) + 24
    ctx.profile(True)
    ctx.kernel_times(reset=True)
    _, pst, _ = ctx.selfplay(R, games, cpuct=CPUCT, seed=0, uid_base=uid_base, want_samples=False)
    kt = ctx.kernel_times()
    ctx.close()
    del bufs
    (dev_ms, e2e_wall), (sims, e2e_sims, positions) = parallel.reduce_max_sum([st["device_ms"], e2e_wall], [st["sims"], est["sims"], st["positions"]],
                                                                               device="cuda" if world > 1 else None)
    hbm, tflops, peak_src = peaks_
    A = spec.maxActions
    dbar = kt["nodes_traversed"] / max(1, kt["descents"])
    per_sim = algorithmic_bytes_per_sim(A, cfg["S"], dbar)
    classes = {kk: v for kk, v in kt.items() if isinstance(v, dict) and v["launches"]}
    total_ms = sum(v["ms"] for v in classes.values())
    search_ms = sum(classes.get(c, {"ms": 0.0})["ms"] for c in ("select", "expand_backup"))
    nn_ms = classes.get("nn", {"ms": 0.0})["ms"]
    f_sim = 2 * (2 * spec.VectorizedState * n + k * n * n + n * (A + 1))
    out = {"name": cfg["name"], "games_per_gpu": games, "as_if_gpus": ngpu, "rollouts": R, "value": sims / (dev_ms * 1e-3), "unit": "sims/s", "ms_per_generation": dev_ms,
           "positions_per_sec": positions / (dev_ms * 1e-3), "d_bar": dbar,
           "e2e": {"value": e2e_sims / e2e_wall, "unit": "sims/s", "h2d_bytes_per_step": net.nbytes, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_wall},
           "kernel_ms": {kk: round(v["ms"], 3) for kk, v in classes.items()}}
    if search_ms > 0:
        A_ = spec.maxActions                                          # SURVEY.md §8(d) without the network's operands: d̄·B_node + B_expand + 16·d̄
        bytes_sim = dbar * (14 * A_ + 8) + (2 * cfg["S"] + 8 + 14 * A_) + 16 * dbar
        ach = bytes_sim * pst["sims"] / (search_ms * 1e-3) / 1e9
        out["roofline"] = {"kernel": "search kernels (step = expand + backUp + descent)", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                           "traffic": None, "peak_source": peak_src, "bytes_per_sim": bytes_sim, "share_of_step": search_ms / total_ms}
    if nn_ms > 0:
        tf = f_sim * pst["sims"] / (nn_ms * 1e-3) / 1e12
        out["roofline_nn"] = {"kernel": "tc_mlp512 (tcgen05 chain)", "bound": "tensor", "achieved": tf, "peak": tflops, "unit": "TFLOP/s", "frac": tf / tflops,
                              "flop_per_sim": f_sim, "share_of_step": nn_ms / total_ms}
    return out


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (oracle port of mcts_gpu.jl, all host threads), bounded sample per step."""
    if rank != 0:
        return
    import oracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import alphagpu_b200 as ag
    # every host thread the process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, and only rank 0 works here
    oracle.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    spec = oracle.Spec(oracle.CONNECT4)
    p = ag.ressimplesf(2 * spec.VS, spec.A, WIDTH, BLOCKS, seed=0)
    net = oracle.Net(p.base, p.res, p.policy, p.policy_bias, p.value, p.value_bias)
    sample_games = 4096                                                          # ~4 s per step on 16 cores: every host thread busy to the end
    times, sims = [], 0
    for i in range(args.warmup + args.steps):
        t = time.perf_counter()
        res, st = oracle.selfplay(spec, net, ROLLOUT, sample_games, cpuct=CPUCT, seed=i)
        dt = time.perf_counter() - t
        if i >= args.warmup:
            times.append(dt)
            sims += st["sims"]
    total = sum(times)
    v = sims / total
    cores = oracle.num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "sims/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "sample": f"{sample_games} of {GAMES} games per step, played to the end"},
            "cpu_baseline": {"value": v, "unit": "sims/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} x {sample_games} Connect4 games, 128x6 fp32 net, 64 rollouts, to completion ({sims} sims)"},
            "e2e": {"value": v, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nn-mode", type=int, default=2, help="2 = fp16 tcgen05 chain (default), 0 = bf16 tcgen05 chain, 1 = fp32 CUDA cores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip configs 3-5, strong scaling and the gathered generation")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    def stage(msg):                                                              # AGPU_BENCH_DEBUG=1: progress markers on stderr
        if os.environ.get("AGPU_BENCH_DEBUG"):
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import alphagpu_b200 as ag

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    stage("process group up")
    spec = ag.GameSpec.named("connect4")
    net = ag.ressimplesf(2 * spec.VectorizedState, spec.maxActions, WIDTH, BLOCKS, seed=0)
    ctx = ag.Context(spec, ROLLOUT, GAMES, WIDTH, BLOCKS, device=local_rank, nn_mode=args.nn_mode)
    ctx.set_weights(net)
    uid_base = rank * GAMES                                                      # weak scaling: every rank plays its own 32768 games

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up + device-resident timing (value) ----
    for i in range(args.warmup):
        ctx.selfplay(ROLLOUT, GAMES, cpuct=CPUCT, seed=1000 + i, uid_base=uid_base, want_samples=False)
    sampler = ClockSampler(local_rank)
    sync()
    sampler.start()
    dev_ms, sims, positions, launches, plies = 0.0, 0, 0, 0, 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        res, st, _ = ctx.selfplay(ROLLOUT, GAMES, cpuct=CPUCT, seed=i, uid_base=uid_base, want_samples=False)
        assert st["faults"] == 0 and int(res.sum()) == GAMES
        dev_ms += st["device_ms"]; sims += st["sims"]; positions += st["positions"]; launches += st["kernel_launches"]; plies += st["plies"]
    sync()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()

    # ---- end to end through the public API with host buffers (e2e) ----
    cap = GAMES * spec.maxLengthGame
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
    bufs = dict(state=pin((cap, 2 * spec.VectorizedState), torch.int8), policy=pin((cap, spec.maxActions), torch.float32), player=pin((cap,), torch.int8),
                value=pin((cap,), torch.float32), fstate=pin((cap, spec.FeatureSize), torch.int8), game=pin((cap,), torch.int32), ply=pin((cap,), torch.int32))
    ctx.set_weights(net); ctx.selfplay(ROLLOUT, GAMES, cpuct=CPUCT, seed=77, uid_base=uid_base, out=bufs)      # warm
    sync()
    e2e_t0 = time.perf_counter()
    e2e_sims, d2h, h2d = 0, 0, 0
    for i in range(args.steps):
        ctx.set_weights(net)                                                     # H2D: the actor's weights, every step
        res, st, smp = ctx.selfplay(ROLLOUT, GAMES, cpuct=CPUCT, seed=i, uid_base=uid_base, out=bufs)   # D2H: every sample
        e2e_sims += st["sims"]
        d2h += sum(v.nbytes for v in smp.values()) + 24
        h2d += net.nbytes
    sync()
    e2e_wall = time.perf_counter() - e2e_t0

    # ---- per-kernel CUDA-event times on the launching stream (roofline) ----
    ctx.profile(True)
    ctx.kernel_times(reset=True)
    _, pst, _ = ctx.selfplay(ROLLOUT, GAMES, cpuct=CPUCT, seed=0, uid_base=uid_base, want_samples=False)
    kt = ctx.kernel_times()
    ctx.profile(False)
    lay = ctx.layout()

    # ---- extras (never allowed to break the headline line) ----
    from alphagpu_b200 import parallel
    strong, gen_e2e, extra_configs = None, None, []
    if not args.no_extras:
        try:
            # strong scaling: the 32768 games of the metric split over the ranks (uids rank*G/N ...), time = max over ranks
            gl = GAMES // world
            ctx.selfplay(ROLLOUT, gl, cpuct=CPUCT, seed=55, uid_base=rank * gl, want_samples=False)
            sync()
            _, sst, _ = ctx.selfplay(ROLLOUT, gl, cpuct=CPUCT, seed=0, uid_base=rank * gl, want_samples=False)
            (sms,), (ssims,) = parallel.reduce_max_sum([sst["device_ms"]], [sst["sims"]], device="cuda" if world > 1 else None)
            strong = {"workload": WORKLOAD, "total_games": GAMES, "games_per_gpu": gl, "value": ssims / (sms * 1e-3), "unit": "sims/s", "ms_per_generation": sms,
                      "scaling": "strong"}
            # one generation end to end INCLUDING the sample gather over the ranks (NCCL all_gather of the padded blocks when world > 1)
            if world > 1:                                                        # communicator set-up and allocator warm-up are not part of a generation
                parallel.gather_samples({kk: bufs[kk] for kk in ("state", "policy", "player", "value", "fstate")}, device="cuda", reuse_buffers=True)
            sync()
            t0g = time.perf_counter()
            ctx.set_weights(net)
            _, gst, gsmp = ctx.selfplay(ROLLOUT, gl, cpuct=CPUCT, seed=0, uid_base=rank * gl, out=bufs)
            t1g = time.perf_counter()
            gathered = parallel.gather_samples({kk: gsmp[kk] for kk in ("state", "policy", "player", "value", "fstate")}, device="cuda" if world > 1 else None,
                                               reuse_buffers=True)
            sync()
            t2g = time.perf_counter()
            (gwall, ggather), (gsims,) = parallel.reduce_max_sum([t2g - t0g, t2g - t1g], [gst["sims"]], device="cuda" if world > 1 else None)
            gen_e2e = {"workload": WORKLOAD, "total_games": GAMES, "value": gsims / gwall, "unit": "sims/s", "ms_per_step": 1e3 * gwall, "gather_ms": 1e3 * ggather,
                       "gathered_samples": int(len(gathered["player"])), "gather": "nccl all_gather_into_tensor of padded per-rank blocks, compacted on the device, one D2H into page-locked memory per field" if world > 1 else "single rank: none"}
        except Exception as e:                                                   # pragma: no cover
            stage(f"strong-scaling / gather leg failed: {e!r}")
            strong = strong or {"error": repr(e)}
    ctx.close()
    del bufs
    if not args.no_extras:
        for cfg in EXTRA_CONFIGS:
            try:
                extra_configs.append(run_extra_config(cfg, ag, torch, parallel, rank, world, local_rank, args.nn_mode, sync, peaks()))
                stage(f"extra config {cfg['name']} done")
            except Exception as e:                                               # pragma: no cover
                stage(f"extra config {cfg['name']} failed: {e!r}")
                extra_configs.append({"name": cfg["name"], "error": repr(e)})

    stage("search legs done")
    # ---- the training step that follows a generation (train.jl): batch 8192 (main4IARow.jl:102-105) split over the ranks,
    #      gradient all-reduce over NCCL when world > 1.  An extra, not part of the headline metric; never allowed to break the line. ----
    train_info = None
    try:
        tnet = ag.ressimplesf_full(2 * spec.VectorizedState, spec.maxActions, spec.FeatureSize, WIDTH, BLOCKS, seed=0)
        TB = 8192
        trainer = ag.Trainer.for_network(tnet, TB // world, device=local_rank)
        rng = np.random.default_rng(1)
        st = (rng.random((TB, 2 * spec.VectorizedState)) < 0.3).astype(np.int8)
        pol = rng.random((TB, spec.maxActions)).astype(np.float32); pol /= pol.sum(1, keepdims=True)
        val = rng.choice(np.array([0, 0.5, 1], np.float32), size=TB)
        fst = rng.integers(-1, 2, size=(TB, spec.FeatureSize)).astype(np.int8)
        sl = ag.train.dp_slice(TB, rank, world)
        tb = [a[sl] for a in (st, pol, val, fst)]
        tstep = (lambda: trainer.step_dp(*tb)) if world > 1 else (lambda: trainer.step(*tb))
        stage("trainer created")
        for _ in range(3):
            tstep()
        stage("train warm-up done")
        sync()
        t0 = time.perf_counter()
        tdev = []
        for _ in range(20):
            tstep()
            tdev.append(sum(trainer.last_ms()))
        sync()
        twall = (time.perf_counter() - t0) / 20
        train_info = {"net": "networkf 128x6 + feature head, fp32", "global_batch": TB, "ms_per_step_e2e": 1e3 * twall, "ms_per_step_device": float(np.median(tdev)),
                              "samples_per_s": TB / twall, "allreduce": "nccl sum of one flat fp32 gradient" if world > 1 else None}
        trainer.close()
        stage("train leg done")
    except Exception as e:                                                       # pragma: no cover
        stage(f"train leg failed: {e!r}")
        train_info = {"error": repr(e)}

    # ---- reduce over ranks: time = max, work = sum ----
    (dev_ms, wall, e2e_wall), w = parallel.reduce_max_sum([dev_ms, wall, e2e_wall], [sims, positions, launches, e2e_sims, d2h, h2d],
                                                          device="cuda" if world > 1 else None)
    sims, positions, launches, e2e_sims, d2h, h2d = [int(x) for x in w]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm, tflops, peak_src = peaks()
    dbar = kt["nodes_traversed"] / max(1, kt["descents"])
    S = 18                                                                       # packed Connect4 state: 2 x u64 + player + round
    per_sim = algorithmic_bytes_per_sim(spec.maxActions, S, dbar)
    classes = {k: v for k, v in kt.items() if isinstance(v, dict) and v["launches"]}
    total_ms = sum(v["ms"] for v in classes.values())
    f_sim = 2 * (2 * spec.VectorizedState * WIDTH + BLOCKS * WIDTH * WIDTH + WIDTH * (spec.maxActions + 1))
    fused = "ply_fused" in classes and classes["ply_fused"]["ms"] > 0.5 * total_ms
    if fused:
        # one persistent kernel per ply runs the whole rollout loop (search phases + tcgen05 chain): both rooflines refer to its duration
        dom, dom_ms = "ply_fused", classes["ply_fused"]["ms"]
        bytes_sim = survey_bytes_per_sim(spec.maxActions, S, spec.VectorizedState, dbar)
        nn_ms, nn_name = dom_ms, "ply_fused (tcgen05 chain inside the per-ply kernel)"
    else:
        dom = max(("select", "expand_backup"), key=lambda k: classes.get(k, {"ms": 0})["ms"])
        dom_ms = classes[dom]["ms"]
        bytes_sim = per_sim[dom] + (per_sim["expand_backup"] if dom == "select" else 0)   # the step kernel fuses expand+backup into select
        nn_ms, nn_name = classes["nn"]["ms"], "nn (tcgen05 chain)"
    achieved = bytes_sim * pst["sims"] / (dom_ms * 1e-3) / 1e9
    traffic, traffic_note = None, "no ncu capture on file"
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_dominant_kernel.json")) as f:
            cap = json.load(f)
        if cap.get("kernel_src_sha16") != kernel_src_sha16():
            traffic_note = "stale: profiles/ncu_dominant_kernel.json was captured on other kernel sources (kernel_src_sha16 differs)"
        else:
            per = cap.get(dom, {}).get("dram_bytes_per_sim")                  # ncu --set full capture of one launch, scaled to the average launch
            traffic = None if per is None else per * pst["sims"] / classes[dom]["launches"]
            traffic_note = "ncu --set full capture of a full-load launch (profiles/ncu_dominant_kernel.json), bytes per sim x sims per average launch"
    except Exception:
        pass
    nn_tf = f_sim * pst["sims"] / (nn_ms * 1e-3) / 1e12

    line = {
        "metric": METRIC, "value": sims / (dev_ms * 1e-3), "unit": "sims/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {2: "f16 operands / f32 accumulate (tcgen05)", 0: "bf16 operands / f32 accumulate (tcgen05)", 1: "f32"}[args.nn_mode] + "; search f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "games_per_gpu": GAMES, "rollouts": ROLLOUT, "net": "DenseNet 128x6 random-init (Glorot), seed 0", "cpuct": CPUCT,
                   "l2": "inputs_larger_than_l2 (tree arrays 268 MB per GPU)", "timing": "cuda events on the library stream, max over ranks",
                   "positions_per_sec": positions / (dev_ms * 1e-3), "mean_game_length": positions / (GAMES * args.steps * world), "d_bar": dbar,
                   "wall_ms_per_step": 1e3 * wall / args.steps, "node_bytes": lay["node_bytes"], "lanes_per_game": lay["lanes_per_game"]},
        "clocks": clocks,
        "e2e": {"value": e2e_sims / e2e_wall, "unit": "sims/s", "h2d_bytes_per_step": h2d // args.steps // world, "d2h_bytes_per_step": d2h // args.steps // world,
                "ms_per_step": 1e3 * e2e_wall / args.steps},
        "gpu_launches": launches,
        "roofline": {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic, "traffic_source": traffic_note,
                     "peak_source": peak_src, "bytes_per_sim": bytes_sim, "avg_launch_us": 1e3 * dom_ms / classes[dom]["launches"],
                     "share_of_step": dom_ms / total_ms, "sims_per_launch": pst["sims"] / classes[dom]["launches"],
                     "bytes_per_sim_basis": "SURVEY.md §8(d) B_sim at the measured d_bar (leaf operands and network outputs counted as HBM traffic although the "
                                            "per-ply kernel keeps them on chip: %d B/sim without them)" % round(per_sim["select"] + per_sim["expand_backup"] + per_sim["nn"]),
                     "note": "not HBM-bound: see DESIGN.md §6 and profiles/ncu_dominant_kernel.json (full-load launch: issue slots 40 % busy, tensor pipe 14 %, "
                             "DRAM 745 B per simulation; the search phases are bound by the number of load/store requests of the SM (one per lane and 16 bytes) and "
                             "by where they hit, the network phase by epilogue issue slots and tcgen05.mma issue; 49 % of a generation is spent in plies with more "
                             "than 128 games per SM, 24 % between 32 and 128, 27 % in the tail, where a rollout is a dependent chain whose length does not shrink "
                             "with the number of live games"},
        "roofline_nn": {"kernel": nn_name, "bound": "tensor", "achieved": nn_tf, "peak": tflops, "unit": "TFLOP/s", "frac": nn_tf / tflops,
                        "flop_per_sim": f_sim, "share_of_step": nn_ms / total_ms},
        "kernel_ms": {k: round(v["ms"], 3) for k, v in classes.items()},
        "strong_scaling": strong,
        "generation_e2e": gen_e2e,
        "extra_configs": extra_configs,
        "train_step": train_info,
    }

    # ---- CPU baseline: the oracle port on the host cores, bounded sample (N=1 only) ----
    if world == 1 and not args.no_cpu_baseline:
        import oracle
        oracle.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
        ospec = oracle.Spec(oracle.CONNECT4)
        onet = oracle.Net(net.base, net.res, net.policy, net.policy_bias, net.value, net.value_bias)
        oracle.selfplay(ospec, onet, ROLLOUT, 32, cpuct=CPUCT, seed=0)           # warm the threads
        games = 8192                                                             # ~10 s of CPU work on the box's 16 cores
        t = time.perf_counter()
        _, ost = oracle.selfplay(ospec, onet, ROLLOUT, games, cpuct=CPUCT, seed=0)
        dt = time.perf_counter() - t
        line["cpu_baseline"] = {"value": ost["sims"] / dt, "unit": "sims/s", "cores": oracle.num_threads(), "kind": "port",
                                "sample": f"{games} of {GAMES} Connect4 games played to the end, same net in fp32, {ost['sims']} sims in {dt:.1f} s"}
    else:
        line["cpu_baseline"] = None
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
