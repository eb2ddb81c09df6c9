"""The bench contract, checked on the committed lines of the last B200 runs (profiles/) and on the reference arm run here."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_line(d, n_gpus):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "sims/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["n_gpus"] == n_gpus and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["unit"] == "sims/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert "traffic" in r
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n_gpus == 1:
        b = d["cpu_baseline"]
        assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]
    else:
        assert d["cpu_baseline"] is None


def test_committed_bench_lines_follow_the_contract():
    _check_line(json.load(open(os.path.join(ROOT, "profiles", "r01_bench_fused.json"))), 1)
    d2 = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_2gpu.json")))
    _check_line(d2, 2)
    d1 = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_fused.json")))
    assert d2["value"] > 1.8 * d1["value"] * 0.9          # weak scaling: two GPUs play twice the games in about the same time
    # the lines of the last session of round 1 (1, 2 and 4 GPUs of one box, same commit)
    e1, e2, e4 = (json.load(open(os.path.join(ROOT, "profiles", f"r01e_bench{s}.json"))) for s in ("", "_2gpu", "_4gpu"))
    _check_line(e1, 1); _check_line(e2, 2); _check_line(e4, 4)
    assert e2["value"] > 1.9 * e1["value"] and e4["value"] > 3.8 * e1["value"]


def test_reference_arm_runs_here_and_uses_the_host_threads():
    env = dict(os.environ, OMP_NUM_THREADS="1")            # what torchrun exports to its workers
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert len(out.stdout.strip().splitlines()) == 1          # stdout is the one JSON line (library banners on fd 1 are sent to stderr)
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "sims/s" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    assert d["cpu_baseline"]["cores"] == ncpu and d["cpu_baseline"]["kind"] == "port"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
