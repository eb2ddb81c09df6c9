"""Round 2: turn the files a scripts/r02_final*.sh / r02_ncu.sh run left in gpurun_out/ into the summaries kept under profiles/.

    python scripts/r02_collect.py r03v"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

T = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def summ(rep, sims):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summarize.py"), os.path.join(G, rep), "--sims", str(sims)], capture_output=True, text=True)
    d = json.loads(out.stdout)[0]
    if "gpu_time_ns" in d:
        d["gpu_time_ms"] = round(d["gpu_time_ns"] / 1e6, 6)
    return d


def live_games(ply):
    for line in open(os.path.join(G, f"{T}_ply_profile.txt")):
        f = line.split()
        if len(f) >= 3 and f[0] == str(ply):
            return int(f[1])
    raise SystemExit(f"ply {ply} not in the per-ply profile")


head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
sha = bench.kernel_src_sha16()
L_mid, L_tail = live_games(18), live_games(30)
dom = {
    "_source": "ncu --set full --clock-control none --import-source on -k regex:^ply_kernel -s {3|18|30} -c 1 python scripts/quick_bench.py --games 32768 --reps 1 --profile 0 "
               f"(gpurun, B200, round 2 final; scripts/r02_ncu.sh, scripts/r02_collect.py): ply_fused = launch 4 (ply 3, 32768 live games, 223 games per CTA, two-tile kernel, 2097152 sims); "
               f"ply_fused_mid = launch 19 (ply 18, {L_mid} live games, one-tile kernel, ordinary orientation); ply_fused_tail = launch 31 (ply 30, {L_tail} live games, small-batch kernel: "
               "swapped orientation, weights resident in tensor memory); cold caches, serialised",
    "head": head,
    "ply_fused": summ(f"{T}_ply_full.ncu-rep", 32768 * 64),
    "ply_fused_mid": summ(f"{T}_ply_mid.ncu-rep", L_mid * 64),
    "ply_fused_tail": summ(f"{T}_ply_tail.ncu-rep", L_tail * 64),
    "kernel_src_sha16": sha,
    "_stamp": "kernel_src_sha16 = sha256 over alphagpu_b200/csrc/* (file names + contents, sorted) at capture time; bench.py reports roofline.traffic from this file only while the sources still hash to it",
    "_round2_progress": "full-load launch: 2.79 ms (round 1) -> 2.30 ms (mid round 2) -> see ply_fused.gpu_time_ms; warp instructions per launch 9.55e8 (mid round 2); global load "
                        "requests 3.45e6 (mid round 2); shared memory per CTA 198.7 KB (mid round 2: L1 60 KB) -> 165.9 KB (L1 92 KB)",
}
json.dump(dom, open(os.path.join(P, "ncu_dominant_kernel.json"), "w"), indent=1)
lb = {
    "_source": "ncu --set full --clock-control none --import-source on, Hex 7 512x8, 16384 games, R = 64 (gpurun, B200, round 2 final; scripts/r02_ncu.sh): step_seg_kernel = one slice (4096 games) "
               "of the four that replay their rollout loops concurrently from CUDA graphs (expand + backUp of rollout k-1 and the descent of rollout k); step_kernel_16384 = the same work unsegmented "
               "(AGPU_SEGMENTS=1), all 16384 games in one launch; tc_mlp512_kernel = DenseNet 512x8 on the 4096 leaves of a slice (32 CTAs of 128 rows)",
    "head": head,
    "step_seg_kernel_hex7": summ(f"{T}_step_hex7.ncu-rep", 4096),
    "step_kernel_hex7_16384": summ(f"{T}_step_hex7_16384.ncu-rep", 16384),
    "tc_mlp512_kernel_hex7": summ(f"{T}_mlp512_hex7.ncu-rep", 4096),
    "kernel_src_sha16": sha,
}
json.dump(lb, open(os.path.join(P, "r02_ncu_large_boards.json"), "w"), indent=1)

# launch list
shutil.copy(os.path.join(G, f"{T}_ncu_launch_list.csv"), os.path.join(P, "r02_ncu_launch_list.csv"))
rows = list(csv.reader(open(os.path.join(P, "r02_ncu_launch_list.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0].split("<")[0].replace("void ", "").replace("ag::", "").replace("fused::", "")
    t = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1e-3)
    agg[name][0] += 1
    agg[name][1] += t
tot = sum(v[1] for v in agg.values())
lines = ["ncu --metrics gpu__time_duration.sum --clock-control none -s 640 -c 260 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras",
         "(gpurun, B200, round 2 final; 260 launches from the timed region; cold caches, serialised)", ""]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k:40s} launches {n:4d}  total {t / 1e3:9.3f} ms  share {100 * t / tot:5.1f}%  avg {t / n:9.2f} us")
lines.append(f"{'total':40s} launches {sum(v[0] for v in agg.values()):4d}  total {tot / 1e3:9.3f} ms")
open(os.path.join(P, "r02_ncu_launch_list_summary.txt"), "w").write("\n".join(lines) + "\n")

shutil.copy(os.path.join(G, f"{T}_bench.json"), os.path.join(P, "r02_bench_1gpu.json"))
shutil.copy(os.path.join(G, f"{T}_bench_reference.json"), os.path.join(P, "r02_bench_reference_arm.json"))
for k in ("ply_fused", "ply_fused_mid", "ply_fused_tail"):
    x = dom[k]
    print(k, x.get("gpu_time_ms"), "ms", x.get("dram_bytes_per_sim"), "B/sim", "issue", round(x["issue_active_pct"], 1), "tensor", round(x["tensor_pipe_active_pct"], 1),
          "inst", x["warp_inst_executed"], "gld", x["global_load_requests"], "smem", x["dyn_smem_per_block"], "L1 hit", round(x["l1_hit_pct"], 1), "L2 hit", round(x["l2_hit_pct"], 1))
for k in ("step_seg_kernel_hex7", "step_kernel_hex7_16384", "tc_mlp512_kernel_hex7"):
    x = lb[k]
    print(k, x.get("gpu_time_ms"), "ms issue", round(x["issue_active_pct"], 1), "tensor", round(x["tensor_pipe_active_pct"], 1), "inst", x["warp_inst_executed"])
print("stamp", sha, "head", head)
