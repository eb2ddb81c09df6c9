"""Summarise an `ncu --set full` report (read here, without a GPU) into the JSON kept under profiles/.

    python scripts/ncu_summarize.py gpurun_out/r01b_ply_full.ncu-rep [--sims N] > profiles/...json

Uses `ncu -i <rep> --page raw --csv`: one row per captured launch, one column per metric."""
import argparse
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu_time_ns": "gpu__time_duration.sum",
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "issue_active_pct": "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "issue_active_pct_alt": "smsp__issue_active.avg.pct",
    "tensor_pipe_active_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "avg_active_threads_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "registers_per_thread": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "occupancy_limit_smem": "launch__occupancy_limit_shared_mem",
    "occupancy_limit_regs": "launch__occupancy_limit_registers",
    "achieved_occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dyn_smem_per_block": "launch__shared_mem_per_block_dynamic",
    "smem_config_size": "launch__shared_mem_config_size",
    "warp_inst_executed": "smsp__inst_executed.sum",
    "local_load_requests": "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
    "global_load_requests": "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
}
STALL_PREFIX = "smsp__average_warps_issue_stalled_"          # ..._per_issue_active.ratio (warp-state statistics)


def num(x):
    try:
        return float(x.replace(",", ""))
    except (ValueError, AttributeError):
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--sims", type=int, default=None, help="simulations processed by the captured launch (adds dram_bytes_per_sim)")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    if out.returncode != 0:
        sys.exit(out.stderr[-2000:])
    text = out.stdout[out.stdout.index('"ID"'):]
    rows = list(csv.reader(io.StringIO(text)))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(header)}
    res = []
    for r in data:
        d = {"kernel": r[col["Kernel Name"]] if "Kernel Name" in col else None}
        for k, m in WANT.items():
            if m in col:
                v = num(r[col[m]])
                if v is not None:
                    d[k] = v
                    if k == "gpu_time_ns" and units[col[m]] in ("us", "usecond"):
                        d[k] = v * 1e3
                    if k == "gpu_time_ns" and units[col[m]] in ("ms", "msecond"):
                        d[k] = v * 1e6
                    for kk in ("dram_bytes_read", "dram_bytes_write"):
                        if k == kk:
                            d[k] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[col[m]], 1)
        stalls = {}
        for h, i in col.items():
            if h.startswith(STALL_PREFIX) and h.endswith("_per_issue_active.ratio") or (h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")):
                v = num(r[i])
                if v:
                    stalls[h.replace(STALL_PREFIX, "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("_per_issue_active.ratio", "").replace(".ratio", "")] = v
        d["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        if a.sims and "dram_bytes_read" in d:
            d["sims_in_captured_launch"] = a.sims
            d["dram_bytes_per_sim"] = round((d["dram_bytes_read"] + d["dram_bytes_write"]) / a.sims, 1)
        res.append(d)
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
