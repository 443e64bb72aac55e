#!/usr/bin/env python3
"""Turn ``ncu --set full`` reports (gpurun_out/<tag>_<name>.ncu-rep) into the committed summaries under profiles/ and register
their DRAM traffic in profiles/traffic.json under the digest of the kernel sources they were captured from.

    python scripts/ncu_summarise.py <tag> <name>:<kernel>:<algo>:<batch> [...]
e.g. python scripts/ncu_summarise.py r02 fir_algo5_b4096:fir_bank_kernel:5:4096 stream_algo2_b1024:norm_stream_kernel:2:1024
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]
UNIT_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units, vals = rows[0], rows[1], rows[2]
    return {n: (u, v) for n, u, v in zip(names, units, vals)}


def main():
    from bench import build_id
    tag = sys.argv[1]
    table_path = os.path.join(ROOT, "profiles", "traffic.json")
    table = json.load(open(table_path)) if os.path.exists(table_path) else {}
    table = {k: v for k, v in table.items() if isinstance(v, dict)}  # entries are keyed by build digest
    bid = build_id()
    for spec in sys.argv[2:]:
        name, kernel, algo, batch = spec.split(":")
        rep = os.path.join(ROOT, "gpurun_out", f"{tag}_{name}.ncu-rep")
        page = raw_page(rep)
        lines = [f"# {tag} -- {page.get('Kernel Name', ('', kernel))[1]}, algo {algo}, B={batch}, L=64600 (build {bid})", "",
                 "`ncu --set full --clock-control none --import-source on`, one launch after 3 warm-up steps (`scripts/gpu_profile.sh "
                 f"{tag}`). Report file: gpurun_out/{tag}_{name}.ncu-rep (scratch, not committed); summarised by scripts/ncu_summarise.py.",
                 "", "| metric | unit | value |", "|---|---|---|"]
        for m in METRICS:
            if m in page:
                lines.append(f"| {m} | {page[m][0]} | {page[m][1]} |")

        def as_bytes(metric):
            u, v = page[metric]
            return float(v.replace(",", "")) * UNIT_BYTES.get(u, 1.0)

        traffic = as_bytes("dram__bytes_read.sum") + as_bytes("dram__bytes_write.sum")
        lines += ["", f"DRAM traffic per launch (read + write): {traffic / 1e9:.3f} GB."]
        with open(os.path.join(ROOT, "profiles", f"{tag}_{name}_ncu_full.md"), "w") as f:
            f.write("\n".join(lines) + "\n")
        table.setdefault(bid, {})[f"{kernel}:algo{algo}:b{batch}"] = traffic
        print(name, f"{traffic / 1e9:.3f} GB", page["gpu__time_duration.sum"])
    with open(table_path, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
