"""Turn ncu outputs (read here, no GPU needed) into the small text/JSON summaries committed under profiles/.

    python tools/ncu_summarize.py launches gpurun_out/r1_launches_b.csv > profiles/r1_launch_summary.txt
    python tools/ncu_summarize.py kernels  gpurun_out/r1_layer_b.ncu-rep > profiles/r1_vit_layer_ncu_summary.txt
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = rows[0]
    ci = {n: i for i, n in enumerate(hdr)}
    kn, mv = ci["Kernel Name"], ci["Metric Value"]
    agg = {}
    for r in rows[1:]:
        if r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[kn]
        v = float(r[mv].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        ms = v / 1e6 if unit.startswith("ns") else (v / 1e3 if unit.startswith("us") else v)
        e = agg.setdefault(name, [0, 0.0])
        e[0] += 1
        e[1] += ms
    total = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"total {total:.1f} ms over {n} launches")
    for name, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {ms:9.3f} ms {100 * ms / total:5.1f}%  x {c:4d}  {name[:110]}")


def kernels(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("Kernel Name =", d.get("Kernel Name"))
        for k in KEYS:
            if k in d:
                print(f"  {k} = {d[k]} {units[hdr.index(k)]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels}[sys.argv[1]](sys.argv[2])
