#!/usr/bin/env python
"""Turn an ncu report / launch list (brought back in gpurun_out/) into the small text summaries kept under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1e.csv  > profiles/r1e_launches_summary.txt
    python profiles/summarize.py full     gpurun_out/prof_r1e.ncu-rep  > profiles/r1e_mi_scan_full.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        k = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e6
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: gpu__time_duration.sum per kernel (ncu --clock-control none; serialised, cold-cache launches)")
    print(f"{'kernel':64s} {'n':>5s} {'total ms':>10s} {'avg ms':>9s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:64]:64s} {v[0]:5d} {v[1]:10.3f} {v[1] / v[0]:9.4f} {v[1] / tot:6.3f}")
    print(f"{'total':64s} {sum(v[0] for v in agg.values()):5d} {tot:10.3f}")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        print(f"# {path}: {d['Kernel Name']}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d:
                print(f"{k:90s} {d[k]:>18s} {u[k]}")
        print("warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active):")
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                print(f"  {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:24s} {float(d[k]):8.3f}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
