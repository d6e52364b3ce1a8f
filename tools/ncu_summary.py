"""Turn `ncu --page raw --csv` exports / launch lists into the markdown summaries kept under profiles/.
    python tools/ncu_summary.py raw <raw.csv> [metric-substring ...]     -> table of selected metrics
    python tools/ncu_summary.py launches <launches.csv>                 -> per-kernel share of device time"""
import csv
import sys
from collections import defaultdict

DEFAULT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__block_size",
           "launch__cluster_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma",
           "sm__inst_executed_pipe_lsu", "sm__inst_executed_pipe_tma", "sm__inst_executed_pipe_tmem", "sm__inst_executed_pipe_xu",
           "sm__inst_executed_pipe_uniform", "sm__pipe_tensor", "sm__inst_executed_pipe_tensor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__issue_active.avg.pct", "sm__inst_issued.avg.pct", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct",
           "smsp__cycles_active.avg", "sm__inst_executed.sum ", "smsp__inst_executed.sum"]


def raw(path, keys):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    print(f"kernel: `{vals[names.index('Kernel Name')]}`\n")
    print("| metric | unit | value |\n|---|---|---:|")
    for n, u, v in sorted(zip(names, units, vals)):
        if any(k in n for k in keys):
            print(f"| {n} | {u} | {v} |")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names = rows[hdr]
    kn, mv, mn = names.index("Kernel Name"), names.index("Metric Value"), names.index("Metric Name")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[hdr + 1:]:
        if len(r) > mv and r[mn] == "gpu__time_duration.sum":
            k = r[kn].split("(")[0]
            tot[k] += float(r[mv].replace(",", "")); cnt[k] += 1
    s = sum(tot.values())
    unit = rows[hdr + 1][names.index("Metric Unit")] if len(rows) > hdr + 1 else "ns"
    print(f"| kernel | launches | total {unit} | share |\n|---|---:|---:|---:|")
    for k in sorted(tot, key=tot.get, reverse=True):
        print(f"| `{k}` | {cnt[k]} | {tot[k]:.1f} | {100 * tot[k] / s:.2f}% |")


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2], sys.argv[3:] or DEFAULT)
    else:
        launches(sys.argv[2])
