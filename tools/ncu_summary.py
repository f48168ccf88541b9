"""Markdown summary of an `ncu --page raw --csv` export: one table per kernel (the longest launch of each name).

    python tools/ncu_summary.py export.csv [name-filter ...] > profiles/<name>.md
"""
import collections
import csv
import sys

METRICS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.max.per_second",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum",
    "sm__warps_active.avg.per_cycle_active",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block",
    "launch__grid_size",
    "launch__block_size",
    "launch__waves_per_multiprocessor",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    filters = sys.argv[2:]
    hdr, units = rows[0], rows[1]
    ik, it = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
    by = collections.OrderedDict()
    for r in rows[2:]:
        if len(r) <= it:
            continue
        name = r[ik].split("(")[0].replace("void ", "")
        if filters and not any(f in name for f in filters):
            continue
        by.setdefault(name, []).append(r)
    for name, rs in by.items():
        r = max(rs, key=lambda x: float(x[it] or 0))
        print(f"## `{name}`  ({len(rs)} launch{'es' if len(rs) > 1 else ''} captured; the longest shown)\n")
        print("| metric | value |\n|---|---|")
        for m in METRICS:
            if m in hdr and r[hdr.index(m)] not in ("", "n/a"):
                print(f"| `{m}` | {r[hdr.index(m)]} {units[hdr.index(m)]} |")
        print()


if __name__ == "__main__":
    main()
