"""Select the roofline-relevant metrics from `ncu -i X.ncu-rep --page raw --csv` (one block per captured launch).
usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py [label ...]"""
import csv
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__cycles_active.avg", "smsp__cycles_active.avg"]


def main():
    labels = sys.argv[1:]
    rows = list(csv.reader(l for l in sys.stdin if l.strip() and not l.startswith("==")))
    if len(rows) < 3:
        print("no data")
        return
    head, units = rows[0], rows[1]
    for n, r in enumerate(rows[2:]):
        print(f"## launch {n}" + (f": {labels[n]}" if n < len(labels) else ""))
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        for k in KEEP:
            if k in d:
                print(f"{k} [{u.get(k, '')}] = {d[k]}")
        print()


if __name__ == "__main__":
    main()
