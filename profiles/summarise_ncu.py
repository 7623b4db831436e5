#!/usr/bin/env python
"""Turn `ncu -i X.ncu-rep --page raw --csv` into the short text summary kept under profiles/.

    python profiles/summarise_ncu.py gpurun_out/X_raw.csv [residues_per_launch] > profiles/X_summary.txt
"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h, u, v = rows[0], rows[1], rows[-1]
    print("kernel:", v[h.index("Kernel Name")] if "Kernel Name" in h else "?")
    for k in KEYS:
        if k in h:
            print(f"{k} [{u[h.index(k)]}] {v[h.index(k)]}")
    for i, k in enumerate(h):
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            try:
                if float(v[i]) >= 0.05:
                    print(f"{k} {v[i]}")
            except ValueError:
                pass
    if len(sys.argv) > 2:
        res = float(sys.argv[2])
        inst = float(v[h.index("smsp__inst_executed.sum")].replace(",", ""))
        ratio = float(v[h.index("smsp__thread_inst_executed_per_inst_executed.ratio")])
        rd = float(v[h.index("dram__bytes_read.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u[h.index("dram__bytes_read.sum")]]
        wr = float(v[h.index("dram__bytes_write.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u[h.index("dram__bytes_write.sum")]]
        print(f"residues per launch {res:.0f}")
        print(f"thread-instructions per residue {inst * ratio / res:.1f}")
        print(f"DRAM bytes per residue {(rd + wr) / res:.2f}  (read {rd / res:.2f}, write {wr / res:.2f})")


if __name__ == "__main__":
    main()
