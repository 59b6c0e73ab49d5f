"""Summarise .ncu-rep captures (read on the CPU box with `ncu -i`) into a small tracked text file.

    python tools/ncu_summary.py gpurun_out/sweep_transe_fb.ncu-rep [...] > profiles/r01_xxx.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
]


def main():
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            print(f"## {path}: no kernels captured")
            continue
        head, units = rows[0], rows[1]
        print(f"## {path}")
        for r in rows[2:]:
            print(f"kernel: {r[head.index('Kernel Name')]}")
            for k in KEYS:
                cols = [i for i, h in enumerate(head) if h == k or h.startswith(k)]
                for i in cols[:1]:
                    print(f"  {head[i]} [{units[i]}] = {r[i]}")
            print()


if __name__ == "__main__":
    main()
