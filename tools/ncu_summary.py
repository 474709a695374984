"""Summarise ncu --set full reports (raw page) into a small text file for profiles/."""
import csv
import subprocess
import sys

KEYS = ['Grid Size', 'Block Size', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__shared_mem_per_block_static',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"],
                             capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(out.splitlines()))
        head, unit = rows[0], rows[1]
        print(f"## {path}")
        for r in rows[2:]:
            print("---")
            if "Kernel Name" in head:
                print(f"kernel = {r[head.index('Kernel Name')][:90]}")
            for i, name in enumerate(head):
                if name in KEYS:
                    print(f"{name} [{unit[i]}] = {r[i]}")


if __name__ == "__main__":
    main()
