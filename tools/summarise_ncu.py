"""ncu --page raw --csv  ->  profiles/<name>.csv (the columns that matter) and, optionally, traffic.json entries.

usage: ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
       python tools/summarise_ncu.py /tmp/raw.csv profiles/r01_ncu_full_tail_step.csv
"""
import csv
import sys

COLS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed l1tex__throughput.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
sm__warps_active.avg.pct_of_peak_sustained_active launch__registers_per_thread launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem launch__grid_size launch__block_size launch__shared_mem_per_block_dynamic
launch__waves_per_multiprocessor l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active""".split()


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    ki = names.index("Kernel Name")
    idx = [names.index(c) if c in names else None for c in COLS]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + COLS)
        w.writerow([""] + [units[i] if i is not None else "" for i in idx])
        for r in rows[hdr + 2:]:
            if len(r) > ki:
                name = r[ki].replace("se::", "")
                w.writerow([name] + [r[i].replace(",", "") if i is not None and i < len(r) else "" for i in idx])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
