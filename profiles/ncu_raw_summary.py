"""Summarise `ncu -i <rep> --page raw --csv` tables (what capture_r2.sh keeps of a --set full capture) per kernel launch.
usage: python profiles/ncu_raw_summary.py gpurun_out/prof_<tag>_*.raw.csv > profiles/r2/ncu_full_summary_<tag>.txt"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__icc_request_hit_rate.pct",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def main(paths):
    for p in paths:
        rows = list(csv.reader(open(p)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        print(f"== {p.split('/')[-1].replace('.raw.csv', '')} (ncu --set full --clock-control none, one eager c2 step; cold cache, serialised)")
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].replace("void ", "").split("(")[0][:48]
            vals = [f"{w.split('.')[0].replace('__', ':')}={r[idx[w]]}{units[idx[w]]}" for w in WANT if w in idx]
            print(f"{name:48s} |", " ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1:])
