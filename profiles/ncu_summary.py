"""Summarise an ncu --set full report (.ncu-rep) per kernel launch: duration, DRAM bytes, DRAM %, tensor-pipe %,
achieved occupancy, registers.  usage: python profiles/ncu_summary.py <file.ncu-rep> [...]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_uniform.sum"]


def main(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        print(f"== {p}")
        for r in rows[2:]:
            name = r[idx["Kernel Name"]][:60]
            vals = []
            for w in WANT:
                if w in idx:
                    vals.append(f"{w.split('.')[0].replace('__', ':')}={r[idx[w]]}{units[idx[w]]}")
            print(name, "|", " ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1:])
