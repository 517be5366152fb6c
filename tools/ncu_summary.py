"""Summarise .ncu-rep captures into a markdown table (run in the authoring container; ncu -i works without a GPU)."""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__inst_executed.avg.per_cycle_active", "ipc"),
    ("sm__instruction_throughput.avg.pct_of_peak_sustained_active", "issue_%"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def main():
    for path in sys.argv[1:]:
        hdr, units, data = load(path)
        idx = {h: i for i, h in enumerate(hdr)}
        print(f"\n### {path}\n")
        cols = [m for m in METRICS if m[0] in idx]
        print("| kernel | " + " | ".join(f"{n} ({units[idx[m]]})" if units[idx[m]] else n for m, n in cols) + " |")
        print("|---|" + "---|" * len(cols))
        for r in data:
            name = r[idx["Kernel Name"]].split("(")[0].replace("gg::", "")
            print(f"| {name} | " + " | ".join(r[idx[m]] for m, _ in cols) + " |")


if __name__ == "__main__":
    main()
