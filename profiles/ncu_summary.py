#!/usr/bin/env python
"""Key counters of every kernel in an .ncu-rep (ncu --set full), as a small table.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv, io, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__cycles_elapsed.avg.per_second", "sm clock"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instr"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read (SM)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("smsp__inst_executed.sum", "warp instr"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]

def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    for r in data:
        print("==", r[name_i][:110])
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print(f"   {label:28s} {r[i]:>18s} {units[i]}")

if __name__ == "__main__":
    main(sys.argv[1])
