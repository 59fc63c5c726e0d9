#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of
numbers the design discussion uses.  Usage: scripts/ncu_summary.py file.ncu-rep [...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("launch__registers_per_thread", "regs"), ("launch__block_size", "block"), ("launch__grid_size", "grid"),
    ("launch__shared_mem_per_block_dynamic", "dsmem"), ("launch__waves_per_multiprocessor", "waves"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ld_requests"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
]


def main():
    for path in sys.argv[1:]:
        if path.endswith(".csv"):
            out = open(path).read()
        else:
            out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        print("==", path)
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")][:60]
            vals = []
            for key, short in KEYS:
                if key in hdr:
                    i = hdr.index(key)
                    vals.append("%s=%s%s" % (short, r[i], units[i] if units[i] not in ("%", "") and short in ("time", "dram_rd", "dram_wr", "dsmem") else ""))
            print(name, " ".join(vals))
            st = []
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                    try:
                        v = float(r[i].replace(",", ""))
                    except ValueError:
                        continue
                    if v >= 0.2:
                        st.append((v, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            print("   stalls/issue:", " ".join("%s=%.2f" % (n, v) for v, n in sorted(st, reverse=True)))


if __name__ == "__main__":
    main()
