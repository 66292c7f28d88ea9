#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one line per captured launch with the metrics the design cites."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__issue_active.avg.pct", "issue%"), ("smsp__inst_executed.sum", "winst"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_disp"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg")]
idx = [(hdr.index(k) if k in hdr else -1, n) for k, n in want]
units = rows[1]
print("|".join(n for _, n in idx))
for r in rows[2:]:
    out = []
    for i, n in idx:
        if i < 0:
            out.append("")
            continue
        v = r[i]
        if n == "kernel":
            v = v.replace("void <unnamed>::", "").split("(")[0]
        elif n == "us":
            v = f"{float(v.replace(',', '')) * (1000 if units[i] == 'ms' else (1e-3 if units[i]=='ns' else 1)):.1f}"
        elif n in ("dram_rd", "dram_wr"):
            f = float(v.replace(',', ''))
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(units[i], 1)
            v = f"{f * mult / 1e6:.1f}MB"
        else:
            try:
                v = f"{float(v.replace(',', '')):.2f}"
            except ValueError:
                pass
        out.append(v)
    print("|".join(out))
