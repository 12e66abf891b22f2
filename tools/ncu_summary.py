#!/usr/bin/env python
"""Text summary of an `ncu --set full` report (.ncu-rep), one block per profiled launch: duration,
DRAM bytes, pipe utilisation, issue rate, occupancy, warp-stall breakdown.  The numbers committed
under profiles/ come from this script.

usage: ncu_summary.py <report.ncu-rep> [reads_per_launch]
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe (LOP3/SHF/IADD3) %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe cycles active %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe (IMAD) %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per instruction"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "global RED requests"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["long_scoreboard", "short_scoreboard", "math_pipe_throttle", "wait", "not_selected", "selected", "lg_throttle",
               "mio_throttle", "branch_resolving", "dispatch_stall", "no_instruction", "barrier", "drain", "membar", "tex_throttle",
               "sleeping", "misc"]


def main():
    path = sys.argv[1]
    reads = float(sys.argv[2]) if len(sys.argv) > 2 else None
    text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("source: %s (ncu --set full --clock-control none)" % path.split("/")[-1])
    for r in rows[2:]:
        print()
        print("kernel: %s" % r[col["Kernel Name"]])
        vals = {}
        for key, label in KEYS:
            if key in col and r[col[key]] != "":
                vals[key] = (r[col[key]], units[col[key]])
                print("  %-45s %s %s" % (label, r[col[key]], units[col[key]]))
        stalls = []
        for name in STALL_NAMES:
            k = STALLS % name
            if k in col and r[col[k]] not in ("", "0"):
                try:
                    stalls.append((float(r[col[k]]), name))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  warp stalls per issue (warps waiting, by reason): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:8]))
        if reads:
            def num(key):
                v, u = vals.get(key, ("0", ""))
                v = float(v.replace(",", ""))
                scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
                         "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
                return v * scale
            dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
            dur = num("gpu__time_duration.sum")
            inst = num("smsp__inst_executed.sum")
            print("  per read (%d reads in this launch): DRAM traffic %.1f B, %.1f warp-instructions per 32-read tile, %.2f G reads/s under ncu"
                  % (reads, dram / reads, inst / (reads / 32.0), reads / dur / 1e9 if dur else 0))


if __name__ == "__main__":
    main()
