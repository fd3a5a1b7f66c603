"""The handful of ncu raw-page counters the round's notes quote, per captured kernel: python tools/ncu_raw_summary.py rep.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        # L2 atomic / reduction traffic (the linked-list allocator is the one global hot address) and the L1 side of it
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "lts__t_sectors_op_atom.sum.per_second", "lts__t_sectors_op_red.sum.per_second",
        "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_atom.sum", "smsp__inst_executed_op_global_red.sum"] + \
       ["smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k for k in
        ("barrier", "long_scoreboard", "short_scoreboard", "wait", "not_selected", "math_pipe_throttle", "lg_throttle", "mio_throttle", "branch_resolving", "no_instruction")]
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("----")
    print(f"{'Kernel Name':<85} {r[hdr.index('Kernel Name')]}")
    for k in WANT:
        if k in hdr:
            print(f"{k:<85} {r[hdr.index(k)]:>20} {units[hdr.index(k)]}")
