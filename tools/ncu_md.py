"""Markdown summary of an .ncu-rep for profiles/: headline metrics and the L1 / LSU pipe breakdown of
every captured launch, and (with --lines, when /tmp/cub holds the matching disassembly: cuobjdump -xelf +
nvdisasm -g -c of the library that was profiled) the top source lines by executed instructions.
usage: python tools/ncu_md.py report.ncu-rep "title" [--lines mangled_substring:block ...] > profiles/x.md"""
import csv
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2]
HEAD = [
    ("gpu__time_duration.sum", "duration"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes per instruction (of 32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data-pipe (LSU) wavefronts % of peak"),
    ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "L1 LSU writeback active %"),
    ("lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "L2 sectors % of peak"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit %"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "local load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "local load sectors"),
    ("l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "local load L1 hit %"),
    ("l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "local store requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "local store sectors"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local load instructions"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local store instructions"),
    ("smsp__sass_inst_executed_op_shared.sum", "shared-memory instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch resolving / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected / issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no instruction / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe throttle / issue"),
]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print(f"# {title}\n")
print(f"source: `{rep.split('/')[-1]}` (`ncu --set full --clock-control none --import-source on`, read with `ncu -i ... --page raw --csv`)\n")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].replace("(bool)", "").split("(")[0]
    print(f"## launch {r[hdr.index('ID')]}: `{name}` grid {r[hdr.index('launch__grid_size')] if 'launch__grid_size' in hdr else ''}\n")
    print("| metric | value |\n|---|---|")
    for m, label in HEAD:
        if m in hdr:
            i = hdr.index(m)
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:,.0f}" if abs(f) >= 1000 else f"{f:.3g}"
            except ValueError:
                pass
            print(f"| {label} | {v} {units[i]} |")
    print()
for a in sys.argv[3:]:
    if a == "--lines":
        continue
    sub, blk = a.rsplit(":", 1)
    o = subprocess.run([sys.executable, "tools/ncu_source_profile.py", rep, sub, blk, "40"], capture_output=True, text=True).stdout
    print(f"## top source lines by executed instructions: block {blk}\n\n```\n{o}```\n")
