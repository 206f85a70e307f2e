#!/usr/bin/env python
"""Key metrics of `ncu --page raw --csv` exports:  python scripts/ncu_brief.py gpurun_out/kernels_r01b/prof_*.csv"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("sm__inst_executed.sum", "warp_inst"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uni%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("smsp__inst_executed_op_shared_ld.sum", "lds"),
    ("smsp__inst_executed_op_shared_st.sum", "sts"),
    ("smsp__inst_executed_op_global_st.sum", "stg"),
    ("smsp__inst_executed_op_global_ld.sum", "ldg"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "st_sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "st_requests"),
]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", path.split("/")[-1], r[hdr.index("Kernel Name")][:110])
        for k, name in KEYS:
            if k in hdr:
                print(f"   {name:22s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
        st = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
              if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
        st.sort(reverse=True)
        print("   stalls/issue:", ", ".join(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.2f}" for v, h in st[:6]))
