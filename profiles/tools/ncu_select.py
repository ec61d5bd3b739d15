"""Selected counters of an `ncu --page raw --csv` export, one JSON object per captured kernel launch:
python profiles/tools/ncu_select.py raw.csv [kernel-name regex] > summary.json"""
import csv, json, re, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h, u = rows[hi], rows[hi + 1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = []
for r in rows[hi + 2:]:
    if len(r) < len(h): continue
    name = r[h.index("Kernel Name")]
    if pat and not pat.search(name): continue
    d = {"kernel": name}
    for i, n in enumerate(h):
        if n in KEYS or ("issue_stalled" in n and n.endswith("per_issue_active.ratio") and "not_issued" not in n):
            try: d[n] = {"value": float(r[i].replace(",", "")), "unit": u[i]}
            except ValueError: pass
    out.append(d)
json.dump(out, sys.stdout, indent=1)
