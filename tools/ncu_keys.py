"""Key counters of the first kernel of an .ncu-rep (`ncu --set full` capture) -> JSON, the format of profiles/*_ncu_full.json."""
import csv, json, subprocess, sys
rep, dst = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
allres = []
for vals in rows[2:]:
    if len(vals) != len(hdr):
        continue
    r = {"kernel": vals[col["Kernel Name"]]}
    for k in KEYS:
        if k in col:
            r[k] = (vals[col[k]] + " " + units[col[k]]).strip()
    allres.append(r)
res = allres[0] if len(allres) == 1 else {"kernels": allres}
print(json.dumps(res, indent=1))
if dst:
    json.dump(res, open(dst, "w"), indent=1)
