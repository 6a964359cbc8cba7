"""ncu `--page raw --csv` dump (tools/ncu_profile.sh writes gpurun_out/<tag>_prof_raw.csv) -> the tracked per-launch summary.
usage: python profiles/summarize_raw.py gpurun_out/r02c_prof_raw.csv profiles/r02_ncu_summary.json "what was captured" """
import csv
import json
import sys

src, out, what = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
KEEP = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utcimma_src_int8.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
launches = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    e = {"kernel": d["Kernel Name"][:80]}
    for k in KEEP:
        if k in d and d[k] != "":
            e[k] = float(d[k])
            e[k + "__unit"] = units[hdr.index(k)]
    st = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(float(d[k]), 3)
          for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and d[k] != ""}
    e["stall_cycles_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:6])
    launches.append(e)
json.dump({"what": what, "launches": launches}, open(out, "w"), indent=1)
print(len(launches), "launches ->", out)
