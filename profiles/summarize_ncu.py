"""Turns an `ncu --set full` report (gpurun_out/*.ncu-rep, scratch) into the small tracked summaries under profiles/.
usage: python profiles/summarize_ncu.py gpurun_out/prof_mac_r01.ncu-rep profiles/r01_mac_gemm"""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEEP = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
summ = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    e = {k: (d[k] + " " + units[hdr.index(k)]).strip() for k in KEEP if k in d}
    st = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(d[k])
          for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")}
    e["stalls_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:8])
    summ.append(e)
json.dump(summ, open(out + "_ncu_summary.json", "w"), indent=1)


def gbytes(s):
    v, u = s.split()[:2]
    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]


tr = [gbytes(e["dram__bytes_read.sum"]) + gbytes(e["dram__bytes_write.sum"]) for e in summ]
json.dump({"kernel": "mac_gemm_tma_kernel", "source": rep, "per_launch_dram_bytes": tr, "dram_bytes_per_launch": sum(tr) / len(tr),
           "note": "dram__bytes_read.sum + dram__bytes_write.sum of the mac_gemm launches of ONE bench step "
                   "(c1, c2, decrypt chunks), averaged per launch like roofline.achieved"}, open(out + "_traffic.json", "w"), indent=1)
print(json.dumps(summ[1] if len(summ) > 1 else summ[0], indent=1)[:1500])
