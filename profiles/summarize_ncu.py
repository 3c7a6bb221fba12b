"""Turn gpurun_out/*.ncu-rep + launch-list CSVs into the small tracked summaries under profiles/.
Usage: python profiles/summarize_ncu.py <tag> <report.ncu-rep> <launches.csv>"""
import csv, json, subprocess, sys, os
from collections import defaultdict

tag, rep, launches = sys.argv[1], sys.argv[2], sys.argv[3]
out_dir = os.path.dirname(os.path.abspath(__file__))

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_tma_ld.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_membar",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle"]
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return x
def to_bytes(v, u):
    v = num(v); u = u.lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0) if isinstance(v, float) else None
kern = []
for r in rows[2:]:
    d = {}
    for k in keep:
        if k in hdr:
            i = hdr.index(k); d[k] = {"value": num(r[i]), "unit": units[i]}
    kern.append(d)
summary = {"tag": tag, "report": os.path.basename(rep), "kernels": kern}
fused = [k for k in kern if "sfh_fg_fused" in str(k["Kernel Name"]["value"])]
if fused:
    k = fused[-1]
    rd = to_bytes(str(k["dram__bytes_read.sum"]["value"]), k["dram__bytes_read.sum"]["unit"])
    wr = to_bytes(str(k["dram__bytes_write.sum"]["value"]), k["dram__bytes_write.sum"]["unit"])
    summary["fused_kernel_dram_bytes_per_launch"] = rd + wr
    summary["fused_kernel_ncu_duration_us"] = k["gpu__time_duration.sum"]["value"]
json.dump(summary, open(os.path.join(out_dir, f"{tag}_fused_ncu.json"), "w"), indent=1)
ncu_sum_path = os.path.join(out_dir, "ncu_summary.json")
try:
    prev = json.load(open(ncu_sum_path))     # keep hand-added keys (steady-state traffic from an application-replay run)
except Exception:
    prev = {}
prev.update({"fused_kernel_dram_bytes_per_launch": summary.get("fused_kernel_dram_bytes_per_launch"), "from": f"profiles/{tag}_fused_ncu.json"})
json.dump(prev, open(ncu_sum_path, "w"), indent=1)

# launch list -> per-kernel shares
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
d = defaultdict(list)
for r in rows[1:]:
    try: d[r[ik]].append(float(r[iv].replace(",", "")))
    except ValueError: pass
tot = sum(sum(v) for v in d.values())
with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
    f.write("| kernel | launches | mean us | total us | share |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| `{k[:90]}` | {len(v)} | {sum(v)/len(v)/1e3:.2f} | {sum(v)/1e3:.1f} | {100*sum(v)/tot:.1f}% |\n")
print(open(os.path.join(out_dir, f"{tag}_launches.md")).read())
print(json.dumps({k: v for k, v in summary.items() if k != "kernels"}, indent=1))
