"""Turn an .ncu-rep (read here, no GPU needed) into profiles/<tag>_ncu_raw.csv + <tag>_ncu_summary.json."""
import csv, json, subprocess, sys, collections
rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
open("profiles/%s_ncu_raw.csv" % tag, "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return x
per = collections.defaultdict(list)
for d in data:
    name = d[idx["Kernel Name"]]
    short = name.split("::")[-1].split("(")[0]
    if "kernel" not in short:
        cand = [p for p in name.split("::") if "kernel" in p]
        if cand: short = cand[0].split("(")[0]
    per[short].append({w: num(d[idx[w]]) for w in WANT if w in idx})
out = {"report": rep, "units": {w: units[idx[w]] for w in WANT if w in idx}, "kernels": {}}
for k, lst in per.items():
    avg = {}
    for w in lst[0]:
        vals = [r[w] for r in lst if isinstance(r[w], float)]
        avg[w] = sum(vals) / len(vals) if vals else lst[0][w]
    avg["launches_profiled"] = len(lst)
    to_bytes = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    r_u, w_u = out["units"]["dram__bytes_read.sum"], out["units"]["dram__bytes_write.sum"]
    avg["dram_traffic_bytes_per_launch"] = avg["dram__bytes_read.sum"] * to_bytes.get(r_u, 1) + avg["dram__bytes_write.sum"] * to_bytes.get(w_u, 1)
    out["kernels"][k] = avg
json.dump(out, open("profiles/%s_ncu_summary.json" % tag, "w"), indent=1)
for k, v in out["kernels"].items():
    print("%-32s n=%d  %.2f us  dram R %.1f W %.1f %s  dram%% %.1f  alu%% %.1f  issue%% %.1f  regs %d  grid %d x %d" % (
        k, v["launches_profiled"], v["gpu__time_duration.sum"], v["dram__bytes_read.sum"], v["dram__bytes_write.sum"], out["units"]["dram__bytes_write.sum"],
        v["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"], v["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"],
        v["smsp__issue_active.avg.pct_of_peak_sustained_active"], v["launch__registers_per_thread"], v["launch__grid_size"], v["launch__block_size"]))
