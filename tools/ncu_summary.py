"""One ncu capture -> the numbers DESIGN.md / bench.py quote.  usage: python tools/ncu_summary.py REPORT.ncu-rep KEY "capture description" [kernel_substr]
Prints a text summary (duration, DRAM traffic, L2 hit rate, issue slots, stall mix) and merges
{KEY: {dram_bytes_per_launch, read, write, warp_instructions_per_launch, issue_active_pct, ...}} into
profiles/r02_ncu_traffic.json (what bench.py copies into roofline.traffic)."""
import csv, io, json, os, subprocess, sys
rep, key, desc = sys.argv[1], sys.argv[2], sys.argv[3]
kern = sys.argv[4] if len(sys.argv) > 4 else ""
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
r = next(x for x in rows[2:] if kern in x[h.index("Kernel Name")])
def val(k, scale_units=True):
    i = h.index(k)
    v = float(r[i].replace(",", ""))
    unit = u[i]
    if scale_units:
        mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-3, "ms": 1, "ns": 1e-6, "s": 1e3}
        if unit in mult: v *= mult[unit]
    return v
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
out = {
    "dram_bytes_per_launch": int(rd + wr), "read": int(rd), "write": int(wr),
    "duration_ms_under_ncu": val("gpu__time_duration.sum"),
    "warp_instructions_per_launch": int(val("smsp__inst_executed.sum")),
    "issue_active_pct": val("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    "l2_hit_pct": val("lts__t_sector_hit_rate.pct"),
    "warps_active_per_sm": val("sm__warps_active.avg.per_cycle_active"),
    "lanes_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "cycles_per_issued_instruction_per_warp": val("smsp__average_warp_latency_per_inst_issued.ratio"),
    "registers_per_thread": int(val("launch__registers_per_thread")),
    "grid": int(val("launch__grid_size")),
    "kernel": r[h.index("Kernel Name")].split("(")[0],
    "capture": desc,
}
stalls = {}
for k in h:
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        stalls[k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(val(k, False), 3)
out["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
path = os.path.join(root, "profiles", "r02_ncu_traffic.json")
try:
    allv = json.load(open(path))
except Exception:
    allv = {"_comment": "per-launch numbers of `ncu --set full --clock-control none` captures of the round-2 build "
                        "(written by tools/ncu_summary.py); bench.py copies dram_bytes_per_launch into roofline.traffic"}
allv[key] = out
json.dump(allv, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
