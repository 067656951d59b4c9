"""Per-instruction stall samples of the hottest loop of a kernel in an ncu report (source page).
usage: python tools/ncu_loop.py REPORT.ncu-rep KERNEL_REGEX [launch_index] [max_lines]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
maxl = int(sys.argv[4]) if len(sys.argv) > 4 else 120
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# split per launch: each launch starts with a "Kernel Name" row
launches, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        launches.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
L = launches[which]
h = L["rows"][0]
ix = {c: i for i, c in enumerate(h)}
body = [r for r in L["rows"][1:] if r and r[0].startswith("0x")]
base = int(body[0][0], 16)
cols = ["stall_dispatch", "stall_math", "stall_wait", "stall_short_sb", "stall_not_selected", "stall_selected",
        "stall_barrier", "stall_long_sb", "stall_mio", "stall_branch_resolving", "stall_no_inst"]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
print(L["name"][:100])
print("total samples", tot, {c[6:]: sum(int(r[ix[c]] or 0) for r in body) for c in cols})
ex = [int(r[ix["Instructions Executed"]] or 0) for r in body]
mx = max(ex)
sel = [k for k, e in enumerate(ex) if e > 0.9 * mx]
print("hot loop: %d instructions, %x..%x, executed %d each" % (len(sel), int(body[sel[0]][0], 16) - base,
                                                               int(body[sel[-1]][0], 16) - base, mx))
for k in sel[:maxl]:
    r = body[k]
    print("%5x %-50s smp=%5s " % (int(r[0], 16) - base, r[1].strip()[:50], r[ix["# Samples"]]) +
          " ".join("%s=%s" % (c[6:10], r[ix[c]]) for c in cols if int(r[ix[c]] or 0) > 0))
