"""Attribute an ncu report's per-SASS counters to the CALL PATH of inlined device functions (needs -lineinfo).
usage: python tools/ncu_funcs.py REPORT.ncu-rep [kernel_substr] [depth]
Every instruction's `nvdisasm -gi` inlining chain is mapped to function names (by definition line ranges of
locityper_b200/csrc/solver.cu) and the warp-instructions / pc samples / long-scoreboard samples are summed per path."""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "k_solve_stage"
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 2
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src_path = os.path.join(root, "locityper_b200/csrc/solver.cu")
defs = []
for i, ln in enumerate(open(src_path).read().split("\n"), 1):
    m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__)[^(]*?\b(\w+)\s*\(", ln)
    if m:
        defs.append((i, m.group(1)))
    elif ln.startswith("k_solve_stage("):
        defs.append((i, "k_solve_stage"))
def fn_of(line):
    name = "?"
    for s, n in defs:
        if s <= line: name = n
        else: break
    return name
from ncu_common import sass_sections
sass_lines = sass_sections(rep, kern, ("-gi", "-c"))
chain, off2path, infunc, pending = [], {}, False, []
infunc = True
for ln in sass_lines:
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        pending.append(int(m.group(2)) if m.group(1).endswith("solver.cu") else -1)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and infunc:
        if pending:
            chain, pending = pending, []
        names = []
        for l in reversed(chain):          # outermost first
            n = fn_of(l) if l > 0 else "<hdr>"
            if not names or names[-1] != n: names.append(n)
        off2path[int(m.group(1), 16)] = tuple(names)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]
ci, smp, lsb = h.index("Instructions Executed"), h.index("# Samples"), h.index("stall_long_sb")
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
ti = ts = tl = 0
for r in rows[2:]:
    try:
        a, n, sm, ls = int(r[0], 16), int(r[ci]), int(r[smp]), int(r[lsb])
    except Exception:
        continue
    if base is None: base = a
    p = off2path.get(a - base, ("?",))
    key = "/".join(p[1:1 + depth]) if len(p) > 1 else p[0]
    agg[key][0] += n; agg[key][1] += sm; agg[key][2] += ls
    ti += n; ts += sm; tl += ls
print(f"total warp-instructions {ti}, samples {ts}, long_sb {tl}")
for k, (n, sm, ls) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if sm / ts < 0.004: continue
    print(f"{sm/ts*100:5.1f}% time  {n/ti*100:5.1f}% inst  {ls/max(tl,1)*100:5.1f}% long_sb   {k}")
