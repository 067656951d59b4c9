"""Attribute an ncu report's per-SASS counters to CUDA source lines (needs -lineinfo).
usage: python tools/ncu_lines.py REPORT.ncu-rep [kernel_substr] [iterations]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "k_solve_stage"
iters = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from ncu_common import sass_sections
sass_lines = sass_sections(rep, kern, ("-g", "-c"))
cur, off2line, infunc = None, {}, False
infunc = True
for ln in sass_lines:
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and infunc:
        off2line[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]
ci, smp = h.index("Instructions Executed"), h.index("# Samples")
base, byi, bys, ti, ts = None, collections.Counter(), collections.Counter(), 0, 0
for r in rows[2:]:
    try:
        a, n, sm = int(r[0], 16), int(r[ci]), int(r[smp])
    except Exception:
        continue
    if base is None:
        base = a
    l = off2line.get(a - base)
    byi[l] += n; bys[l] += sm; ti += n; ts += sm
src = {}
def text(l):
    if not l: return "?"
    path = os.path.join(root, "locityper_b200/csrc", l[0])
    if l[0] not in src:
        src[l[0]] = open(path).read().split("\n") if os.path.exists(path) else None
    return (src[l[0]][l[1] - 1].strip()[:100] if src[l[0]] else l[0])
print(f"total warp-instructions {ti} ({ti/iters:.1f} per iteration), samples {ts}")
print("--- by samples (time)")
for l, n in bys.most_common(28):
    print(f"{n/ts*100:5.1f}% smp {byi[l]/iters:7.1f} inst/iter  {l[1] if l else 0:>4} {text(l)}")

# ---- aggregate by enclosing function of solver.cu
import re as _re
_src = open(os.path.join(root, "locityper_b200/csrc/solver.cu")).read().split("\n")
_marks = []
for _i, _l in enumerate(_src, 1):
    if _re.match(r"^(__device__|__global__|static|template)", _l) or _l.startswith("k_solve_stage("):
        for _c in (_l, _src[_i] if _i < len(_src) else ""):
            _m = _re.search(r"\b([a-z_0-9]+)\(", _c)
            if _m and _m.group(1) not in ("__launch_bounds__",):
                _marks.append((_i, _m.group(1)))
                break
def _func(line):
    name = "?"
    for s, n in _marks:
        if s <= line: name = n
        else: break
    return name
bf_i, bf_s = collections.Counter(), collections.Counter()
for l, n in byi.items():
    f = _func(l[1]) if (l and l[0] == "solver.cu") else (l[0] if l else "?")
    bf_i[f] += n; bf_s[f] += bys[l]
print("--- by function")
for f, n in bf_s.most_common(24):
    print(f"{n/ts*100:5.1f}% time {bf_i[f]/iters:7.1f} inst/iter  {f}")
if os.environ.get("BY_INST"):
    print("--- by instructions")
    for l, n in byi.most_common(60):
        print(f"{n/ti*100:5.1f}% inst {n/iters:9.1f} inst/iter {bys[l]/ts*100:5.1f}% smp  {l[1] if l else 0:>4} {text(l)}")
