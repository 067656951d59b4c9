"""Per-source-line stall breakdown of an ncu report (needs -lineinfo). usage: REPORT [kernel_substr]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep = sys.argv[1]; kern = sys.argv[2] if len(sys.argv) > 2 else "k_solve_stage"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from ncu_common import sass_sections
sass_lines = sass_sections(rep, kern, ("-g", "-c"))
cur, off2line, infunc = None, {}, False
infunc = True
for ln in sass_lines:
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2)))
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and infunc: off2line[int(m.group(1), 16)] = (cur, m.group(2))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]
cols = {c: h.index(c) for c in ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_branch_resolving", "stall_math", "stall_barrier", "stall_not_selected", "stall_selected", "stall_no_inst", "stall_lg", "stall_mio", "stall_dispatch", "stall_membar", "stall_sleep", "stall_tex", "stall_drain", "stall_misc", "# Samples"]}
tot = collections.Counter(); byline = collections.defaultdict(collections.Counter); base = None
for r in rows[2:]:
    try: a = int(r[0], 16)
    except Exception: continue
    if base is None: base = a
    l = off2line.get(a - base, (None, ""))
    for c, i in cols.items():
        try: v = int(r[i])
        except Exception: v = 0
        tot[c] += v; byline[(l[0], l[1][:60])][c] += v
print({k: v for k, v in tot.most_common()})
for reason in ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_barrier", "stall_branch_resolving"]:
    print("---", reason, tot[reason])
    for k, c in sorted(byline.items(), key=lambda kv: -kv[1][reason])[:12]:
        print(f"{c[reason]:8d}  {k[0]} {k[1]}")
