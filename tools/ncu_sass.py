"""Top stalled SASS instructions of an ncu report with their CUDA source lines and preceding context.
usage: python tools/ncu_sass.py REPORT.ncu-rep [kernel_substr] [n_top] [context]"""
import csv, io, os, re, subprocess, sys, tempfile
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "k_solve_stage"
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 12
nctx = int(sys.argv[4]) if len(sys.argv) > 4 else 4
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "locityper_b200/_lib/liblctp.so")], cwd=tmp, capture_output=True)
sass = ""
for f in os.listdir(tmp):
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kern in out:
        sass = out
cur, off2line, infunc = None, {}, False
for ln in sass.split("\n"):
    if ".text." in ln and kern in ln:
        infunc = True
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and infunc:
        off2line[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]
ix = {c: i for i, c in enumerate(h)}
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
body = rows[2:]
base = int(body[0][0], 16)
tot = sum(int(r[4] or 0) for r in body)
order = sorted(range(len(body)), key=lambda k: -int(body[k][4] or 0))[:ntop]
for k in order:
    r = body[k]
    st = sorted([(int(r[ix[c]] or 0), c) for c in stalls], reverse=True)[:2]
    print(f"{int(r[4])/tot*100:5.1f}%  line {off2line.get(int(r[0],16)-base)}  {st}")
    for kk in range(max(0, k - nctx), k + 1):
        rr = body[kk]
        print(f"        {int(rr[0],16)-base:6x} {str(off2line.get(int(rr[0],16)-base, ('',0))[1]):>5} {rr[1][:90]}")
