"""Shared by the ncu_* tools: the SASS (nvdisasm text) of exactly the kernel instantiation an ncu report captured.
k_solve_stage is a template (k_solve_stage<WIDE, BIG>); the report names the instantiation (`k_solve_stage<0, 0>`), the
cubin holds all of them, so the section is selected by the mangled template arguments."""
import csv, io, os, re, subprocess, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernel_name(rep, kern):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    i = rows[0].index("Kernel Name")
    for r in rows[2:]:
        if kern in r[i]:
            return r[i]
    return kern


def mangled_filter(name):
    """`k_solve_stage<0, 0>(...)` -> regex matching `k_solve_stageILb0ELb0EE` in the section name."""
    m = re.search(r"(\w+)<([^>]*)>", name)
    if not m:
        return re.compile(re.escape(re.match(r"(?:void\s+)?(\w+)", name).group(1)))
    args = [a.strip() for a in m.group(2).split(",")]
    enc = "".join((r"L[bij]%sE" % a) if re.fullmatch(r"\d+", a) else r"\w+" for a in args)
    return re.compile(re.escape(m.group(1)) + "I" + enc + "E")


def sass_sections(rep, kern, flags=("-g", "-c")):
    """-> list of lines of the one .text section of the captured instantiation."""
    flt = mangled_filter(kernel_name(rep, kern))
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "locityper_b200/_lib/liblctp.so")], cwd=tmp,
                   capture_output=True)
    for f in sorted(os.listdir(tmp)):
        out = subprocess.run(["nvdisasm", *flags, os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kern not in out:
            continue
        lines, keep, got = out.split("\n"), [], False
        for ln in lines:
            if re.match(r"\s*\.text\.", ln) or re.match(r"\s*\.section\s+\.text\.", ln):
                got = bool(flt.search(ln))
            if got:
                keep.append(ln)
        if keep:
            return keep
    return []
