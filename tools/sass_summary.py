#!/usr/bin/env python3
"""Per-kernel counts of the Blackwell-specific SASS opcodes in libscir_b200.so, written to
profiles/r02_sass_opcodes.txt -- the committed evidence that the hot kernels are tcgen05 / TMEM / TMA / FFMA2 code.

    python tools/sass_summary.py [--lib PATH] [--out PATH]

Opcode families (B200_PROFILING.md): UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, LDTM / STTM =
tcgen05.ld / st (TMEM), UBLKCP = cp.async.bulk (1-D TMA bulk copy), SYNCS = mbarrier, FFMA2 = packed FP32 FMA,
FFMA / DFMA = scalar FP32 / FP64 FMA.  `cuobjdump -sass` prints mangled names; c++filt demangles them.
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "FFMA2", "FFMA", "DFMA", "LDS", "STS", "LDG", "STG")


def summarise(lib):
    p = subprocess.Popen(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, errors="replace")
    counts = collections.OrderedDict()
    cur, arch = None, set()
    fn_re = re.compile(r"^\s*Function : (\S+)")
    op_re = re.compile(r"^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)")
    for line in p.stdout:
        m = fn_re.match(line)
        if m:
            cur = counts.setdefault(m.group(1), collections.Counter())
            continue
        if line.startswith("arch = "):
            arch.add(line.split("=")[1].strip())
            continue
        if cur is None:
            continue
        m = op_re.match(line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            if op in OPS:
                cur[op] += 1
    p.wait()
    if p.returncode != 0:
        raise RuntimeError("cuobjdump failed")
    names = list(counts)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = {}
    for n, d in zip(names, dem):
        d = re.sub(r"^void ", "", d)
        d = re.sub(r"\(anonymous namespace\)::", "", d)
        d = re.sub(r"scir_b200::", "", d)
        short[n] = re.sub(r"\(.*$", "", d)           # drop the argument list, keep the template arguments
    return counts, short, arch


def family(counts, short):
    fam = collections.OrderedDict()
    for n, c in counts.items():
        base = re.sub(r"<.*$", "", short[n])
        f = fam.setdefault(base, collections.Counter())
        f.update(c)
        f["_instances"] += 1
    return fam


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "scir_b200", "lib", "libscir_b200.so"))
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_sass_opcodes.txt"))
    args = ap.parse_args()
    counts, short, arch = summarise(args.lib)
    fam = family(counts, short)
    lines = ["# SASS opcode summary of scir_b200/lib/libscir_b200.so (tools/sass_summary.py; cuobjdump -sass)",
             f"# cubin architectures: {', '.join(sorted(arch)) or 'n/a'}",
             "# UTCHMMA = tcgen05.mma  UTCBAR = tcgen05.commit  LDTM/STTM = tcgen05.ld/st (TMEM)  UBLKCP = cp.async.bulk (TMA)",
             "# SYNCS = mbarrier  FFMA2 = packed FP32 FMA  DFMA = FP64 FMA",
             "",
             "## per kernel family (all template instances summed)",
             f"{'kernel':<34}{'inst':>5}" + "".join(f"{o:>9}" for o in OPS) + f"{'total':>10}"]
    for base, c in fam.items():
        lines.append(f"{base:<34}{c['_instances']:>5}" + "".join(f"{c[o]:>9}" for o in OPS) + f"{c['_total']:>10}")
    tot = collections.Counter()
    for c in fam.values():
        tot.update(c)
    lines.append(f"{'ALL':<34}{tot['_instances']:>5}" + "".join(f"{tot[o]:>9}" for o in OPS) + f"{tot['_total']:>10}")
    lines += ["", "## per kernel instance (only instances with tensor-core, TMEM or TMA opcodes)",
              f"{'kernel':<70}" + "".join(f"{o:>9}" for o in OPS[:7])]
    for n, c in counts.items():
        if c["UTCHMMA"] or c["LDTM"] or c["STTM"] or c["UBLKCP"]:
            lines.append(f"{short[n][:69]:<70}" + "".join(f"{c[o]:>9}" for o in OPS[:7]))
    text = "\n".join(lines) + "\n"
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
