"""Summarise the SASS of libmindrec_b200.so (cuobjdump -sass): per kernel, the counts of the instructions that show the
Blackwell / bulk-copy / vector paths (UBLKCP = cp.async.bulk on the TMA engine, SYNCS = mbarrier, LDGSTS = cp.async,
LDG/STG.E.*.128/.256 vector accesses, UCGABAR = cluster barrier, MATCH / REDUX warp primitives), plus short excerpts.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mindrec_b200", "libmindrec_b200.so")
PATTERNS = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDGSTS", r"\bLDGSTS"), ("LDG.256", r"\bLDG\.E[.\w]*\.256"),
            ("STG.256", r"\bSTG\.E[.\w]*\.256"), ("LDG.128", r"\bLDG\.E[.\w]*\.128"), ("STG.128", r"\bSTG\.E[.\w]*\.128"),
            ("UCGABAR", r"\bUCGABAR"), ("MATCH", r"\bMATCH"), ("ATOM/RED", r"\b(ATOMG|ATOMS|RED|REDG)\b"),
            ("HMMA/UTC*MMA", r"\b(HMMA|UTC\w*MMA)")]
EXCERPT = {"gather_rows_kernel": r"UBLKCP|SYNCS|LDG\.E.*128", "segsum_stage_kernel": r"LDGSTS",
           "rows_update_kernel": r"LDG\.E.*256|STG\.E.*256", "onesweep_pass_kernel": r"MATCH|LDG\.E.*STRONG|ATOMG"}


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(.*", "", cur)
            kernels[cur] = []
        elif cur is not None and re.search(r"/\*[0-9a-f]{4}\*/", line):
            kernels[cur].append(line.strip())
    print("libmindrec_b200.so: %d kernels, cubin architectures: %s" % (len(kernels), ", ".join(arch)))
    print("%-78s %6s  %s" % ("kernel", "instr", "  ".join(n for n, _ in PATTERNS)))
    tot = collections.Counter()
    for name, lines in kernels.items():
        counts = [sum(1 for l in lines if re.search(p, l)) for _, p in PATTERNS]
        for (n, _), c in zip(PATTERNS, counts):
            tot[n] += c
        if any(counts):
            print("%-78s %6d  %s" % (name[:78], len(lines), "  ".join("%*d" % (len(n), c) for (n, _), c in zip(PATTERNS, counts))))
    print("%-78s %6s  %s" % ("TOTAL", "", "  ".join("%*d" % (len(n), tot[n]) for n, _ in PATTERNS)))
    print("\nNo tensor-core instruction is expected: nothing on the product path is a dense contraction (the DenseLayer GEMMs"
          " are cuBLAS).\n")
    for key, pat in EXCERPT.items():
        for name, lines in kernels.items():
            if key in name:
                hits = [l for l in lines if re.search(pat, l)][:6]
                if hits:
                    print("---- %s" % name[:110])
                    for h in hits:
                        print("    " + h[:150])
                    break


if __name__ == "__main__":
    sys.exit(main())
