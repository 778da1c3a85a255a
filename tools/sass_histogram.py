"""Static SASS opcode histogram of the hot kernels in libmcp_b200.so (cuobjdump -sass; no GPU needed):
    python tools/sass_histogram.py > profiles/<tag>_sass_opcodes.md
Shows which Blackwell features each kernel really uses: UBLKCP (cp.async.bulk = TMA bulk copies), SYNCS (mbarrier),
FFMA2 / FADD2 / FMUL2 (packed fp32x2), IMAD.WIDE (Philox multiplies), MUFU (SFU), DFMA (fp64), LDS / STS / LDG / STG widths."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "montecarlooptionspricer_b200", "libmcp_b200.so")
HOT = ["rbergomi_paths_n256pair_kernelILb0", "rbergomi_paths_n256x2_kernelILb1ELb0", "rbergomi_paths_kernelILi32ELb0ELb0", "rbergomi_rows_kernel", "gbm_paths_kernelILb0ELb0",
       "lsm_sweep_tma_kernelILi3ELb0", "lsm_sweep_tma64_kernelILi3", "lsm_persist_kernelILi3ELb0", "lsm_multi_kernelILi3", "lsm_sweep_kernelIfdLi3",
       "lsm_policy_kernelIf", "rows_price_kernelILi2", "dual_nested_kernel", "asym_kernelIf", "mart_dual_kernelIf", "branch_upper_kernelIf"]
KEYS = ["UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "IMAD.WIDE", "IMAD", "LOP3", "MUFU", "DFMA", "DADD", "DMUL", "LDS", "STS", "LDG", "STG", "ATOM", "RED", "BAR", "SHFL"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            op = m.group(1)
            per[cur]["total"] += 1
            per[cur][op.split(".")[0]] += 1
            if op.startswith("IMAD.WIDE"):
                per[cur]["IMAD.WIDE"] += 1
            if op.startswith(("LDG", "STG", "LDS", "STS")):
                w = re.search(r"\.(64|128|256)", op)
                per[cur][op.split(".")[0] + "." + (w.group(1) if w else "32")] += 1
    print("# SASS opcode histogram of the hot kernels (static instruction counts, `cuobjdump -sass libmcp_b200.so`)\n")
    print("| kernel | total | " + " | ".join(KEYS) + " | widest global / shared access |")
    print("|---|---|" + "---|" * (len(KEYS) + 1))
    for want in HOT:
        fn = next((f for f in per if want in f), None)
        if not fn:
            continue
        c = per[fn]
        wide = ", ".join(k for k in ("LDG.256", "STG.256", "LDG.128", "STG.128", "LDS.128", "STS.128", "LDS.64", "STS.64", "STG.64") if c.get(k))
        name = re.sub(r"^_ZN\d+_GLOBAL__N__\w+?_cu_[0-9a-f]+\d*", "", fn)
        print(f"| `{want}` | {c['total']} | " + " | ".join(str(c.get(k, 0)) for k in KEYS) + f" | {wide} |")
    print("\nNo `UTCHMMA` / `UTCQMMA` / `LDTM` / `UTMALDG` anywhere: the path has no dense contraction (p + 1 = 4 basis columns) and its tiles are 1-D rows,")
    print("so tcgen05 / TMEM / tensor-map TMA are deliberately unused; `UBLKCP` is the 1-D bulk form of TMA.")


if __name__ == "__main__":
    main()
