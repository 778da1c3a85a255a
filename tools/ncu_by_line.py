"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm line info of the same cubin, and
aggregate executed instructions / stall samples per CUDA source line.

    python tools/ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-substring> [cubin-substring] [top=40]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def disasm_lines(lib, kernel_sub, cubin_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    out = {}
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin") or (cubin_sub and cubin_sub not in f) or f.count("-") > 2:
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_fn, cur_line, inl = None, None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur_fn = m.group(1)
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur_fn:
                out.setdefault(cur_fn, []).append((int(m.group(1), 16), m.group(2).strip(), cur_line))
    for fn, ins in out.items():
        if kernel_sub in fn:
            return fn, ins
    raise SystemExit(f"kernel {kernel_sub} not found")


def main():
    rep, lib, ksub = sys.argv[1:4]
    cubin_sub = sys.argv[4] if len(sys.argv) > 4 else ""
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    fn, ins = disasm_lines(lib, ksub, cubin_sub)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # several kernels may be in the report: take blocks whose kernel name matches
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    want = re.sub(r"[^A-Za-z0-9_]", "", ksub.split("IL")[0])
    blk = [b for b in blocks if want.split("kernel")[0] in b["name"].replace(" ", "")]
    blk = blk[0] if blk else blocks[0]
    h = blk["hdr"]
    iE, iS, iSrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
    n = min(len(ins), len(blk["rows"]))
    per_line = collections.defaultdict(lambda: [0, 0])
    per_file = collections.defaultdict(lambda: [0, 0])
    tot_e = tot_s = 0
    for k in range(n):
        off, sass, line = ins[k]
        r = blk["rows"][k]
        e, s = int(r[iE] or 0), int(r[iS] or 0)
        per_line[line][0] += e
        per_line[line][1] += s
        per_file[line[0] if line else None][0] += e
        per_file[line[0] if line else None][1] += s
        tot_e += e
        tot_s += s
    print(f"kernel {blk['name'][:100]}\n  SASS instructions {len(ins)} (ncu rows {len(blk['rows'])}), executed warp-instr {tot_e}, samples {tot_s}")
    print("  per file:")
    for f, (e, s) in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
        print(f"    {str(f):22s} instr {100.0 * e / max(1, tot_e):5.1f}%  samples {100.0 * s / max(1, tot_s):5.1f}%")
    print(f"  top {top} lines by executed instructions:")
    for line, (e, s) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"    {str(line):34s} instr {100.0 * e / max(1, tot_e):5.1f}%  samples {100.0 * s / max(1, tot_s):5.1f}%")
    return per_line, tot_e, tot_s


if __name__ == "__main__":
    main()
