"""Turn the ncu artefacts of a gpurun call (gpurun_out/) into the small tracked summaries under profiles/.

    python tools/summarize_profiles.py <round-tag> <launches.csv> <title>=<report.ncu-rep> ... [--out DIR]

Runs where the reports are (on the GPU box, inside the gpurun call: the reports themselves are too large to travel back)
and writes <DIR>/<round-tag>_summary.md (default DIR = profiles/).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|dram__cycles_active\.avg\.pct_of_peak_sustained_elapsed|"
    r"lts__t_sector_hit_rate\.pct|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__inst_executed\.sum|sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|"
    r"launch__block_size|launch__shared_mem_per_block_dynamic|launch__shared_mem_per_block_static|launch__occupancy_limit_\w+|"
    r"sm__inst_executed_pipe_(alu|fma|fmaheavy|fmalite|xu|lsu|fp64|tc|tensor\w*|tma|tmem|uniform)\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_(fma|fmaheavy|alu|fp64|shared|tensor)\w*_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|smsp__cycles_active\.avg|sm__cycles_elapsed\.max)$")


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = collections.OrderedDict()
        for h, u, v in zip(hdr, units, r):
            if h == "Kernel Name":
                d["kernel"] = v
            elif KEEP.match(h):
                d[h] = f"{v} {u}".strip()
        res.append(d)
    return res


def stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    tot = collections.Counter()
    ops = collections.Counter()
    for r in rows:
        if r and r[0] == "Address":
            if hdr is not None:
                break  # first kernel only
            hdr = r
            cols = [(i, c) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
            iE, iSrc = hdr.index("Instructions Executed"), hdr.index("Source")
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        for i, c in cols:
            tot[c] += int(r[i] or 0)
        parts = r[iSrc].split()
        op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
        ops[op.split(".")[0]] += int(r[iE] or 0)
    return tot, ops


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    iK, iV, iU = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[iV].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iU], 1.0)  # -> microseconds
        name = re.sub(r"\(.*", "", r[iK]).replace("void <unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += v * scale
    return agg


def main():
    argv = sys.argv[1:]
    out_dir = os.path.join(ROOT, "profiles")
    if "--out" in argv:
        k = argv.index("--out")
        out_dir = argv[k + 1]
        del argv[k:k + 2]
    tag, launch_csv, reps = argv[0], argv[1], [a.split("=", 1) for a in argv[2:]]
    os.makedirs(out_dir, exist_ok=True)
    lines = [f"# ncu summary {tag}", ""]
    if os.path.exists(launch_csv) and os.path.getsize(launch_csv) > 0:
        agg = launches(launch_csv)
        total = sum(v[1] for v in agg.values())
        lines += [f"## launch list ({os.path.basename(launch_csv)}; cold-cache, serialised: compare SHARES)", "",
                  "| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"| {name} | {n} | {us:.1f} | {100 * us / total:.1f}% | {us / n:.2f} |")
        lines.append("")
    for title, rep in reps:
        if not os.path.exists(rep):
            lines += [f"## {title}: capture missing ({os.path.basename(rep)})", ""]
            continue
        ms = raw_metrics(rep)
        lines += [f"## {title}: `ncu --set full` ({os.path.basename(rep)})", ""]
        for d in ms[:3]:
            lines.append(f"### {d.get('kernel', '?')[:120]}")
            lines.append("")
            lines.append("| metric | value |")
            lines.append("|---|---|")
            for k, v in d.items():
                if k != "kernel":
                    lines.append(f"| {k} | {v} |")
            lines.append("")
        st, ops = stalls(rep)
        s = sum(st.values()) or 1
        lines.append("stall samples: " + ", ".join(f"{c.replace('stall_', '')} {100 * v / s:.1f}%" for c, v in st.most_common(8)))
        e = sum(ops.values()) or 1
        lines.append("")
        lines.append("executed warp instructions by opcode: " + ", ".join(f"{o} {100 * v / e:.1f}%" for o, v in ops.most_common(14)))
        lines.append("")
    out = os.path.join(out_dir, f"{tag}_summary.md")
    open(out, "w").write("\n".join(lines) + "\n")
    print(out)


if __name__ == "__main__":
    main()
