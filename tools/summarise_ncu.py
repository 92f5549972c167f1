#!/usr/bin/env python
"""Turn ncu outputs into the committed summaries under profiles/ (run here, no GPU needed):
  launches <launches.csv> <out.md> <title>          per-kernel totals and shares of a `--metrics gpu__time_duration.sum` list
  full <report.ncu-rep> <out.md> <out_raw.csv> <title>   selected columns of every launch of a `--set full` capture"""
import collections
import csv
import io
import re
import subprocess
import sys


def short(name: str) -> str:
    name = name.replace("vdf::", "")
    m = re.search(r"functor_kernel<\(?(?:int\))?\d+, \(?(?:int\))?\d+, (\w+)", name)
    if m:
        return m.group(1)
    m = re.search(r"(\w+)<", name) or re.search(r"(\w+)\(", name)
    return m.group(1) if m else name[:40]


def launches(path, out, title):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        ns = float(r[v].replace(",", ""))
        tot[short(r[k])] += ns
        cnt[short(r[k])] += 1
    total = sum(tot.values())
    lines = [f"# {title}", "", f"{len(rows)} launches captured, {total / 1e6:.2f} ms of device time in total "
             "(cold-cache, serialised: compare SHARES, not absolutes).", "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for name, ns in tot.most_common():
        lines.append(f"| {name} | {cnt[name]} | {ns / 1e6:.3f} | {100 * ns / total:.1f} % |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:30]))


COLS = [("gpu__time_duration.sum", "time ms", 1e-6 if False else None),
        ("launch__grid_size", "grid", None), ("launch__registers_per_thread", "regs", None),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", None),
        ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy busy %", None),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu busy %", None),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue active %", None),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %", None),
        ("dram__bytes_read.sum", "dram read", None), ("dram__bytes_write.sum", "dram write", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak", None),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard", None),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe", None),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait", None),
        ("lts__t_sector_hit_rate.pct", "L2 hit %", None)]


def full(rep, out, out_raw, title):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, rows = rows[0], rows[1], rows[2:]

    def col(metric):
        for i, h in enumerate(hdr):
            if h == metric:
                return i
        for i, h in enumerate(hdr):
            if h.endswith("." + metric):
                return i
        return None

    sel = [(m, label, col(m)) for m, label, _ in COLS]
    kcol = hdr.index("Kernel Name")
    lines = [f"# {title}", "", "| # | kernel | " + " | ".join(f"{label}{' (' + units[c] + ')' if c is not None and units[c] and units[c] not in label else ''}" for _, label, c in sel) + " |",
             "|---|---|" + "---|" * len(sel)]
    raw = [["#", "kernel"] + [m for m, _, _ in sel]]
    dram = 0.0
    for i, r in enumerate(rows, 1):
        vals = [r[c] if c is not None else "n/a" for _, _, c in sel]
        lines.append(f"| {i} | `{short(r[kcol])}` | " + " | ".join(vals) + " |")
        raw.append([str(i), short(r[kcol])] + vals)
    open(out, "w").write("\n".join(lines) + "\n")
    with open(out_raw, "w", newline="") as f:
        csv.writer(f).writerows(raw)
    print("\n".join(lines))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(*sys.argv[2:5])
    else:
        full(*sys.argv[2:6])
