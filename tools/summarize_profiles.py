"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_profiles.py launches <csv> <out.md> <title> [last_n]
  python tools/summarize_profiles.py full <ncu-rep> <out.md> <title>
  python tools/summarize_profiles.py multi <csv> <out.md> <title> <first kernel of the sequence to keep>
      (csv of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
       sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed --csv`: one row per launch and metric)
  python tools/summarize_profiles.py traffic <csv> <key> <first kernel> <kernel-name regex> <msm|ntt> [note]
      (same csv plus sm__inst_executed_pipe_fmaheavy.sum and smsp__thread_inst_executed_per_inst_executed.ratio): sums DRAM
      bytes and fmaheavy thread-instructions (= executed wide MACs) over the matching launches of the LAST sequence that
      starts with <first kernel> and files them in profiles/traffic.json under <key> together with the hash of the
      kernel sources -- bench.py reports them only while that hash still matches.
"""
import collections
import csv
import subprocess
import sys


def launches(path, out, title, last_n=None):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
    if last_n:
        rows = rows[-int(last_n):]
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "").strip()
        t = float(r[-1]) / 1e6
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        total += t
    with open(out, "w") as f:
        f.write("# %s\n\n" % title)
        f.write("Source: `ncu --metrics gpu__time_duration.sum --clock-control none` (per-launch times are cold-cache and "
                "serialised: compare SHARES, not absolutes).  %d launches, %.3f ms in total.\n\n" % (len(rows), total))
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f %% |\n" % (k, v[0], v[1], 100 * v[1] / total))
        f.write("\nLaunch order (last sequence):\n\n```\n")
        for r in rows:
            f.write("%-64s %10.3f ms\n" % (r[4][:64], float(r[-1]) / 1e6))
        f.write("```\n")


def multi(path, out, title, first_kernel, take=""):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    ix = {h: i for i, h in enumerate(rows[0])}
    data = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("kzg::", "").strip()
        data.setdefault((int(r[ix["ID"]]), name), {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    items = list(data.items())
    starts = [i for i, ((_, k), _m) in enumerate(items) if k.startswith(first_kernel)]
    seq = items[starts[0]:starts[0] + int(take)] if take else items[starts[-1]:]
    T, RD, WR, FM = "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", \
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"
    total = sum(m[T] for _, m in seq) / 1e6
    agg = collections.OrderedDict()
    for (_, k), m in seq:
        a = agg.setdefault(k, [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += m[T] / 1e6
        a[2] += m[RD]
        a[3] += m[WR]
        a[4] += m[FM] * m[T]
    with open(out, "w") as f:
        f.write("# %s\n\n" % title)
        f.write("Source: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active"
                " --clock-control none` (per-launch times are cold-cache and serialised: compare SHARES, not absolutes).  "
                "%d launches, %.3f ms in total.\n\n" % (len(seq), total))
        f.write("| kernel | launches | total ms | share | DRAM read GB | DRAM written GB | fmaheavy % (time-weighted) |\n|---|---:|---:|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f %% | %.2f | %.2f | %.1f |\n" % (k, v[0], v[1], 100 * v[1] / total, v[2] / 1e9, v[3] / 1e9,
                                                                          v[4] / (v[1] * 1e6) if v[1] else 0))
        f.write("\nLaunch order:\n\n```\n")
        for (_, k), m in seq:
            f.write("%-34s %9.3f ms  fmaheavy %5.1f %%  DRAM read %8.1f MB  written %8.1f MB\n" % (k[:34], m[T] / 1e6, m[FM], m[RD] / 1e6, m[WR] / 1e6))
        f.write("```\n")


def traffic(path, key, first_kernel, pattern, family, note="", take=""):
    import json
    import os
    import re
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from kzg_grandsums_study_b200 import build as b
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    ix = {h: i for i, h in enumerate(rows[0])}
    data = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("kzg::", "").strip()
        data.setdefault((int(r[ix["ID"]]), name), {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    items = list(data.items())
    starts = [i for i, ((_, k), _m) in enumerate(items) if k.startswith(first_kernel)]
    seq = items[starts[0]:starts[0] + int(take)] if take else items[starts[-1]:]   # take = N: the first N launches from the first start
    rx = re.compile(pattern)
    sel = [(k, m) for (_, k), m in seq if rx.search(k)]
    dram = sum(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"] for _, m in sel)
    macs = sum(m["sm__inst_executed_pipe_fmaheavy.sum"] * m["smsp__thread_inst_executed_per_inst_executed.ratio"] for _, m in sel)
    ms = sum(m["gpu__time_duration.sum"] for _, m in sel) / 1e6
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    table = json.load(open(out)) if os.path.exists(out) else {}
    table[key] = {"src_sha16": b.source_digest(b.MSM_SOURCES if family == "msm" else b.NTT_SOURCES),
                  "dram_bytes": dram, "fmaheavy_thread_instructions": macs, "kernel_ms_under_ncu": ms, "launches": len(sel),
                  "kernels": sorted(set(k for k, _ in sel)), "source_csv": os.path.basename(path), "note": note}
    json.dump(table, open(out, "w"), indent=1, sort_keys=True)
    print(key, json.dumps(table[key]))


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def full(path, out, title):
    p = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    rows = list(csv.reader(p.stdout.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# %s\n\nSource: `ncu --set full --clock-control none --import-source on` (`%s`).\n\n" % (title, path))
        for val in rows[2:]:
            name = val[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write("## `%s`\n\n| metric | unit | value |\n|---|---|---:|\n" % name.split("(")[0])
            for h, u, v in zip(hdr, units, val):
                if h in WANT:
                    f.write("| `%s` | %s | %s |\n" % (h, u, v))
            f.write("\n")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "traffic":
    traffic(*sys.argv[2:])
    sys.exit(0)
if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "multi":
    multi(*sys.argv[2:])
    sys.exit(0)
if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(*sys.argv[2:])
    else:
        full(*sys.argv[2:])
