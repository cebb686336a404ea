"""Summarise an .ncu-rep (read here, no GPU needed) into a small markdown file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_name.md [kernel-substring]"""
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.per_cycle_active",
       "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
       "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
       "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed_op_shared_atom.sum",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "local_load_bytes", ]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True, check=True).stdout


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = ["# ncu summary of `%s`" % rep.split("/")[-1], "",
             "Captured with `ncu --set full --clock-control none --import-source on` (one launch, replayed passes);",
             "times under the profiler are not bench values.", ""]
    want = sys.argv[3] if len(sys.argv) > 3 else None
    seen = set()
    picked = []
    for r in data:  # one table per distinct kernel (the first launch of each), or only the ones matching argv[3]
        name = r[hdr.index("Kernel Name")]
        if name in seen or (want and want not in name):
            continue
        seen.add(name)
        picked.append(r)
    for r in picked:
        lines += ["## %s" % r[hdr.index("Kernel Name")], "", "| metric | value | unit |", "|---|---|---|"]
        for m in RAW:
            if m in hdr:
                lines.append("| `%s` | %s | %s |" % (m, r[hdr.index(m)], units[hdr.index(m)]))
        stalls = sorted(((num(r[i]) or 0.0, h) for i, h in enumerate(hdr)
                         if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")),
                        reverse=True)
        if stalls:
            lines += ["", "Warps per issue slot by state (`smsp__average_warps_issue_stalled_*_per_issue_active.ratio`, top 8):", ""]
            lines += ["* %s: %.2f" % (h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), v)
                      for v, h in stalls[:8]]
    # per-source-line view
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    recs, hdr2, fname = [], None, ""
    for r in src:
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr2 = r
        elif hdr2 and r[0].isdigit():
            try:
                recs.append((fname, int(r[0]), r[hdr2.index("Source")].strip(), int(r[hdr2.index("Instructions Executed")] or 0),
                             int(r[hdr2.index("# Samples")] or 0)))
            except ValueError:
                pass
    ti, ts = sum(x[3] for x in recs) or 1, sum(x[4] for x in recs) or 1
    lines += ["", "## Hottest source lines (share of executed warp instructions / of stall samples)", "",
              "| file:line | inst % | samples % | source |", "|---|---|---|---|"]
    for f, ln, s, i, smp in sorted(recs, key=lambda x: -(x[3] / ti + x[4] / ts))[:28]:
        lines.append("| %s:%d | %.1f | %.1f | `%s` |" % (f, ln, 100.0 * i / ti, 100.0 * smp / ts, s[:90].replace("|", "\\|")))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
