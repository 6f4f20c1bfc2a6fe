"""Summarise an .ncu-rep here (no GPU needed): key raw metrics + per-file / per-line aggregation of
the source page. Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel_index] [--md out.md]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__sass_average_branch_targets_threads_uniform.pct', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers', 'lts__t_sectors.sum',
        'smsp__sass_branch_targets_threads_divergent.sum', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'local_load_bytes', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    kidx = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
    md = sys.argv[sys.argv.index("--md") + 1] if "--md" in sys.argv else None
    out = []
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, r = rows[0], rows[1], rows[2 + kidx]
    out.append("## raw metrics (%s)" % r[hdr.index("Kernel Name")])
    out.append("")
    out.append("| metric | value | unit |")
    out.append("|---|---|---|")
    for k in KEYS:
        if k in hdr:
            out.append("| %s | %s | %s |" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    stall = [(h, float(r[i])) for i, h in enumerate(hdr) if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct")]
    for h, v in sorted(stall, key=lambda x: -x[1])[:8]:
        out.append("| %s | %.2f | %% |" % (h, v))
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    agg, cur_file, h2, kern, kcount = collections.OrderedDict(), None, None, None, -1
    for x in rows:
        if not x:
            continue
        if x[0] == "File Path":
            cur_file = x[1].split("/")[-1]
            continue
        if x[0] == "Function Name":
            continue
        if x[0] == "Kernel Name":
            continue
        if x[0] == "Line No":
            h2 = x
            continue
        if h2 is None:
            continue
        try:
            ln = int(x[0])
        except ValueError:
            continue
        if x[2] != "-":
            continue
        d = dict(zip(h2[4:], x[4:]))
        a = agg.setdefault((cur_file, ln, x[1][:100]), [0, 0, 0])
        a[0] += int(d["Instructions Executed"])
        a[1] += int(d["Thread Instructions Executed"])
        a[2] += int(d["# Samples"])
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[2] for a in agg.values()) or 1
    pf = collections.defaultdict(lambda: [0, 0, 0])
    for (f, ln, src), a in agg.items():
        for i in range(3):
            pf[f][i] += a[i]
    out += ["", "## per file (all captured launches of the report)", "", "| file | warp instructions | avg active threads | stall samples |", "|---|---|---|---|"]
    for f, a in sorted(pf.items(), key=lambda kv: -kv[1][0]):
        out.append("| %s | %.1f %% | %.1f | %.1f %% |" % (f, a[0] / tot * 100, a[1] / max(a[0], 1), a[2] / tots * 100))
    out += ["", "## top source lines by stall samples", "", "| file:line | inst %% | avg threads | samples %% | source |".replace("%%", "%"), "|---|---|---|---|---|"]
    for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:45]:
        out.append("| %s:%d | %.2f | %.1f | %.2f | `%s` |" % (f, ln, a[0] / tot * 100, a[1] / max(a[0], 1), a[2] / tots * 100, src.replace("|", "/")))
    txt = "\n".join(out) + "\n"
    if md:
        open(md, "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
