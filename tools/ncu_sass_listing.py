"""Address-ordered SASS listing of one kernel from an .ncu-rep with the source line each instruction
is attributed to, executed counts, active threads and stall samples (incl. the no-instruction and
long-scoreboard reasons). Usage: python tools/ncu_sass_listing.py rep.ncu-rep [kernel-substring] > out.txt
A per-region summary (regions = runs of instructions between main-file lines given with
--marks file.cu:name:line,...) goes to stderr."""
import csv
import io
import re
import subprocess
import sys


def load(rep, kernel_sub=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    f, fn, hdr, cur_line, recs = None, None, None, None, {}
    for x in rows:
        if not x:
            continue
        if x[0] == "File Path":
            f = x[1].split("/")[-1]
            continue
        if x[0] == "Function Name":
            fn = x[1]
            continue
        if x[0] == "Line No":
            hdr = x
            continue
        if hdr is None or len(x) < 10:
            continue
        if kernel_sub and fn and kernel_sub not in fn:
            continue
        if x[0].isdigit():
            cur_line = int(x[0])
            continue
        if x[0] == "" and re.fullmatch(r"0x[0-9a-f]+", x[2].strip()):
            d = dict(zip(hdr[4:], x[4:]))

            def g(k):
                try:
                    return int(d.get(k, "0") or 0)
                except ValueError:
                    return 0

            recs[int(x[2], 16)] = dict(file=f, line=cur_line, sass=x[3].strip(), inst=g("Instructions Executed"), thr=g("Thread Instructions Executed"),
                                       smp=g("# Samples"), noinst=g("stall_no_inst"), longsb=g("stall_long_sb"), wait=g("stall_wait"),
                                       branch=g("stall_branch_resolving"), math=g("stall_math"), shortsb=g("stall_short_sb"))
    return recs


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else None
    recs = load(rep, sub)
    base = min(recs)
    T = sum(r["inst"] for r in recs.values()) or 1
    S = sum(r["smp"] for r in recs.values()) or 1
    print("# %d SASS instructions, %d warp instructions executed, %d samples" % (len(recs), T, S))
    print("# off  file:line  inst%  thr  smp%  noinst longsb wait branch | sass")
    for a in sorted(recs):
        r = recs[a]
        print("%5x %-22s %5.2f %4.1f %5.2f %5d %5d %5d %5d | %s" % (a - base, "%s:%s" % (r["file"], r["line"]), r["inst"] / T * 100, r["thr"] / max(r["inst"], 1),
                                                                   r["smp"] / S * 100, r["noinst"], r["longsb"], r["wait"], r["branch"], r["sass"]))


if __name__ == "__main__":
    main()
