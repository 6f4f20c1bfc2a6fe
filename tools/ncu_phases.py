"""Attribute SASS instructions of a kernel to phases of its main source file: inlined helper code
(vrt_common.cuh, vrt_bsdf.cuh, ...) is charged to the phase whose code precedes it in address
order. Usage: python tools/ncu_phases.py rep.ncu-rep main_file.cu "name:first_line,name:first_line,..." """
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, main_file, spec = sys.argv[1], sys.argv[2], sys.argv[3]
    bounds = [(int(b.split(":")[1]), b.split(":")[0]) for b in spec.split(",")]
    bounds.sort()

    def phase(ln):
        name = bounds[0][1]
        for first, n in bounds:
            if ln >= first:
                name = n
        return name

    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    f, hdr, cur_line, recs = None, None, None, {}
    for x in rows:
        if not x:
            continue
        if x[0] == "File Path":
            f = x[1].split("/")[-1]
            continue
        if x[0] == "Line No":
            hdr = x
            continue
        if hdr is None or len(x) < 10:
            continue
        if x[0].isdigit():
            cur_line = int(x[0])
            continue
        if x[0] == "" and re.fullmatch(r"0x[0-9a-f]+", x[2].strip()):
            d = dict(zip(hdr[4:], x[4:]))
            try:
                recs[int(x[2], 16)] = (f, cur_line, int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["# Samples"]))
            except (KeyError, ValueError):
                pass
    tot, last = collections.OrderedDict(), bounds[0][1]
    for addr in sorted(recs):
        f, ln, inst, thr, smp = recs[addr]
        if f == main_file:
            last = phase(ln)
        t = tot.setdefault(last, [0, 0, 0])
        t[0] += inst
        t[1] += thr
        t[2] += smp
    T = sum(v[0] for v in tot.values()) or 1
    S = sum(v[2] for v in tot.values()) or 1
    print("| phase | warp instructions | avg active threads | stall samples |")
    print("|---|---|---|---|")
    for k, v in tot.items():
        print("| %s | %.1f %% | %.1f | %.1f %% |" % (k, v[0] / T * 100, v[1] / max(v[0], 1), v[2] / S * 100))


if __name__ == "__main__":
    main()
