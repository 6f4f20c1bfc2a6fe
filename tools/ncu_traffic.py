"""DRAM traffic of the path kernel from an `ncu --set full` report -> profiles/k_path_traffic.json, the figure bench.py
reports as roofline.traffic. The JSON records the launch shape and a digest of the kernel sources; bench.py reports
the figure only while both still match (a stale constant is worse than none).

    python tools/ncu_traffic.py gpurun_out/prof_r2j_path.ncu-rep --spp 8 --sky-res 3840 --sky-format f16
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--spp", type=int, default=8)
    ap.add_argument("--sky-res", type=int, default=3840)
    ap.add_argument("--sky-format", default="f16")
    ap.add_argument("--kernel", default="k_path")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    launches = [r for r in rows[2:] if a.kernel in r[hdr.index("Kernel Name")]]
    if not launches:
        raise SystemExit("no %s launch in %s" % (a.kernel, a.report))

    def val(r, name):
        v, u = float(r[hdr.index(name)]), units[hdr.index(name)].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]

    rd = sum(val(r, "dram__bytes_read.sum") for r in launches) / len(launches)
    wr = sum(val(r, "dram__bytes_write.sum") for r in launches) / len(launches)
    from bench import kernel_source_digest

    j = {"kernel": launches[0][hdr.index("Kernel Name")], "report": os.path.basename(a.report), "launches_averaged": len(launches),
         "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr, "dram_bytes_per_launch": rd + wr, "spp": a.spp,
         "sky_res": a.sky_res, "sky_format": a.sky_format, "kernel_source_digest": kernel_source_digest(),
         "note": "ncu --set full --clock-control none, 1920x1080, dense random 256^3 (bench.py workload config3)"}
    with open(os.path.join(ROOT, "profiles", "k_path_traffic.json"), "w") as f:
        json.dump(j, f, indent=1)
    print(json.dumps(j))


if __name__ == "__main__":
    main()
