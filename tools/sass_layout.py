"""Static code layout of one kernel: for each 1 KB of SASS, which source lines it comes from.
Usage: python tools/sass_layout.py lib.so vrt_render k_pathILb0ELi0 ; reads the cubin with
nvdisasm --print-line-info (no GPU needed). Used to see where cold code sits inside the hot loop."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    lib, tu, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if tu in f and f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], capture_output=True, text=True).stdout
    cur_fn, cur_src, rows = None, None, []
    for ln in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln) or re.search(r"//-+ \.text\.(\S+)", ln)
        if m:
            cur_fn = m.group(1)
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_src = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(\S.*?);", ln)
        if m and cur_fn and kernel in cur_fn:
            rows.append((int(m.group(1), 16), cur_src, m.group(2)))
    print("%s: %d instructions, %.1f KB" % (kernel, len(rows), len(rows) * 16 / 1024))
    for base in range(0, (rows[-1][0] // 1024 + 1) * 1024, 1024):
        c = collections.Counter()
        for a, src, _ in rows:
            if base <= a < base + 1024 and src:
                c["%s:%d" % (src[0].replace("vrt_", "").replace(".cuh", "").replace(".cu", ""), src[1] // 10 * 10)] += 1
        print("%6x  %s" % (base, "  ".join("%s(%d)" % kv for kv in c.most_common(5))))


if __name__ == "__main__":
    main()
