"""Attribute the warp-stall samples of an ncu report (--page source --csv) to CUDA source lines.
Usage: python tools/ncu_lines.py <source.csv> <kernel-substring> [top]   (needs the built .so)"""
import csv
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    src_csv, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    tmp = "/tmp/sass_lines"
    os.makedirs(tmp, exist_ok=True)
    subprocess.run("cd %s && rm -f *.cubin && cuobjdump -xelf all %s/caro-ai_b200/libcaro_b200.so >/dev/null 2>&1" % (tmp, ROOT), shell=True)
    off2line = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kern not in out:
            continue
        cur, infn = None, False
        for line in out.split("\n"):
            if line.startswith(".text.") or line.strip().startswith(".section"):
                infn = kern in line
            m = re.search(r'//## File ".*?/([^/"]+)", line (\d+)', line)
            if m:
                cur = (m.group(1), int(m.group(2)))
            m = re.search(r"/\*([0-9a-f]{4,5})\*/", line)
            if m and infn and cur:
                off2line.setdefault(int(m.group(1), 16), cur)
    allrows = list(csv.reader(open(src_csv)))
    # the csv holds one section per profiled kernel: "Kernel Name",<name> / header / rows
    starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"]
    want = kern.split("INS_")[0].replace("_ZN4caro", "").lstrip("0123456789")
    sec = None
    for n, i in enumerate(starts):
        if want in allrows[i][1] or len(starts) == 1:
            sec = (i, starts[n + 1] if n + 1 < len(starts) else len(allrows))
            break
    rows = allrows[sec[0]:sec[1]]
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    base = int(data[0][idx["Address"]], 16)
    c, tot = Counter(), 0
    for r in data:
        s = int(r[idx["# Samples"]] or 0)
        c[off2line.get(int(r[idx["Address"]], 16) - base, ("?", 0))] += s
        tot += s
    cache = {}
    for (f, ln), s in c.most_common(top):
        text = ""
        for d in ("caro-ai_b200/csrc",):
            p = os.path.join(ROOT, d, f)
            if os.path.exists(p):
                cache.setdefault(p, open(p).read().split("\n"))
                text = cache[p][ln - 1].strip()[:100] if ln > 0 else ""
        print("%7d %5.1f%% %s:%d  %s" % (s, 100.0 * s / max(1, tot), f, ln, text))


if __name__ == "__main__":
    main()
