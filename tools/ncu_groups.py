"""Executed warp-instructions and stall samples of one kernel of an ncu report, grouped by source file / line ranges and
by opcode. usage: python tools/ncu_groups.py <report.ncu-rep> <object.o> <mangled-kernel-prefix> [file:lo-hi=name ...]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def load(rep, obj, kern):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][int(os.environ.get("NCU_TABLE", "0"))]
    hdr = rows[hi]
    ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = [(r[1], int(r[ie] or 0), int(r[isamp] or 0)) for r in rows[hi + 1:] if len(r) >= len(hdr)]
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
    body = next(s for s in dis.split(".text.") if s.startswith(kern) and "/*0000*/" in s)
    cur, lines = ("?", 0), []
    for ln in body.split("\n"):
        m = re.search(r'//## File "(.*?)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.search(r'/\*[0-9a-f]{4,5}\*/', ln):
            lines.append(cur)
    return data, lines


def main():
    rep, obj, kern = sys.argv[1:4]
    ranges = []
    for a in sys.argv[4:]:
        loc, name = a.split("=")
        f, r = loc.split(":")
        lo, hi = r.split("-")
        ranges.append((f, int(lo), int(hi), name))
    data, lines = load(rep, obj, kern)
    tot = sum(x[1] for x in data)
    tots = sum(x[2] for x in data)
    groups, samp, ops = collections.Counter(), collections.Counter(), collections.Counter()
    for (src, ex, sm), (f, l) in zip(data, lines):
        g = f
        for rf, lo, hi, name in ranges:
            if f == rf and lo <= l <= hi:
                g = name
        groups[g] += ex
        samp[g] += sm
        op = src.split()[1] if src.startswith("@") else src.split()[0]
        ops[(g, op.split(".")[0])] += ex
    print(f"total executed {tot / 1e9:.3f} G warp-instructions, {tots} samples")
    for g, v in groups.most_common():
        print(f"{g:28s} exec {100 * v / tot:6.2f}%  {v / 1e9:7.3f} G   samples {100 * samp[g] / max(1, tots):6.2f}%")
    print()
    for (g, op), v in sorted(ops.items(), key=lambda kv: -kv[1])[:int(os.environ.get("TOP", "40"))]:
        print(f"{g:24s} {op:12s} {100 * v / tot:5.2f}%")


if __name__ == "__main__":
    main()
