"""Join the SASS page of an ncu report (per-instruction executed counts + stall samples) with nvdisasm line info.
usage: python tools/ncu_lines.py <report.ncu-rep> <object.o> <kernel-mangled-substring> [top]
Prints executed warp-instructions and stall samples aggregated per source file:line and per file."""
import csv, io, re, subprocess, sys, collections, os, tempfile

def main(rep, obj, kern, top=45):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # a report with several kernels prints one table per kernel: NCU_TABLE picks it (0-based, default the first)
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    which = int(os.environ.get("NCU_TABLE", "0"))
    hi = heads[which]
    rows = rows[: heads[which + 1]] if which + 1 < len(heads) else rows
    hdr = rows[hi]
    ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr): continue
        data.append((r[1], int(r[ie] or 0), int(r[isamp] or 0), {hdr[i]: int(r[i] or 0) for i in stall_cols}))
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
    sec = dis.split(".text." )
    body = next(s for s in sec if s.startswith(kern) and ":\n" in s[:400] and "/*0000*/" in s)
    cur, lines = ("?", 0), []
    for ln in body.split("\n"):
        m = re.search(r'//## File "(.*?)", line (\d+)', ln)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.search(r'/\*[0-9a-f]{4,5}\*/', ln): lines.append(cur)
    n = min(len(lines), len(data))
    print(f"sass instructions: report {len(data)}, disasm {len(lines)}")
    per_line, per_file = collections.defaultdict(lambda: [0, 0, collections.Counter()]), collections.defaultdict(lambda: [0, 0])
    for (src, ex, sm, st), loc in zip(data[:n], lines[:n]):
        per_line[loc][0] += ex; per_line[loc][1] += sm; per_line[loc][2].update(st)
        per_file[loc[0]][0] += ex; per_file[loc[0]][1] += sm
    tot_ex = sum(v[0] for v in per_line.values()); tot_s = sum(v[1] for v in per_line.values())
    print(f"total executed {tot_ex}, samples {tot_s}")
    for f, v in sorted(per_file.items(), key=lambda kv: -kv[1][1]):
        print(f"  {f:24s} exec {100*v[0]/tot_ex:5.1f}%  samples {100*v[1]/tot_s:5.1f}%")
    for loc, v in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
        st = ", ".join(f"{k[6:]}:{c}" for k, c in v[2].most_common(3))
        print(f"  {loc[0]}:{loc[1]:<4d} exec {100*v[0]/tot_ex:5.2f}%  samples {100*v[1]/tot_s:5.2f}%  [{st}]")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 45)
