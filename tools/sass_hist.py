"""SASS opcode histogram of one kernel of a built object: proves which tensor-core path a kernel uses (UTC*MMA / LDTM / STTM
= tcgen05 + TMEM, HMMA = legacy mma.sync). usage: python tools/sass_hist.py <object.o> <mangled-kernel-substring> [out.json]"""
import collections
import json
import re
import subprocess
import sys


def main(obj, kern, out=None):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    hist, cur, name = collections.Counter(), False, None
    for ln in sass.split("\n"):
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = kern in m.group(1)
            name = m.group(1) if cur else name
            continue
        if cur:
            m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", ln)
            if m:
                hist[m.group(1)] += 1
    total = sum(hist.values())
    keys = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "LDSM", "MOVM", "FHFMA", "LDG", "STS", "LDS", "SHFL", "BAR", "MUFU", "F2FP"]
    res = {"kernel": name, "sass_instructions": total, "tensor_core_path": {k: hist.get(k, 0) for k in keys},
           "top": dict(hist.most_common(25))}
    print(json.dumps(res, indent=1))
    if out:
        json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:])
