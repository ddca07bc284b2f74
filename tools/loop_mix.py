"""Instruction mix of the hot loop of every kernel matching a pattern in an object / library (no GPU needed).
usage: python tools/loop_mix.py file.o|.so name_regex [min_dpx]   -- the smallest backward-branch loop holding >= min_dpx DPX ops"""
import collections
import re
import subprocess
import sys


def main():
    path, pat = sys.argv[1], sys.argv[2]
    min_dpx = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            cur = funcs.setdefault(name, []) if re.search(pat, name) else None
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if cur is not None and m:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    for name, ins in sorted(funcs.items()):
        short = re.search(pat + r"[A-Za-z0-9_]*", name).group(0)[:60]
        best = None
        for a, t in ins:
            if "BRA" not in t:
                continue
            m = re.search(r"0x([0-9a-f]+)\s*$", t)
            if not m or int(m.group(1), 16) >= a:
                continue
            tgt = int(m.group(1), 16)
            c = collections.Counter()
            for x, s in ins:
                if tgt <= x <= a:
                    f = s.split()
                    op = (f[1] if f[0].startswith("@") else f[0])
                    c["IMAD.MOV" if op.startswith("IMAD.MOV") else op.split(".")[0]] += 1
            dpx = c["VIADDMNMX"] + c["VIMNMX"] + c["VIMNMX3"]
            if dpx >= min_dpx and (best is None or sum(c.values()) < sum(best.values())):
                best = c
        if best:
            print(short, sum(best.values()), dict(best.most_common(14)))


if __name__ == "__main__":
    main()
