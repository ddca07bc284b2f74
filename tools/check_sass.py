"""SASS lint for the packed fill kernels (no GPU needed): for every instantiation in libCUDAKernel.so find
the row loop (the innermost backward-branch loop that holds the DPX recurrence) and report its instruction
mix.  A loop whose LOP3 count explodes is ptxas parking the direction predicates in a register (DESIGN.md
4.4): that build is ~1.7x slower, so it fails the check.
usage: python tools/check_sass.py [path/to/libCUDAKernel.so]   (exit 1 on a bad loop)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_LIB = os.path.join(ROOT, "versalignlib_b200", "lib", "libCUDAKernel.so")
DPX = ("VIADDMNMX", "VIMNMX", "VIMNMX3")


def loops_of(lib: str, pattern: str = r"fill_(nw|fast)_kernel"):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            cur = funcs.setdefault(name, []) if re.search(pattern, name) else None
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if cur is not None and m:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    report = []
    for name, ins in funcs.items():
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
                    c[(f[1] if f[0].startswith("@") else f[0]).split(".")[0]] += 1
            dpx = sum(c[k] for k in DPX)
            # the row loop: the SMALLEST loop that still holds a whole row pair of the recurrence
            if dpx >= 60 and (best is None or sum(c.values()) < sum(best[1].values())):
                best = ((tgt, a), c)
        if best:
            report.append((name, best[0], best[1]))
    return report


def main() -> int:
    lib = sys.argv[1] if len(sys.argv) > 1 else DEFAULT_LIB
    bad = 0
    for name, (lo, hi), c in sorted(loops_of(lib)):
        short = re.sub(r".*(fill_(?:nw|fast)_kernelI[^E]*(?:E[A-Za-z0-9_]*?)?)EvNS.*", r"\1", name)[:60]
        total, lop3, dpx = sum(c.values()), c["LOP3"], sum(c[k] for k in DPX)
        flag = "BAD" if lop3 > dpx else "ok"  # a storm is 2 LOP3 per predicate = 4+ per cell-pair; mild cases pass
        bad += flag == "BAD"
        print(f"{flag:3s} {short:60s} loop 0x{lo:x}..0x{hi:x}: {total:4d} instr, {dpx:3d} DPX, {c['PRMT']:3d} PRMT, {c['FADD']:3d} FADD, {lop3:3d} LOP3")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
