"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under
profiles/: the launch list, a per-kernel metric summary and traffic.json (DRAM bytes per launch of
the fill kernel, read by bench.py for roofline.traffic).
usage: python tools/make_profiles.py <round tag> <launches.csv> <full.ncu-rep>"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ncu_summary import WANT  # noqa: E402


def main():
    tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
    out_dir = os.path.join(ROOT, "profiles")
    os.makedirs(out_dir, exist_ok=True)
    # 1. launch list (strip ncu's ==PROF== chatter)
    rows = [r for r in csv.reader(l for l in open(launches) if not l.startswith("=="))]
    head = rows[0]
    ki, vi = head.index("Kernel Name"), head.index("Metric Value")
    with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
        f.write("launch,kernel,gpu__time_duration_ns\n")
        for n, r in enumerate(rows[1:]):
            f.write(f"{n},\"{r[ki][:120]}\",{r[vi]}\n")
    # shares over one resident step: the last complete run of meta..traceback before the e2e chunks
    # 2. full capture summary
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(out)))
    h, units = rr[0], rr[1]
    kk = h.index("Kernel Name")
    lines = [f"# ncu --set full summary ({tag})", "",
             "Command: `python bench.py --steps 2 --warmup 1 --no-cpu` under "
             "`ncu --set full --clock-control none --import-source on`; per-launch times are cold-cache and serialised.", ""]
    traffic = {}
    done = set()
    for row in rr[2:]:
        name = row[kk]
        lines.append(f"## {name[:110]}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---:|---|")
        vals = {}
        for m in WANT:
            if m in h:
                i = h.index(m)
                lines.append(f"| {m} | {row[i]} | {units[i]} |")
                vals[m] = (row[i], units[i])
        lines.append("")
        if ("fill_nw_kernel" in name or "fill_fast" in name) and "traffic" not in done and "dram__bytes_read.sum" in vals:
            done.add("traffic")
            def to_bytes(v, u):
                x = float(v.replace(",", ""))
                return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            traffic["fill_dram_bytes_per_launch"] = to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"])
            traffic["fill_kernel"] = name[:110]
            traffic["source"] = f"profiles/{tag}_ncu_summary.md"
    open(os.path.join(out_dir, f"{tag}_ncu_summary.md"), "w").write("\n".join(lines))
    if traffic:
        json.dump(traffic, open(os.path.join(out_dir, "traffic.json"), "w"), indent=1)
    print("wrote", os.listdir(out_dir))


if __name__ == "__main__":
    main()
