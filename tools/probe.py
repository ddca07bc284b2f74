"""Print host/GPU facts and the measured integer-pipe peaks (roofline denominators)."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from versalignlib_b200 import capi  # noqa: E402


def main():
    info = {"nproc": os.cpu_count()}
    try:
        info["cpu"] = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    try:
        info["nvidia_smi"] = subprocess.run(
            ["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit", "--format=csv,noheader"],
            capture_output=True, text=True).stdout.strip().splitlines()
    except Exception:
        pass
    with capi.CudaContext(devices=[0]) as ctx:
        names = {0: "viaddmax_s32", 1: "viaddmax_s16x2", 2: "viaddmax_s16x2_relu", 3: "vimax3_s16x2"}
        info["int_peak_lane_ops_per_s"] = {names[k]: ctx.int_peak(k) for k in names}
    print(json.dumps(info, indent=1))


if __name__ == "__main__":
    main()
