"""DRAM traffic of cull_bake_kernel from one `ncu --set full` capture, recorded with the identity of the kernel source it was taken from:
    python tools/capture_traffic.py [objects]        (on the GPU box: gpurun -- python tools/capture_traffic.py)
writes profiles/ncu_cull_bake_traffic.json + profiles/r2_ncu_cull_bake_10M.txt.  bench.py reports `roofline.traffic` from this file only
when `kernel_source_sha16` still matches the sources in the tree — no literal lives in bench.py."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_sha  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    rep = os.path.join(ROOT, "gpurun_out", "cullbake_traffic.ncu-rep")
    os.makedirs(os.path.dirname(rep), exist_ok=True)
    env = dict(os.environ, R3_OBJECTS=str(n))
    cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", "regex:cull_bake_kernel", "-s", "2", "-c", "1", "-f", "-o", rep[:-8],
           sys.executable, os.path.join(ROOT, "tools", "profile_workloads.py"), "cullbake", "4"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if r.returncode != 0 or not os.path.exists(rep):
        sys.exit(f"ncu failed: {r.stdout[-1000:]} {r.stderr[-1000:]}")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]

    def metric(name):
        v, u = float(vals[hdr.index(name)].replace(",", "")), units[hdr.index(name)]
        return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1.0)
    cap = {"kernel": vals[hdr.index("Kernel Name")], "objects": n, "dram_bytes_read": metric("dram__bytes_read.sum"), "dram_bytes_write": metric("dram__bytes_write.sum"),
           "duration_s_under_ncu": metric("gpu__time_duration.sum"), "kernel_source_sha16": kernel_source_sha(),
           "command": "ncu --set full --clock-control none -k regex:cull_bake_kernel -s 2 -c 1 python tools/profile_workloads.py cullbake 4"}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(cap, open(os.path.join(ROOT, "profiles", "ncu_cull_bake_traffic.json"), "w"), indent=1)
    summary = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(ROOT, "profiles", "r2_ncu_cull_bake_10M.txt"), "w").write(summary)
    # gpurun only brings gpurun_out/ back: leave copies there too
    json.dump(cap, open(os.path.join(ROOT, "gpurun_out", "ncu_cull_bake_traffic.json"), "w"), indent=1)
    open(os.path.join(ROOT, "gpurun_out", "r2_ncu_cull_bake_10M.txt"), "w").write(summary)
    print(json.dumps(cap))


if __name__ == "__main__":
    main()
