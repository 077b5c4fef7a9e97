#!/usr/bin/env python
"""Turns the ncu captures of a gpurun call (gpurun_out/TAG_cells.ncu-rep from `ncu --set full`, TAG_launches.csv from the
`--metrics gpu__time_duration.sum` pass) into the tracked summaries under profiles/:
  profiles/ROUND_ncu_full_pa_rdf_cells_kernel.csv   every raw metric of the captured launches (ncu --page raw --csv)
  profiles/ROUND_launches_bench_1M.csv              the launch list
  profiles/ROUND_kernel_share.json                  time per kernel name over the launch list (share of the step)
  profiles/r2_traffic.json                          dram bytes per launch of the dominant kernel (read by bench.py)
usage: python tools/ncu_summary.py TAG ROUND"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

tag, rnd = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(root, "gpurun_out", f"{tag}_cells.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
out_csv = os.path.join(root, "profiles", f"{rnd}_ncu_full_pa_rdf_cells_kernel.csv")
open(out_csv, "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[0], rows[2:]
col = {n: i for i, n in enumerate(hdr)}


def num(r, name):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return None


for r in data:
    name = r[col["Kernel Name"]]
    if ", 1>" in name.split("(")[0]:            # PART 1 = the float32 launch
        unit_r, unit_w = rows[1][col["dram__bytes_read.sum"]], rows[1][col["dram__bytes_write.sum"]]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tr = num(r, "dram__bytes_read.sum") * scale[unit_r] + num(r, "dram__bytes_write.sum") * scale[unit_w]
        json.dump({"dram_bytes_per_launch": tr, "kernel": name.split("(")[0],
                   "source": f"profiles/{rnd}_ncu_full_pa_rdf_cells_kernel.csv: dram__bytes_read.sum + dram__bytes_write.sum of the captured "
                             "float32 launch (merged cross-type pass, 500k x 500k molecules) of `bench.py --steps 1 --warmup 3`",
                   "duration_ms_under_ncu": num(r, "gpu__time_duration.sum")},
                  open(os.path.join(root, "profiles", "r2_traffic.json"), "w"), indent=1)
        break
launch_src = os.path.join(root, "gpurun_out", f"{tag}_launches.csv")
if os.path.exists(launch_src):
    shutil.copy(launch_src, os.path.join(root, "profiles", f"{rnd}_launches_bench_1M.csv"))
    t = collections.defaultdict(float)
    n = collections.Counter()
    for r in csv.reader(open(launch_src)):
        if len(r) > 5 and r[-3] == "gpu__time_duration.sum":
            k = r[4].split("(")[0]
            v = float(r[-1].replace(",", ""))
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[-2], 1e-6)
            t[k] += v
            n[k] += 1
    total = sum(t.values())
    json.dump({"total_ms": total, "kernels": [{"name": k, "launches": n[k], "ms": v, "share": v / total} for k, v in sorted(t.items(), key=lambda kv: -kv[1])]},
              open(os.path.join(root, "profiles", f"{rnd}_kernel_share.json"), "w"), indent=1)
print("wrote", out_csv)
