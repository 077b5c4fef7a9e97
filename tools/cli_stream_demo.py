#!/usr/bin/env python
"""End-to-end demonstration of the drop-in CLI on BASELINE's 1M-atom case as FILES: writes an xyz trajectory of
11 frames x 1M atoms (two one-atom types, same generator as bench.py; ~385 MB of text), a molecular.inp, a general.inp
with `sequential_read T|F` and the two-reference RDF input, runs `prealpha_b200_cli` and prints its timing lines and the
wall time: text file in, `_rdf` / `_numberintegral` files out."""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import pandas as pd

sys.path.insert(0, ".")
import bench  # noqa: E402
from prealpha_b200 import capi  # noqa: E402

n_half, L, nf = bench.N_PER_TYPE, float(bench.L_BOX), 11
wd = tempfile.mkdtemp(prefix="cli_stream_")
os.makedirs(os.path.join(wd, "out"))
t0 = time.time()
with open(os.path.join(wd, "traj.xyz"), "w") as f:
    els = np.array(["Na"] * n_half + ["Cl"] * n_half)
    for s in range(nf):
        xyz = bench.synth_frame(s, 2 * n_half, L)
        f.write(f"{2 * n_half}\nstep {s}\n")
        pd.DataFrame({"e": els, "x": xyz[:, 0], "y": xyz[:, 1], "z": xyz[:, 2]}).to_csv(f, sep=" ", header=False, index=False, float_format="%.5f")
print(f"wrote {os.path.getsize(os.path.join(wd, 'traj.xyz')) / 1e6:.0f} MB in {time.time() - t0:.0f} s", flush=True)
open(os.path.join(wd, "molecular.inp"), "w").write(f" {nf}\n 2\n +1 1 {n_half}\n -1 1 {n_half}\nquit\n")
open(os.path.join(wd, "rdf.inp"), "w").write("pdf 2\n+0 +0 +1 1 -1\n+0 +0 +1 2 -1\nmaxdist_optimize\nsampling_interval 1\nbin_count_a 1000\nbin_count_b 1\nquit\n")
cli = os.path.join(os.path.dirname(capi.LIB_PATH), "prealpha_b200_cli")
for seq in ("T", "F"):
    open(os.path.join(wd, "general.inp"), "w").write(
        f' "traj.xyz"\n "molecular.inp"\n "./"\n "./"\n "./out/"\n cubic_box_edge 0.0 {bench.L_BOX}\n sequential_read {seq}\n distribution "rdf.inp"\n quit\n')
    t0 = time.time()
    out = subprocess.run([cli, "general.inp"], cwd=wd, capture_output=True, timeout=1200).stdout.decode()
    wall = time.time() - t0
    keep = [ln for ln in out.splitlines() if "reading the trajectory" in ln or "B200 execution" in ln or "accelerated section" in ln or "within bin" in ln]
    print(f"sequential_read {seq}: wall {wall:.2f} s (text file -> output files, {nf} frames of {2 * n_half} atoms)")
    print("\n".join(keep), flush=True)
