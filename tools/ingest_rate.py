#!/usr/bin/env python
"""Measures the drop-in CLI's trajectory ingestion (threaded xyz text -> float32) on the box's host cores:
writes a synthetic xyz file (200k atoms x 11 frames, 68 MB), runs `prealpha_b200_cli` on a general.inp that only loads it,
and prints the CLI's own "reading the trajectory...done" line."""
import os
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, ".")
from prealpha_b200 import capi  # noqa: E402

n, nf, L = 200_000, 11, 158.74
wd = tempfile.mkdtemp(prefix="ingest_")
rng = np.random.default_rng(0)
with open(os.path.join(wd, "traj.xyz"), "w") as f:
    for s in range(nf):
        xyz = rng.random((n, 3)) * L
        f.write(f"{n}\nstep {s}\n")
        f.write("".join(f"Na {a:.5f} {b:.5f} {c:.5f}\n" for a, b, c in xyz))
open(os.path.join(wd, "molecular.inp"), "w").write(f" {nf}\n 1\n +1 1 {n}\nquit\n")
open(os.path.join(wd, "general.inp"), "w").write(' "traj.xyz"\n "molecular.inp"\n "./"\n "./"\n "./"\n cubic_box_edge 0.0 158.74\n sequential_read F\n quit\n')
cli = os.path.join(os.path.dirname(capi.LIB_PATH), "prealpha_b200_cli")
out = subprocess.run([cli, "general.inp"], cwd=wd, capture_output=True, timeout=600).stdout.decode()
print(f"host threads: {os.cpu_count()}; file: {os.path.getsize(os.path.join(wd, 'traj.xyz')) / 1e6:.0f} MB")
print("\n".join(line for line in out.splitlines() if "reading the trajectory" in line) or out[-1500:])
