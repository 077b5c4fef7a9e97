#!/usr/bin/env python
"""Secondary measurements on the other BASELINE.json configurations (not the bench.py contract line):
  C3a  ionic-liquid box, 100k atoms: molecular RDFs (2500 + 2500 molecules of 25 / 15 atoms), 4 ordered type pairs
  C3b  same box, atom-resolved: 6 one-atom types, 36 ordered type pairs, 1e10 pair evaluations per frame
  C5   4M-atom box, 2-D distance-angle histogram 200x90, maxdist 40 A
One JSON line per configuration (host buffers in, histograms out: the end-to-end C-ABI call)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from prealpha_b200 import api, capi  # noqa: E402


def run(name, traj, jobs, note):
    api.histogram_multi(traj, jobs[:1], api.make_exec(frames_per_batch=1))          # context, module load
    t0 = time.perf_counter()
    api.histogram_multi(traj, jobs)                                                  # first call of this shape: the pools allocate staging
    cold = time.perf_counter() - t0
    t0 = time.perf_counter()
    h, w, o, st = api.histogram_multi(traj, jobs)
    dt = time.perf_counter() - t0
    print(json.dumps({"config": name, "note": note, "seconds_e2e": dt, "kernel_ms": st.kernel_ms, "frames": int(st.frames_done),
                      "pair_evals": int(st.pairs_evaluated), "pair_evals_per_s_e2e": st.pairs_evaluated / dt,
                      "binned": int(w.sum()), "binned_per_s_e2e": float(w.sum()) / dt, "launches": int(st.kernel_launches),
                      "h2d_bytes": int(st.h2d_bytes), "seconds_e2e_first_call_of_this_shape": cold}), flush=True)


def c3(frames_mol=200, frames_atom=20):
    rng = np.random.default_rng(1)
    L = round((100000 / 0.05) ** (1 / 3), 4)
    lo, hi = capi.cubic_box_edge(0.0, L)
    cat = ["C"] * 8 + ["H"] * 15 + ["N"] * 2
    ani = ["C"] * 2 + ["F"] * 6 + ["N"] + ["O"] * 4 + ["S"] * 2

    def mol_frames(nf, nm, na):
        c = rng.random((nf, nm, 1, 3)) * L
        return np.round(c + (rng.random((nf, nm, na, 3)) - 0.5) * 5.0, 5).astype(np.float32)
    tl = [dict(charge=-1, natoms=15, nmol=2500, elements=ani, xyz=mol_frames(frames_mol, 2500, 15)),
          dict(charge=+1, natoms=25, nmol=2500, elements=cat, xyz=mol_frames(frames_mol, 2500, 25))]
    traj = capi.Trajectory(tl, lo, hi)
    jobs = [api.job_setup(traj, capi.make_request(mode="pdf", ref_type=r, obs_type=o, bin_a=500, bin_b=1, maxdist_optimize=True,
                                                  sampling_interval=1)) for r in (1, 2) for o in (1, 2)]
    run("C3a_molecular_100k_atoms", traj, jobs, f"{frames_mol} frames, 4 type-pair RDFs, COM of 15/25-atom molecules on the GPU")
    counts = dict(C=25000, H=37500, N=7500, F=15000, O=10000, S=5000)
    tl = [dict(charge=0, natoms=1, nmol=n, elements=[e], xyz=np.round(rng.random((frames_atom, n, 1, 3)) * L, 5).astype(np.float32))
          for e, n in counts.items()]
    traj = capi.Trajectory(tl, lo, hi)
    jobs = [api.job_setup(traj, capi.make_request(mode="pdf", ref_type=r, obs_type=-1, bin_a=500, bin_b=1, maxdist_optimize=True,
                                                  sampling_interval=1)) for r in range(1, 7)]
    run("C3b_atom_resolved_100k_atoms", traj, jobs, f"{frames_atom} frames, 6 element types, all 36 ordered type pairs")


def c5(natoms=4_000_000, maxdist=40.0):
    rng = np.random.default_rng(2)
    L = 430.887
    lo, hi = capi.cubic_box_edge(0.0, L)
    half = natoms // 2
    tl = [dict(charge=+1, natoms=1, nmol=half, elements=["Na"], xyz=np.round(rng.random((1, half, 1, 3)) * L, 5).astype(np.float32)),
          dict(charge=-1, natoms=1, nmol=half, elements=["Cl"], xyz=np.round(rng.random((1, half, 1, 3)) * L, 5).astype(np.float32))]
    traj = capi.Trajectory(tl, lo, hi)
    jobs = [api.job_setup(traj, capi.make_request(mode="pdf", vec=(0, 0, 1), ref_type=1, obs_type=-1, bin_a=200, bin_b=90,
                                                  maxdist=maxdist, sampling_interval=1))]
    run(f"C5_4M_atoms_2D_maxdist{int(maxdist)}", traj, jobs, "1 frame, reference type 1 around all, 200x90 distance-angle bins")


def neigh(natoms=100_000, cutoff=3.0):
    """8f-4 primitive: all-atom neighbour pairs below `cutoff` in a 100k-atom box (what one frame of a GLOBAL cluster
    analysis tests, Cluster:1092-1100), host lists in, sorted pair list out."""
    rng = np.random.default_rng(3)
    L = round((natoms / 0.05) ** (1 / 3), 4)
    lo, hi = capi.cubic_box_edge(0.0, L)
    half = natoms // 2
    tl = [dict(charge=+1, natoms=1, nmol=half, elements=["Na"], xyz=np.round(rng.random((1, half, 1, 3)) * L, 5).astype(np.float32)),
          dict(charge=-1, natoms=1, nmol=half, elements=["Cl"], xyz=np.round(rng.random((1, half, 1, 3)) * L, 5).astype(np.float32))]
    traj = capi.Trajectory(tl, lo, hi)
    inst = np.array([(t, m, 1) for t in (1, 2) for m in range(1, half + 1)], dtype=np.int32)
    rq, keep = capi.make_neigh_request(1, a=inst, b=None, cutoff_sq=np.float32(cutoff) ** 2, same_list=True)
    api.neighbour_pairs(traj, rq, capacity=1 << 22)                      # warm-up
    t0 = time.perf_counter()
    pairs, tests = api.neighbour_pairs(traj, rq, capacity=1 << 22)
    dt = time.perf_counter() - t0
    print(json.dumps({"config": "neighbour_pairs_100k_atoms", "note": f"all-atom list against itself, cutoff {cutoff} A, sorted pair list on the host",
                      "seconds_e2e": dt, "distance_tests": int(tests), "tests_per_s_e2e": tests / dt, "pairs_found": int(len(pairs))}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c3", "c5"]
    if "c3" in which:
        c3()
    if "c5" in which:
        c5()
    if "neigh" in which:
        neigh()
    if "c5big" in which:
        c5(maxdist=140.0)
