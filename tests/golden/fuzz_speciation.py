#!/usr/bin/env python
"""Differential check of the `speciation` replay against the reference binary on random inputs (test infrastructure like
make_golden.py: needs oracle/_ref, i.e. /root/reference, so it is run by hand and not collected by pytest).  For every random speciation input on the first
12 frames of Research_Example_4 the binary writes its statistics / table files; the drop-in's bookkeeping runs on the
oracle's neighbour pairs through the test seam, and the files are compared byte for byte.
usage: python tests/golden/fuzz_speciation.py [ncases] [seed]"""
import ctypes as C
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle_lib as orc  # noqa: E402
from prealpha_b200 import api, capi  # noqa: E402
import fixtures  # noqa: E402
from test_speciation_neighbours import _parse_input, _types_in_order  # noqa: E402

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
orc.load()
assert orc.ref_available()
fx = fixtures.Fixture("re4")
traj = fx.trajectory()
natoms = {1: 5, 2: 6, 3: 1}


def random_input():
    def side(avoid=()):
        ents = []
        for _ in range(int(rng.integers(1, 5))):
            t = int(rng.choice([1, 2, 3]))
            a = int(rng.integers(1, natoms[t] + 1))
            if (t, a) not in ents and (t, a) not in avoid:
                ents.append((t, a))
        return ents or [(3, 1)]
    acc = side()
    don = side(avoid=acc) if rng.random() < 0.8 else side()
    text = f" {len(acc)} acceptors\n" + "".join(f" {t} {a} {rng.uniform(0.6, 2.4):.2f}\n" for t, a in acc)
    text += f" {len(don)} donors\n" + "".join(f" {t} {a} {rng.uniform(0.6, 2.4):.2f}\n" for t, a in don)
    for who, ents in (("acceptor", acc), ("donor", don)):
        for t in sorted({t for t, _ in ents}):
            mine = [a for tt, a in ents if tt == t]
            if len(mine) >= 2 and rng.random() < 0.6:
                for g in range(int(rng.integers(1, 3))):
                    text += f" new_{who}_group {t}\n"
                    for a in mine:
                        if rng.random() < 0.5:
                            text += f" assign_to_{who}_group {a}\n"
    text += f" N_species {int(rng.integers(2, 40))}\n N_neighbours {int(rng.integers(2, 14))}\n nsteps {int(rng.integers(2, 13))}\n"
    text += f" sampling_interval {int(rng.integers(1, 4))}\n" + (" use_ccc T\n" if rng.random() < 0.7 else "") + " quit\n"
    return text


def connections(ordinal, text):
    acc, don, nsteps, dt = _parse_input(text)
    rows = []
    don_types = _types_in_order(don)
    b, b_cls, b_meta = [], [], []
    for nd, t in enumerate(don_types, start=1):
        entries = [(e, at) for e, (t2, at, _) in enumerate(don) if t2 == t]
        for m in range(1, fx.types[t - 1][2] + 1):
            for k, (e, at) in enumerate(entries, start=1):
                b.append((t, m, at)); b_cls.append(e); b_meta.append((nd, m, k))
    b = np.array(b, dtype=np.int32); b_cls = np.array(b_cls, dtype=np.int32)
    cut = np.array([[(ra + rd) ** 2 for _, _, rd in don] for _, _, ra in acc], dtype=np.float32)
    for n_acc, ty in enumerate(_types_in_order(acc), start=1):
        entries = [(e, at) for e, (t2, at, _) in enumerate(acc) if t2 == ty]
        a, a_cls, a_meta = [], [], []
        for m in range(1, fx.types[ty - 1][2] + 1):
            for k, (e, at) in enumerate(entries, start=1):
                a.append((ty, m, at)); a_cls.append(e); a_meta.append((m, k))
        a = np.array(a, dtype=np.int32); a_cls = np.array(a_cls, dtype=np.int32)
        for step in range(1, min(nsteps, fx.nsteps) + 1, dt):
            rq, keep = capi.make_neigh_request(step, a=a, b=b, cutoff_sq=cut, a_class=a_cls, b_class=b_cls, skip_same_molecule=True)
            for ia, ib in orc.neighbour_pairs(traj, rq):
                rows.append((ordinal, step, n_acc, a_meta[ia][0], b_meta[ib][0], b_meta[ib][1], b_meta[ib][2], a_meta[ia][1]))
    return rows


lib = capi.load()
fn = lib.pa_debug_run_general_input_with_connections
fn.restype = C.c_int
fn.argtypes = [C.c_char_p, C.POINTER(capi.pa_exec), C.POINTER(C.c_int32), C.c_int64]
bad = 0
for case in range(ncases):
    text = random_input()
    dirs = []
    for who in ("ref", "mine"):
        wd = tempfile.mkdtemp(prefix=f"fuzz_{who}_")
        os.makedirs(os.path.join(wd, "output"))
        orc.write_xyz(os.path.join(wd, "traj.xyz"), fx.elements, fx.coords.astype(np.float64), fmt="%.9g")
        open(os.path.join(wd, "molecular.inp"), "w").write(f" {fx.nsteps}\n {len(fx.types)}\n" + "".join(f" {c:+d} {a} {m}\n" for c, a, m in fx.types) + "quit\n")
        open(os.path.join(wd, "spec.inp"), "w").write(text)
        open(os.path.join(wd, "general.inp"), "w").write(' "traj.xyz"\n "molecular.inp"\n "./"\n "./"\n "./output/"\n cubic_box_edge 0.0 34.7853\n parallel_operation F\n'
                                                          ' sequential_read F\n trajectory_type xyz\n set_prefix "f_"\n speciation "spec.inp"\n quit\n')
        dirs.append(wd)
    out, _ = orc.run_reference(dirs[0])
    rows = connections(0, text)
    table = np.ascontiguousarray(np.array(rows, dtype=np.int32).reshape(-1, 8))
    cwd = os.getcwd(); os.chdir(dirs[1])
    devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)
    try:
        rc = fn(b"general.inp", C.byref(api.make_exec()), table.ctypes.data_as(C.POINTER(C.c_int32)), len(table))
    finally:
        os.dup2(saved, 1); os.close(devnull); os.close(saved); os.chdir(cwd)
    ref_files = sorted(f for f in os.listdir(os.path.join(dirs[0], "output")) if "statistics" in f)
    my_files = sorted(f for f in os.listdir(os.path.join(dirs[1], "output")) if "statistics" in f)
    ok = ref_files == my_files and all(open(os.path.join(dirs[0], "output", f), "rb").read() == open(os.path.join(dirs[1], "output", f), "rb").read() for f in ref_files)
    if not ok:
        bad += 1
        keep = f"/tmp/fuzz_fail_{case}"
        shutil.rmtree(keep, ignore_errors=True); os.makedirs(keep)
        shutil.copytree(dirs[0], keep + "/ref"); shutil.copytree(dirs[1], keep + "/mine")
        open(keep + "/ref_stdout.txt", "w").write(out)
        print(f"case {case}: MISMATCH (kept in {keep}); files ref={ref_files} mine={my_files}\n{text}")
    else:
        print(f"case {case}: ok ({len(ref_files)} files, {len(rows)} connections)")
    for d in dirs:
        shutil.rmtree(d)
print("mismatches:", bad, "of", ncases)
